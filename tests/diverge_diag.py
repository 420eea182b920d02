"""Diagnostic: where and when does the bench workload's device trajectory leave the oracle's?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
from oracle.oracle import OracleModel
np.set_printoptions(linewidth=220, precision=2)
raw = open(bench.GOLDEN, "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
om = OracleModel(raw); om.set_options(enable_lidar=False)
nenv, nv, nq = 64, dm.nv, dm.nq
lo, hi = A["actuator_ctrlrange"][:, 0].copy(), A["actuator_ctrlrange"][:, 1].copy()
f64 = lambda t: t.cpu().numpy().astype(np.float64)
B = engine.Batch(dm, nenv, maxcon=32)
q, v, w = f64(B.qpos), f64(B.qvel), f64(B.qacc_warmstart); t = np.zeros(nenv)
c = bench.ctrl_np(0, 0, nenv, 0, lo, hi)
B.ctrl.copy_(torch.tensor(c, dtype=torch.float32)); cc = f64(B.ctrl)
for step in range(1, 101):
    B.step(1)
    o = om.step(q, v, cc, w, t, nsteps=1, want=("contact_geom", "ncon", "qacc", "solver_iter"), maxcon=32)
    if step in (1, 2, 3, 5, 10, 20, 35, 50, 75, 100):
        dq = np.abs(f64(B.qpos) - q); dv = np.abs(f64(B.qvel) - v)
        da = np.abs(f64(B.qacc) - o["qacc"]).max(1) / (np.abs(o["qacc"]).max(1) + 1e-3)
        eq = (B.contact_geom.cpu().numpy() == o["contact_geom"]).all((1, 2))
        worst = int(dq.max(1).argmax())
        print(f"step {step:3d}: median max|dqpos| {np.median(dq.max(1)):.1e} max {dq.max():.1e} (env {worst}, qpos idx {int(dq[worst].argmax())}) | median max|dqvel| {np.median(dv.max(1)):.1e} "
              f"| qacc rel err median {np.median(da):.1e} max {da.max():.1e} | contact lists equal {int(eq.sum())}/64 | iters dev {B.solver_iter.float().mean():.2f} orcl {o['solver_iter'].mean():.2f}")
dq = np.abs(f64(B.qpos) - q)
print("per-qpos-index median |dqpos| at step 100:", np.median(dq, 0))
print("per-qpos-index max    |dqpos| at step 100:", dq.max(0))
