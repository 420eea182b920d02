"""Scratch probe: calibrate parity tolerances from the settled home pose."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from stretch_mujoco_b200 import engine, blob
from oracle.oracle import OracleModel
raw = open(os.path.join(os.path.dirname(__file__), "golden", "stretch_empty_floor.ssm"), "rb").read()
A, _ = blob.unpack(raw)
om = OracleModel(raw); om.set_options(enable_lidar=False)
dm = engine.DeviceModel(raw, 0)
nenv = 64
rng = np.random.default_rng(0)
qpos = A["qpos0"][None].copy(); qvel = np.zeros((1, om.nv)); warm = np.zeros((1, om.nv)); t = np.zeros(1)
home = A["key_ctrl"][0][None].copy()
om.step(qpos, qvel, home, warm, t, nsteps=1500)
qpos = np.tile(qpos, (nenv, 1)); qvel = np.tile(qvel, (nenv, 1)); warm = np.tile(warm, (nenv, 1)); t = np.zeros(nenv)
lo, hi = A["actuator_ctrlrange"][:, 0], A["actuator_ctrlrange"][:, 1]
# moderate targets around home: lift 0.3..1.0, arm 0..0.4, wrist within limits, head anywhere, wheels +-3
ctrl = np.tile(home, (nenv, 1))
ctrl[:, 0:2] = rng.uniform(-3, 3, (nenv, 2)); ctrl[:, 2] = rng.uniform(0.3, 1.0, nenv); ctrl[:, 3] = rng.uniform(0, 0.4, nenv)
ctrl[:, 4] = rng.uniform(-1, 3, nenv); ctrl[:, 5] = rng.uniform(-1, 0.5, nenv); ctrl[:, 6] = rng.uniform(-2, 2, nenv)
ctrl[:, 7] = rng.uniform(-0.02, 0.04, nenv); ctrl[:, 8] = rng.uniform(-3, 1.5, nenv); ctrl[:, 9] = rng.uniform(-1.4, 0.7, nenv)
ctrl[0] = home[0]
B = engine.Batch(dm, nenv, debug=True)
B.qpos.copy_(torch.tensor(qpos, dtype=torch.float32)); B.qvel.copy_(torch.tensor(qvel, dtype=torch.float32))
B.qacc_warmstart.copy_(torch.tensor(warm, dtype=torch.float32)); B.ctrl.copy_(torch.tensor(ctrl, dtype=torch.float32))
qpos = B.qpos.cpu().numpy().astype(np.float64); qvel = B.qvel.cpu().numpy().astype(np.float64); warm = B.qacc_warmstart.cpu().numpy().astype(np.float64)
ctrl = B.ctrl.cpu().numpy().astype(np.float64)
def relrow(a, b): return np.abs(a - b).max(axis=1) / (np.abs(b).max(axis=1) + 1e-9)
B.forward(); torch.cuda.synchronize()
o = om.forward(qpos, qvel, ctrl, warm, want=("qacc", "ncon", "contact_geom", "solver_iter"), maxcon=24)
r = relrow(B.qacc.cpu().numpy(), o["qacc"])
print("forward qacc rel: med %.2e max %.2e" % (np.median(r), r.max()), "pairs equal", np.array_equal(B.contact_geom.cpu().numpy(), o["contact_geom"]))
scale_q = np.maximum(np.abs(qpos).max(), 1.0)
for k in range(10):
    B.step(100); torch.cuda.synchronize()
    oo = om.step(qpos, qvel, ctrl, warm, t, nsteps=100, want=("ncon", "contact_geom", "solver_iter"), maxcon=24)
    gq = B.qpos.cpu().numpy(); gv = B.qvel.cpu().numpy()
    eq = np.abs(gq - qpos); ev = np.abs(gv - qvel)
    req = (eq / np.maximum(np.abs(qpos), 1e-2)).max(axis=1)
    worst = int(np.argmax(eq.max(axis=1)))
    print(f"step {100*(k+1)} |dq| med {np.median(eq.max(axis=1)):.2e} max {eq.max():.2e} (env {worst} idx {int(np.argmax(eq[worst]))}) rel med {np.median(req):.2e} max {req.max():.2e} | "
          f"|dv| med {np.median(ev.max(axis=1)):.2e} max {ev.max():.2e} | pairs eq {np.array_equal(B.contact_geom.cpu().numpy(), oo['contact_geom'])} "
          f"ncon {B.ncon.cpu().numpy().max()} iters gpu {B.solver_iter.cpu().numpy().max()} cpu {oo['solver_iter'].max()}")
# timing at home steady state vs random
for n in (64, 1024, 4096):
    Bn = engine.Batch(dm, n)
    Bn.reset(key=0); Bn.step(100); torch.cuda.synchronize()
    t0 = time.time(); Bn.step(200); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"home steady: {n} envs x 200: {dt*1e3:.1f} ms -> {n*200/dt:.0f} env-steps/s; iters {Bn.solver_iter.max().item()}")
