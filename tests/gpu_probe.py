"""Scratch probe (not a test): GPU kernel vs oracle on the empty-floor scene."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from stretch_mujoco_b200 import engine
from oracle.oracle import OracleModel

blob = open(os.path.join(os.path.dirname(__file__), "golden", "stretch_empty_floor.ssm"), "rb").read()
om = OracleModel(blob)
dm = engine.DeviceModel(blob, 0)
nenv = int(os.environ.get("NENV", 64))
B = engine.Batch(dm, nenv, debug=True)
torch.cuda.synchronize()
print("smem/env", "ok; nq", dm.nq, "nv", dm.nv)
rng = np.random.default_rng(0)
lo = dm.get("actuator_ctrlrange").reshape(-1, 2)
ctrl = rng.uniform(lo[:, 0], lo[:, 1], size=(nenv, dm.nu))
ctrl[0] = dm.get("key_ctrl").reshape(-1, dm.nu)[0]
B.ctrl.copy_(torch.tensor(ctrl, dtype=torch.float32))
qpos = B.qpos.cpu().numpy().astype(np.float64); qvel = np.zeros((nenv, dm.nv)); warm = np.zeros((nenv, dm.nv))
ctrl64 = B.ctrl.cpu().numpy().astype(np.float64)
# forward parity
B.forward(); torch.cuda.synchronize()
o = om.forward(qpos, qvel, ctrl64, warm, want=("M", "qacc", "qacc_smooth", "ncon", "nefc", "contact_geom", "contact_dist", "solver_iter", "qfrc_constraint"), maxcon=24)
def rel(a, b): return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
print("M rel", rel(B.dbg["M"].cpu().numpy(), o["M"]))
print("qacc_smooth rel", rel(B.dbg["qacc_smooth"].cpu().numpy(), o["qacc_smooth"]))
print("ncon", B.ncon.cpu().numpy()[:8], o["ncon"][:8], "nefc", B.dbg["nefc"].cpu().numpy()[:8], o["nefc"][:8])
print("geoms eq", np.array_equal(B.contact_geom.cpu().numpy(), o["contact_geom"]))
print("dist err", np.abs(B.contact_dist.cpu().numpy() - o["contact_dist"]).max())
print("iters", B.solver_iter.cpu().numpy()[:8], o["solver_iter"][:8])
print("qfrc_con rel", rel(B.dbg["qfrc_constraint"].cpu().numpy(), o["qfrc_constraint"]))
print("qacc rel", rel(B.qacc.cpu().numpy(), o["qacc"]), "per-env max", [round(rel(B.qacc[i].cpu().numpy(), o["qacc"][i]), 6) for i in range(4)])
# rollout parity
t = np.zeros(nenv)
for k in range(10):
    B.step(100); torch.cuda.synchronize()
    oo = om.step(qpos, qvel, ctrl64, warm, t, nsteps=100, want=("ncon", "flags", "solver_iter"))
    gq = B.qpos.cpu().numpy(); gv = B.qvel.cpu().numpy()
    eq = np.abs(gq - qpos).max(axis=1); ev = np.abs(gv - qvel).max(axis=1)
    print(f"step {100*(k+1)} qpos err med {np.median(eq):.2e} max {eq.max():.2e} | qvel err med {np.median(ev):.2e} max {ev.max():.2e} | "
          f"ncon gpu {B.ncon.cpu().numpy()[:6]} cpu {oo['ncon'][:6]} flags {B.env_flags.cpu().numpy().max()} {oo['flags'].max()} iters {B.solver_iter.cpu().numpy().max()} {oo['solver_iter'].max()}")
# timing
for n in (nenv,):
    torch.cuda.synchronize(); t0 = time.time(); B.step(200); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{n} envs x 200 steps: {dt*1e3:.1f} ms -> {n*200/dt:.0f} env-steps/s")
