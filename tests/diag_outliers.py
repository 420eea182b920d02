"""Which qacc outliers of the random-rollout forward parity test are explained by contact differences? (diagnostic, GPU box)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
from oracle.oracle import OracleModel
raw = open(bench.GOLDEN, "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0); om = OracleModel(raw); om.set_options(enable_lidar=False)
nenv = 1024
B = engine.Batch(dm, nenv, debug=True)
dev = B.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
for p in range(4):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)); B.step(50)
B.step(7); torch.cuda.synchronize()
f = lambda t: t.cpu().numpy().astype(np.float64)
q, v, w, c = f(B.qpos), f(B.qvel), f(B.qacc_warmstart), f(B.ctrl)
B.forward(); torch.cuda.synchronize()
o = om.forward(q, v, c, w, maxcon=B.maxcon, want=("qacc", "ncon", "contact_geom", "contact_dist", "contact_frame", "contact_pos", "nefc", "flags", "solver_iter", "qacc_smooth"))
qa = f(B.qacc)
err = np.abs(qa - o["qacc"]).max(1) / (np.abs(o["qacc"]).max(1) + 1e-3)
live = np.arange(B.maxcon)[None, :] < o["ncon"][:, None]
ndiff = np.where(live, np.abs(f(B.dbg["contact_normal"]) - o["contact_frame"]).max(2), 0.0).max(1)
ddist = np.where(live, np.abs(f(B.contact_dist) - o["contact_dist"]), 0.0).max(1)
dpos = np.where(live, np.abs(f(B.dbg["contact_pos"]) - o["contact_pos"]).max(2), 0.0).max(1)
es = np.abs(f(B.dbg["qacc_smooth"]) - o["qacc_smooth"]).max(1) / (np.abs(o["qacc_smooth"]).max(1) + 1e-3)
print("median err", np.median(err), "p99", np.quantile(err, 0.99))
for e in np.argsort(-err)[:16]:
    print(f"env {e:4d} err {err[e]:.2e} ndiff {ndiff[e]:.2e} ddist {ddist[e]:.2e} dpos {dpos[e]:.2e} smooth_err {es[e]:.1e} ncon {o['ncon'][e]} nefc {o['nefc'][e]} iters dev {int(B.solver_iter[e])} oracle {o['solver_iter'][e]} maxacc {np.abs(o['qacc'][e]).max():.1f}")
