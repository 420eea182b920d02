"""Diagnostic (not a test): Newton iteration counts of the fp32 device solver vs the fp64 oracle on the same states."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
from oracle.oracle import OracleModel
raw = open(bench.GOLDEN, "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = 2048
B = engine.Batch(dm, nenv)
dev = B.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
for p in range(4):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)); B.step(50)
B.step(7)
torch.cuda.synchronize()
q, v, w, c = (t.cpu().numpy().astype(np.float64) for t in (B.qpos, B.qvel, B.qacc_warmstart, B.ctrl))
B.forward(); torch.cuda.synchronize()
git = B.solver_iter.cpu().numpy()
om = OracleModel(raw); om.set_options(enable_lidar=False)
o = om.forward(q, v, c, w, want=("solver_iter", "nefc", "qacc"))
oit = o["solver_iter"]
print("gpu  iter hist", np.bincount(git, minlength=10)[:12].tolist(), "mean", git.mean())
print("orcl iter hist", np.bincount(oit, minlength=10)[:12].tolist(), "mean", oit.mean())
print("gpu - oracle hist", np.bincount(np.clip(git - oit + 6, 0, 12), minlength=13).tolist())
qa = B.qacc.cpu().numpy()
err = np.abs(qa - o["qacc"]).max(1) / (np.abs(o["qacc"]).max(1) + 1e-3)
print("qacc rel err: median %.2e p99 %.2e max %.2e" % (np.median(err), np.quantile(err, 0.99), err.max()))
print("frac err>1e-3: %.4f  >1e-2: %.4f" % ((err > 1e-3).mean(), (err > 1e-2).mean()))
onefc = o["nefc"]
o2 = om.forward(q, v, c, w, want=("ncon", "contact_geom", "contact_dist"))
gn = B.ncon.cpu().numpy()
print("ncon mismatch envs:", int((gn != o2["ncon"]).sum()))
bad = np.argsort(-err)[:12]
for e in bad:
    print(f"env {e}: err {err[e]:.2e} iters gpu {git[e]} orcl {oit[e]} ncon gpu {gn[e]} orcl {o2['ncon'][e]} nefc {onefc[e]} |qacc| {np.abs(o['qacc'][e]).max():.2f}")
# same solve but with the device forced to iterate longer (tolerance 0 -> only the fp32-noise stops apply)
om.set_options(max_iter=200, tolerance=1e-14, enable_lidar=False)
t = om.forward(q, v, c, w, want=("solver_iter", "qacc"))
den = np.abs(t["qacc"]).max(1) + 1e-3
eg = np.abs(qa - t["qacc"]).max(1) / den; eo = np.abs(o["qacc"] - t["qacc"]).max(1) / den
print("vs tight fp64 solve: gpu median %.2e p99 %.2e max %.2e | oracle(default tol) median %.2e p99 %.2e max %.2e" % (
    np.median(eg), np.quantile(eg, 0.99), eg.max(), np.median(eo), np.quantile(eo, 0.99), eo.max()))
for e in bad[:8]:
    print(f"env {e}: gpu-vs-true {eg[e]:.2e} oracle-vs-true {eo[e]:.2e} true iters {t['solver_iter'][e]}")

names = None
for e in bad[:5]:
    d = np.abs(qa[e] - t["qacc"][e]); i = int(d.argmax())
    print(f"env {e}: worst dof {i} gpu {qa[e, i]:.3f} true {t['qacc'][e, i]:.3f}; top dofs by err {np.argsort(-d)[:5].tolist()} errs {np.sort(d)[::-1][:5].round(2).tolist()}")
