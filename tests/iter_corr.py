"""Diagnostic: how predictable is an env's Newton iteration count from its previous step(s)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
raw = open(bench.GOLDEN, "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = 4096
B = engine.Batch(dm, nenv)
dev = B.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
for p in range(3):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)); B.step(50)
B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, 3, lo, hi, dev))
its = []
for s in range(50):
    B.step(1); its.append(B.solver_iter.cpu().numpy().copy())
its = np.array(its)          # [step, env]
print("mean iters per step (first 10):", its[:10].mean(1).round(2), "last 5:", its[-5:].mean(1).round(2))
c1 = np.corrcoef(its[:-1].ravel(), its[1:].ravel())[0, 1]
print("corr(step t, t+1) = %.3f" % c1, " corr(sum over first 25, sum over last 25) = %.3f" % np.corrcoef(its[:25].sum(0), its[25:].sum(0))[0, 1])
def mm(order, step):
    x = its[step][order][: (nenv // 7) * 7].reshape(-1, 7); return x.max(1).mean()
ident = np.arange(nenv)
print("mean of max-of-7, env order: %.2f" % np.mean([mm(ident, s) for s in range(25, 50)]))
print("mean of max-of-7, sorted by sum of first 25 steps: %.2f" % np.mean([mm(np.argsort(-its[:25].sum(0), kind='stable'), s) for s in range(25, 50)]))
print("mean of max-of-7, sorted by previous step: %.2f" % np.mean([mm(np.argsort(-its[s - 1], kind='stable'), s) for s in range(25, 50)]))
print("mean iters: %.2f" % its[25:].mean())
