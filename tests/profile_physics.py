"""Profiling driver (not a test): random-ctrl rollout, last launch is the one to capture.
    ncu --set full --clock-control none --import-source on -k regex:ss_physics -s 3 -c 1 -o gpurun_out/phys python tests/profile_physics.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
raw = blob.read_bytes(os.path.join(bench.GOLDEN_DIR, os.environ.get("BLOB", "stretch_empty_floor.ssm")))
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = int(os.environ.get("NENV", 4096)); nsteps = int(os.environ.get("NSTEPS", 10))
B = engine.Batch(dm, nenv, maxcon=int(os.environ.get("MAXCON", 32)), maxefc=int(os.environ.get("MAXEFC", 0)))
dev = B.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
for p in range(3):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)); B.step(50)
B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, 3, lo, hi, dev)); B.step(nsteps)
torch.cuda.synchronize()
print("iters mean", B.solver_iter.float().mean().item(), "ncon mean", B.ncon.float().mean().item(), "max", B.ncon.max().item(), "flags", B.env_flags.max().item())
it = B.solver_iter.cpu().numpy(); print("iter hist", np.bincount(it, minlength=12)[:16].tolist())
g = it[: (len(it) // 7) * 7].reshape(-1, 7); print("mean", it.mean(), "mean of max-of-7 (env order)", g.max(1).mean())
