"""Diagnostic: two batches fed the same ctrl stream must stay bit-identical (scheduling must not leak into results)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
raw = open(bench.GOLDEN, "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = 4096
B1, B2 = engine.Batch(dm, nenv), engine.Batch(dm, nenv)
dev = B1.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
nper = int(os.environ.get("NPER", 24))
for p in range(nper):
    c = bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)
    B1.ctrl.copy_(c); B2.ctrl.copy_(c)
    B1.step(50); B2.step(50)
    torch.cuda.synchronize()
    bad = (B1.qpos != B2.qpos).any(1) | (B1.qvel != B2.qvel).any(1)
    if bad.any():
        ids = torch.nonzero(bad).flatten().cpu().numpy()
        e = int(ids[0])
        print(f"period {p}: {len(ids)} envs differ, first env {e}: max |dqpos| {float((B1.qpos[e]-B2.qpos[e]).abs().max()):.3e} ncon {int(B1.ncon[e])}/{int(B2.ncon[e])} iters {int(B1.solver_iter[e])}/{int(B2.solver_iter[e])} flags {int(B1.env_flags[e])}/{int(B2.env_flags[e])}")
        print("  differing envs:", ids[:20])
        break
else:
    print(f"bit-identical over {nper} periods x 50 steps x {nenv} envs")
import hashlib
print("final state sha1", hashlib.sha1(B1.qpos.cpu().numpy().tobytes() + B1.qvel.cpu().numpy().tobytes()).hexdigest()[:16], "checksum", float(B1.qpos.abs().sum()))
