"""Tuning driver (not a test): times the cfg2 random-ctrl workload for the current SS_WPB / SS_SYNC."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
raw = blob.read_bytes(os.path.join(bench.GOLDEN_DIR, os.environ.get("BLOB", "stretch_empty_floor.ssm")))
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = int(os.environ.get("NENV", 4096))
B = engine.Batch(dm, nenv, maxcon=int(os.environ.get("MAXCON", 32)), maxefc=int(os.environ.get("MAXEFC", 0)), debug=bool(os.environ.get("DEBUG")))
dev = B.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
for p in range(3):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)); B.step(50)
ms = []
for p in range(3, 9):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); B.step(50); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print(f"ncon max {int(B.ncon.max())} mean {float(B.ncon.float().mean()):.2f} p99 {float(torch.quantile(B.ncon.float(), 0.99)):.0f} | flags2={int((B.env_flags & 2).ne(0).sum())} SS_WPB={os.environ.get('SS_WPB','-')} SS_SYNC={os.environ.get('SS_SYNC','-')}: {np.mean(ms):.1f} ms/50 steps -> {nenv*50/np.mean(ms)*1e3:.0f} env-steps/s; checksum {B.qpos.abs().sum().item():.3f}")
if os.environ.get("DEBUG"):
    B.forward(); torch.cuda.synchronize()
    ne = B.dbg["nefc"].float(); nc = B.ncon.float()
    print(f"nefc mean {ne.mean():.1f} p50 {ne.median():.0f} p99 {torch.quantile(ne, 0.99):.0f} max {ne.max():.0f} | ncon mean {nc.mean():.1f} p99 {torch.quantile(nc, 0.99):.0f} max {nc.max():.0f} | iters mean {B.solver_iter.float().mean():.2f} max {int(B.solver_iter.max())}")
