"""Which capacity do the envs of the bench workload run into, and when?  (diagnostic, GPU box)
Per control period of bench.py's cfg2 ctrl stream: max / p99 of ncon, nefc and broadphase candidates, overflow flags."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
raw = blob.read_bytes(os.path.join(bench.GOLDEN_DIR, os.environ.get("BLOB", "stretch_empty_floor.ssm")))
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = 4096
B = engine.Batch(dm, nenv, maxcon=int(os.environ.get("MAXCON", 96)), maxefc=int(os.environ.get("MAXEFC", 400)), debug=True)
dev = B.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
mc = torch.zeros(nenv, device=dev); me = torch.zeros(nenv, device=dev)
for p in range(int(os.environ.get("PERIODS", 23))):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev))
    pc = torch.zeros(nenv, device=dev); pe = torch.zeros(nenv, device=dev)
    for s in range(50):
        B.step(1)
        pc = torch.maximum(pc, B.ncon.float()); pe = torch.maximum(pe, B.dbg["nefc"].float())
    mc = torch.maximum(mc, pc); me = torch.maximum(me, pe)
    print(f"period {p:2d}: ncon max {int(pc.max())} p99.9 {torch.quantile(pc, 0.999):.0f} p99 {torch.quantile(pc, 0.99):.0f} mean {pc.mean():.1f} | nefc max {int(pe.max())} p99.9 {torch.quantile(pe, 0.999):.0f} p99 {torch.quantile(pe, 0.99):.0f} | overflow flags {int((B.env_flags & 2).ne(0).sum())} resets {int((B.env_flags & 1).ne(0).sum())}")
ns = int(me.max()); print("rollout max ncon", int(mc.max()), "nefc", ns, "| envs with ncon > 32:", int((mc > 32).sum()), " nefc-contact rows > 96:", "n/a")
for thr in (100, 120, 140, 160, 200): print(f"envs whose nefc ever exceeded {thr}: {int((me > thr).sum())}")
for thr in (24, 32, 40, 48, 64): print(f"envs whose ncon ever exceeded {thr}: {int((mc > thr).sum())}")
