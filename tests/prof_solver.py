"""Per-stage cycle budget of the physics kernel.  Needs the profiling build:
    make -C stretch_mujoco_b200/csrc clean && make -C stretch_mujoco_b200/csrc PROFILE=1 && python tests/prof_solver.py
(rebuild without PROFILE afterwards; the counters cost 1-2 %)."""
import os, sys, ctypes as C
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
torch.cuda.synchronize()
L = engine.lib()
buf = (C.c_ulonglong * 32)()
if L.ss_debug_prof(buf, 1) != 0:
    raise SystemExit("libstretchsim.so was built without PROFILE=1")
B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, 3, lo, hi, dev))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); B.step(50); e1.record(); torch.cuda.synchronize()
L.ss_debug_prof(buf, 0)
v = np.array(list(buf), dtype=np.float64)
names = ["A mul_JT+grad+checks", "B hessian build + barrier + chol/solve (alive)", "C gs/symv/mul_J", "D line search probes", "E update+eval write+cost",
         "barrier wait in hessian_solve (working warps)", "barrier wait (done warps)", "done warps: loop overhead"]
tot_ms = e0.elapsed_time(e1)
warp_cycles = tot_ms * 1e-3 * 1.965e9 * 148 * 7
print(f"50 steps: {tot_ms:.1f} ms; total warp-cycles available {warp_cycles:.3e}")
for n, x in zip(names, v[:8]):
    print(f"{x / warp_cycles * 100:6.2f}% of warp time  {n}")
print(f"{v[:8].sum() / warp_cycles * 100:6.2f}% solver loop total (B includes the barrier rows)")

stage = {16: "group fetch + state load + stage barrier before kinematics", 17: "kinematics", 18: "barrier + CRB / mass matrix", 19: "stage barrier before collision",
         20: "collision (broadphase + narrowphase)", 21: "stage barrier before velocity stage", 22: "velocity / RNE", 23: "smooth forces + actuation",
         24: "constraint rows", 25: "barrier + M factor / qacc_smooth", 26: "Newton solver (incl. its barriers)", 27: "observations + IMU + end-of-step barrier",
         28: "integrate (incl. its barrier)", 29: "state store"}
print()
for k, n in stage.items():
    print(f"{v[k] / warp_cycles * 100:6.2f}% of warp time  {n}")
print(f"{v[16:30].sum() / warp_cycles * 100:6.2f}% accounted (the rest: SMs without a resident CTA between / at the end of launches, model-pack load)")
print(f"{v[30] / warp_cycles * 100:6.2f}% of warp time  narrowphase (inside collision); broadphase = collision - this")
