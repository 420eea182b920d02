"""BASELINE config 1 (1 env, default scene.xml, `home` keyframe then idle, no cameras): single-env step rate of the device
path and of the CPU restatement on one core, next to the reference's published figures (the
reference caps itself at 500 steps/s real time; 55-59 steps/s with 5 cameras)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import blob, engine
from oracle.oracle import OracleModel
raw = open(os.path.join(os.path.dirname(bench.GOLDEN), "stretch_default_scene.ssm"), "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
B = engine.Batch(dm, 1, maxcon=32)
B.reset(key=0); B.step(200); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); B.step(1000); e1.record(); torch.cuda.synchronize()
ms_fused = e0.elapsed_time(e1)
t0 = time.perf_counter()
for _ in range(1000):
    B.step(1)
torch.cuda.synchronize()
ms_calls = (time.perf_counter() - t0) * 1e3
om = OracleModel(raw); om.set_options(enable_lidar=False)
q = A["key_qpos"][0][None].copy() if "key_qpos" in A else A["qpos0"][None].copy()
v = np.zeros((1, om.nv)); w = np.zeros((1, om.nv)); t = np.zeros(1); c = A["key_ctrl"][0][None].copy()
om.step(q, v, c, w, t, nsteps=200, nthreads=1)
t0 = time.perf_counter(); om.step(q, v, c, w, t, nsteps=1000, nthreads=1); cpu_s = time.perf_counter() - t0
print(json.dumps({"workload": "cfg1: 1 env, default scene, home keyframe, no sensors", "device_steps_per_s_one_call_of_1000": 1000 / (ms_fused * 1e-3),
                  "device_steps_per_s_1000_calls_of_1": 1000 / (ms_calls * 1e-3), "oracle_1_core_steps_per_s": 1000 / cpu_s,
                  "reference_published": "<= 500 steps/s (real-time cap), 55-59 steps/s with 5 cameras (SURVEY.md §6)",
                  "real_time_factor_device": 0.002 * 1000 / (ms_fused * 1e-3)}))
