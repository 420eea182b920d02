"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): physics with contacts + lidar + render."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
raw = blob.read_bytes(os.path.join(os.path.dirname(bench.GOLDEN), "stretch_default_scene_render.ssm.z"))
rawE = open(bench.GOLDEN, "rb").read()
A, _ = blob.unpack(rawE)
for r, nenv, nsteps in ((rawE, 24, int(os.environ.get("NSTEPS", 6))), (raw, 5, 3)):
    dm = engine.DeviceModel(r, 0)
    B = engine.Batch(dm, nenv, maxcon=32 if r is rawE else 72, maxefc=0 if r is rawE else 288)
    if r is rawE:
        lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device="cuda"); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device="cuda")
        B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, 0, lo, hi, "cuda"))
    else:
        B.reset(key=0)
    B.step(nsteps); B.forward()
    torch.cuda.synchronize()
    if r is raw:
        d = B.lidar(); cam = dm.name2id(engine.OBJ_CAMERA, "d405_rgb")
        rgb = torch.zeros(nenv, 27, 48, 3, dtype=torch.uint8, device="cuda"); dep = torch.zeros(nenv, 27, 48, device="cuda")
        B.render(cam, 48, 27, 58.0, rgb, dep); torch.cuda.synchronize()
        rgb2 = torch.zeros(nenv, 120, 160, 3, dtype=torch.uint8, device="cuda"); dep2 = torch.zeros(nenv, 120, 160, device="cuda")
        B.render(dm.name2id(engine.OBJ_CAMERA, "d435i_camera_rgb"), 160, 120, 42.0, rgb2, dep2, 10.0); torch.cuda.synchronize()   # raster path incl. large boxes
    print("ok", nenv, float(B.qpos.abs().sum()))
