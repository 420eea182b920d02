import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from stretch_mujoco_b200 import engine, blob
raw = blob.read_bytes(os.path.join(os.path.dirname(bench.GOLDEN), "stretch_kitchen_proxy_render.ssm.z"))
dm = engine.DeviceModel(raw, 0)
B = engine.Batch(dm, 6, maxcon=32)
B.reset(key=0); B.step(4); B.forward(); d = B.lidar()
cam = dm.name2id(engine.OBJ_CAMERA, "d435i_camera_rgb")
rgb = torch.zeros(6, 40, 30, 3, dtype=torch.uint8, device="cuda"); dep = torch.zeros(6, 40, 30, device="cuda")
B.render(cam, 30, 40, 42.0, rgb, dep, 10.0, rot90=-1, bgr=True)
torch.cuda.synchronize(); print("ok", float(B.qpos.abs().sum()), float(d.mean()))
