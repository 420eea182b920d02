import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from stretch_mujoco_b200 import engine, blob
raw = blob.read_bytes(os.path.join(os.path.dirname(bench.GOLDEN), "stretch_default_scene_render.ssm.z"))
dm = engine.DeviceModel(raw, 0)
rawE = open(bench.GOLDEN, "rb").read()
dmE = engine.DeviceModel(rawE, 0)
for model, nm in ((dmE, "E"), (dm, "D")):
    for nenv in (1, 7, 8, 15, 1183, 1184, 1185, 5000):
        if nm == "D" and nenv > 1200: continue
        B = engine.Batch(model, nenv)
        B.reset(key=0)
        for n in (1, 2, 3, 7):
            B.step(n)
        torch.cuda.synchronize()
        ok = bool(torch.isfinite(B.qpos).all()) and abs(float(B.time[0]) - 13 * 0.002) < 1e-6 and abs(float(B.time[-1]) - 13 * 0.002) < 1e-6
        extra = ""
        if nm == "D" and nenv <= 16:
            d = B.lidar(); torch.cuda.synchronize(); extra = f" lidar ok {bool(torch.isfinite(d).all())}"
            cam = model.name2id(engine.OBJ_CAMERA, "nav_camera_rgb")
            rgb = torch.zeros(nenv, 37, 51, 3, dtype=torch.uint8, device="cuda"); dep = torch.zeros(nenv, 37, 51, device="cuda")
            B.render(cam, 51, 37, 90.0, rgb, dep); torch.cuda.synchronize(); extra += f" render mean depth {float(dep.mean()):.3f}"
        print(nm, nenv, "ok" if ok else "FAIL", "flags", int(B.env_flags.max()), extra)
        del B
print("done")
