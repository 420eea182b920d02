"""Sensor-path timing driver (not a test, not the headline bench): lidar (S2/L1) and pinhole
RGB+depth camera (C3/C4) kernels on the BASELINE config-3 shapes, timed with CUDA events.

    python tests/bench_sensors.py [--nenv 4096] [--chunk 256] [--scene empty|default] [--reps 5]

Prints one JSON line per kernel with the roofline fields of SURVEY.md 8(d):
  lidar : 4*nray + 1.2 kB algorithmic bytes per env
  render: W*H*7 algorithmic bytes per env-frame (uint8 RGB + f32 depth)
The robot poses come from a short random-ctrl rollout so that the envs differ.
"""
import argparse, json, os, sys
import numpy as np, torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import blob, engine

ap = argparse.ArgumentParser()
ap.add_argument("--nenv", type=int, default=4096)
ap.add_argument("--chunk", type=int, default=256)
ap.add_argument("--scene", default="empty")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--W", type=int, default=640)
ap.add_argument("--H", type=int, default=480)
ap.add_argument("--cams", default="d435i_camera_rgb")
ap.add_argument("--skip-lidar", action="store_true")
args = ap.parse_args()

name = {"empty": "stretch_empty_floor_render.ssm.z", "default": "stretch_default_scene_render.ssm.z",
        "kitchen": "stretch_kitchen_proxy_render.ssm.z"}[args.scene]
raw = blob.read_bytes(os.path.join(os.path.dirname(bench.GOLDEN), name))
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = args.nenv
B = engine.Batch(dm, nenv, maxcon=32)
dev = B.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev)
hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
for p in range(4):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)); B.step(50)
torch.cuda.synchronize()
peak = 6538.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(bench.GOLDEN), "..", "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    ms = []
    for k in range(reps):
        flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), float(np.min(ms))


if dm.nrange > 0 and not args.skip_lidar:
    out = torch.empty(nenv, dm.nrange, device=dev)
    med, best = timed(lambda: B.lidar(out), args.reps)
    algo = nenv * (4 * dm.nrange + 1200)
    hit = float((out >= 0).float().mean())
    print(json.dumps({"kernel": "lidar_kernel (+ray_prepare)", "scene": args.scene, "nenv": nenv, "nray": dm.nrange, "ms": med, "ms_best": best,
                      "rays_per_s": nenv * dm.nrange / (med * 1e-3), "env_scans_per_s": nenv / (med * 1e-3), "hit_frac": hit,
                      "roofline": {"bound": "hbm", "achieved": algo / (med * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": algo / (med * 1e-3) / 1e9 / peak}}), flush=True)

W, H, chunk = args.W, args.H, min(args.chunk, nenv)
rgb = torch.empty(chunk, H, W, 3, dtype=torch.uint8, device=dev)
depth = torch.empty(chunk, H, W, dtype=torch.float32, device=dev)
for cname in args.cams.split(","):
    cam = dm.name2id(engine.OBJ_CAMERA, cname)
    fovy = float(A["cam_fovy"][cam])

    def frame_all():
        for e0 in range(0, nenv, chunk):
            B.render(cam, W, H, fovy, rgb, depth, 10.0, env_begin=e0, env_count=min(chunk, nenv - e0))
    med, best = timed(frame_all, args.reps)
    algo = nenv * W * H * 7
    cover = float((depth > 0).float().mean())
    print(json.dumps({"kernel": "render_kernel (+ray_prepare)", "scene": args.scene, "camera": cname, "nenv": nenv, "W": W, "H": H,
                      "chunk": chunk, "ms": med, "ms_best": best, "env_frames_per_s": nenv / (med * 1e-3),
                      "rays_per_s": nenv * W * H / (med * 1e-3), "depth_cover_last_chunk": cover,
                      "roofline": {"bound": "hbm", "achieved": algo / (med * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": algo / (med * 1e-3) / 1e9 / peak}}), flush=True)

# BASELINE config 3: physics step + head RGB+depth render of every env, every mj_step
cam = dm.name2id(engine.OBJ_CAMERA, "d435i_camera_rgb")
fovy = float(A["cam_fovy"][cam])


def cfg3_step():
    B.step(1)
    for e0 in range(0, nenv, chunk):
        B.render(cam, W, H, fovy, rgb, depth, 10.0, env_begin=e0, env_count=min(chunk, nenv - e0))
med, best = timed(cfg3_step, args.reps)
print(json.dumps({"workload": "cfg3: physics step + %dx%d head RGB+depth render of every env each mj_step" % (W, H), "scene": args.scene,
                  "nenv": nenv, "ms_per_step": med, "env_steps_per_s": nenv / (med * 1e-3),
                  "roofline": {"bound": "hbm", "achieved": nenv * (W * H * 7 + 828) / (med * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": nenv * (W * H * 7 + 828) / (med * 1e-3) / 1e9 / peak}}), flush=True)

if args.scene == "kitchen":
    # BASELINE config 4 (kitchen proxy): physics step + 1000-ray lidar + head 424x240 + wrist 480x270 RGB+depth each mj_step
    cams = [(dm.name2id(engine.OBJ_CAMERA, "d435i_camera_rgb"), 424, 240, 42.0, 10.0), (dm.name2id(engine.OBJ_CAMERA, "d405_rgb"), 480, 270, 58.0, 1.0)]
    bufs = [(torch.empty(chunk, h, w, 3, dtype=torch.uint8, device=dev), torch.empty(chunk, h, w, device=dev)) for _, w, h, _, _ in cams]
    scan = torch.empty(nenv, dm.nrange, device=dev)

    def cfg4_step():
        B.step(1)
        B.lidar(scan)
        for (cid, w, h, fv, lim), (c, d) in zip(cams, bufs):
            for e0 in range(0, nenv, chunk):
                B.render(cid, w, h, fv, c, d, lim, env_begin=e0, env_count=min(chunk, nenv - e0))
    med, best = timed(cfg4_step, args.reps)
    algo = nenv * (828 + 4 * dm.nrange + 1200 + 7 * (424 * 240 + 480 * 270))
    print(json.dumps({"workload": "cfg4 (kitchen proxy): physics step + %d-ray lidar + head 424x240 + wrist 480x270 RGB+depth each mj_step" % dm.nrange,
                      "nenv": nenv, "ms_per_step": med, "env_steps_per_s": nenv / (med * 1e-3),
                      "roofline": {"bound": "hbm", "achieved": algo / (med * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": algo / (med * 1e-3) / 1e9 / peak}}), flush=True)
