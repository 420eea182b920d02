#!/usr/bin/env python
"""Headline benchmark: env-steps/sec of parallel Stretch envs on N B200s, next to the CPU path timed on
the host cores.  The default run is BASELINE.json config 2 (4096 envs, empty-floor scene, physics only).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config cfg2|default|cfg3|cfg4|cfg5] [--scaling weak|strong] [--nenv 4096]

Workloads; every env gets a fresh uniform-random ctrl row per control period from a
counter-based RNG keyed (seed, env, period, actuator):
  cfg2     empty floor, physics only; one bench step = one control period of 50 mj_steps
  default  the reference's default scene.xml (dock, table, two free objects; nv = 44), physics only, 50 mj_steps
  cfg3     cfg2 + head RGB+depth render at 640x480 after every mj_step; one bench step = 5 mj_steps
  cfg4     kitchen proxy + 1000-ray lidar + head 424x240 + wrist 480x270 RGB+depth after every mj_step; 5 mj_steps
  cfg5     cfg4 + nav camera 800x600 RGB (all five cameras), 4096 envs per GPU (32768 on 8 GPUs)
`--scaling weak` (default): --nenv envs PER GPU; `--scaling strong`: --nenv envs in TOTAL, split over the GPUs.

`--impl reference` times the CPU restatement of the reference's path (oracle/, all host threads) on a bounded
sample of the same workload -- the real mujoco==3.2.6 wheel cannot be installed in this image (no wheel in
/opt/wheelhouse, no network).  If `import mujoco` ever succeeds on the box, the reference arm additionally
times `mujoco.mj_step` on one MjData per thread and reports that as `kind: "reference"`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN = os.path.join(GOLDEN_DIR, "stretch_empty_floor.ssm")
MJ_STEPS_PER_STEP = 50
ALGO_BYTES_PER_ENV_STEP = 828  # 89 floats read + 118 floats written
METRIC = "env-steps/sec (4096 parallel Stretch envs) at 1/2/4/8 B200 vs reference CPU"

# camera tuple: (MJCF camera, W, H, fovy, depth limit, rgb?, depth?)
HEAD640 = ("d435i_camera_rgb", 640, 480, 42.0, 10.0, True, True)
HEAD = ("d435i_camera_rgb", 424, 240, 42.0, 10.0, True, True)
WRIST = ("d405_rgb", 480, 270, 58.0, 1.0, True, True)
NAV = ("nav_camera_rgb", 800, 600, 102.0, 0.0, True, False)
CONFIGS = {
    "cfg2": dict(blob="stretch_empty_floor.ssm", mj_steps=50, cams=[], lidar=False, maxcon=32,
                 workload="cfg2: {n} parallel Stretch envs per GPU, empty-floor scene, physics only (no sensors)"),
    "default": dict(blob="stretch_default_scene.ssm", mj_steps=50, cams=[], lidar=False, maxcon=72, maxefc=320,
                    workload="default scene.xml (dock, table, two free objects; nv = 44): {n} envs per GPU, physics only"),
    "cfg3": dict(blob="stretch_empty_floor_render.ssm.z", mj_steps=5, cams=[HEAD640], lidar=False, maxcon=32,
                 workload="cfg3: {n} envs per GPU, empty floor, physics + 640x480 head RGB+depth render after every mj_step"),
    "cfg4": dict(blob="stretch_kitchen_proxy_render.ssm.z", mj_steps=5, cams=[HEAD, WRIST], lidar=True, maxcon=40,
                 workload="cfg4: {n} envs per GPU, kitchen proxy (Robocasa assets are download-only), physics + 1000-ray lidar "
                          "+ head 424x240 + wrist 480x270 RGB+depth after every mj_step"),
    "cfg5": dict(blob="stretch_kitchen_proxy_render.ssm.z", mj_steps=5, cams=[HEAD, WRIST, NAV], lidar=True, maxcon=40,
                 workload="cfg5: {n} envs per GPU, kitchen proxy, physics + 1000-ray lidar + all five cameras "
                          "(head / wrist RGB+depth, nav 800x600 RGB) after every mj_step"),
}


# ----------------------------------------------------------------------------- counter-based ctrl stream
def _mix_np(x):
    M = np.uint64(0xFFFFFFFF)
    x = x & M
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & M
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & M
    x ^= x >> np.uint64(16)
    return x


def ctrl_np(seed, env0, nenv, period, lo, hi):
    env = np.arange(env0, env0 + nenv, dtype=np.uint64)[:, None]
    act = np.arange(len(lo), dtype=np.uint64)[None, :]
    key = (np.uint64(seed) * np.uint64(0x9E3779B1) + env * np.uint64(0x85EBCA77) + np.uint64(period) * np.uint64(0xC2B2AE3D)
           + act * np.uint64(0x27D4EB2F))
    u = _mix_np(_mix_np(key)).astype(np.float64) / 4294967296.0
    return (lo[None, :] + u * (hi - lo)[None, :])


def ctrl_torch(seed, env0, nenv, period, lo, hi, device):
    import torch
    M = 0xFFFFFFFF

    def mix(x):
        x = x & M
        x = x ^ (x >> 16); x = (x * 0x7FEB352D) & M
        x = x ^ (x >> 15); x = (x * 0x846CA68B) & M
        x = x ^ (x >> 16)
        return x
    env = torch.arange(env0, env0 + nenv, dtype=torch.int64, device=device)[:, None]
    act = torch.arange(lo.numel(), dtype=torch.int64, device=device)[None, :]
    key = (seed * 0x9E3779B1 + env * 0x85EBCA77 + period * 0xC2B2AE3D + act * 0x27D4EB2F) & 0xFFFFFFFFFFFF
    u = mix(mix(key)).to(torch.float64) / 4294967296.0
    return (lo[None, :] + u * (hi - lo)[None, :]).to(torch.float32)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) > 2 + k and r[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def _load_blob(cfg):
    from stretch_mujoco_b200 import blob
    raw = blob.read_bytes(os.path.join(GOLDEN_DIR, cfg["blob"]))
    A, names = blob.unpack(raw)
    return raw, A, names


def try_mujoco_baseline(nenv, nsteps_mj, periods):
    """if the real wheel is ever importable, time mujoco.mj_step on one MjData per host thread
    (the binding releases the GIL) on the reference's own MJCF.  Returns None when mujoco or the MJCF is absent."""
    try:
        import mujoco  # noqa: F401
    except Exception:
        return None
    from concurrent.futures import ThreadPoolExecutor
    from stretch_mujoco_b200 import scenes
    d = scenes.models_dir()
    if d is None:
        return None
    cwd = os.getcwd()
    try:
        os.chdir(d)                       # <include file="stretch.xml"/> and its assetdir resolve against the models directory
        xml = scenes.EMPTY_FLOOR_XML
        model = mujoco.MjModel.from_xml_string(xml)
    except Exception:
        return None
    finally:
        os.chdir(cwd)
    cores = os.cpu_count() or 1
    datas = [mujoco.MjData(model) for _ in range(nenv)]
    lo, hi = model.actuator_ctrlrange[:, 0].copy(), model.actuator_ctrlrange[:, 1].copy()

    def run(chunk):
        for d in chunk:
            for _ in range(nsteps_mj):
                mujoco.mj_step(model, d)
    rates = []
    with ThreadPoolExecutor(cores) as ex:
        for p in range(periods):
            c = ctrl_np(0, 0, nenv, p, lo, hi)
            for e, d in enumerate(datas):
                d.ctrl[:] = c[e]
            t0 = time.perf_counter()
            list(ex.map(run, [datas[k::cores] for k in range(cores)]))
            rates.append(nenv * nsteps_mj / (time.perf_counter() - t0))
    return np.array(rates), cores


def cpu_sample(cfg, nenv, periods, threads=0):
    """Times the CPU path (oracle) on `nenv` envs for `periods` bench steps of config `cfg`; env-steps/s per period."""
    from oracle.oracle import OracleModel, lib
    raw, A, names = _load_blob(cfg)
    om = OracleModel(raw)
    om.set_options(enable_lidar=cfg["lidar"])          # the oracle evaluates <rangefinder> sensors inside its step, like mj_step
    lo, hi = A["actuator_ctrlrange"][:, 0].copy(), A["actuator_ctrlrange"][:, 1].copy()
    qpos = np.tile(A["qpos0"], (nenv, 1)); qvel = np.zeros((nenv, om.nv)); warm = np.zeros((nenv, om.nv)); t = np.zeros(nenv)
    cores = threads or lib().om_max_threads()
    cam_names = names[4]
    times = []
    for p in range(periods):
        c = ctrl_np(0, 0, nenv, p, lo, hi)
        t0 = time.perf_counter()
        if not cfg["cams"]:
            om.step(qpos, qvel, c, warm, t, nsteps=cfg["mj_steps"], nthreads=cores)
        else:
            for _ in range(cfg["mj_steps"]):
                o = om.step(qpos, qvel, c, warm, t, nsteps=1, nthreads=cores, want=("xpos", "xquat"))
                for (cname, W, H, fovy, lim, _, _) in cfg["cams"]:
                    om.render(o["xpos"], o["xquat"], cam_names.index(cname), W, H, fovy, nthreads=cores)
        times.append(time.perf_counter() - t0)
    return nenv * cfg["mj_steps"] / np.array(times), cores


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    nenv = args.cpu_nenv if not cfg["cams"] else min(args.cpu_nenv, 16)
    kind, note = "port", "CPU restatement of the mujoco==3.2.6 mj_step path (oracle/), not the MuJoCo binary"
    got = try_mujoco_baseline(nenv, cfg["mj_steps"], args.warmup + args.steps) if args.config == "cfg2" else None
    if got is not None:
        rates, cores = got
        kind, note = "reference", "mujoco.mj_step, one MjData per host thread"
    else:
        rates, cores = cpu_sample(cfg, nenv, args.warmup + args.steps)
    timed = rates[args.warmup:]
    total_t = float((nenv * cfg["mj_steps"] / timed).sum())
    value = nenv * cfg["mj_steps"] * len(timed) / total_t
    sample = f"{nenv} envs x {cfg['mj_steps']} mj_steps per step, {args.steps} timed steps of the same ctrl stream"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / len(timed),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"].format(n=args.nenv), "config": args.config,
                       "mj_steps_per_step": cfg["mj_steps"], "nenv_timed": nenv},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample, "note": note},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from stretch_mujoco_b200 import engine
    from stretch_mujoco_b200.simulator import StretchMujocoSimulator

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: CUDA is required (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    raw, A, _ = _load_blob(cfg)
    # env sharding, no data-path collective: weak = --nenv per GPU, strong = --nenv in total
    if args.scaling == "strong":
        nenv = args.nenv // world
        env0 = rank * nenv
    else:
        nenv = args.nenv
        env0 = rank * nenv
    sim = StretchMujocoSimulator(model_blob=raw, nenv=nenv, device=local, maxcon=cfg["maxcon"], maxefc=cfg.get("maxefc", 0))
    sim.start(home=False)
    B = sim.batch
    dm = sim.dmodel
    S = cfg["mj_steps"]
    lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev)
    hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
    lo_np, hi_np = A["actuator_ctrlrange"][:, 0].copy(), A["actuator_ctrlrange"][:, 1].copy()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # sensor buffers: cameras are rendered in env chunks into reused buffers (a full 4096 x 640x480 RGB+depth frame set is 8.8 GB)
    chunk = min(nenv, 256)
    cams = []
    for (cname, W, H, fovy, lim, want_rgb, want_depth) in cfg["cams"]:
        cid = dm.name2id(engine.OBJ_CAMERA, cname)
        rgb = torch.empty(chunk, H, W, 3, dtype=torch.uint8, device=dev) if want_rgb else None
        depth = torch.empty(chunk, H, W, dtype=torch.float32, device=dev) if want_depth else None
        cams.append((cid, W, H, fovy, lim, rgb, depth))
    scan = torch.empty(nenv, dm.nrange, device=dev) if cfg["lidar"] and dm.nrange > 0 else None
    sensor_bytes_per_env_step = sum(W * H * ((3 if r is not None else 0) + (4 if d is not None else 0)) for (_, W, H, _, _, r, d) in cams) \
        + (4 * dm.nrange + 1200 if scan is not None else 0)

    def sensors():
        if scan is not None:
            B.lidar(scan)
        for (cid, W, H, fovy, lim, rgb, depth) in cams:
            for e0 in range(0, nenv, chunk):
                B.render(cid, W, H, fovy, rgb, depth, lim, env_begin=e0, env_count=min(chunk, nenv - e0))

    def bench_step_device():
        if not cams and scan is None:
            B.step(S)
        else:
            for _ in range(S):
                B.step(1)
                sensors()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W_ = args.steps, args.warmup
    # ---- leg 1: device-resident inputs (ctrl drawn on the device before the timed region)
    for p in range(W_):
        B.ctrl.copy_(ctrl_torch(0, env0, nenv, p, lo, hi, dev)); bench_step_device()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = B.launches
    with ClockSampler(local) as clocks:
        barrier()
        for k in range(K):
            B.ctrl.copy_(ctrl_torch(0, env0, nenv, W_ + k, lo, hi, dev))
            flush.fill_(k & 0xFF)           # L2 flush between timed iterations (outside the event pair)
            ev[k][0].record()
            bench_step_device()
            ev[k][1].record()
        barrier()
    launches = B.launches - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * nenv * S * K / (total_ms * 1e-3)
    step_ms = float(np.mean(ms))
    flags_leg1 = B.env_flags.clone()
    ncon_leg1 = B.ncon.sum().float()
    qsum_leg1 = B.qpos.abs().sum()

    # ---- leg 2: end to end through the public API with host buffers (H2D ctrl, D2H status [+ lidar + env-0 frames])
    pin_ctrl = torch.empty(nenv, dm.nu, dtype=torch.float32).pin_memory()
    pin_status = torch.empty(nenv, 24, dtype=torch.float32).pin_memory()
    pin_scan = torch.empty(nenv, dm.nrange, dtype=torch.float32).pin_memory() if scan is not None else None
    pin_img = [(torch.empty(H, W, 3, dtype=torch.uint8).pin_memory() if r is not None else None,
                torch.empty(H, W, dtype=torch.float32).pin_memory() if d is not None else None) for (_, W, H, _, _, r, d) in cams]
    d2h = nenv * 24 * 4 + (nenv * dm.nrange * 4 if scan is not None else 0) \
        + sum((H * W * 3 if r is not None else 0) + (H * W * 4 if d is not None else 0) for (_, W, H, _, _, r, d) in cams)

    def e2e_step():
        sim.set_ctrl(pin_ctrl)                                    # H2D from pinned host memory
        if not cams and scan is None:
            sim.step(S)                                           # command kernel + physics kernels
        else:
            for _ in range(S):
                sim.step(1)
                sensors()
        pin_status.copy_(sim.pull_status().raw, non_blocking=True)  # status kernel + D2H
        if scan is not None:
            pin_scan.copy_(scan, non_blocking=True)
        for (c, p) in zip(cams, pin_img):                          # frames of the last chunk's first env (see config.e2e_d2h)
            if p[0] is not None:
                p[0].copy_(c[5][0], non_blocking=True)
            if p[1] is not None:
                p[1].copy_(c[6][0], non_blocking=True)

    sim.batch.reset()
    for p in range(W_):
        pin_ctrl.copy_(torch.from_numpy(ctrl_np(0, env0, nenv, p, lo_np, hi_np).astype(np.float32)))
        e2e_step()
    host_ctrl = [torch.from_numpy(ctrl_np(0, env0, nenv, W_ + k, lo_np, hi_np).astype(np.float32)) for k in range(K)]
    barrier()
    e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ms = 0.0
    for k in range(K):
        pin_ctrl.copy_(host_ctrl[k])
        flush.fill_(k & 0xFF)
        e0_.record()
        e2e_step()
        e1_.record()
        e1_.synchronize()
        e2e_ms += e0_.elapsed_time(e1_)
    barrier()
    e2e_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * nenv * S * K / (float(e2e_t.item()) * 1e-3)

    # ---- leg 3 (physics-only configs): the reference's own control cadence -- H2D ctrl, ONE mj_step, D2H status, every step.
    # Continues the rollout of leg 2 (same workload mix as the other legs, not the start-up transient), ctrl redrawn every 50 steps.
    per_step = None
    if not cams and scan is None:
        nper = 4
        ctrl3 = [torch.from_numpy(ctrl_np(0, env0, nenv, W_ + K + k, lo_np, hi_np).astype(np.float32)) for k in range(nper + 1)]
        pin_ctrl.copy_(ctrl3[0])
        for _ in range(10):
            sim.set_ctrl(pin_ctrl); sim.step(1); pin_status.copy_(sim.pull_status().raw, non_blocking=True)
        barrier()
        t3_ms = 0.0
        for k in range(nper):
            pin_ctrl.copy_(ctrl3[k + 1])
            e0_.record()
            for _ in range(S):
                sim.set_ctrl(pin_ctrl); sim.step(1); pin_status.copy_(sim.pull_status().raw, non_blocking=True)
            e1_.record(); e1_.synchronize()
            t3_ms += e0_.elapsed_time(e1_)
        t3 = torch.tensor([t3_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        per_step = world * nenv * nper * S / (float(t3.item()) * 1e-3)

    # ---- leg 4: per-kernel durations of the physics pipeline (CUDA events inside the library, kernels serialised on one stream)
    kms = None
    if hasattr(B, "profile_step"):
        acc = np.zeros(3)
        nprof = 50
        for _ in range(5):
            B.profile_step()
        for _ in range(nprof):
            acc += np.array(B.profile_step())
        kms = acc / nprof

    # ---- end-of-rollout metrics: the path's only collective
    metrics = torch.stack([torch.tensor(float(nenv * S * K), device=dev), torch.tensor(total_ms, device=dev, dtype=torch.float32),
                           qsum_leg1, ncon_leg1, (flags_leg1 & 1).ne(0).sum().float(), (flags_leg1 & 2).ne(0).sum().float()]).float()
    if world > 1:
        gathered = torch.empty(world * metrics.numel(), device=dev)
        dist.all_gather_into_tensor(gathered, metrics)
        gathered = gathered.view(world, -1)
    else:
        gathered = metrics[None]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # committed ncu captures of the same kernels (profiles/physics_r2.json): issue-slot / lane utilisation, DRAM traffic
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "physics_r2.json")))
    except Exception:
        pass
    if cams or scan is not None:
        algo = (ALGO_BYTES_PER_ENV_STEP + sensor_bytes_per_env_step) * nenv * S
        kernel = "render_kernel" if cams else "lidar_kernel"
        note = ("per bench step; algorithmic bytes = 828 B physics + W*H*7 (RGB u8 + depth f32) per camera frame + 4*nray + 1.2 kB lidar "
                "per env-step (SURVEY.md 8(d)); the camera kernel is the HBM-write-bound one in the limit")
    else:
        algo = ALGO_BYTES_PER_ENV_STEP * nenv * S
        kernel = "ss_solve_kernel"
        note = ("achieved = 828 algorithmic B per env-step (SURVEY.md 8(d)) x envs / mean duration of one ss_solve_kernel launch over all envs, "
                "measured with CUDA events inside the library (ss_batch_profile_step, kernels serialised); step_frac_of_roofline is the same "
                "figure for the whole bench step (launches_per_step launches: schedule + ss_smooth + ss_narrow + ss_solve per mj_step and env "
                "set, overlapped on three streams).  Physics is instruction-issue / latency bound, not HBM bound: the binding figures are "
                "issue_slot_util and lane_util (ncu capture of the same kernels, profiles/physics_r2.json)")
    achieved = algo / (step_ms * 1e-3) / 1e9
    kernel_ms = step_ms
    if kms is not None and not cams and scan is None:
        # the dominant kernel alone: its mean launch duration over 50 serialised mj_steps of all envs, against the
        # 828 algorithmic bytes of an env-step (the whole step's state traffic is attributed to it)
        kernel_ms = float(kms[2])
        achieved = ALGO_BYTES_PER_ENV_STEP * nenv / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture, scaled from the captured
    # launch (envs_per_launch envs) to this launch (all nenv envs of one profile_step launch); null for the sensor configs
    traffic = prof.get(args.config, {}).get("dram_bytes_per_launch")
    if traffic is not None and prof.get("envs_per_launch"):
        traffic = traffic * nenv / prof["envs_per_launch"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel, "kernel_ms": kernel_ms,
                "kernels_ms_per_mj_step": None if kms is None else {"ss_smooth_kernel": float(kms[0]), "ss_narrow_kernel": float(kms[1]),
                                                                    "ss_solve_kernel": float(kms[2]), "note": "serialised, one env set"},
                "step_frac_of_roofline": algo / (step_ms * 1e-3) / 1e9 / peak,
                "algorithmic_bytes_per_step": algo, "launches_per_step": int(launches) // K,
                "issue_slot_util": prof.get("issue_slot_util"), "lane_util": prof.get("lane_util"),
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback", "note": note}
    line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"].format(n=nenv), "config": args.config, "nenv_per_gpu": nenv, "nenv_total": nenv * world,
                       "mj_steps_per_step": S, "solver": "Newton (reference default)",
                       "ctrl": "uniform over ctrlrange, counter-based RNG (seed 0, env, period, actuator), redrawn every step",
                       "l2": "flushed between timed iterations (256 MiB fill outside the timed event pair)",
                       "parallelism": f"env-sharded x{world}, no data-path collective",
                       "e2e_d2h": "status rows" + (" + lidar scans" if scan is not None else "")
                                  + (" + one env's frames per camera (all frames stay on the device as torch tensors)" if cams else "")},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": nenv * dm.nu * 4, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "rollout_metrics": {"env_steps": float(gathered[:, 0].sum()), "sum_abs_qpos": float(gathered[:, 2].sum()),
                                "contacts_last_step": float(gathered[:, 3].sum()), "envs_reset": float(gathered[:, 4].sum()),
                                "envs_contact_overflow": float(gathered[:, 5].sum())}}
    if per_step is not None:
        line["e2e_per_mj_step"] = {"value": per_step, "unit": "env-steps/s",
                                   "note": "H2D ctrl + one mj_step + D2H status per host round trip (the reference's _ctrl_callback cadence)"}
    if world == 1 and not args.no_cpu:
        ncpu = args.cpu_nenv if not cfg["cams"] else min(args.cpu_nenv, 16)
        periods = 11 if not cfg["cams"] else 3
        rates, cores = cpu_sample(cfg, ncpu, periods)
        line["cpu_baseline"] = {"value": float(np.median(rates[1:])), "unit": "env-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{ncpu} envs x {S} mj_steps x {periods - 1} periods of the same ctrl stream (median period)",
                                "note": "CPU restatement of the mujoco==3.2.6 path (oracle/: mj_step"
                                        + (", ray-cast camera" if cfg["cams"] else "") + (", rangefinders" if cfg["lidar"] else "")
                                        + "), not the MuJoCo binary"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--nenv", type=int, default=4096)
    ap.add_argument("--cpu-nenv", type=int, default=4096)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.steps is None:
        args.steps = 20 if not cfg["cams"] else 4
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
