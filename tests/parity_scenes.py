"""Rollout parity of the reference's default scene (scene.xml, nv = 44) and of the kitchen proxy (nv = 32):
16 envs x 1000 steps with the `home` command plus per-env random arm / head targets, device fp32 vs oracle fp64.
    python tests/parity_scenes.py >> profiles/parity_r2.md"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
from oracle.oracle import OracleModel

f64 = lambda t: t.cpu().numpy().astype(np.float64)


def rel(a, b):
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-2)


def run(title, name, maxcon, maxefc, nenv=16, nsteps=1000):
    raw = blob.read_bytes(os.path.join(bench.GOLDEN_DIR, name))
    A, _ = blob.unpack(raw)
    dm = engine.DeviceModel(raw, 0)
    om = OracleModel(raw); om.set_options(enable_lidar=False)
    B = engine.Batch(dm, nenv, maxcon=maxcon, maxefc=maxefc)
    # settle the `home` keyframe for 3 s in the oracle (the stowed start pose of qpos0 interpenetrates the base)
    q0 = A["key_qpos"][0][None].copy(); v0 = np.zeros((1, om.nv)); w0 = np.zeros((1, om.nv))
    om.step(q0, v0, A["key_ctrl"][0][None].copy(), w0, np.zeros(1), nsteps=1500)
    B.qpos.copy_(torch.tensor(np.tile(q0, (nenv, 1)), dtype=torch.float32)); B.qvel.copy_(torch.tensor(np.tile(v0, (nenv, 1)), dtype=torch.float32))
    B.qacc_warmstart.copy_(torch.tensor(np.tile(w0, (nenv, 1)), dtype=torch.float32))
    rng = np.random.default_rng(3)
    ctrl = np.tile(A["key_ctrl"][0], (nenv, 1))
    ctrl[:, 2] = rng.uniform(0.3, 1.0, nenv); ctrl[:, 3] = rng.uniform(0, 0.3, nenv); ctrl[:, 8] = rng.uniform(-2, 1, nenv); ctrl[:, 9] = rng.uniform(-1, 0.5, nenv)
    B.ctrl.copy_(torch.tensor(ctrl, dtype=torch.float32))
    q, v, w, c = f64(B.qpos), f64(B.qvel), f64(B.qacc_warmstart), f64(B.ctrl)
    t = np.zeros(nenv)
    same = np.ones(nenv, bool)
    print(f"\n### {title}\n")
    print("| step | median rel qpos | max rel qpos | median rel qvel | envs with equal contact list now | ... at every step so far | max ncon |")
    print("|---|---|---|---|---|---|---|")
    for blk in range(nsteps // 100):
        for s in range(100):
            B.step(1)
            o = om.step(q, v, c, w, t, nsteps=1, want=("contact_geom", "ncon"), maxcon=maxcon)
            eq = (B.contact_geom.cpu().numpy() == o["contact_geom"]).all((1, 2))
            same &= eq
        e, ev = rel(f64(B.qpos), q), rel(f64(B.qvel), v)
        if blk == 9:
            dq = np.abs(f64(B.qpos) - q); dv = np.abs(f64(B.qvel) - v)
            print("<!-- per-index max |dqpos|", np.array2string(dq.max(0), precision=1, max_line_width=400), "max |dqvel|", np.array2string(dv.max(0), precision=1, max_line_width=400), "max|qvel|", np.abs(v).max(), "-->")
        print(f"| {blk * 100 + 100} | {np.median(e):.1e} | {e.max():.1e} | {np.median(ev):.1e} | {int(eq.sum())}/{nenv} | {int(same.sum())}/{nenv} | {int(o['ncon'].max())} |")
    fl = B.env_flags.cpu().numpy()
    print(f"\nflags: envs reset {int((fl & 1).sum())}, envs with contact overflow {int(((fl >> 1) & 1).sum())}")


print("\n## Other scenes (16 envs x 1000 steps, `home` command + per-env random lift / arm / head targets)")
run("default scene.xml (dock, table, two free objects; nv = 44)", "stretch_default_scene.ssm", 72, 288)
run("kitchen proxy (box fixtures + one free box; nv = 32)", "stretch_kitchen_proxy_render.ssm.z", 32, 0)
