import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from stretch_mujoco_b200 import engine, blob
from oracle.oracle import OracleModel
from test_physics_gpu import _targets, _load
raw = open(os.path.join(os.path.dirname(__file__), "golden", "stretch_empty_floor.ssm"), "rb").read()
A, names = blob.unpack(raw)
om = OracleModel(raw); om.set_options(enable_lidar=False)
dm = engine.DeviceModel(raw, 0)
nenv = 64
rng = np.random.default_rng(0)
qpos = A["qpos0"][None].copy(); qvel = np.zeros((1, om.nv)); warm = np.zeros((1, om.nv)); t = np.zeros(1)
home = A["key_ctrl"][0][None].copy()
om.step(qpos, qvel, home, warm, t, nsteps=1500)
B = engine.Batch(dm, nenv)
tg = _targets(rng, home[0], nenv)
if os.environ.get("GENTLE"):
    tg[:, 0:2] = rng.uniform(-0.5, 0.5, (nenv, 2)); tg[:, 7] = rng.uniform(0.0, 0.04, nenv)
qpos, qvel, warm, ctrl = _load(B, np.tile(qpos, (nenv, 1)), np.tile(qvel, (nenv, 1)), np.tile(warm, (nenv, 1)), tg)
q1, v1, w1 = qpos.copy(), qvel.copy(), warm.copy()
B.step(1); torch.cuda.synchronize(); om.step(q1, v1, ctrl, w1, np.zeros(nenv), nsteps=1)
print("one step: |dqvel| med %.2e max %.2e  |dqpos| max %.2e" % (np.median(np.abs(B.qvel.cpu().numpy()-v1).max(axis=1)), np.abs(B.qvel.cpu().numpy()-v1).max(), np.abs(B.qpos.cpu().numpy()-q1).max()))
print("worst dof per env", np.bincount(np.argmax(np.abs(B.qvel.cpu().numpy()-v1), axis=1), minlength=26))
qpos, qvel, warm, ctrl = _load(B, qpos, qvel, warm, ctrl)
t = np.zeros(nenv)
gtype = A["geom_type"]; gn = names[2]
first_div = -np.ones(nenv, int)
for k in range(100):
    B.step(10); torch.cuda.synchronize()
    o = om.step(qpos, qvel, ctrl, warm, t, nsteps=10, maxcon=24, want=("contact_geom", "ncon"))
    err = np.abs(B.qpos.cpu().numpy() - qpos).max(axis=1)
    for e in range(nenv):
        if first_div[e] < 0 and err[e] > 1e-3:
            first_div[e] = k * 10
            cg = o["contact_geom"][e]; gg = B.contact_geom[e].cpu().numpy()
            print("env", e, "diverged by step", k * 10 + 10, "err", err[e], "idx", int(np.argmax(np.abs(B.qpos[e].cpu().numpy() - qpos[e]))),
                  "oracle contacts", [(gn[a] or a, gn[b] or b) for a, b in cg if a >= 0][:8], "gpu", [(int(a), int(b)) for a, b in gg if a >= 0][:8])
rel = (np.abs(B.qpos.cpu().numpy() - qpos) / np.maximum(np.abs(qpos), 1.0)).max(axis=1)
print("n diverged", (first_div >= 0).sum(), "rel sorted", np.sort(rel)[-12:])
