"""Where does the fp32 error of qacc_smooth come from?  (diagnostic, GPU box)
e_total : device qacc_smooth vs oracle
e_inputs: fp64 solve of the DEVICE's M and qfrc_smooth vs oracle  (rounding of kinematics / CRB / RNE)
e_solve : device qacc_smooth vs fp64 solve of the device's own M and qfrc_smooth (Cholesky + substitutions in fp32)
all in the energy norm of the oracle's M."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from stretch_mujoco_b200 import engine, blob
from oracle.oracle import OracleModel
raw = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stretch_empty_floor.ssm"), "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0); om = OracleModel(raw)
nenv = 32; rng = np.random.default_rng(0)
B = engine.Batch(dm, nenv, debug=True)
lo, hi = A["actuator_ctrlrange"][:, 0], A["actuator_ctrlrange"][:, 1]
for vs in (0.0, 0.05):
    qpos = np.tile(A["qpos0"], (nenv, 1)); qvel = rng.normal(scale=vs, size=(nenv, dm.nv)); ctrl = rng.uniform(lo, hi, size=(nenv, dm.nu))
    B.qpos.copy_(torch.tensor(qpos, dtype=torch.float32)); B.qvel.copy_(torch.tensor(qvel, dtype=torch.float32))
    B.qacc_warmstart.zero_(); B.ctrl.copy_(torch.tensor(ctrl, dtype=torch.float32)); B.env_flags.zero_()
    f = lambda t: t.cpu().numpy().astype(np.float64)
    B.forward(); torch.cuda.synchronize()
    o = om.forward(f(B.qpos), f(B.qvel), f(B.ctrl), f(B.qacc_warmstart), maxcon=B.maxcon, want=("M", "qacc_smooth", "qfrc_bias", "qfrc_passive", "qfrc_actuator"))
    Md, fd, xd = f(B.dbg["M"]), f(B.dbg["qfrc_smooth"]), f(B.dbg["qacc_smooth"])
    Mo, fo, xo = o["M"], o["qfrc_passive"] - o["qfrc_bias"] + o["qfrc_actuator"], o["qacc_smooth"]
    def en(e, ref): return np.sqrt(np.einsum("ei,eij,ej->e", e, Mo, e) / np.einsum("ei,eij,ej->e", ref, Mo, ref))
    x_in = np.stack([np.linalg.solve(Md[e], fd[e]) for e in range(nenv)])
    x_M = np.stack([np.linalg.solve(Md[e], fo[e]) for e in range(nenv)])
    x_f = np.stack([np.linalg.solve(Mo[e], fd[e]) for e in range(nenv)])
    print(f"qvel scale {vs}: e_total {en(xd - xo, xo).max():.2e}  e_inputs {en(x_in - xo, xo).max():.2e}  (M only {en(x_M - xo, xo).max():.2e}, qfrc only {en(x_f - xo, xo).max():.2e})  e_solve {en(xd - x_in, xo).max():.2e}")
    rel = np.abs(fd - fo) / np.abs(fo).max(axis=1, keepdims=True)
    print("   qfrc_smooth rel err per dof (max over envs):", np.array2string(rel.max(axis=0), precision=1))
    print("   M rel err (per-entry, vs sqrt(Mii Mjj)):", (np.abs(Md - Mo) / np.sqrt(np.einsum('eii->ei', Mo)[:, :, None] * np.einsum('eii->ei', Mo)[:, None, :])).max())
