"""Diagnostic: where do the device's qacc outliers come from (M, qfrc_smooth, qacc_smooth, solver)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
from oracle.oracle import OracleModel
np.set_printoptions(linewidth=200, precision=4, suppress=False)
raw = open(bench.GOLDEN, "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = 2048
B = engine.Batch(dm, nenv, debug=True)
dev = B.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
for p in range(4):
    B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)); B.step(50)
B.step(7)
torch.cuda.synchronize()
q, v, w, c = (t.cpu().numpy().astype(np.float64) for t in (B.qpos, B.qvel, B.qacc_warmstart, B.ctrl))
B.forward(); torch.cuda.synchronize()
om = OracleModel(raw); om.set_options(max_iter=200, tolerance=1e-14, enable_lidar=False)
o = om.forward(q, v, c, w, want=("qacc", "M", "qacc_smooth", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qfrc_constraint", "nefc"))
qa = B.qacc.cpu().numpy(); M = B.dbg["M"].cpu().numpy(); qas = B.dbg["qacc_smooth"].cpu().numpy(); qfs = B.dbg["qfrc_smooth"].cpu().numpy()
qfc = B.dbg["qfrc_constraint"].cpu().numpy()
ofs = o["qfrc_passive"] - o["qfrc_bias"] + o["qfrc_actuator"]
err = np.abs(qa - o["qacc"]).max(1) / (np.abs(o["qacc"]).max(1) + 1e-3)
es = np.abs(qas - o["qacc_smooth"]).max(1) / (np.abs(o["qacc_smooth"]).max(1) + 1e-3)
print("qacc err  median %.2e p99 %.2e max %.2e" % (np.median(err), np.quantile(err, .99), err.max()))
print("qacc_smooth err median %.2e p99 %.2e max %.2e" % (np.median(es), np.quantile(es, .99), es.max()))
eM = np.abs(M - o["M"]).max((1, 2)) / np.abs(o["M"]).max((1, 2)); print("M err (rel to max entry) median %.2e max %.2e" % (np.median(eM), eM.max()))
dM = np.abs(np.diagonal(M, axis1=1, axis2=2) - np.diagonal(o["M"], axis1=1, axis2=2)) / np.diagonal(o["M"], axis1=1, axis2=2)
print("M diag rel err per dof (max over envs):", dM.max(0))
ef = np.abs(qfs - ofs); print("qfrc_smooth abs err per dof (max over envs):", ef.max(0))
print("qvel abs max per dof:", np.abs(v).max(0))
for e in np.argsort(-err)[:4]:
    print("env", e, "err", err[e], "qacc_smooth err", es[e])
    print("  qacc gpu ", qa[e, 16:23]); print("  qacc true", o["qacc"][e, 16:23])
    print("  qas  gpu ", qas[e, 16:23]); print("  qas  true", o["qacc_smooth"][e, 16:23])
    print("  qfs  gpu ", qfs[e, 16:23]); print("  qfs  true", ofs[e, 16:23])
    print("  qfc  gpu ", qfc[e, 16:23]); print("  qfc  true", o["qfrc_constraint"][e, 16:23])
    print("  qvel", v[e, 16:23])
    # residual of the device's answer in fp64: M qacc - qfs - qfc
    r = o["M"][e] @ qa[e].astype(np.float64) - ofs[e] - qfc[e]; print("  residual(gpu qacc, true M, gpu qfc)", r[16:23])
print("=== contacts of outlier envs")
o3 = om.forward(q, v, c, w, want=("ncon", "contact_geom", "contact_dist", "contact_pos", "contact_frame"))
gd = B.contact_dist.cpu().numpy(); gp = B.dbg["contact_pos"].cpu().numpy(); gnrm = B.dbg["contact_normal"].cpu().numpy(); gg = B.contact_geom.cpu().numpy()
gnames = _[2] if False else None
for e in np.argsort(-err)[:5]:
    n = int(o3["ncon"][e])
    print("env", e, "err %.2e" % err[e], "ncon", n)
    for k in range(n):
        print("   geoms", gg[e, k], o3["contact_geom"][e, k], "dist gpu %.7f orcl %.7f" % (gd[e, k], o3["contact_dist"][e, k]),
              "| dpos %.2e" % np.abs(gp[e, k] - o3["contact_pos"][e, k]).max(), "| normal gpu", gnrm[e, k].round(5), "orcl", o3["contact_frame"][e, k].round(5))
# global contact statistics
nc = o3["ncon"]; m_ = np.arange(gd.shape[1])[None, :] < nc[:, None]
dd = np.abs(gd[:, :o3["contact_dist"].shape[1]] - o3["contact_dist"][:, :gd.shape[1]])[m_[:, :o3["contact_dist"].shape[1]]] if gd.shape[1] <= o3["contact_dist"].shape[1] else None
