"""Diagnostic: which stage of the pipeline differs first between two identically fed batches?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from stretch_mujoco_b200 import engine, blob
raw = open(bench.GOLDEN, "rb").read()
A, _ = blob.unpack(raw)
dm = engine.DeviceModel(raw, 0)
nenv = int(os.environ.get("NENV", 4096))
B1, B2 = engine.Batch(dm, nenv, debug=True), engine.Batch(dm, nenv, debug=True)
dev = B1.qpos.device
lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev); hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
c = bench.ctrl_torch(0, 0, nenv, 0, lo, hi, dev)
B1.ctrl.copy_(c); B2.ctrl.copy_(c)
nst = int(os.environ.get("NST", 2))
for s in range(60):
    B1.step(nst); B2.step(nst); torch.cuda.synchronize()
    d = {k: int((getattr(B1, k) != getattr(B2, k)).reshape(nenv, -1).any(1).sum()) for k in ("xpos", "ncon", "contact_geom", "contact_dist", "qacc", "qpos", "qvel", "solver_iter", "act_length", "sensordata")}
    d.update({"dbg_" + k: int((B1.dbg[k] != B2.dbg[k]).reshape(nenv, -1).any(1).sum()) for k in ("M", "qfrc_smooth", "qacc_smooth", "nefc", "contact_pos", "contact_normal", "qfrc_constraint")})
    if any(d.values()):
        print("step", (s + 1) * nst, {k: v for k, v in d.items() if v})
        bad = (B1.qacc != B2.qacc).any(1)
        e = int(torch.nonzero(bad).flatten()[0]) if bad.any() else 0
        print(" env", e, "ncon", int(B1.ncon[e]), int(B2.ncon[e]), "iter", int(B1.solver_iter[e]), int(B2.solver_iter[e]))
        print(" dist1", B1.contact_dist[e, :8].cpu().numpy()); print(" dist2", B2.contact_dist[e, :8].cpu().numpy())
        print(" geom1", B1.contact_geom[e, :8].cpu().numpy().tolist())
        break
else:
    print("identical for", 60 * nst, "steps; nan envs", int(torch.isnan(B1.qpos).any(1).sum()), "flags1", int((B1.env_flags & 1).ne(0).sum()))
