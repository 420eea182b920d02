"""Mesh asset loading and mass-property integration for the MJCF scene loader.

Covers what the reference's model files need:
Wavefront OBJ (triangles, quads, n-gons; ``v``, ``v/vt``, ``v//vn``, ``v/vt/vn``) and
binary STL.  The reference hands these files to MuJoCo's compiler
(`stretch_mujoco/mujoco_server.py:252`); this module restates the compiler's per-mesh steps:
volume / centre of mass / inertia integration, recentring on the COM, alignment with the
principal axes, and the qhull convex hull used for collision.

UPSTREAM-ASSUMPTION tags mark MuJoCo 3.2.6 behaviours restated from memory.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


def load_obj(path: str):
    """Return (verts[n,3] f64, faces[m,3] i32, texcoords[n_t,2] or None, face_tc[m,3] or None)."""
    verts, tcs, faces, face_tc = [], [], [], []
    with open(path, "r", errors="ignore") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("vt "):
                p = line.split()
                tcs.append((float(p[1]), float(p[2])))
            elif line.startswith("f "):
                idx, tidx = [], []
                for tok in line.split()[1:]:
                    parts = tok.split("/")
                    vi = int(parts[0])
                    idx.append(vi - 1 if vi > 0 else len(verts) + vi)
                    if len(parts) > 1 and parts[1]:
                        ti = int(parts[1])
                        tidx.append(ti - 1 if ti > 0 else len(tcs) + ti)
                    else:
                        tidx.append(-1)
                for k in range(1, len(idx) - 1):  # fan triangulation
                    faces.append((idx[0], idx[k], idx[k + 1]))
                    face_tc.append((tidx[0], tidx[k], tidx[k + 1]))
    v = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    f = np.asarray(faces, dtype=np.int32).reshape(-1, 3)
    # CAD exports repeat every vertex once per incident face: merge exact duplicates
    v, inv = np.unique(v, axis=0, return_inverse=True)
    f = inv.reshape(-1)[f].astype(np.int32)
    if tcs and f.size and np.all(np.asarray(face_tc) >= 0):
        return v, f, np.asarray(tcs, dtype=np.float64), np.asarray(face_tc, dtype=np.int32)
    return v, f, None, None


def load_stl(path: str):
    """Binary STL → (verts, faces); duplicate vertices are merged (exact match)."""
    with open(path, "rb") as fh:
        data = fh.read()
    ntri = struct.unpack_from("<I", data, 80)[0]
    if 84 + 50 * ntri != len(data):
        raise ValueError(f"{path}: not a binary STL (size mismatch)")
    rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                        count=ntri, offset=84)
    allv = rec["v"].reshape(-1, 3)
    uniq, inv = np.unique(allv, axis=0, return_inverse=True)
    return uniq.astype(np.float64), inv.reshape(-1, 3).astype(np.int32), None, None


def load_mesh_file(path: str):
    low = path.lower()
    if low.endswith(".obj"):
        return load_obj(path)
    if low.endswith(".stl"):
        return load_stl(path)
    raise ValueError(f"unsupported mesh format: {path}")


def _eig_frame(inertia: np.ndarray):
    """Principal moments (descending) and a right-handed rotation whose columns are the axes."""
    w, v = np.linalg.eigh(inertia)
    order = np.argsort(-w)
    w, v = w[order], v[:, order]
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
    return w, v


def volume_props(verts: np.ndarray, faces: np.ndarray, shell: bool = False):
    """(measure, com[3], inertia_about_com[3,3] per unit density).

    Volume mode — UPSTREAM-ASSUMPTION (3.2.6 default ``inertia="legacy"``): tetrahedra are built
    from each face to the area-weighted surface centroid and their volumes enter with their
    absolute value.  Shell mode integrates over the surface (``shellinertia="true"``).
    """
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    nrm = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(nrm, axis=1)
    tot_area = float(area.sum())
    if tot_area <= 0:
        raise ValueError("mesh has zero surface area")
    cen = (a + b + c) / 3.0
    surf_c = (cen * area[:, None]).sum(0) / tot_area
    if shell:
        # surface integral of a linear triangle: second moments with the (1/12) rule
        com = surf_c
        pa, pb, pc = a - com, b - com, c - com
        S = np.zeros((3, 3))
        for p, q, wgt in ((pa, pa, 2), (pb, pb, 2), (pc, pc, 2), (pa, pb, 1), (pb, pa, 1), (pa, pc, 1),
                          (pc, pa, 1), (pb, pc, 1), (pc, pb, 1)):
            S += wgt * np.einsum("n,ni,nj->ij", area / 12.0, p, q)
        inertia = np.trace(S) * np.eye(3) - S
        return tot_area, com, inertia
    pa, pb, pc = a - surf_c, b - surf_c, c - surf_c
    vol6 = np.einsum("ni,ni->n", pa, np.cross(pb, pc))
    vol = np.abs(vol6) / 6.0
    tot = float(vol.sum())
    if tot <= 1e-18:
        return volume_props(verts, faces, shell=True)
    com = surf_c + ((pa + pb + pc) / 4.0 * vol[:, None]).sum(0) / tot
    pa, pb, pc = a - com, b - com, c - com
    pd = -com + surf_c  # apex relative to com
    pd = np.broadcast_to(pd, pa.shape)
    # second moment of a tetrahedron with vertices p0..p3: V/20 * (sum_i p_i p_i^T + (sum p)(sum p)^T)
    s = pa + pb + pc + pd
    S = np.zeros((3, 3))
    for p in (pa, pb, pc, pd):
        S += np.einsum("n,ni,nj->ij", vol / 20.0, p, p)
    S += np.einsum("n,ni,nj->ij", vol / 20.0, s, s)
    inertia = np.trace(S) * np.eye(3) - S
    return tot, com, inertia


@dataclass
class MeshAsset:
    """A compiled mesh: geometry is stored in the COM-centred principal-axes frame."""
    name: str
    verts: np.ndarray            # [n,3] centred/aligned render vertices
    faces: np.ndarray            # [m,3]
    texcoord: np.ndarray | None  # [nt,2]
    face_tc: np.ndarray | None   # [m,3]
    pos: np.ndarray              # offset of the mesh frame in the file frame
    quat: np.ndarray             # orientation of the mesh frame in the file frame (w,x,y,z)
    volume: float
    inertia_unit: np.ndarray     # principal moments per unit density (volume mode)
    shell_area: float = 0.0
    shell_inertia_unit: np.ndarray = field(default_factory=lambda: np.zeros(3))
    shell_pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
    shell_quat: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))
    hull_verts: np.ndarray | None = None   # [h,3] convex-hull vertices (same frame as verts)
    hull_faces: np.ndarray | None = None


def mat2quat(R: np.ndarray) -> np.ndarray:
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def compile_mesh(name: str, path: str, scale=(1.0, 1.0, 1.0), want_hull: bool = False) -> MeshAsset:
    verts, faces, tc, ftc = load_mesh_file(path)
    scale = np.asarray(scale, dtype=np.float64)
    verts = verts * scale
    if np.prod(scale) < 0:
        faces = faces[:, ::-1].copy()
        if ftc is not None:
            ftc = ftc[:, ::-1].copy()
    # drop degenerate faces
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    ok = np.linalg.norm(np.cross(b - a, c - a), axis=1) > 0
    faces = faces[ok]
    if ftc is not None:
        ftc = ftc[ok]
    vol, com, inertia = volume_props(verts, faces, shell=False)
    w, R = _eig_frame(inertia)
    area, scom, sin = volume_props(verts, faces, shell=True)
    sw, sR = _eig_frame(sin)
    cverts = (verts - com) @ R  # coordinates in the principal frame
    asset = MeshAsset(name=name, verts=cverts, faces=faces, texcoord=tc, face_tc=ftc, pos=com,
                      quat=mat2quat(R), volume=vol, inertia_unit=w, shell_area=area,
                      shell_inertia_unit=sw, shell_pos=scom, shell_quat=mat2quat(sR))
    if want_hull:
        from scipy.spatial import ConvexHull
        hull = ConvexHull(cverts)  # qhull, the library MuJoCo's compiler uses
        remap = -np.ones(len(cverts), dtype=np.int64)
        remap[hull.vertices] = np.arange(len(hull.vertices))
        asset.hull_verts = cverts[hull.vertices].copy()
        asset.hull_faces = remap[hull.simplices].astype(np.int32)
    return asset
