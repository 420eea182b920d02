"""Model compiler: parsed MJCF (`mjcf.Scene`) → flat numeric model arrays (`Model`).

Restates what ``MjModel.from_xml_path`` (`stretch_mujoco/mujoco_server.py:252`) produces for the
MJCF subset of SURVEY.md Appendix A.3.  Array names follow ``mjModel`` so that a dump of a real
``mjModel`` could be loaded through the same blob format (`blob.py`).  All arrays are fp64 / int32
masters; the CUDA side converts to fp32 when it uploads.

UPSTREAM-ASSUMPTION tags mark behaviours of MuJoCo 3.2.6's compiler restated from memory (the
library's source is not vendored in the reference).
"""
from __future__ import annotations

import math

import numpy as np

from . import meshio
from .mjcf import Scene, _floats, quat2mat, quat_mul, quat_norm, quat_rot

GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4, "cylinder": 5, "box": 6,
              "mesh": 7}
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
JNT_TYPES = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}
SENS_GYRO, SENS_ACCEL, SENS_RANGE = 0, 1, 2
OBJ_BODY, OBJ_JOINT, OBJ_GEOM, OBJ_SITE, OBJ_CAMERA, OBJ_ACTUATOR, OBJ_SENSOR, OBJ_KEY, OBJ_MESH, OBJ_TENDON = range(10)
MINVAL = 1e-15


def _solimp(s: str):
    """solimp may be given with 3 values; the remaining two default to midpoint 0.5, power 2."""
    v = _floats(s)
    return (v + [0.9, 0.95, 0.001, 0.5, 2.0][len(v):])[:5]


class Model:
    """Bag of named numpy arrays + name tables. ``arrays`` is what goes into the blob."""

    def __init__(self):
        self.arrays: dict[str, np.ndarray] = {}
        self.names: dict[int, list[str]] = {}

    def __getattr__(self, k):
        arrays = self.__dict__.get("arrays", {})
        if k in arrays:
            return arrays[k]
        raise AttributeError(k)

    def name2id(self, objtype: int, name: str) -> int:
        try:
            return self.names[objtype].index(name)
        except ValueError:
            return -1

    def id2name(self, objtype: int, i: int) -> str:
        return self.names[objtype][i]


def _primitive_props(gtype: str, size):
    """(volume, principal inertia per unit mass) for primitive geoms in their own frame."""
    if gtype == "sphere":
        r = size[0]
        return 4.0 / 3.0 * math.pi * r ** 3, np.array([0.4 * r * r] * 3)
    if gtype == "box":
        a, b, c = size[:3]
        return 8 * a * b * c, np.array([(b * b + c * c) / 3, (a * a + c * c) / 3, (a * a + b * b) / 3])
    if gtype == "cylinder":
        r, h = size[0], size[1]
        return math.pi * r * r * 2 * h, np.array([(3 * r * r + 4 * h * h) / 12] * 2 + [r * r / 2])
    if gtype == "ellipsoid":
        a, b, c = size[:3]
        return 4.0 / 3.0 * math.pi * a * b * c, np.array([(b * b + c * c) / 5, (a * a + c * c) / 5, (a * a + b * b) / 5])
    if gtype == "capsule":
        r, h = size[0], size[1]
        vc, vs = math.pi * r * r * 2 * h, 4.0 / 3.0 * math.pi * r ** 3
        v = vc + vs
        izz = (vc * r * r / 2 + vs * 0.4 * r * r) / v
        ixx = (vc * (3 * r * r + 4 * h * h) / 12 + vs * (0.4 * r * r + h * h + 0.75 * r * h)) / v
        return v, np.array([ixx, ixx, izz])
    raise ValueError(gtype)


def compile_scene(sc: Scene, with_render: bool = True, verbose: bool = False) -> Model:
    m = Model()
    A = m.arrays
    nbody = len(sc.bodies)
    body_names = [b.name for b in sc.bodies]

    # ------------------------------------------------------------------ meshes (lazy, cached)
    mesh_cache: dict[str, meshio.MeshAsset] = {}
    collide_meshes = set()
    for b in sc.bodies:
        for g in b.geoms:
            if g["type"] == "mesh" and (int(g["contype"]) or int(g["conaffinity"])):
                collide_meshes.add(g["mesh"])

    def get_mesh(name: str) -> meshio.MeshAsset:
        if name not in mesh_cache:
            spec = sc.meshes[name]
            mesh_cache[name] = meshio.compile_mesh(name, spec["path"], spec["scale"],
                                                   want_hull=name in collide_meshes)
        return mesh_cache[name]

    # ------------------------------------------------------------------ bodies / joints / dofs
    body_parent = np.array([max(b.parent, 0) for b in sc.bodies], dtype=np.int32)
    jnt_type, jnt_body, jnt_pos, jnt_axis, jnt_qposadr, jnt_dofadr = [], [], [], [], [], []
    jnt_stiffness, jnt_range, jnt_limited, jnt_margin, jnt_solref, jnt_solimp, jnt_names = [], [], [], [], [], [], []
    dof_body, dof_jnt, dof_parent, dof_armature, dof_damping, dof_floss, dof_solref, dof_solimp = ([] for _ in range(8))
    qpos0, qpos_spring = [], []
    body_jntnum, body_jntadr, body_dofnum, body_dofadr = (np.zeros(nbody, np.int32) for _ in range(4))
    body_lastdof = -np.ones(nbody, dtype=np.int64)  # last dof of the nearest jointed ancestor-or-self
    for bi, b in enumerate(sc.bodies):
        body_jntadr[bi] = len(jnt_type) if b.joints else -1
        body_dofadr[bi] = len(dof_body) if b.joints else -1
        body_jntnum[bi] = len(b.joints)
        last = body_lastdof[b.parent] if bi > 0 else -1
        for j in b.joints:
            t = JNT_TYPES[j["type"]]
            jid = len(jnt_type)
            jnt_type.append(t); jnt_body.append(bi); jnt_names.append(j["name"])
            jnt_pos.append(j["_pos"]); ax = np.asarray(j["_axis"], dtype=np.float64)
            jnt_axis.append(ax / max(np.linalg.norm(ax), 1e-14))
            jnt_qposadr.append(len(qpos0)); jnt_dofadr.append(len(dof_body))
            jnt_stiffness.append(float(j["stiffness"]))
            jnt_margin.append(float(j["margin"]))
            jnt_solref.append(_floats(j["solreflimit"], 2)); jnt_solimp.append(_solimp(j["solimplimit"]))
            has_range = "range" in j
            rng = [sc._ang(x) if t == JNT_HINGE else x for x in _floats(j["range"], 2)] if has_range else [0.0, 0.0]
            lim = j.get("limited", "auto")
            limited = (lim == "true") or (lim == "auto" and sc.autolimits and has_range and rng[0] < rng[1])
            jnt_range.append(rng); jnt_limited.append(int(limited and t in (JNT_SLIDE, JNT_HINGE)))
            ndof = {JNT_FREE: 6, JNT_BALL: 3, JNT_SLIDE: 1, JNT_HINGE: 1}[t]
            if t == JNT_FREE:
                qpos0 += list(b.pos) + list(b.quat); qpos_spring += list(b.pos) + list(b.quat)
            elif t == JNT_BALL:
                qpos0 += [1, 0, 0, 0]; qpos_spring += [1, 0, 0, 0]
            else:
                ref = float(j["ref"]); sref = float(j["springref"])
                if t == JNT_HINGE:
                    ref, sref = sc._ang(ref), sc._ang(sref)
                qpos0.append(ref); qpos_spring.append(sref)
            for k in range(ndof):
                dof_body.append(bi); dof_jnt.append(jid); dof_parent.append(last)
                last = len(dof_body) - 1
                dof_armature.append(float(j["armature"])); dof_damping.append(float(j["damping"]))
                dof_floss.append(float(j["frictionloss"]))
                dof_solref.append(_floats(j["solreffriction"], 2)); dof_solimp.append(_solimp(j["solimpfriction"]))
        body_dofnum[bi] = len(dof_body) - (body_dofadr[bi] if b.joints else len(dof_body))
        body_lastdof[bi] = last
    nq, nv, njnt = len(qpos0), len(dof_body), len(jnt_type)
    dof_parent = np.array(dof_parent, dtype=np.int32).reshape(nv)
    # sparse-M addressing: row i holds M[i, i], M[i, parent(i)], M[i, parent(parent(i))], ...
    dof_Madr = np.zeros(nv, np.int32)
    nM = 0
    for i in range(nv):
        dof_Madr[i] = nM
        k = i
        while k >= 0:
            nM += 1
            k = dof_parent[k]
    body_weldid = np.zeros(nbody, np.int32)
    body_rootid = np.zeros(nbody, np.int32)
    for bi in range(1, nbody):
        body_weldid[bi] = bi if sc.bodies[bi].joints else body_weldid[body_parent[bi]]
        body_rootid[bi] = bi if body_parent[bi] == 0 else body_rootid[body_parent[bi]]

    # ------------------------------------------------------------------ geoms
    G = dict(type=[], contype=[], conaffinity=[], condim=[], bodyid=[], dataid=[], priority=[], size=[], rbound=[],
             aabb=[], pos=[], quat=[], friction=[], solref=[], solimp=[], solmix=[], margin=[], gap=[], group=[],
             rgba=[], matid=[], names=[])
    mesh_ids: dict[str, int] = {}
    mesh_list: list[meshio.MeshAsset] = []
    body_mass = np.zeros(nbody); body_ipos = np.zeros((nbody, 3)); body_iquat = np.tile([1.0, 0, 0, 0], (nbody, 1))
    body_inertia = np.zeros((nbody, 3))
    body_geomadr = -np.ones(nbody, np.int32); body_geomnum = np.zeros(nbody, np.int32)
    mat_ids = {n: i for i, n in enumerate(sc.material_order)}
    for bi, b in enumerate(sc.bodies):
        contrib = []  # (mass, com_in_body[3], inertia_about_com_in_body_axes[3,3])
        if b.geoms:
            body_geomadr[bi] = len(G["type"])
        body_geomnum[bi] = len(b.geoms)
        for g in b.geoms:
            gtype = g["type"]
            size = _floats(g.get("size", "0 0 0"), 3)
            pos, quat = np.asarray(g["_pos"], dtype=np.float64), quat_norm(g["_quat"])
            shell = g.get("shellinertia", "false") == "true"
            mass_attr = float(g["mass"]) if "mass" in g else None
            density = float(g["density"])
            dataid = -1
            if gtype == "mesh":
                ma = get_mesh(g["mesh"])
                # MuJoCo >= 3.3 moved the switch from the geom (shellinertia="true") to the mesh asset (inertia="shell"):
                # models/stretch_mj_3.3.0.xml:129-222
                shell = shell or sc.meshes.get(g["mesh"], {}).get("inertia", "") == "shell"
                if g["mesh"] not in mesh_ids:
                    mesh_ids[g["mesh"]] = len(mesh_list); mesh_list.append(ma)
                dataid = mesh_ids[g["mesh"]]
                gpos_file, gquat_file = pos, quat
                # geom frame = mesh's COM/principal frame expressed in the body frame
                pos = gpos_file + quat_rot(gquat_file, ma.pos)
                quat = quat_norm(quat_mul(gquat_file, ma.quat))
                if shell:
                    measure, iunit = ma.shell_area, ma.shell_inertia_unit
                    ipos = gpos_file + quat_rot(gquat_file, ma.shell_pos)
                    iquat = quat_norm(quat_mul(gquat_file, ma.shell_quat))
                else:
                    measure, iunit, ipos, iquat = ma.volume, ma.inertia_unit, pos, quat
                mass = mass_attr if mass_attr is not None else density * measure
                inertia_p = iunit * (mass / measure) if measure > 0 else np.zeros(3)
                ext = np.abs(ma.verts).max(0)
                lo, hi = ma.verts.min(0), ma.verts.max(0)
                aabb = np.concatenate([(lo + hi) / 2, (hi - lo) / 2])
                rbound = float(np.linalg.norm(ma.verts, axis=1).max())
                size = list(ext)
            elif gtype == "plane":
                mass, inertia_p, ipos, iquat = 0.0, np.zeros(3), pos, quat
                aabb = np.array([0, 0, 0, 1e10, 1e10, 1e10]); rbound = 0.0
            else:
                vol, iunit = _primitive_props(gtype, size)
                mass = mass_attr if mass_attr is not None else density * vol
                inertia_p, ipos, iquat = iunit * mass, pos, quat
                if gtype == "sphere":
                    half = np.array([size[0]] * 3); rbound = size[0]
                elif gtype == "box" or gtype == "ellipsoid":
                    half = np.array(size[:3]); rbound = float(np.linalg.norm(size[:3])) if gtype == "box" else max(size[:3])
                elif gtype == "cylinder":
                    half = np.array([size[0], size[0], size[1]]); rbound = math.hypot(size[0], size[1])
                else:  # capsule
                    half = np.array([size[0], size[0], size[1] + size[0]]); rbound = size[0] + size[1]
                aabb = np.concatenate([np.zeros(3), half])
            if bi > 0 and mass > 0:
                R = quat2mat(iquat)
                contrib.append((mass, ipos, R @ np.diag(inertia_p) @ R.T))
            matname = g.get("material")
            if "rgba" in g.get("_explicit", ()) or matname is None:
                rgba = _floats(g["rgba"], 4)
            else:
                rgba = _floats(sc.materials[matname]["rgba"], 4)
            G["type"].append(GEOM_TYPES[gtype]); G["contype"].append(int(g["contype"]))
            G["conaffinity"].append(int(g["conaffinity"])); G["condim"].append(int(g["condim"]))
            G["bodyid"].append(bi); G["dataid"].append(dataid); G["priority"].append(int(g["priority"]))
            G["size"].append(size); G["rbound"].append(rbound); G["aabb"].append(aabb)
            G["pos"].append(pos); G["quat"].append(quat); G["friction"].append(_floats(g["friction"], 3))
            G["solref"].append(_floats(g["solref"], 2)); G["solimp"].append(_solimp(g["solimp"]))
            G["solmix"].append(float(g["solmix"])); G["margin"].append(float(g["margin"])); G["gap"].append(float(g["gap"]))
            G["group"].append(int(g["group"])); G["rgba"].append(rgba)
            G["matid"].append(mat_ids.get(matname, -1) if matname else -1); G["names"].append(g["name"])
        if bi == 0:
            continue
        if b.inertial is not None:
            ine = b.inertial
            body_mass[bi] = ine["mass"]; body_ipos[bi] = ine["pos"]
            if ine["fullinertia"] is not None:
                f = ine["fullinertia"]
                I = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                w, V = meshio._eig_frame(I)
                body_inertia[bi] = w
                body_iquat[bi] = quat_norm(quat_mul(ine["quat"], meshio.mat2quat(V)))
            else:
                body_inertia[bi] = ine["diaginertia"]; body_iquat[bi] = ine["quat"]
        elif contrib:
            mtot = sum(c[0] for c in contrib)
            com = sum(c[0] * c[1] for c in contrib) / mtot
            I = np.zeros((3, 3))
            for mass, p, Ic in contrib:
                d = p - com
                I += Ic + mass * ((d @ d) * np.eye(3) - np.outer(d, d))
            w, V = meshio._eig_frame(I)
            body_mass[bi] = mtot; body_ipos[bi] = com; body_inertia[bi] = w; body_iquat[bi] = meshio.mat2quat(V)
        if sc.bodies[bi].joints and body_mass[bi] <= 0:
            raise ValueError(f"moving body '{b.name}' has no mass")
    ngeom = len(G["type"])

    # ------------------------------------------------------------------ sites / cameras / lights
    site_body, site_pos, site_quat, site_names = [], [], [], []
    cam_body, cam_pos, cam_quat, cam_fovy, cam_names = [], [], [], [], []
    light_body, light_pos, light_dir, light_directional, light_ambient, light_diffuse, light_specular = ([] for _ in range(7))
    for bi, b in enumerate(sc.bodies):
        for s in b.sites:
            site_body.append(bi); site_pos.append(s["_pos"]); site_quat.append(quat_norm(s["_quat"])); site_names.append(s["name"])
        for c in b.cameras:
            cam_body.append(bi); cam_pos.append(c["_pos"]); cam_quat.append(quat_norm(c["_quat"]))
            fov = float(c["fovy"]); cam_fovy.append(fov if sc.angle == "degree" else fov)  # fovy is always degrees
            cam_names.append(c["name"])
        for l in b.lights:
            if l.get("active", "true") != "true":
                continue
            light_body.append(bi); light_pos.append(l["_pos"])
            d = np.asarray(l["_dir"], dtype=np.float64); light_dir.append(d / max(np.linalg.norm(d), 1e-14))
            light_directional.append(int(l["directional"] == "true"))
            light_ambient.append(_floats(l["ambient"], 3)); light_diffuse.append(_floats(l["diffuse"], 3))
            light_specular.append(_floats(l["specular"], 3))

    # ------------------------------------------------------------------ tendons / equality / actuators
    ten_adr, ten_num, wrap_jnt, wrap_coef, ten_names = [], [], [], [], []
    for t in sc.tendons:
        ten_adr.append(len(wrap_jnt)); ten_num.append(len(t["joints"])); ten_names.append(t["name"])
        for jn, coef in t["joints"]:
            wrap_jnt.append(jnt_names.index(jn)); wrap_coef.append(coef)
    eq_obj1, eq_obj2, eq_data, eq_solref, eq_solimp, eq_active = [], [], [], [], [], []
    for e in sc.equalities:
        eq_obj1.append(jnt_names.index(e["joint1"]))
        eq_obj2.append(jnt_names.index(e["joint2"]) if e["joint2"] else -1)
        eq_data.append(e["polycoef"]); eq_solref.append(e["solref"]); eq_solimp.append(e["solimp"])
        eq_active.append(int(e["active"]))
    nu = len(sc.actuators)
    act_trntype = np.zeros(nu, np.int32); act_trnid = np.zeros(nu, np.int32); act_gear = np.zeros(nu)
    act_gain = np.zeros((nu, 3)); act_bias = np.zeros((nu, 3)); act_ctrlrange = np.zeros((nu, 2))
    act_ctrllimited = np.zeros(nu, np.int32); act_forcerange = np.zeros((nu, 2)); act_forcelimited = np.zeros(nu, np.int32)
    for i, a in enumerate(sc.actuators):
        if a["trntype"] == "joint":
            act_trntype[i] = 0; act_trnid[i] = jnt_names.index(a["target"])
            if jnt_type[act_trnid[i]] not in (JNT_SLIDE, JNT_HINGE):
                raise ValueError("actuators on free/ball joints are not supported")
        else:
            act_trntype[i] = 1; act_trnid[i] = ten_names.index(a["target"])
        act_gear[i] = a["gear"][0]; act_gain[i] = a["gainprm"]; act_bias[i] = a["biasprm"] if a["biastype"] == "affine" else 0
        for key, rng, lim in (("ctrl", act_ctrlrange, act_ctrllimited), ("force", act_forcerange, act_forcelimited)):
            r = a[key + "range"]; l = a[key + "limited"]
            if r is not None:
                rng[i] = r
            lim[i] = int(l == "true" or ((l is None or l == "auto") and sc.autolimits and r is not None and r[0] < r[1]))

    # ------------------------------------------------------------------ sensors (implicit replication)
    sens_type, sens_obj, sens_adr, sens_dim, sens_cutoff, sens_names = [], [], [], [], [], []
    adr = 0
    for s in sc.sensors:
        kind = {"gyro": SENS_GYRO, "accelerometer": SENS_ACCEL, "rangefinder": SENS_RANGE}.get(s["type"])
        if kind is None:
            raise ValueError(f"sensor <{s['type']}> is not supported")
        site = s["site"]
        if site in site_names:
            targets = [(s.get("name", ""), site_names.index(site))]
        else:  # the site was replicated: the sensor is replicated with the same suffixes
            targets = [(s.get("name", "") + n[len(site):], i) for i, n in enumerate(site_names)
                       if n.startswith(site) and n[len(site):].isdigit()]
            if not targets:
                raise ValueError(f"sensor site '{site}' not found")
        dim = 1 if kind == SENS_RANGE else 3
        for nm, sid in targets:
            sens_type.append(kind); sens_obj.append(sid); sens_adr.append(adr); sens_dim.append(dim)
            sens_cutoff.append(float(s.get("cutoff", 0))); sens_names.append(nm); adr += dim
    nsensordata = adr

    # ------------------------------------------------------------------ keyframes
    key_names = [k.get("name", "") for k in sc.keys]
    key_ctrl = np.zeros((len(sc.keys), nu)); key_qpos = np.tile(np.asarray(qpos0, dtype=np.float64), (len(sc.keys), 1))
    for i, k in enumerate(sc.keys):
        if "ctrl" in k:
            key_ctrl[i] = _floats(k["ctrl"], nu)
        if "qpos" in k:
            key_qpos[i] = _floats(k["qpos"], nq)

    # ------------------------------------------------------------------ store
    f64 = lambda x, shape=None: np.asarray(x, dtype=np.float64).reshape(shape) if shape is not None else np.asarray(x, dtype=np.float64)
    i32 = lambda x: np.asarray(x, dtype=np.int32)
    o = sc.option
    if o["integrator"] != "implicitfast":
        raise ValueError("only integrator='implicitfast' is supported (stretch.xml:7)")
    A["opt_timestep"] = f64([o["timestep"]]); A["opt_gravity"] = f64(o["gravity"]); A["opt_impratio"] = f64([o["impratio"]])
    A["opt_tolerance"] = f64([o["tolerance"]]); A["opt_ls_tolerance"] = f64([o["ls_tolerance"]])
    A["opt_iterations"] = i32([o["iterations"]]); A["opt_ls_iterations"] = i32([o["ls_iterations"]])
    A["opt_cone"] = i32([1 if o["cone"] == "elliptic" else 0]); A["opt_multiccd"] = i32([int(o["multiccd"])])
    A["opt_solver"] = i32([{"PGS": 0, "CG": 1, "Newton": 2}[o["solver"]]])
    A["body_parentid"] = body_parent; A["body_rootid"] = body_rootid; A["body_weldid"] = body_weldid
    A["body_jntnum"] = body_jntnum; A["body_jntadr"] = body_jntadr; A["body_dofnum"] = body_dofnum; A["body_dofadr"] = body_dofadr
    A["body_geomnum"] = body_geomnum; A["body_geomadr"] = body_geomadr
    A["body_pos"] = f64([b.pos for b in sc.bodies], (nbody, 3)); A["body_quat"] = f64([b.quat for b in sc.bodies], (nbody, 4))
    A["body_ipos"] = body_ipos; A["body_iquat"] = body_iquat; A["body_mass"] = body_mass; A["body_inertia"] = body_inertia
    A["body_gravcomp"] = f64([b.gravcomp for b in sc.bodies])
    A["jnt_type"] = i32(jnt_type); A["jnt_bodyid"] = i32(jnt_body); A["jnt_qposadr"] = i32(jnt_qposadr); A["jnt_dofadr"] = i32(jnt_dofadr)
    A["jnt_pos"] = f64(jnt_pos, (njnt, 3)); A["jnt_axis"] = f64(jnt_axis, (njnt, 3)); A["jnt_stiffness"] = f64(jnt_stiffness)
    A["jnt_range"] = f64(jnt_range, (njnt, 2)); A["jnt_limited"] = i32(jnt_limited); A["jnt_margin"] = f64(jnt_margin)
    A["jnt_solref"] = f64(jnt_solref, (njnt, 2)); A["jnt_solimp"] = f64(jnt_solimp, (njnt, 5))
    A["qpos0"] = f64(qpos0); A["qpos_spring"] = f64(qpos_spring)
    A["dof_bodyid"] = i32(dof_body); A["dof_jntid"] = i32(dof_jnt); A["dof_parentid"] = dof_parent; A["dof_Madr"] = dof_Madr
    A["dof_armature"] = f64(dof_armature); A["dof_damping"] = f64(dof_damping); A["dof_frictionloss"] = f64(dof_floss)
    A["dof_solref"] = f64(dof_solref, (nv, 2)); A["dof_solimp"] = f64(dof_solimp, (nv, 5))
    for k in ("type", "contype", "conaffinity", "condim", "bodyid", "dataid", "priority", "group", "matid"):
        A["geom_" + k] = i32(G[k])
    for k, w in (("size", 3), ("aabb", 6), ("pos", 3), ("quat", 4), ("friction", 3), ("solref", 2), ("solimp", 5), ("rgba", 4)):
        A["geom_" + k] = f64(G[k], (ngeom, w))
    for k in ("rbound", "solmix", "margin", "gap"):
        A["geom_" + k] = f64(G[k])
    A["site_bodyid"] = i32(site_body); A["site_pos"] = f64(site_pos, (len(site_body), 3)); A["site_quat"] = f64(site_quat, (len(site_body), 4))
    ncam = len(cam_body)
    A["cam_bodyid"] = i32(cam_body); A["cam_pos"] = f64(cam_pos, (ncam, 3)); A["cam_quat"] = f64(cam_quat, (ncam, 4)); A["cam_fovy"] = f64(cam_fovy)
    nl = len(light_body)
    A["light_bodyid"] = i32(light_body); A["light_pos"] = f64(light_pos, (nl, 3)); A["light_dir"] = f64(light_dir, (nl, 3))
    A["light_directional"] = i32(light_directional); A["light_ambient"] = f64(light_ambient, (nl, 3))
    A["light_diffuse"] = f64(light_diffuse, (nl, 3)); A["light_specular"] = f64(light_specular, (nl, 3))
    A["tendon_adr"] = i32(ten_adr); A["tendon_num"] = i32(ten_num); A["wrap_objid"] = i32(wrap_jnt); A["wrap_prm"] = f64(wrap_coef)
    neq = len(eq_obj1)
    A["eq_obj1id"] = i32(eq_obj1); A["eq_obj2id"] = i32(eq_obj2); A["eq_data"] = f64(eq_data, (neq, 5))
    A["eq_solref"] = f64(eq_solref, (neq, 2)); A["eq_solimp"] = f64(eq_solimp, (neq, 5)); A["eq_active0"] = i32(eq_active)
    A["actuator_trntype"] = act_trntype; A["actuator_trnid"] = act_trnid; A["actuator_gear"] = act_gear
    A["actuator_gainprm"] = act_gain; A["actuator_biasprm"] = act_bias; A["actuator_ctrlrange"] = act_ctrlrange
    A["actuator_ctrllimited"] = act_ctrllimited; A["actuator_forcerange"] = act_forcerange; A["actuator_forcelimited"] = act_forcelimited
    A["sensor_type"] = i32(sens_type); A["sensor_objid"] = i32(sens_obj); A["sensor_adr"] = i32(sens_adr)
    A["sensor_dim"] = i32(sens_dim); A["sensor_cutoff"] = f64(sens_cutoff)
    A["key_ctrl"] = key_ctrl; A["key_qpos"] = key_qpos
    ex = []
    for b1, b2 in sc.excludes:
        i1, i2 = body_names.index(b1), body_names.index(b2)
        ex.append((min(i1, i2) << 16) + max(i1, i2))
    A["exclude_signature"] = i32(sorted(ex))
    # convex hulls of collision meshes (per mesh id; -1 adr when the mesh never collides)
    hull_adr, hull_num, hull_verts = [], [], []
    for ma in mesh_list:
        if ma.hull_verts is not None:
            hull_adr.append(sum(len(h) for h in hull_verts)); hull_num.append(len(ma.hull_verts)); hull_verts.append(ma.hull_verts)
        else:
            hull_adr.append(-1); hull_num.append(0)
    A["mesh_hulladr"] = i32(hull_adr); A["mesh_hullnum"] = i32(hull_num)
    A["hull_vert"] = f64(np.concatenate(hull_verts) if hull_verts else np.zeros((0, 3)), (-1, 3))
    # hull vertex adjacency ("mesh graph"): hull_edge[hull_edgeadr[v] : hull_edgeadr[v + 1]] are the hull-local ids
    # of the neighbours of hull vertex v (v indexes hull_vert globally).  UPSTREAM-ASSUMPTION (mjCMesh::MakeGraph):
    # a vertex's neighbours are listed in the order of the qhull facets that contain it, deduplicated; the
    # plane-mesh collider walks this list to add up to three more contacts (oracle/ss_oracle_collision.c:plane_mesh).
    edge_adr, edges = [0], []
    for ma in mesh_list:
        if ma.hull_verts is None:
            continue
        nbr: list[list[int]] = [[] for _ in range(len(ma.hull_verts))]
        for f in ma.hull_faces:
            for k in range(3):
                v = int(f[k])
                for w in (int(f[(k + 1) % 3]), int(f[(k + 2) % 3])):
                    if w not in nbr[v]:
                        nbr[v].append(w)
        for lst in nbr:
            edges += lst
            edge_adr.append(len(edges))
    A["hull_edgeadr"] = i32(edge_adr); A["hull_edge"] = i32(edges)
    vis = sc.visual
    stat_extent = sc.statistic.get("extent", [None])[0]
    A["vis_headlight"] = f64([vis["headlight"]["ambient"], vis["headlight"]["diffuse"], vis["headlight"]["specular"]], (3, 3))
    A["vis_headlight_active"] = i32([vis["headlight"]["active"]])
    A["vis_haze"] = f64(vis["haze"]); A["vis_map"] = f64([vis["znear"], vis["zfar"]])
    sky = [t for t in sc.textures.values() if t.get("type") == "skybox"]
    if sky:
        A["skybox_rgb"] = f64([_floats(sky[0].get("rgb1", "0.8 0.8 0.8"), 3), _floats(sky[0].get("rgb2", "0.5 0.5 0.5"), 3)], (2, 3))
    else:
        A["skybox_rgb"] = np.zeros((0, 3))
    mat_rows = []
    for n in sc.material_order:
        mt = sc.materials[n]
        mat_rows.append(_floats(mt["rgba"], 4) + [float(mt["specular"]), float(mt["shininess"]), float(mt["reflectance"]),
                                                float(mt["emission"])])
    A["mat_prm"] = f64(mat_rows, (len(mat_rows), 8))

    m.names = {OBJ_BODY: body_names, OBJ_JOINT: jnt_names, OBJ_GEOM: G["names"], OBJ_SITE: site_names,
               OBJ_CAMERA: cam_names, OBJ_ACTUATOR: [a["name"] for a in sc.actuators], OBJ_SENSOR: sens_names,
               OBJ_KEY: key_names, OBJ_MESH: list(mesh_ids.keys()), OBJ_TENDON: ten_names}
    m.mesh_assets = mesh_list
    m.scene = sc

    _set_const(m, stat_extent)
    _collision_pairs(m)
    A["sizes"] = i32([nq, nv, nu, nbody, njnt, ngeom, len(site_body), ncam, len(ten_adr), neq, len(sens_type),
                      nsensordata, len(sc.keys), nM, len(A["pair_geom1"]), len(mesh_list)])
    if with_render:
        from .raygeom import build_ray_geometry
        build_ray_geometry(m, verbose=verbose)
    return m


# ----------------------------------------------------------------------------- compile-time dynamics

def fk_numpy(m: Model, qpos: np.ndarray):
    """Plain forward kinematics (used for compile-time constants and as a test cross-check)."""
    A = m.arrays
    nbody = len(A["body_parentid"])
    xpos = np.zeros((nbody, 3)); xquat = np.tile([1.0, 0, 0, 0], (nbody, 1))
    anchors, axes = {}, {}
    for b in range(1, nbody):
        p = A["body_parentid"][b]
        jn, ja = A["body_jntnum"][b], A["body_jntadr"][b]
        if jn == 1 and A["jnt_type"][ja] == JNT_FREE:
            qa = A["jnt_qposadr"][ja]
            xpos[b] = qpos[qa:qa + 3]; xquat[b] = quat_norm(qpos[qa + 3:qa + 7])
            continue
        pos = xpos[p] + quat_rot(xquat[p], A["body_pos"][b]); quat = quat_mul(xquat[p], A["body_quat"][b])
        for j in range(ja, ja + jn):
            qa = A["jnt_qposadr"][j]
            anchor = pos + quat_rot(quat, A["jnt_pos"][j]); axis = quat_rot(quat, A["jnt_axis"][j])
            if A["jnt_type"][j] == JNT_SLIDE:
                pos = pos + axis * (qpos[qa] - A["qpos0"][qa])
            elif A["jnt_type"][j] == JNT_HINGE:
                ang = qpos[qa] - A["qpos0"][qa]
                quat = quat_mul(quat, np.concatenate([[math.cos(ang / 2)], A["jnt_axis"][j] * math.sin(ang / 2)]))
                pos = anchor - quat_rot(quat, A["jnt_pos"][j])
            else:
                raise ValueError("ball joints are not supported")
            anchors[j], axes[j] = anchor, axis
        xpos[b] = pos; xquat[b] = quat_norm(quat)
    return xpos, xquat, anchors, axes


def mass_matrix_numpy(m: Model, qpos: np.ndarray):
    """Dense joint-space inertia from body Jacobians (independent of the CRBA used at run time)."""
    A = m.arrays
    nbody, nv = len(A["body_parentid"]), len(A["dof_bodyid"])
    xpos, xquat, anchors, axes = fk_numpy(m, qpos)
    xipos = np.array([xpos[b] + quat_rot(xquat[b], A["body_ipos"][b]) for b in range(nbody)])
    Jp = np.zeros((nbody, 3, nv)); Jr = np.zeros((nbody, 3, nv))
    for b in range(1, nbody):
        k = b
        while k > 0:
            for j in range(A["body_jntadr"][k], A["body_jntadr"][k] + A["body_jntnum"][k]):
                d = A["jnt_dofadr"][j]
                t = A["jnt_type"][j]
                if t == JNT_FREE:
                    R = quat2mat(xquat[k])
                    Jp[b, :, d:d + 3] = np.eye(3)
                    for a in range(3):
                        Jr[b, :, d + 3 + a] = R[:, a]
                        Jp[b, :, d + 3 + a] = np.cross(R[:, a], xipos[b] - xpos[k])
                elif t == JNT_SLIDE:
                    Jp[b, :, d] = axes[j]
                else:
                    Jr[b, :, d] = axes[j]; Jp[b, :, d] = np.cross(axes[j], xipos[b] - anchors[j])
            k = A["body_parentid"][k]
    M = np.diag(A["dof_armature"]).astype(np.float64)
    for b in range(1, nbody):
        R = quat2mat(quat_mul(xquat[b], A["body_iquat"][b]))
        Iw = R @ np.diag(A["body_inertia"][b]) @ R.T
        M += A["body_mass"][b] * Jp[b].T @ Jp[b] + Jr[b].T @ Iw @ Jr[b]
    return M, Jp, Jr, xpos, xquat, xipos


def _set_const(m: Model, stat_extent):
    """Compile-time constants at qpos0: invweight0 (constraint regularisation), meaninertia."""
    A = m.arrays
    nbody, nv = len(A["body_parentid"]), len(A["dof_bodyid"])
    M, Jp, Jr, xpos, xquat, xipos = mass_matrix_numpy(m, A["qpos0"])
    Minv = np.linalg.inv(M) if nv else np.zeros((0, 0))
    biw = np.zeros((nbody, 2))
    for b in range(1, nbody):
        if A["body_weldid"][b] == 0:
            continue
        biw[b, 0] = np.trace(Jp[b] @ Minv @ Jp[b].T) / 3
        biw[b, 1] = np.trace(Jr[b] @ Minv @ Jr[b].T) / 3
    diw = np.zeros(nv)
    for j in range(len(A["jnt_type"])):
        d = A["jnt_dofadr"][j]
        t = A["jnt_type"][j]
        if t == JNT_FREE:
            diw[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3]); diw[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        elif t == JNT_BALL:
            diw[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
        else:
            diw[d] = Minv[d, d]
    A["body_invweight0"] = biw; A["dof_invweight0"] = diw
    A["stat_meaninertia"] = np.array([np.mean(np.diag(M)) if nv else 1.0])
    sub = A["body_mass"].copy()
    for b in range(nbody - 1, 0, -1):
        sub[A["body_parentid"][b]] += sub[b]
    A["body_subtreemass"] = sub
    if stat_extent is None:
        # half-diagonal of the AABB over geom centres ± rbound (planes excluded), floor at 2x the largest rbound
        lo, hi = np.full(3, 1e30), np.full(3, -1e30)
        for g in range(len(A["geom_type"])):
            if A["geom_type"][g] == 0:
                continue
            b = A["geom_bodyid"][g]
            c = xpos[b] + quat_rot(xquat[b], A["geom_pos"][g])
            lo = np.minimum(lo, c - A["geom_rbound"][g]); hi = np.maximum(hi, c + A["geom_rbound"][g])
        stat_extent = max(0.5 * float(np.linalg.norm(hi - lo)), 1e-5) if np.all(hi > lo) else 1.0
    A["stat_extent"] = np.array([float(stat_extent)])


def _collision_pairs(m: Model):
    """Statically filtered candidate geom pairs in the reference's enumeration order.

    UPSTREAM-ASSUMPTION (SURVEY.md Appendix B "Pair filtering"): body pairs ascending
    (body1 < body2), geoms in id order inside a pair; filters = same weld group, weld
    parent–child (unless one side is world-welded), contype/conaffinity mask, <exclude>.
    The pair is stored with the lower geom-type enum first (that geom is ``geom1`` of the contact).
    """
    A = m.arrays
    nbody = len(A["body_parentid"])
    weld = A["body_weldid"]; par = A["body_parentid"]
    excl = set(int(x) for x in A["exclude_signature"])
    g1s, g2s = [], []
    for b1 in range(nbody):
        if A["body_geomnum"][b1] == 0:
            continue
        for b2 in range(b1 + 1, nbody):
            if A["body_geomnum"][b2] == 0:
                continue
            w1, w2 = weld[b1], weld[b2]
            if w1 == w2:
                continue
            wp1, wp2 = weld[par[w1]], weld[par[w2]]
            if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
                continue
            if ((b1 << 16) + b2) in excl:
                continue
            for ga in range(A["body_geomadr"][b1], A["body_geomadr"][b1] + A["body_geomnum"][b1]):
                for gb in range(A["body_geomadr"][b2], A["body_geomadr"][b2] + A["body_geomnum"][b2]):
                    if not ((A["geom_contype"][ga] & A["geom_conaffinity"][gb]) or
                            (A["geom_contype"][gb] & A["geom_conaffinity"][ga])):
                        continue
                    if A["geom_type"][ga] <= A["geom_type"][gb]:
                        g1s.append(ga); g2s.append(gb)
                    else:
                        g1s.append(gb); g2s.append(ga)
    A["pair_geom1"] = np.asarray(g1s, dtype=np.int32); A["pair_geom2"] = np.asarray(g2s, dtype=np.int32)
    # mixed contact parameters per pair (mj_contactParam): condim, friction[5], solref[2], solimp[5], margin, gap
    n = len(g1s)
    condim = np.zeros(n, np.int32); fr = np.zeros((n, 5)); sref = np.zeros((n, 2)); simp = np.zeros((n, 5))
    margin = np.zeros(n); gap = np.zeros(n)
    for k, (a, b) in enumerate(zip(g1s, g2s)):
        p1, p2 = A["geom_priority"][a], A["geom_priority"][b]
        if p1 != p2:
            w = a if p1 > p2 else b
            condim[k] = A["geom_condim"][w]; f3 = A["geom_friction"][w]; sref[k] = A["geom_solref"][w]; simp[k] = A["geom_solimp"][w]
        else:
            condim[k] = max(A["geom_condim"][a], A["geom_condim"][b])
            s1, s2 = A["geom_solmix"][a], A["geom_solmix"][b]
            if s1 >= MINVAL and s2 >= MINVAL:
                mix = s1 / (s1 + s2)
            elif s1 < MINVAL and s2 < MINVAL:
                mix = 0.5
            else:
                mix = 0.0 if s1 < MINVAL else 1.0
            r1, r2 = A["geom_solref"][a], A["geom_solref"][b]
            sref[k] = mix * r1 + (1 - mix) * r2 if (r1[0] > 0 and r2[0] > 0) else np.minimum(r1, r2)
            simp[k] = mix * A["geom_solimp"][a] + (1 - mix) * A["geom_solimp"][b]
            f3 = np.maximum(A["geom_friction"][a], A["geom_friction"][b])
        f3 = np.maximum(f3, 1e-5)
        fr[k] = [f3[0], f3[0], f3[1], f3[2], f3[2]]
        margin[k] = max(A["geom_margin"][a], A["geom_margin"][b]); gap[k] = max(A["geom_gap"][a], A["geom_gap"][b])
    A["pair_condim"] = condim; A["pair_friction"] = fr; A["pair_solref"] = sref; A["pair_solimp"] = simp
    A["pair_margin"] = margin; A["pair_gap"] = gap


# ----------------------------------------------------------------------------- real-mjModel import hook
# the day a MuJoCo install is reachable, compile parity (this compiler vs
# MjModel.from_xml_path) and step parity can be checked separately by loading a DUMP of the real mjModel through the
# same blob format.  `dump_mjmodel_npz` runs where `import mujoco` works; `load_mjmodel_npz` runs anywhere.

# mjModel attributes that map one-to-one onto this compiler's arrays
_MJMODEL_FIELDS = [
    "body_parentid", "body_rootid", "body_weldid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr", "body_geomnum",
    "body_geomadr", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_gravcomp",
    "body_invweight0", "body_subtreemass", "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_pos", "jnt_axis",
    "jnt_stiffness", "jnt_range", "jnt_limited", "jnt_margin", "jnt_solref", "jnt_solimp", "qpos0", "qpos_spring",
    "dof_bodyid", "dof_jntid", "dof_parentid", "dof_Madr", "dof_armature", "dof_damping", "dof_frictionloss", "dof_solref",
    "dof_solimp", "dof_invweight0", "geom_type", "geom_contype", "geom_conaffinity", "geom_condim", "geom_bodyid", "geom_dataid",
    "geom_priority", "geom_group", "geom_matid", "geom_size", "geom_aabb", "geom_pos", "geom_quat", "geom_friction", "geom_solref",
    "geom_solimp", "geom_rgba", "geom_rbound", "geom_solmix", "geom_margin", "geom_gap", "site_bodyid", "site_pos", "site_quat",
    "cam_bodyid", "cam_pos", "cam_quat", "cam_fovy", "light_bodyid", "light_pos", "light_dir", "light_directional", "light_ambient",
    "light_diffuse", "light_specular", "tendon_adr", "tendon_num", "wrap_objid", "wrap_prm", "eq_obj1id", "eq_obj2id", "eq_data",
    "eq_solref", "eq_solimp", "eq_active0", "actuator_trntype", "actuator_trnid", "actuator_gainprm", "actuator_biasprm",
    "actuator_ctrlrange", "actuator_ctrllimited", "actuator_forcerange", "actuator_forcelimited", "sensor_type", "sensor_objid",
    "sensor_adr", "sensor_dim", "sensor_cutoff", "key_ctrl", "key_qpos", "exclude_signature"]
_MJ_SENSOR = {39: SENS_GYRO, 1: SENS_ACCEL, 7: SENS_RANGE}   # placeholders, overwritten by the dump's own enum table
_OBJ_NAMES = {OBJ_BODY: "body", OBJ_JOINT: "jnt", OBJ_GEOM: "geom", OBJ_SITE: "site", OBJ_CAMERA: "cam", OBJ_ACTUATOR: "actuator",
              OBJ_SENSOR: "sensor", OBJ_KEY: "key", OBJ_MESH: "mesh", OBJ_TENDON: "tendon"}


def dump_mjmodel_npz(mjmodel, path: str) -> None:
    """Run where `import mujoco` works: flat .npz of the named mjModel arrays this engine consumes."""
    import mujoco
    out = {}
    for f in _MJMODEL_FIELDS:
        if hasattr(mjmodel, f):
            out[f] = np.asarray(getattr(mjmodel, f))
    o = mjmodel.opt
    out.update(opt_timestep=o.timestep, opt_gravity=np.asarray(o.gravity), opt_impratio=o.impratio, opt_tolerance=o.tolerance,
               opt_ls_tolerance=o.ls_tolerance, opt_iterations=o.iterations, opt_ls_iterations=o.ls_iterations, opt_cone=int(o.cone),
               opt_solver=int(o.solver), opt_multiccd=int(bool(o.enableflags & mujoco.mjtEnableBit.mjENBL_MULTICCD)),
               stat_meaninertia=mjmodel.stat.meaninertia, stat_extent=mjmodel.stat.extent)
    out["sensor_enum"] = np.array([int(mujoco.mjtSensor.mjSENS_GYRO), int(mujoco.mjtSensor.mjSENS_ACCELEROMETER),
                                   int(mujoco.mjtSensor.mjSENS_RANGEFINDER)])
    for k in ("mesh_vert", "mesh_vertadr", "mesh_vertnum", "mesh_graph", "mesh_graphadr", "actuator_gear"):
        out[k] = np.asarray(getattr(mjmodel, k))
    for objtype, prefix in _OBJ_NAMES.items():
        mjobj = {"body": mujoco.mjtObj.mjOBJ_BODY, "jnt": mujoco.mjtObj.mjOBJ_JOINT, "geom": mujoco.mjtObj.mjOBJ_GEOM,
                 "site": mujoco.mjtObj.mjOBJ_SITE, "cam": mujoco.mjtObj.mjOBJ_CAMERA, "actuator": mujoco.mjtObj.mjOBJ_ACTUATOR,
                 "sensor": mujoco.mjtObj.mjOBJ_SENSOR, "key": mujoco.mjtObj.mjOBJ_KEY, "mesh": mujoco.mjtObj.mjOBJ_MESH,
                 "tendon": mujoco.mjtObj.mjOBJ_TENDON}[prefix]
        n = getattr(mjmodel, "n" + prefix)
        out["names_" + prefix] = np.array([mujoco.mj_id2name(mjmodel, mjobj, i) or "" for i in range(n)])
    np.savez_compressed(path, **out)


def load_mjmodel_npz(path: str) -> Model:
    """mjModel dump (dump_mjmodel_npz) -> `Model` with this compiler's array names; the candidate pair list, the
    per-pair contact parameters and the hull tables are derived here the same way compile_scene derives them.
    Physics only (no ray geometry)."""
    Z = np.load(path, allow_pickle=False)
    m = Model()
    A = m.arrays
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))
    ints = {"body_parentid", "body_rootid", "body_weldid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr", "body_geomnum",
            "body_geomadr", "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_limited", "dof_bodyid", "dof_jntid",
            "dof_parentid", "dof_Madr", "geom_type", "geom_contype", "geom_conaffinity", "geom_condim", "geom_bodyid", "geom_dataid",
            "geom_priority", "geom_group", "geom_matid", "site_bodyid", "cam_bodyid", "light_bodyid", "light_directional", "tendon_adr",
            "tendon_num", "wrap_objid", "eq_obj1id", "eq_obj2id", "eq_active0", "actuator_trntype", "actuator_trnid",
            "actuator_ctrllimited", "actuator_forcelimited", "sensor_type", "sensor_objid", "sensor_adr", "sensor_dim", "exclude_signature"}
    for f in _MJMODEL_FIELDS:
        if f in Z.files:
            A[f] = i32(Z[f]) if f in ints else f64(Z[f])
    for k in ("opt_timestep", "opt_impratio", "opt_tolerance", "opt_ls_tolerance", "stat_meaninertia", "stat_extent"):
        A[k] = f64([float(Z[k])])
    A["opt_gravity"] = f64(Z["opt_gravity"])
    for k in ("opt_iterations", "opt_ls_iterations", "opt_cone", "opt_solver", "opt_multiccd"):
        A[k] = i32([int(Z[k])])
    # mjModel keeps 3 gain / 3 bias parameters of 10, a 6-vector gear, row vectors for wrap_prm ... : trim to this engine's shapes
    if "actuator_gear" in Z.files:
        A["actuator_gear"] = f64(np.asarray(Z["actuator_gear"]).reshape(len(A["actuator_trnid"].reshape(-1, 2) if A["actuator_trnid"].ndim > 1 else A["actuator_trnid"]), -1)[:, 0])
    if A["actuator_trnid"].ndim > 1:
        A["actuator_trnid"] = i32(A["actuator_trnid"][:, 0])
    for k in ("actuator_gainprm", "actuator_biasprm"):
        A[k] = f64(A[k][:, :3])
    if "sensor_enum" in Z.files:
        gy, ac, rf = [int(x) for x in Z["sensor_enum"]]
        A["sensor_type"] = i32([{gy: SENS_GYRO, ac: SENS_ACCEL, rf: SENS_RANGE}.get(int(t), -1) for t in A["sensor_type"]])
    # convex hulls from the mesh graphs (graph = numvert, numface, vert_edgeadr[nv], vert_globalid[nv], edge_localid[nv + 3 nf], face_globalid[3 nf])
    nmesh = len(Z["mesh_vertadr"]) if "mesh_vertadr" in Z.files else 0
    hull_adr, hull_num, hull_verts, edge_adr, edges = [], [], [], [0], []
    for k in range(nmesh):
        ga = int(Z["mesh_graphadr"][k])
        if ga < 0:
            hull_adr.append(-1); hull_num.append(0)
            continue
        G = np.asarray(Z["mesh_graph"][ga:])
        nvh, nfh = int(G[0]), int(G[1])
        vert_edgeadr = G[2:2 + nvh]; vert_globalid = G[2 + nvh:2 + 2 * nvh]; edge_localid = G[2 + 2 * nvh:2 + 3 * nvh + 3 * nfh]
        V = np.asarray(Z["mesh_vert"]).reshape(-1, 3)[int(Z["mesh_vertadr"][k]):][vert_globalid]
        hull_adr.append(sum(len(h) for h in hull_verts)); hull_num.append(nvh); hull_verts.append(V.astype(np.float64))
        for v in range(nvh):
            e = int(vert_edgeadr[v])
            while edge_localid[e] >= 0:
                edges.append(int(edge_localid[e])); e += 1
            edge_adr.append(len(edges))
    A["mesh_hulladr"] = i32(hull_adr); A["mesh_hullnum"] = i32(hull_num)
    A["hull_vert"] = f64(np.concatenate(hull_verts) if hull_verts else np.zeros((0, 3))).reshape(-1, 3)
    A["hull_edgeadr"] = i32(edge_adr); A["hull_edge"] = i32(edges)
    m.names = {objtype: [str(s) for s in Z["names_" + prefix]] if "names_" + prefix in Z.files else [] for objtype, prefix in _OBJ_NAMES.items()}
    _collision_pairs(m)
    nq, nv, nu = len(A["qpos0"]), len(A["dof_bodyid"]), len(A["actuator_trnid"])
    nM = 0
    for d in range(nv):
        k = d
        while k >= 0:
            nM += 1
            k = int(A["dof_parentid"][k])
    A["sizes"] = i32([nq, nv, nu, len(A["body_parentid"]), len(A["jnt_type"]), len(A["geom_type"]), len(A["site_bodyid"]),
                      len(A["cam_bodyid"]), len(A["tendon_adr"]), len(A["eq_obj1id"]), len(A["sensor_type"]),
                      int(A["sensor_dim"].sum()) if len(A["sensor_dim"]) else 0, len(A["key_ctrl"]), nM,
                      len(A["pair_geom1"]), nmesh])
    return m


def compare_models(ours: Model, theirs: Model, rtol: float = 1e-9) -> dict:
    """{array name: max abs difference (or 'shape a vs b')} for every array the two models share: compile parity report."""
    out = {}
    for k, a in ours.arrays.items():
        if k not in theirs.arrays:
            continue
        b = theirs.arrays[k]
        if a.shape != b.shape:
            out[k] = f"shape {a.shape} vs {b.shape}"
        elif a.size and not np.allclose(a, b, rtol=rtol, atol=rtol):
            out[k] = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
    return out
