"""MJCF scene loader: XML → flat element lists (bodies, joints, geoms, ...).

This is row T1 of SURVEY.md §8(a): the input contract is the reference's own model files
(`stretch_mujoco/models/stretch.xml`, `scene.xml`, `docking_station.xml`), which the reference
passes to ``MjModel.from_xml_path`` (`stretch_mujoco/mujoco_server.py:252`) or
``from_xml_string`` (`stretch_mujoco/robocasa_gen.py:232`).  The feature subset is the one listed
in SURVEY.md Appendix A.3.  Nothing here touches the GPU; `compiler.py` turns the element lists
into the numeric model arrays.
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------- small math


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def quat_rot(q, v):
    return quat2mat(q) @ np.asarray(v, dtype=np.float64)


def quat2mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_norm(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    if n < 1e-14:
        return np.array([1.0, 0, 0, 0])
    return q / n


def axisangle_quat(axis, ang):
    axis = np.asarray(axis, dtype=np.float64)
    return np.concatenate([[math.cos(ang / 2)], axis * math.sin(ang / 2)])


def euler_quat(e, seq="xyz"):
    """Intrinsic rotations for lower-case sequence letters (MJCF ``eulerseq`` default "xyz")."""
    q = np.array([1.0, 0, 0, 0])
    for ang, ch in zip(e, seq):
        ax = {"x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1)}[ch.lower()]
        r = axisangle_quat(ax, ang)
        q = quat_mul(q, r) if ch.islower() else quat_mul(r, q)
    return q


def zaxis_quat(vec):
    """Minimal rotation taking +z to ``vec`` (MJCF ``zaxis`` attribute)."""
    vec = np.asarray(vec, dtype=np.float64)
    vec = vec / np.linalg.norm(vec)
    axis = np.cross([0, 0, 1.0], vec)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0, 0])
    else:
        axis = axis / s
    ang = math.atan2(s, vec[2])
    return axisangle_quat(axis, ang)


def _floats(s, n=None):
    v = [float(x) for x in s.replace(",", " ").split()]
    if n is not None and len(v) < n:
        v = v + [0.0] * (n - len(v))
    return v


# ----------------------------------------------------------------------------- data classes

@dataclass
class Body:
    name: str
    parent: int
    pos: np.ndarray
    quat: np.ndarray
    gravcomp: float = 0.0
    childclass: str | None = None
    joints: list = field(default_factory=list)
    geoms: list = field(default_factory=list)
    sites: list = field(default_factory=list)
    cameras: list = field(default_factory=list)
    lights: list = field(default_factory=list)
    inertial: dict | None = None


ACT_DEFAULT = dict(gainprm=[1.0, 0, 0], biasprm=[0.0, 0, 0], gear=[1.0, 0, 0, 0, 0, 0], gaintype="fixed",
                   biastype="none", dyntype="none", ctrlrange=None, forcerange=None, ctrllimited=None,
                   forcelimited=None)

GEOM_DEFAULT = dict(type="sphere", contype="1", conaffinity="1", condim="3", group="0", priority="0",
                    friction="1 0.005 0.0001", solmix="1", solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2",
                    margin="0", gap="0", density="1000", rgba="0.5 0.5 0.5 1")
JOINT_DEFAULT = dict(type="hinge", axis="0 0 1", pos="0 0 0", damping="0", stiffness="0", armature="0",
                     frictionloss="0", springref="0", ref="0", margin="0", solreflimit="0.02 1",
                     solimplimit="0.9 0.95 0.001 0.5 2", solreffriction="0.02 1",
                     solimpfriction="0.9 0.95 0.001 0.5 2")
EQ_DEFAULT = dict(solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", active="true")
MAT_DEFAULT = dict(rgba="1 1 1 1", specular="0.5", shininess="0.5", reflectance="0", emission="0",
                   texrepeat="1 1", texuniform="false")
CAM_DEFAULT = dict(fovy="45", pos="0 0 0")
SITE_DEFAULT = dict(pos="0 0 0", size="0.005", rgba="0.5 0.5 0.5 1", group="0", type="sphere")
LIGHT_DEFAULT = dict(pos="0 0 0", dir="0 0 -1", directional="false", castshadow="true", active="true",
                     ambient="0 0 0", diffuse="0.7 0.7 0.7", specular="0.3 0.3 0.3")


class DefaultClass:
    def __init__(self, name, parent=None):
        self.name = name
        self.parent = parent
        if parent is None:
            self.attrs = {"geom": dict(GEOM_DEFAULT), "joint": dict(JOINT_DEFAULT), "site": dict(SITE_DEFAULT),
                          "camera": dict(CAM_DEFAULT), "light": dict(LIGHT_DEFAULT), "mesh": {},
                          "material": dict(MAT_DEFAULT), "equality": dict(EQ_DEFAULT), "tendon": {}}
            self.act = {k: (list(v) if isinstance(v, list) else v) for k, v in ACT_DEFAULT.items()}
        else:
            self.attrs = {k: dict(v) for k, v in parent.attrs.items()}
            self.act = {k: (list(v) if isinstance(v, list) else v) for k, v in parent.act.items()}


def apply_actuator(tag: str, a: dict, act: dict) -> None:
    """Update an actuator struct from one <general|motor|position|velocity> element.

    UPSTREAM-ASSUMPTION: every actuator shortcut shares ONE default struct per class and the
    shortcut handler runs identically inside <default> and <actuator>; <position> overwrites
    ``biasprm[1] = -kp`` and, only when ``kv`` is given, ``biasprm[2] = -kv`` -- so a class-level
    ``kv`` (finger classes) or ``biasprm[2]`` (lift/telescope, `stretch.xml:22,26`) survives:
    lift force = 400(ctrl - q) - 100 qdot, arm = 150(ctrl - L) - 10 Ldot.
    """
    if "ctrlrange" in a:
        act["ctrlrange"] = _floats(a["ctrlrange"], 2)
    if "forcerange" in a:
        act["forcerange"] = _floats(a["forcerange"], 2)
    if "ctrllimited" in a:
        act["ctrllimited"] = a["ctrllimited"]
    if "forcelimited" in a:
        act["forcelimited"] = a["forcelimited"]
    if "gear" in a:
        g = _floats(a["gear"])
        act["gear"] = g + [0.0] * (6 - len(g))
    if tag == "general":
        if "gainprm" in a:
            g = _floats(a["gainprm"])
            act["gainprm"] = (g + [0.0] * 3)[:3]
        if "biasprm" in a:
            g = _floats(a["biasprm"])
            act["biasprm"] = (g + [0.0] * 3)[:3]
        if "gaintype" in a:
            act["gaintype"] = a["gaintype"]
        if "biastype" in a:
            act["biastype"] = a["biastype"]
        if "dyntype" in a:
            act["dyntype"] = a["dyntype"]
    elif tag == "motor":
        act["gainprm"] = [1.0, 0, 0]
        act["biasprm"] = [0.0, 0, 0]
        act["gaintype"], act["biastype"], act["dyntype"] = "fixed", "none", "none"
    elif tag == "position":
        if "kp" in a:
            act["gainprm"][0] = float(a["kp"])
        act["biasprm"][1] = -act["gainprm"][0]
        if "kv" in a:
            act["biasprm"][2] = -float(a["kv"])
        act["gaintype"], act["biastype"], act["dyntype"] = "fixed", "affine", "none"
    elif tag == "velocity":
        if "kv" in a:
            act["gainprm"][0] = float(a["kv"])
        act["biasprm"][2] = -act["gainprm"][0]
        act["gaintype"], act["biastype"], act["dyntype"] = "fixed", "affine", "none"
    else:
        raise ValueError(f"unsupported actuator element <{tag}>")
    if act["dyntype"] != "none" or act["gaintype"] != "fixed" or act["biastype"] not in ("none", "affine"):
        raise ValueError("only stateless fixed-gain / affine-bias actuators are supported")


class Scene:
    """Parsed MJCF: flat lists in MuJoCo id order (bodies depth-first, elements grouped by body)."""

    def __init__(self):
        self.model_name = ""
        self.angle = "degree"
        self.eulerseq = "xyz"
        self.assetdir = ""
        self.meshdir = None
        self.texturedir = None
        self.autolimits = True
        self.option = dict(timestep=0.002, gravity=[0, 0, -9.81], integrator="Euler", cone="pyramidal",
                           impratio=1.0, solver="Newton", iterations=100, tolerance=1e-8, ls_iterations=50,
                           ls_tolerance=0.01, jacobian="auto", multiccd=False, o_margin=0.0)
        self.statistic = {}
        self.visual = dict(headlight=dict(ambient=[0.1, 0.1, 0.1], diffuse=[0.4, 0.4, 0.4], specular=[0.5, 0.5, 0.5],
                                          active=1), haze=[1, 1, 1, 1], fog=[0, 0, 0, 1], znear=0.01, zfar=50.0,
                           fogstart=3.0, fogend=10.0, offwidth=640, offheight=480)
        self.defaults: dict[str, DefaultClass] = {"main": DefaultClass("main")}
        self.bodies: list[Body] = [Body("world", -1, np.zeros(3), np.array([1.0, 0, 0, 0]))]
        self.meshes: dict[str, dict] = {}
        self.mesh_order: list[str] = []
        self.textures: dict[str, dict] = {}
        self.materials: dict[str, dict] = {}
        self.material_order: list[str] = []
        self.excludes: list[tuple[str, str]] = []
        self.pairs: list[dict] = []
        self.tendons: list[dict] = []
        self.equalities: list[dict] = []
        self.actuators: list[dict] = []
        self.sensors: list[dict] = []
        self.keys: list[dict] = []
        self.base_dir = "."

    # ------------------------------------------------------------------ parse entry points
    @classmethod
    def from_xml_path(cls, path: str) -> "Scene":
        sc = cls()
        sc.base_dir = os.path.dirname(os.path.abspath(path))
        root = ET.parse(path).getroot()
        sc._expand_includes(root, sc.base_dir)
        sc._parse_root(root)
        return sc

    @classmethod
    def from_xml_string(cls, xml: str, base_dir: str = ".", replicate_override: dict | None = None) -> "Scene":
        """``replicate_override`` maps the name of an element inside a <replicate> to ``(count, euler)``: the
        benchmark's 1000-ray lidar re-spins the reference's 360-site ring without
        editing the reference's stretch.xml."""
        sc = cls()
        sc.replicate_override = dict(replicate_override or {})
        sc.base_dir = os.path.abspath(base_dir)
        root = ET.fromstring(xml)
        sc._expand_includes(root, sc.base_dir)
        sc._parse_root(root)
        return sc

    def _expand_includes(self, elem, base_dir):
        """<include file=…/> is replaced in place by the children of the included file's root."""
        i = 0
        while i < len(elem):
            ch = elem[i]
            if ch.tag == "include":
                sub = ET.parse(os.path.join(base_dir, ch.attrib["file"])).getroot()
                self._expand_includes(sub, base_dir)
                elem.remove(ch)
                for k, g in enumerate(list(sub)):
                    elem.insert(i + k, g)
                i += len(sub)
            else:
                self._expand_includes(ch, base_dir)
                i += 1

    # ------------------------------------------------------------------ helpers
    def _ang(self, x):
        return x * math.pi / 180.0 if self.angle == "degree" else x

    def _orient(self, a: dict):
        if "quat" in a:
            return quat_norm(_floats(a["quat"], 4))
        if "euler" in a:
            return euler_quat([self._ang(x) for x in _floats(a["euler"], 3)], self.eulerseq)
        if "zaxis" in a:
            return zaxis_quat(_floats(a["zaxis"], 3))
        if "axisangle" in a:
            v = _floats(a["axisangle"], 4)
            ax = np.asarray(v[:3])
            return axisangle_quat(ax / np.linalg.norm(ax), self._ang(v[3]))
        if "xyaxes" in a:
            v = _floats(a["xyaxes"], 6)
            x = np.asarray(v[:3]); x = x / np.linalg.norm(x)
            y = np.asarray(v[3:]); y = y - x * (x @ y); y = y / np.linalg.norm(y)
            from .meshio import mat2quat
            return mat2quat(np.stack([x, y, np.cross(x, y)], axis=1))
        return np.array([1.0, 0, 0, 0])

    def _resolve(self, tag: str, elem_attrs: dict, childclass: str | None) -> dict:
        cname = elem_attrs.get("class") or childclass or "main"
        if cname not in self.defaults:
            raise ValueError(f"unknown default class '{cname}'")
        merged = dict(self.defaults[cname].attrs.get(tag, {}))
        merged.update({k: v for k, v in elem_attrs.items() if k != "class"})
        return merged

    # ------------------------------------------------------------------ sections
    def _parse_root(self, root):
        self.model_name = root.attrib.get("model", "")
        # compiler/option/size first (they affect interpretation of everything else)
        for sec in root:
            if sec.tag == "compiler":
                a = sec.attrib
                self.angle = a.get("angle", self.angle)
                self.eulerseq = a.get("eulerseq", self.eulerseq)
                self.assetdir = a.get("assetdir", self.assetdir)
                self.meshdir = a.get("meshdir", self.meshdir)
                self.texturedir = a.get("texturedir", self.texturedir)
                if "autolimits" in a:
                    self.autolimits = a["autolimits"] == "true"
        for sec in root:
            t = sec.tag
            if t == "compiler" or t == "size":
                continue
            elif t == "option":
                self._parse_option(sec)
            elif t == "default":
                self._parse_default(sec, None)
            elif t == "asset":
                self._parse_asset(sec)
            elif t == "worldbody":
                self._parse_body_children(sec, 0, None, (np.zeros(3), np.array([1.0, 0, 0, 0])), "")
            elif t == "contact":
                for e in sec:
                    if e.tag == "exclude":
                        self.excludes.append((e.attrib["body1"], e.attrib["body2"]))
                    elif e.tag == "pair":
                        raise ValueError("<contact><pair> is not supported")
            elif t == "tendon":
                for e in sec:
                    if e.tag != "fixed":
                        raise ValueError("only <tendon><fixed> is supported")
                    self.tendons.append(dict(name=e.attrib.get("name", ""),
                                             joints=[(j.attrib["joint"], float(j.attrib["coef"])) for j in e]))
            elif t == "equality":
                for e in sec:
                    if e.tag != "joint":
                        raise ValueError(f"equality <{e.tag}> is not supported")
                    a = self._resolve("equality", e.attrib, None)
                    self.equalities.append(dict(type="joint", name=a.get("name", ""), joint1=a["joint1"],
                                                joint2=a.get("joint2"),
                                                polycoef=_floats(a.get("polycoef", "0 1 0 0 0"), 5),
                                                solref=_floats(a["solref"], 2), solimp=(_floats(a["solimp"]) + [0.9, 0.95, 0.001, 0.5, 2.0][len(_floats(a["solimp"])):])[:5],
                                                active=a.get("active", "true") == "true"))
            elif t == "actuator":
                for e in sec:
                    cname = e.attrib.get("class", "main")
                    base = self.defaults[cname].act
                    act = {k: (list(v) if isinstance(v, list) else v) for k, v in base.items()}
                    apply_actuator(e.tag, e.attrib, act)
                    act["name"] = e.attrib.get("name", "")
                    if "joint" in e.attrib:
                        act["trntype"], act["target"] = "joint", e.attrib["joint"]
                    elif "tendon" in e.attrib:
                        act["trntype"], act["target"] = "tendon", e.attrib["tendon"]
                    else:
                        raise ValueError("actuator transmission must be joint or tendon")
                    self.actuators.append(act)
            elif t == "sensor":
                for e in sec:
                    self.sensors.append(dict(type=e.tag, **e.attrib))
            elif t == "keyframe":
                for e in sec:
                    self.keys.append(dict(e.attrib))
            elif t == "statistic":
                self.statistic = {k: _floats(v) for k, v in sec.attrib.items()}
            elif t == "visual":
                for e in sec:
                    if e.tag == "headlight":
                        for k in ("ambient", "diffuse", "specular"):
                            if k in e.attrib:
                                self.visual["headlight"][k] = _floats(e.attrib[k], 3)
                        if "active" in e.attrib:
                            self.visual["headlight"]["active"] = int(e.attrib["active"])
                    elif e.tag == "rgba":
                        for k in ("haze", "fog"):
                            if k in e.attrib:
                                self.visual[k] = _floats(e.attrib[k], 4)
                    elif e.tag == "map":
                        for k in ("znear", "zfar", "fogstart", "fogend"):
                            if k in e.attrib:
                                self.visual[k] = float(e.attrib[k])
                    elif e.tag == "global":
                        for k in ("offwidth", "offheight"):
                            if k in e.attrib:
                                self.visual[k] = int(e.attrib[k])
            elif t in ("custom", "extension"):
                continue
            else:
                raise ValueError(f"unsupported MJCF section <{t}>")

    def _parse_option(self, sec):
        a = sec.attrib
        o = self.option
        for k in ("timestep", "impratio", "tolerance", "ls_tolerance", "o_margin"):
            if k in a:
                o[k] = float(a[k])
        for k in ("iterations", "ls_iterations"):
            if k in a:
                o[k] = int(a[k])
        for k in ("integrator", "cone", "solver", "jacobian"):
            if k in a:
                o[k] = a[k]
        if "gravity" in a:
            o["gravity"] = _floats(a["gravity"], 3)
        for e in sec:
            if e.tag == "flag" and "multiccd" in e.attrib:
                o["multiccd"] = e.attrib["multiccd"] == "enable"

    def _parse_default(self, sec, parent: DefaultClass | None):
        if parent is None:
            cls_ = self.defaults["main"]  # top-level <default> is always the "main" class
        else:
            name = sec.attrib["class"]
            cls_ = DefaultClass(name, parent)
            self.defaults[name] = cls_
        for e in sec:
            if e.tag == "default":
                continue
            if e.tag in ("general", "motor", "position", "velocity"):
                apply_actuator(e.tag, e.attrib, cls_.act)
            elif e.tag in cls_.attrs:
                cls_.attrs[e.tag].update(e.attrib)
            else:
                raise ValueError(f"unsupported default element <{e.tag}>")
        for e in sec:  # nested classes inherit the finished parent
            if e.tag == "default":
                self._parse_default(e, cls_)

    def _parse_asset(self, sec):
        for e in sec:
            a = dict(e.attrib)
            if e.tag == "mesh":
                a = self._resolve("mesh", a, None)
                name = a.get("name") or os.path.splitext(os.path.basename(a["file"]))[0]
                d = self.meshdir or self.assetdir
                a["path"] = os.path.join(self.base_dir, d, a["file"])
                a["scale"] = _floats(a.get("scale", "1 1 1"), 3)
                self.meshes[name] = a
                self.mesh_order.append(name)
            elif e.tag == "texture":
                name = a.get("name") or (os.path.splitext(os.path.basename(a["file"]))[0] if "file" in a else
                                         a.get("type", "tex"))
                if "file" in a:
                    d = self.texturedir or self.assetdir
                    a["path"] = os.path.join(self.base_dir, d, a["file"])
                self.textures[name] = a
            elif e.tag == "material":
                a = self._resolve("material", a, None)
                self.materials[a["name"]] = a
                self.material_order.append(a["name"])
            else:
                raise ValueError(f"unsupported asset <{e.tag}>")

    def _parse_body_children(self, elem, body_id: int, childclass, frame, suffix: str):
        """``frame`` = (pos, quat) of an enclosing <frame>/<replicate> relative to the body."""
        fpos, fquat = frame
        body = self.bodies[body_id]

        def place(a):
            p = np.asarray(_floats(a.get("pos", "0 0 0"), 3))
            q = self._orient(a)
            return fpos + quat_rot(fquat, p), quat_mul(fquat, q)

        for e in elem:
            t = e.tag
            if t == "body":
                a = e.attrib
                pos, quat = place(a)
                cc = a.get("childclass", childclass)
                b = Body(a.get("name", "") + suffix if a.get("name") else "", body_id, pos, quat_norm(quat),
                         float(a.get("gravcomp", 0.0)), cc)
                self.bodies.append(b)
                self._parse_body_children(e, len(self.bodies) - 1, cc,
                                          (np.zeros(3), np.array([1.0, 0, 0, 0])), suffix)
            elif t in ("joint", "freejoint"):
                if t == "freejoint":
                    j = dict(JOINT_DEFAULT, type="free", name=e.attrib.get("name", ""))
                else:
                    j = self._resolve("joint", e.attrib, childclass)
                j["name"] = e.attrib.get("name", "") + (suffix if e.attrib.get("name") else "")
                p = np.asarray(_floats(j.get("pos", "0 0 0"), 3))
                j["_pos"] = fpos + quat_rot(fquat, p)
                j["_axis"] = quat_rot(fquat, _floats(j["axis"], 3))
                body.joints.append(j)
            elif t == "geom":
                g = self._resolve("geom", e.attrib, childclass)
                g["_explicit"] = set(e.attrib.keys())
                g["name"] = e.attrib.get("name", "") + (suffix if e.attrib.get("name") else "")
                g["_pos"], g["_quat"] = place(g)
                if "fromto" in g:
                    raise ValueError("geom fromto is not supported")
                body.geoms.append(g)
            elif t == "site":
                s = self._resolve("site", e.attrib, childclass)
                s["name"] = e.attrib.get("name", "") + (suffix if e.attrib.get("name") else "")
                s["_pos"], s["_quat"] = place(s)
                body.sites.append(s)
            elif t == "camera":
                c = self._resolve("camera", e.attrib, childclass)
                c["name"] = e.attrib.get("name", "") + (suffix if e.attrib.get("name") else "")
                c["_pos"], c["_quat"] = place(c)
                body.cameras.append(c)
            elif t == "light":
                l = self._resolve("light", e.attrib, childclass)
                p = np.asarray(_floats(l.get("pos", "0 0 0"), 3))
                l["_pos"] = fpos + quat_rot(fquat, p)
                l["_dir"] = quat_rot(fquat, _floats(l.get("dir", "0 0 -1"), 3))
                body.lights.append(l)
            elif t == "inertial":
                a = e.attrib
                body.inertial = dict(pos=np.asarray(_floats(a.get("pos", "0 0 0"), 3)), quat=self._orient(a),
                                     mass=float(a["mass"]),
                                     diaginertia=_floats(a["diaginertia"], 3) if "diaginertia" in a else None,
                                     fullinertia=_floats(a["fullinertia"], 6) if "fullinertia" in a else None)
            elif t == "frame":
                p, q = place(e.attrib)
                self._parse_body_children(e, body_id, e.attrib.get("childclass", childclass), (p, q), suffix)
            elif t == "replicate":
                a = dict(e.attrib)
                for ch in e:
                    ov = getattr(self, "replicate_override", {}).get(ch.attrib.get("name"))
                    if ov:
                        a["count"], a["euler"] = str(ov[0]), ov[1]
                count = int(a["count"])
                off = np.asarray(_floats(a.get("offset", "0 0 0"), 3))
                rq = euler_quat([self._ang(x) for x in _floats(a.get("euler", "0 0 0"), 3)], self.eulerseq)
                sep = a.get("sep", "")
                width = len(str(count))  # naming assumed by the reference: enums/stretch_sensors.py:33-41
                p, q = fpos.copy(), fquat.copy()
                for i in range(count):
                    self._parse_body_children(e, body_id, childclass, (p, q), f"{suffix}{sep}{i:0{width}d}")
                    # accumulate: next frame = this frame ∘ (offset, rotation)
                    p = p + quat_rot(q, off)
                    q = quat_norm(quat_mul(q, rq))
            else:
                raise ValueError(f"unsupported body child <{t}>")
