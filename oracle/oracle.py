"""ctypes front-end for the CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/ss_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libss_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "ss_blob.h"))
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "CC=gcc"], stdout=subprocess.DEVNULL)
    return _LIB


class _Outputs(C.Structure):
    _fields_ = [("xpos", C.c_void_p), ("xquat", C.c_void_p), ("act_length", C.c_void_p), ("act_velocity", C.c_void_p),
                ("sensordata", C.c_void_p), ("qacc", C.c_void_p), ("qfrc_constraint", C.c_void_p), ("ncon", C.c_void_p),
                ("contact_geom", C.c_void_p), ("contact_dist", C.c_void_p), ("contact_pos", C.c_void_p),
                ("contact_frame", C.c_void_p), ("maxcon", C.c_int), ("solver_iter", C.c_void_p), ("nefc", C.c_void_p),
                ("flags", C.c_void_p), ("M", C.c_void_p), ("qacc_smooth", C.c_void_p), ("qfrc_bias", C.c_void_p),
                ("qfrc_passive", C.c_void_p), ("qfrc_actuator", C.c_void_p), ("efc_J", C.c_void_p),
                ("efc_aref", C.c_void_p), ("efc_D", C.c_void_p), ("efc_force", C.c_void_p), ("maxefc_out", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.om_model_load.restype = C.c_void_p
        _lib.om_model_load.argtypes = [C.c_char_p, C.c_size_t]
        _lib.om_model_sizes.argtypes = [C.c_void_p, C.c_void_p]
        _lib.om_set_options.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        _lib.om_batch_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.om_batch_forward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int]
        _lib.om_batch_rays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _lib.om_batch_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_double, C.c_void_p, C.c_void_p, C.c_int]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleModel:
    def __init__(self, blob_bytes: bytes):
        self._blob = blob_bytes
        self._h = lib().om_model_load(blob_bytes, len(blob_bytes))
        if not self._h:
            raise ValueError("oracle could not load the model blob")
        sz = np.zeros(16, np.int32)
        lib().om_model_sizes(self._h, _p(sz))
        (self.nq, self.nv, self.nu, self.nbody, self.njnt, self.ngeom, self.nsite, self.ncam, self.ntendon, self.neq,
         self.nsensor, self.nsensordata, self.nkey, self.nM, self.npair, self.nmesh) = [int(x) for x in sz]

    def set_options(self, max_iter=0, tolerance=0.0, enable_lidar=True):
        lib().om_set_options(self._h, max_iter, tolerance, int(enable_lidar))

    def _outputs(self, nenv, want, maxcon=32, maxefc=0):
        o = _Outputs()
        bufs = {}
        shapes = dict(xpos=(self.nbody, 3), xquat=(self.nbody, 4), act_length=(self.nu,), act_velocity=(self.nu,),
                      sensordata=(self.nsensordata,), qacc=(self.nv,), qfrc_constraint=(self.nv,),
                      contact_dist=(maxcon,), contact_pos=(maxcon, 3), contact_frame=(maxcon, 3), M=(self.nv, self.nv),
                      qacc_smooth=(self.nv,), qfrc_bias=(self.nv,), qfrc_passive=(self.nv,), qfrc_actuator=(self.nv,),
                      efc_J=(maxefc, self.nv), efc_aref=(maxefc,), efc_D=(maxefc,), efc_force=(maxefc,))
        ishapes = dict(ncon=(), contact_geom=(maxcon, 2), solver_iter=(), nefc=(), flags=())
        for k in want:
            if k in shapes:
                bufs[k] = np.zeros((nenv,) + shapes[k], np.float64)
            elif k in ishapes:
                bufs[k] = np.zeros((nenv,) + ishapes[k], np.int32)
            else:
                raise KeyError(k)
            setattr(o, k, bufs[k].ctypes.data)
        o.maxcon = maxcon if any(k.startswith("contact") for k in want) else 0
        o.maxefc_out = maxefc if any(k.startswith("efc") for k in want) else 0
        return o, bufs

    def step(self, qpos, qvel, ctrl, warm, time=None, nsteps=1, want=(), nthreads=0, maxcon=32, maxefc=0):
        """In-place advance of fp64 state arrays [nenv, n]. ctrl: [nenv,nu] or [nsteps,nenv,nu]."""
        nenv = qpos.shape[0]
        for a in (qpos, qvel, warm):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        per_step = int(ctrl.ndim == 3)
        if time is None:
            time = np.zeros(nenv)
        o, bufs = self._outputs(nenv, want, maxcon, maxefc)
        rc = lib().om_batch_step(self._h, nenv, nsteps, _p(qpos), _p(qvel), _p(ctrl), per_step, _p(warm), _p(time),
                                 C.byref(o), nthreads)
        assert rc == 0
        return bufs

    def forward(self, qpos, qvel, ctrl, warm=None, want=(), nthreads=0, maxcon=32, maxefc=0):
        nenv = qpos.shape[0]
        qpos = np.ascontiguousarray(qpos, dtype=np.float64); qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        if warm is not None:
            warm = np.ascontiguousarray(warm, dtype=np.float64)
        o, bufs = self._outputs(nenv, want, maxcon, maxefc)
        rc = lib().om_batch_forward(self._h, nenv, _p(qpos), _p(qvel), _p(ctrl), _p(warm), C.byref(o), nthreads)
        assert rc == 0
        return bufs

    def rays(self, xpos, xquat, origin, direction, groupmask=0, bodyexclude=-1, nthreads=0):
        nenv, nray = origin.shape[0], origin.shape[1]
        xpos = np.ascontiguousarray(xpos, dtype=np.float64); xquat = np.ascontiguousarray(xquat, dtype=np.float64)
        origin = np.ascontiguousarray(origin, dtype=np.float64); direction = np.ascontiguousarray(direction, dtype=np.float64)
        dist = np.zeros((nenv, nray)); geom = np.zeros((nenv, nray), np.int32)
        rc = lib().om_batch_rays(self._h, nenv, _p(xpos), _p(xquat), nray, _p(origin), _p(direction), groupmask,
                                 bodyexclude, _p(dist), _p(geom), nthreads)
        assert rc == 0
        return dist, geom

    def render(self, xpos, xquat, cam_id, W, H, fovy_deg, nthreads=0):
        nenv = xpos.shape[0]
        xpos = np.ascontiguousarray(xpos, dtype=np.float64); xquat = np.ascontiguousarray(xquat, dtype=np.float64)
        rgb = np.zeros((nenv, H, W, 3), np.uint8); depth = np.zeros((nenv, H, W), np.float32)
        rc = lib().om_batch_render(self._h, nenv, _p(xpos), _p(xquat), cam_id, W, H, float(fovy_deg), _p(rgb),
                                   _p(depth), nthreads)
        assert rc == 0
        return rgb, depth
