"""ctypes binding of libstretchsim (include/stretchsim.h) + torch-owned device buffers.

This is the "thin C-ABI/ctypes layer that hands back PyTorch CUDA tensors" of the north star:
torch allocates every per-env array, the C side only sees raw pointers and the current CUDA
stream.  There is NO CPU fallback: importing works without a GPU (so the CPU test-suite can
check the ABI), but creating a model or batch without CUDA raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstretchsim.so")

OBJ_BODY, OBJ_JOINT, OBJ_GEOM, OBJ_SITE, OBJ_CAMERA, OBJ_ACTUATOR, OBJ_SENSOR, OBJ_KEY, OBJ_MESH, OBJ_TENDON = range(10)
STATUS_WIDTH = 24
CMD_WIDTH = 44

EXPORTS = [
    "ss_model_load_blob", "ss_model_free", "ss_model_dims", "ss_name2id", "ss_id2name", "ss_model_get", "ss_model_field_info",
    "ss_model_set", "ss_batch_create", "ss_batch_free", "ss_batch_reset", "ss_batch_step", "ss_batch_forward", "ss_batch_profile_step",
    "ss_batch_launch_count", "ss_batch_set_debug", "ss_batch_pull_status", "ss_batch_apply_commands",
    "ss_batch_step_controlled", "ss_batch_lidar",
    "ss_model_num_rangefinders", "ss_batch_rays", "ss_batch_render", "ss_batch_render_post", "ss_depth_colormap",
    "ss_last_error", "ss_version",
]


class Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "ncam", "ntendon", "neq",
                                       "nsensor", "nsensordata", "nkey", "nM", "npair", "nmesh")]


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("qpos", "qvel", "qacc_warmstart", "time", "ctrl", "xpos", "xquat", "act_length",
                                          "act_velocity", "sensordata", "qacc", "ncon", "contact_geom", "contact_dist",
                                          "solver_iter", "env_flags")]


class DebugBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("M", "qacc_smooth", "qfrc_smooth", "qfrc_constraint", "nefc", "contact_pos",
                                          "contact_normal")]


_lib = None


def lib():
    """Load libstretchsim.so; fails loudly when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.ss_last_error.restype = C.c_char_p
        L.ss_version.restype = C.c_char_p
        L.ss_id2name.restype = C.c_char_p
        L.ss_id2name.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ss_name2id.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.ss_model_load_blob.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        L.ss_model_free.argtypes = [C.c_void_p]
        L.ss_model_dims.argtypes = [C.c_void_p, C.POINTER(Dims)]
        L.ss_model_get.restype = C.c_long
        L.ss_model_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.ss_model_set.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.ss_model_field_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        L.ss_batch_step_controlled.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ss_batch_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Buffers), C.POINTER(C.c_void_p)]
        L.ss_batch_free.argtypes = [C.c_void_p]
        L.ss_batch_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ss_batch_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ss_batch_forward.argtypes = [C.c_void_p, C.c_void_p]
        L.ss_batch_profile_step.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
        L.ss_batch_launch_count.restype = C.c_long
        L.ss_batch_launch_count.argtypes = [C.c_void_p]
        L.ss_batch_set_debug.argtypes = [C.c_void_p, C.POINTER(DebugBuffers)]
        L.ss_batch_pull_status.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ss_batch_apply_commands.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ss_batch_lidar.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ss_model_num_rangefinders.argtypes = [C.c_void_p]
        L.ss_batch_rays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
        L.ss_batch_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_float,
                                      C.c_int, C.c_int, C.c_void_p]
        L.ss_batch_render_post.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_float,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ss_depth_colormap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def depth_colormap(depth):
    """JET colour map of depth images [n, H, W] (device tensor) -> BGR uint8 [n, H, W, 3], normalised per image
    (utils.get_depth_color_map, stretch_mujoco/utils.py:363-373)."""
    import torch
    depth = depth.contiguous()
    n = depth.shape[0]
    out = torch.empty(*depth.shape, 3, dtype=torch.uint8, device=depth.device)
    _check(lib().ss_depth_colormap(C.c_void_p(depth.data_ptr()), n, depth[0].numel(), C.c_void_p(out.data_ptr()), _stream()))
    return out


class StretchSimError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise StretchSimError(lib().ss_last_error().decode())
    return rc


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceModel:
    """A compiled model uploaded to one GPU (handle to `ss_model`)."""

    def __init__(self, blob_bytes: bytes, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise StretchSimError("CUDA is not available: the engine has no CPU fallback")
        self._h = C.c_void_p()
        self.device = device
        _check(lib().ss_model_load_blob(blob_bytes, len(blob_bytes), device, C.byref(self._h)))
        d = Dims()
        _check(lib().ss_model_dims(self._h, C.byref(d)))
        for n, _ in Dims._fields_:
            setattr(self, n, getattr(d, n))
        self.nrange = lib().ss_model_num_rangefinders(self._h)

    def __del__(self):
        try:
            if self._h:
                lib().ss_model_free(self._h)
                self._h = None
        except Exception:
            pass

    def name2id(self, objtype: int, name: str) -> int:
        return lib().ss_name2id(self._h, objtype, name.encode())

    def id2name(self, objtype: int, i: int) -> str | None:
        r = lib().ss_id2name(self._h, objtype, i)
        return r.decode() if r is not None else None

    def get(self, field: str) -> np.ndarray:
        """Host copy of a named model array with the dtype and shape recorded in the model blob."""
        dt, nd, shape = C.c_int(), C.c_int(), (C.c_uint64 * 4)()
        _check(lib().ss_model_field_info(self._h, field.encode(), C.byref(dt), C.byref(nd), shape))
        n = _check(lib().ss_model_get(self._h, field.encode(), None, 0))
        buf = np.zeros(n, np.uint8)
        if n:
            _check(lib().ss_model_get(self._h, field.encode(), buf.ctypes.data_as(C.c_void_p), n))
        dtype = (np.float64, np.int32, np.float32, np.uint8)[dt.value]
        arr = buf.view(dtype).copy()
        if field == "qpos0":        # reflects ss_model_set; always fp64 [nq]
            return arr
        return arr.reshape([int(shape[k]) for k in range(nd.value)])

    def set(self, field: str, value) -> None:
        if field == "opt_iterations":
            v = np.asarray([value], np.int32)
        else:
            v = np.ascontiguousarray(value, dtype=np.float64).ravel()
        _check(lib().ss_model_set(self._h, field.encode(), v.ctypes.data_as(C.c_void_p), v.nbytes))


class Batch:
    """`nenv` independent envs on one GPU; all arrays are torch CUDA tensors, env-major."""

    def __init__(self, model: DeviceModel, nenv: int, maxcon: int = 32, maxefc: int = 0, debug: bool = False):
        import torch
        self.model, self.nenv, self.maxcon = model, nenv, maxcon
        dev = torch.device("cuda", model.device)
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        i = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        m = model
        self.qpos, self.qvel, self.qacc_warmstart = f(nenv, m.nq), f(nenv, m.nv), f(nenv, m.nv)
        self.time, self.ctrl = f(nenv), f(nenv, m.nu)
        self.xpos, self.xquat = f(nenv, m.nbody, 3), f(nenv, m.nbody, 4)
        self.act_length, self.act_velocity = f(nenv, m.nu), f(nenv, m.nu)
        self.sensordata, self.qacc = f(nenv, max(m.nsensordata, 1)), f(nenv, m.nv)
        self.ncon, self.contact_geom, self.contact_dist = i(nenv), i(nenv, maxcon, 2), f(nenv, maxcon)
        self.solver_iter, self.env_flags = i(nenv), i(nenv)
        b = Buffers()
        for n, _ in Buffers._fields_:
            setattr(b, n, getattr(self, n).data_ptr())
        self._h = C.c_void_p()
        with torch.cuda.device(dev):
            _check(lib().ss_batch_create(model._h, nenv, maxcon, maxefc, C.byref(b), C.byref(self._h)))
        self.dbg = None
        if debug:
            self.dbg = dict(M=f(nenv, m.nv, m.nv), qacc_smooth=f(nenv, m.nv), qfrc_smooth=f(nenv, m.nv),
                            qfrc_constraint=f(nenv, m.nv), nefc=i(nenv), contact_pos=f(nenv, maxcon, 3),
                            contact_normal=f(nenv, maxcon, 3))
            d = DebugBuffers()
            for n, _ in DebugBuffers._fields_:
                setattr(d, n, self.dbg[n].data_ptr())
            _check(lib().ss_batch_set_debug(self._h, C.byref(d)))
        self.status = f(nenv, STATUS_WIDTH)
        self.command = f(nenv, CMD_WIDTH)
        self.base_state = f(nenv, 8)
        self.reset()

    def __del__(self):
        try:
            if self._h:
                lib().ss_batch_free(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self, env_mask=None, key: int = -1):
        mp = None if env_mask is None else C.c_void_p(env_mask.data_ptr())
        _check(lib().ss_batch_reset(self._h, mp, key, _stream()))

    def step(self, nsteps: int = 1):
        _check(lib().ss_batch_step(self._h, nsteps, _stream()))

    def forward(self):
        _check(lib().ss_batch_forward(self._h, _stream()))

    def profile_step(self):
        """One mj_step with its kernels serialised and timed by CUDA events: (smooth, narrow, solve) in ms."""
        ms = (C.c_float * 3)()
        _check(lib().ss_batch_profile_step(self._h, ms, _stream()))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def step_controlled(self, nsteps: int = 1):
        """nsteps x (apply_commands, one physics step): the reference's per-step command cadence."""
        _check(lib().ss_batch_step_controlled(self._h, nsteps, C.c_void_p(self.command.data_ptr()),
                                              C.c_void_p(self.base_state.data_ptr()), _stream()))

    def pull_status(self):
        _check(lib().ss_batch_pull_status(self._h, C.c_void_p(self.status.data_ptr()), _stream()))
        return self.status

    def apply_commands(self):
        _check(lib().ss_batch_apply_commands(self._h, C.c_void_p(self.command.data_ptr()),
                                             C.c_void_p(self.base_state.data_ptr()), _stream()))

    def lidar(self, out=None):
        import torch
        if out is None:
            out = torch.empty(self.nenv, self.model.nrange, dtype=torch.float32, device=self.qpos.device)
        _check(lib().ss_batch_lidar(self._h, C.c_void_p(out.data_ptr()), _stream()))
        return out

    def rays(self, origin, direction, groupmask: int = 0, bodyexclude: int = -1):
        import torch
        nray = origin.shape[1]
        dist = torch.empty(self.nenv, nray, dtype=torch.float32, device=origin.device)
        geom = torch.empty(self.nenv, nray, dtype=torch.int32, device=origin.device)
        _check(lib().ss_batch_rays(self._h, nray, C.c_void_p(origin.data_ptr()), C.c_void_p(direction.data_ptr()),
                                   groupmask, bodyexclude, C.c_void_p(dist.data_ptr()), C.c_void_p(geom.data_ptr()),
                                   _stream()))
        return dist, geom

    def render(self, cam_id: int, W: int, H: int, fovy: float, rgb=None, depth=None, depth_limit: float = 0.0,
               env_begin: int = 0, env_count: int | None = None, rot90: int = 0, bgr: bool = False):
        """rot90 / bgr: the reference's client-side np.rot90(img, k) / RGB->BGR fused into the kernel; with
        rot90 != 0 the output tensors are [n, W, H(, 3)]."""
        n = self.nenv - env_begin if env_count is None else env_count
        _check(lib().ss_batch_render_post(self._h, cam_id, W, H, float(fovy),
                                          None if rgb is None else C.c_void_p(rgb.data_ptr()),
                                          None if depth is None else C.c_void_p(depth.data_ptr()), float(depth_limit),
                                          env_begin, n, int(rot90), int(bool(bgr)), _stream()))
        return rgb, depth

    @property
    def launches(self) -> int:
        return int(lib().ss_batch_launch_count(self._h))
