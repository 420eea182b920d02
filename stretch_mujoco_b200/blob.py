"""Binary model blob: the compiled-model interchange format consumed by the C-ABI
(`ss_model_load_blob`, include/stretchsim.h) and by the CPU oracle.

Layout (little endian):
    char[8]  magic "SSMBLOB1"
    u32      n_arrays
    u32      n_nametables
    n_arrays x { char[40] name; u32 dtype (0=f64, 1=i32, 2=f32, 3=u8); u32 ndim; u64 shape[4]; u64 offset; u64 nbytes }
    n_nametables x { u32 objtype; u32 count; u64 offset; u64 nbytes }   (names are NUL-separated)
    payload (each array 16-byte aligned)

The same container can carry a dump of a real ``mjModel`` (arrays named as in this repo), which
is the hook SURVEY.md §7.1 asks for should a MuJoCo install ever become reachable.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

MAGIC = b"SSMBLOB1"
_DT = {np.dtype("float64"): 0, np.dtype("int32"): 1, np.dtype("float32"): 2, np.dtype("uint8"): 3}
_DT_INV = {v: k for k, v in _DT.items()}
_ENTRY = struct.Struct("<40sII4QQQ")
_NT = struct.Struct("<IIQQ")


def pack(arrays: dict[str, np.ndarray], names: dict[int, list[str]] | None = None) -> bytes:
    names = names or {}
    head = 8 + 8 + _ENTRY.size * len(arrays) + _NT.size * len(names)
    off = (head + 15) // 16 * 16
    entries, tables, payload = [], [], []
    cur = off
    for k, a in arrays.items():
        a = np.ascontiguousarray(a)
        if a.dtype not in _DT:
            raise TypeError(f"{k}: unsupported dtype {a.dtype}")
        if len(k) >= 40 or a.ndim > 4:
            raise ValueError(k)
        shape = list(a.shape) + [0] * (4 - a.ndim)
        entries.append(_ENTRY.pack(k.encode(), _DT[a.dtype], a.ndim, *shape, cur, a.nbytes))
        raw = a.tobytes()
        pad = (-len(raw)) % 16
        payload.append(raw + b"\0" * pad)
        cur += len(raw) + pad
    for objtype, lst in names.items():
        raw = b"".join(n.encode() + b"\0" for n in lst)
        tables.append(_NT.pack(objtype, len(lst), cur, len(raw)))
        pad = (-len(raw)) % 16
        payload.append(raw + b"\0" * pad)
        cur += len(raw) + pad
    hdr = MAGIC + struct.pack("<II", len(arrays), len(names)) + b"".join(entries) + b"".join(tables)
    return hdr + b"\0" * (off - len(hdr)) + b"".join(payload)


def unpack(buf: bytes):
    if buf[:8] != MAGIC:
        raise ValueError("not a stretchsim model blob")
    na, nt = struct.unpack_from("<II", buf, 8)
    arrays, names = {}, {}
    p = 16
    for _ in range(na):
        nm, dt, nd, s0, s1, s2, s3, off, nb = _ENTRY.unpack_from(buf, p)
        p += _ENTRY.size
        dtype = _DT_INV[dt]
        shape = (s0, s1, s2, s3)[:nd]
        arrays[nm.rstrip(b"\0").decode()] = np.frombuffer(buf, dtype=dtype, count=nb // dtype.itemsize, offset=off).reshape(shape).copy()
    for _ in range(nt):
        ot, cnt, off, nb = _NT.unpack_from(buf, p)
        p += _NT.size
        raw = buf[off:off + nb].split(b"\0")[:cnt]
        names[ot] = [r.decode() for r in raw]
    return arrays, names


def save(path: str, model) -> None:
    """`.ssm` = raw blob; `.ssm.z` = zlib-compressed blob (used for the fixtures that carry meshes)."""
    raw = pack(model.arrays, model.names)
    with open(path, "wb") as fh:
        fh.write(zlib.compress(raw, 9) if path.endswith(".z") else raw)


def read_bytes(path: str) -> bytes:
    with open(path, "rb") as fh:
        raw = fh.read()
    return zlib.decompress(raw) if path.endswith(".z") else raw


def load(path: str):
    from .compiler import Model
    arrays, names = unpack(read_bytes(path))
    m = Model()
    m.arrays, m.names = arrays, names
    return m
