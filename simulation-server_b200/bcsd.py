"""BCSD named-array container (scene / state / dump files).

Python twin of ``include/bcsd_io.hpp``::

    file   := b"BCSD1\\0\\0\\0" record*
    record := u32 name_len, name bytes, u32 dtype, u64 count, payload
    dtype  := 0 f32 | 1 i32 | 2 u32 | 3 f64 | 4 i64
"""
from __future__ import annotations

import struct
from typing import Dict, Mapping

import numpy as np

_MAGIC = b"BCSD1\0\0\0"
_DTYPES = {0: np.float32, 1: np.int32, 2: np.uint32, 3: np.float64, 4: np.int64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def read(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:5] != _MAGIC[:5]:
        raise ValueError(f"{path}: not a BCSD file")
    out: Dict[str, np.ndarray] = {}
    off = 8
    n = len(buf)
    while off < n:
        (nl,) = struct.unpack_from("<I", buf, off)
        off += 4
        name = buf[off:off + nl].decode()
        off += nl
        dtype, count = struct.unpack_from("<IQ", buf, off)
        off += 12
        dt = np.dtype(_DTYPES[dtype])
        nbytes = count * dt.itemsize
        out[name] = np.frombuffer(buf, dtype=dt, count=count, offset=off).copy()
        off += nbytes
    return out


def write(path: str, arrays: Mapping[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(_MAGIC)
        for name, a in arrays.items():
            a = np.ascontiguousarray(a).reshape(-1)
            if a.dtype not in _CODES:
                raise TypeError(f"{name}: unsupported dtype {a.dtype}")
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<IQ", _CODES[a.dtype], a.size))
            f.write(a.tobytes())
