"""ctypes binding of oracle/liboracle.so (hash_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc if absent or stale."""
    src = os.path.join(_HERE, "hash_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(
            [os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-std=c11", "-shared", "-o", _LIB_PATH, src]
        )
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        u8p, u64p = ctypes.c_void_p, ctypes.c_void_p
        _lib.orc_hash_one.argtypes = [u8p, ctypes.c_uint64, u8p, u8p]
        _lib.orc_hash_one.restype = None
        _lib.orc_trimmed_len.argtypes = [u8p, ctypes.c_uint64]
        _lib.orc_trimmed_len.restype = ctypes.c_uint64
        _lib.orc_hash_batch.argtypes = [u8p, u64p, u64p, ctypes.c_uint64, ctypes.c_int, u8p, u8p, u64p]
        _lib.orc_hash_batch.restype = None
        _lib.orc_multipart_md5.argtypes = [u8p, ctypes.c_uint64, ctypes.c_uint64, u8p, u8p]
        _lib.orc_multipart_md5.restype = ctypes.c_uint64
    return _lib


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        return data
    return np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(0, np.uint8)


def sha256(data) -> bytes:
    a = _as_u8(data)
    out = np.zeros(32, np.uint8)
    lib().orc_hash_one(_ptr(a), a.size, _ptr(out), None)
    return out.tobytes()


def md5(data) -> bytes:
    a = _as_u8(data)
    out = np.zeros(16, np.uint8)
    lib().orc_hash_one(_ptr(a), a.size, None, _ptr(out))
    return out.tobytes()


def trimmed_len(data) -> int:
    a = _as_u8(data)
    return int(lib().orc_trimmed_len(_ptr(a), a.size))


def hash_batch(base, offsets, lengths, *, sha=True, md5=True, trim=False):
    """-> (sha[n,32] | None, md5[n,16] | None, end[n])."""
    a = _as_u8(base)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lengths, dtype=np.uint64)
    n = off.size
    s = np.zeros((n, 32), np.uint8) if sha else None
    m = np.zeros((n, 16), np.uint8) if md5 else None
    e = np.zeros(n, np.uint64)
    lib().orc_hash_batch(_ptr(a), _ptr(off), _ptr(ln), n, int(trim), _ptr(s), _ptr(m), _ptr(e))
    return s, m, e


def multipart_md5(data, part_len: int):
    """-> (part_md5[nparts,16], etag_md5 bytes16)."""
    a = _as_u8(data)
    nparts = -(-a.size // part_len) if a.size else 0
    parts = np.zeros((max(nparts, 1), 16), np.uint8)
    etag = np.zeros(16, np.uint8)
    got = lib().orc_multipart_md5(_ptr(a), a.size, part_len, _ptr(parts), _ptr(etag))
    assert got == nparts
    return parts[:nparts], etag.tobytes()
