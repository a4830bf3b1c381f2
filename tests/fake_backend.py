"""Test scaffolding: an oracle-backed stand-in for ``_lib.Context`` so that host-layer logic
(size classes, stream semantics, multipart planning, sharding) can be exercised on machines without a
GPU.  Lives in tests/ only; the product never falls back to it."""
from __future__ import annotations

import numpy as np

from modal_client_b200 import _lib
from oracle import c_oracle


def _hexed(table, flags):
    """What B200H_HEX_OUT does on the device: rows as lowercase ASCII hex."""
    if table is None or not (flags & _lib.HEX_OUT):
        return table
    t = np.ascontiguousarray(table, np.uint8)
    return np.frombuffer(t.tobytes().hex().encode("ascii"), np.uint8).reshape(t.shape[0], 2 * t.shape[1]) if t.size else \
        np.zeros((t.shape[0], 2 * t.shape[1]), np.uint8)


def _u8(b) -> np.ndarray:
    return b if isinstance(b, np.ndarray) else np.frombuffer(b, dtype=np.uint8)


class FakeStream:
    def __init__(self, flags):
        self.flags, self.buf = flags, bytearray()

    def update(self, data):
        self.buf += bytes(data)

    def digests(self):
        d = bytes(self.buf)
        return (c_oracle.sha256(d) if self.flags & _lib.SHA256 else None, c_oracle.md5(d) if self.flags & _lib.MD5 else None)

    def reset(self):
        self.buf.clear()

    def close(self):
        pass


class FakeContext:
    device = -1

    def __init__(self):
        self.calls = []

    def hash_buffers(self, bufs, flags=_lib.SHA256 | _lib.MD5):
        self.calls.append(("hash_buffers", len(bufs), flags))
        n = len(bufs)
        sha = np.zeros((n, 32), np.uint8) if flags & _lib.SHA256 else None
        md5 = np.zeros((n, 16), np.uint8) if flags & _lib.MD5 else None
        ln = np.zeros(n, np.uint64)
        for i, b in enumerate(bufs):
            a = _u8(b)
            ln[i] = a.size
            if sha is not None:
                sha[i] = np.frombuffer(c_oracle.sha256(a), np.uint8)
            if md5 is not None:
                md5[i] = np.frombuffer(c_oracle.md5(a), np.uint8)
        return _hexed(sha, flags), _hexed(md5, flags), ln

    def hash_batch_host(self, base, offsets, lengths, flags=_lib.SHA256 | _lib.MD5):
        self.calls.append(("hash_batch_host", len(offsets), flags))
        off = np.asarray(offsets, np.uint64)
        ln = np.asarray(lengths, np.uint64)
        if base is None:
            import ctypes

            bufs = [np.frombuffer((ctypes.c_uint8 * int(n)).from_address(int(o)), np.uint8) if n else np.zeros(0, np.uint8)
                    for o, n in zip(off, ln)]
            total = np.concatenate(bufs) if bufs else np.zeros(0, np.uint8)
            pos = np.concatenate([[0], np.cumsum(ln)])[:-1].astype(np.uint64)
            base, off = total, pos
        s, m, e = c_oracle.hash_batch(_u8(base), off, ln, sha=bool(flags & _lib.SHA256), md5=bool(flags & _lib.MD5),
                                      trim=bool(flags & _lib.TRIM_ZEROS))
        return _hexed(s, flags), _hexed(m, flags), e

    def hash_fixed_parts(self, data, part_len, flags=_lib.SHA256 | _lib.MD5, want_etag=False):
        self.calls.append(("hash_fixed_parts", part_len, flags))
        a = _u8(data)
        starts = np.arange(0, a.size, part_len, dtype=np.uint64)
        lens = np.minimum(part_len, a.size - starts).astype(np.uint64)
        s, m, e = c_oracle.hash_batch(a, starts, lens, sha=bool(flags & _lib.SHA256), md5=bool(flags & _lib.MD5),
                                      trim=bool(flags & _lib.TRIM_ZEROS))
        etag = None
        if want_etag:
            etag = c_oracle.md5(m.tobytes() if m is not None and len(m) else b"")
        return s, m, e, etag

    def stream(self, flags=_lib.SHA256 | _lib.MD5):
        return FakeStream(flags)

    def dedupe(self, keys):
        from oracle import ref_port

        self.calls.append(("dedupe", len(keys)))
        first, nd = ref_port.first_occurrence(np.asarray(keys, np.uint8))
        return np.asarray(first, np.uint32), nd

    def stat_files(self, paths):
        import os

        st = [os.stat(p) for p in paths]
        for p, x in zip(paths, st):
            if not os.path.isfile(p):
                raise _lib.B200HashError(f"stat {p}: not a regular file")
        return (np.array([x.st_size for x in st], np.uint64), np.array([x.st_mode & 0o7777 for x in st], np.uint32))

    def hash_files(self, paths, sizes, part_len=0, flags=_lib.SHA256 | _lib.MD5):
        self.calls.append(("hash_files", len(paths), part_len, flags))
        shas, md5s, lens = [], [], []
        for p, size in zip(paths, np.asarray(sizes, np.uint64)):
            data = np.fromfile(p, dtype=np.uint8)
            if data.size < int(size):
                raise _lib.B200HashError(f"reading {p}: file is shorter than the size passed in")
            data = data[: int(size)]
            if part_len == 0:
                starts, ln = np.array([0], np.uint64), np.array([data.size], np.uint64)
            else:
                starts = np.arange(0, data.size, part_len, dtype=np.uint64)
                ln = np.minimum(part_len, data.size - starts).astype(np.uint64)
            s, m, e = c_oracle.hash_batch(data, starts, ln, sha=bool(flags & _lib.SHA256), md5=bool(flags & _lib.MD5),
                                          trim=bool(flags & _lib.TRIM_ZEROS))
            if s is not None:
                shas.append(s)
            if m is not None:
                md5s.append(m)
            lens.append(e)
        cat = lambda xs, w: (np.concatenate(xs) if xs else np.zeros((0, w), np.uint8))  # noqa: E731
        return (_hexed(cat(shas, 32), flags) if flags & _lib.SHA256 else None,
                _hexed(cat(md5s, 16), flags) if flags & _lib.MD5 else None,
                np.concatenate(lens) if lens else np.zeros(0, np.uint64))
