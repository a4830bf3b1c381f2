"""tools/sanitize_cases.py -- small batches that drive every kernel through its synchronisation-heavy paths, sized
for compute-sanitizer (memcheck / racecheck / synccheck slow kernels down 10-100x):
  lane kernel  more messages than resident lanes (time slicing: states saved/restored, ring re-queue with volatile
               spin-waits and cross-warp atomics), unaligned starts, every tail shape
  chain kernel TMA tile ring + mbarrier hand-offs between the expander / SHA-256 / MD5 warps, 3 long messages
  trim kernels probe + wide scan (atomicMax across CTAs, work counter), blank and half-blank blocks
  dedupe       hash-table insert/resolve with CAS
  streams      absorb/digest on a private CUDA stream
Digests are compared with hashlib, so a sanitizer run is also a parity run.
  compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
  compute-sanitizer --tool racecheck python tools/sanitize_cases.py"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modal_client_b200 import _lib
from modal_client_b200.synth import synth_array

BOTH = _lib.SHA256 | _lib.MD5
ctx = _lib.Context(0, pinned_bytes=8 << 20, device_bytes=64 << 20)


def check(buf, offs, lens, sha, md5, idx):
    for i in idx:
        m = buf[int(offs[i]) : int(offs[i] + lens[i])].tobytes()
        assert sha is None or sha[i].tobytes() == hashlib.sha256(m).digest(), i
        assert md5 is None or md5[i].tobytes() == hashlib.md5(m).digest(), i


# 1. lane kernel with time slicing: 120 000 short messages (> 94 720 resident lanes) + mixed tails, unaligned
n = int(os.environ.get("SAN_N", 120_000))
rng = np.random.default_rng(1)
lens = rng.integers(0, 200, n).astype(np.uint64)
lens[:64] = np.arange(64) * 3
lens[64:72] = [4096, 5000, 70_000, 64, 128, 55, 56, 119]
offs = (np.concatenate([[0], np.cumsum(lens + np.uint64(1))])[:-1] + np.uint64(3)).astype(np.uint64)
buf = synth_array(1, int(offs[-1] + lens[-1]) + 8)
sha, md5, _ = ctx.hash_batch_host(buf, offs, lens, BOTH)
check(buf, offs, lens, sha, md5, list(range(0, 80)) + list(range(80, n, 997)))
print("lane kernel (time-sliced) ok", flush=True)

# 2. chain kernel: three outliers among small messages, aligned and unaligned, all flag sets
lens = np.array([200_000, 150_001, 131_072 + 55] + [300] * 400, np.uint64)
offs = (np.concatenate([[0], np.cumsum(lens + np.uint64(5))])[:-1] + np.uint64(16)).astype(np.uint64)
buf = synth_array(2, int(offs[-1] + lens[-1]) + 8)
for flags in (BOTH, _lib.SHA256, _lib.MD5):
    sha, md5, _ = ctx.hash_batch_host(buf, offs, lens, flags)
    assert ctx.last_outlier_count == 3, ctx.last_outlier_count
    check(buf, offs, lens, sha, md5, [0, 1, 2, 3, 402])
print("chain kernel ok", flush=True)

# 3. trim: blank / half-blank / dense blocks of 1 MiB, one with a lone byte deep inside
B = 1 << 20
blk = np.zeros(6 * B + 77, np.uint8)
blk[B : B + 300_000] = synth_array(3, 300_000) | 1
blk[3 * B : 4 * B] = synth_array(4, B) | 1
blk[4 * B + 12345] = 9
blk[5 * B + 5 : 5 * B + 70] = 1
sha, _, trimmed, _ = ctx.hash_fixed_parts(blk, B, _lib.SHA256 | _lib.TRIM_ZEROS)
want = [0, 300_000, 0, B, 12346, 70, 0]
assert list(map(int, trimmed)) == want, list(map(int, trimmed))
for i, t in enumerate(want):
    assert sha[i].tobytes() == hashlib.sha256(blk[i * B : i * B + t]).digest()
print("trim kernels ok", flush=True)

# 4. dedupe
keys = np.frombuffer(b"".join(hashlib.sha256(bytes([i % 50])).digest() for i in range(5000)), np.uint8).reshape(-1, 32)
first, nd = ctx.dedupe(keys)
assert nd == 50 and all(int(first[i]) == i % 50 for i in range(5000))
print("dedupe ok", flush=True)

# 5. stream on its own CUDA stream (small buffer so that several absorbs happen)
os.environ["B200H_STREAM_BUF"] = "65536"
c2 = _lib.Context(0, pinned_bytes=1 << 20, device_bytes=4 << 20)
s = c2.stream(BOTH)
data = synth_array(5, 300_007).tobytes()
for i in range(0, len(data), 50_000):
    s.update(data[i : i + 50_000])
assert s.digests() == (hashlib.sha256(data).digest(), hashlib.md5(data).digest())
s.close()
c2.close()
print("stream ok", flush=True)
ctx.close()
print("ALL OK")
