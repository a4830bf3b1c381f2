"""tools/stress.py -- randomized soak of the persistent/time-sliced scheduler against the oracle.
Random batch shapes (message counts 1..20000, sizes from empty to MiB, skewed / uniform / duplicated lengths,
random misalignment, random digest flags and trim) hashed through the C ABI and compared bit-for-bit.
usage: python tools/stress.py [seconds=90] [seed=0]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modal_client_b200 import _lib
from modal_client_b200.synth import synth_array
from oracle import c_oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ctx = _lib.Context(0, pinned_bytes=64 << 20, device_bytes=256 << 20)
pool = synth_array(123, (96 << 20) + 4096).copy()
pool[rng.integers(0, pool.size - 70000, 400)[:, None] + np.arange(60000)[None, :]] = 0  # long zero runs for trim
t_end = time.time() + budget
rounds = msgs = nbytes = 0
while time.time() < t_end:
    shape = rng.integers(0, 6)
    n = int(rng.integers(1, [20000, 3000, 300, 40, 6000, 2][shape] + 1))
    if shape == 0:
        lens = rng.integers(0, 300, n)  # tiny, ragged
    elif shape == 1:
        lens = rng.integers(0, 20000, n)
    elif shape == 2:
        lens = np.minimum(rng.lognormal(9, 2.0, n), 3 << 20).astype(np.int64)  # long tail
    elif shape == 3:
        lens = rng.integers(200_000, 2_500_000, n)  # few big
    elif shape == 4:
        lens = np.full(n, int(rng.choice([63, 64, 65, 4096, 2048 + 55, 2048 + 56])))  # uniform: time slicing
    else:
        lens = np.array([int(rng.integers(1 << 20, 6 << 20)), 5][:n])
    total = int(lens.sum())
    if total > (90 << 20):
        lens = (lens * ((90 << 20) / total)).astype(np.int64)
    offs = rng.integers(0, pool.size - np.maximum(lens, 1) - 1)
    lens, offs = lens.astype(np.uint64), offs.astype(np.uint64)
    flags = int(rng.choice([3, 3, 3, 1, 2])) | (4 if rng.random() < 0.3 else 0)
    sha, md5, trimmed = ctx.hash_batch_host(pool, offs, lens, flags)
    s, m, e = c_oracle.hash_batch(pool, offs, lens, sha=bool(flags & 1), md5=bool(flags & 2), trim=bool(flags & 4))
    ok = np.array_equal(trimmed, e) and (s is None or np.array_equal(sha, s)) and (m is None or np.array_equal(md5, m))
    if not ok:
        bad = np.flatnonzero((trimmed != e) | ((sha != s).any(1) if s is not None else False) | ((md5 != m).any(1) if m is not None else False))
        print(f"MISMATCH round {rounds} shape {shape} n {n} flags {flags} first bad idx {bad[:5]} len {lens[bad[:5]]} off {offs[bad[:5]]}")
        sys.exit(1)
    rounds += 1
    msgs += n
    nbytes += int(e.sum())
print(f"stress ok: {rounds} random batches, {msgs} messages, {nbytes / 2**30:.2f} GiB, all digests equal to the oracle")
