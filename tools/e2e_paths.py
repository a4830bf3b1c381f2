"""tools/e2e_paths.py -- end-to-end GiB/s of the host entry points on one GPU:
  (a) page-locked contiguous buffer (what bench.py's e2e measures)
  (b) pageable contiguous numpy buffer (staged through the pinned ring by the library's pack threads)
  (c) list of separate Python bytes objects (get_upload_hashes_many: what the map pump passes)
usage: python tools/e2e_paths.py [n_msgs] [msg_bytes]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modal_client_b200 import _lib, hash_utils, _backend

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
size = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
ctx = _lib.Context(0, pinned_bytes=int(os.environ.get("B200H_PINNED", 512 << 20)), device_bytes=8 << 30)
_backend.set_context(ctx)
total = n * size
rng = np.random.default_rng(0)
pageable = rng.integers(0, 256, total, dtype=np.uint8)
off = np.arange(n, dtype=np.uint64) * np.uint64(size)
ln = np.full(n, size, dtype=np.uint64)


def timeit(name, fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:58s} {total / dt / 2**30:8.2f} GiB/s  ({dt * 1e3:.1f} ms)", flush=True)
    return out


pinned = ctx.host_alloc(total)
pinned[:] = pageable
a = timeit("(a) pinned contiguous  -> hash_batch_host", lambda: ctx.hash_batch_host(pinned, off, ln))
b = timeit("(b) pageable contiguous -> hash_batch_host", lambda: ctx.hash_batch_host(pageable, off, ln))
payloads = [pageable[i * size:(i + 1) * size].tobytes() for i in range(n)]
c = timeit("(c) list[bytes] -> hash_utils.get_upload_hashes_many", lambda: hash_utils.get_upload_hashes_many(payloads))
assert np.array_equal(a[0], b[0]) and c[5].sha256_hex() == a[0][5].tobytes().hex()
t0 = time.perf_counter()
views = [np.frombuffer(p, dtype=np.uint8) for p in payloads]
print(f"    (python overhead: building {n} numpy views {1e3 * (time.perf_counter() - t0):.1f} ms)")
