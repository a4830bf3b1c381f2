"""tools/packbench.py -- a host's packing rate, no GPU: the staging step of b200h_hash_batch_host on its own
(b200h_pack_preview: the library's packer team gathering scattered payloads into wave layout, slot by slot), over N
Python ``bytes`` payloads of 256 KiB as the map pump hands them over, from C concurrent callers (the pump keeps two
windows in flight on two contexts).

    python tools/packbench.py [threads=16] [callers=1] [n=16384]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modal_client_b200 import _lib

threads = int(sys.argv[1]) if len(sys.argv) > 1 else 16
callers = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
SZ = 262144
base = np.random.default_rng(0).integers(0, 256, SZ, dtype=np.uint8).tobytes()
payloads = [base[:-1] + bytes([i & 255]) for i in range(n)]  # n distinct objects, scattered over the heap
off = np.fromiter(map(id, payloads), np.uint64, n) + np.uint64(_lib._BYTES_HDR)
ln = np.fromiter(map(len, payloads), np.uint64, n)
dsts = [np.empty(n * SZ, np.uint8) for _ in range(callers)]
for d in dsts:
    d[::4096] = 0  # touch


def one(k):
    _lib.pack_preview(None, off, ln, threads=threads, slot_bytes=256 << 20, dst=dsts[k])


best = 0.0
for _ in range(4):
    ths = [threading.Thread(target=one, args=(k,)) for k in range(callers)]
    t = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    best = max(best, callers * n * SZ / (time.perf_counter() - t) / 2**30)
assert bytes(dsts[0][:SZ]) == payloads[0] and bytes(dsts[-1][(n - 1) * SZ :]) == payloads[-1]
print(f"{_lib.load_library().b200h_stream_copy_isa().decode()}, {callers} caller(s) x {threads} packers, {n} x 256 KiB: {best:.2f} GiB/s")
