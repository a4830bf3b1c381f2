"""tools/pump_probe.py -- where the map pump's hash stage loses time: the same 100 000 x 256 KiB pageable payloads
through hash_utils.get_upload_hashes_many as (A) one call, (B) sequential 1 GiB windows, (C) two threads on two
contexts, (D) like C while the main thread burns Python (GIL contention), then the pump itself with its stage timers."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modal_client_b200 import _backend, _lib, hash_utils
from modal_client_b200.synth import synth_array

N = int(os.environ.get("PROBE_N", 100_000))
SZ = 262144
GiB = float(1 << 30)
ctx = _lib.Context(0, pinned_bytes=512 << 20, device_bytes=8 << 30)
_backend.set_context(ctx)
ctx2 = _backend.context_pool(2)[1]
base = synth_array(3, 64 * SZ)
payloads = [base[(i % 64) * SZ : (i % 64 + 1) * SZ].tobytes() for i in range(N)]  # distinct objects, 24.4 GiB
W = int(os.environ.get("PROBE_WINDOW", 4096))
wins = [payloads[i : i + W] for i in range(0, N, W)]


def report(name, dt, **kw):
    print(json.dumps({"case": name, "seconds": round(dt, 4), "GiBps": round(N * SZ / GiB / dt, 2), **kw}), flush=True)


hash_utils.get_upload_hashes_many(wins[0], ctx=ctx)
hash_utils.get_upload_hashes_many(wins[0], ctx=ctx2)
t = time.perf_counter()
hash_utils.get_upload_hashes_many(payloads, ctx=ctx)
report("A one call, 100k payloads", time.perf_counter() - t)
t = time.perf_counter()
per = []
for w in wins:
    t1 = time.perf_counter()
    hash_utils.get_upload_hashes_many(w, ctx=ctx)
    per.append(time.perf_counter() - t1)
report(f"B sequential windows of {W}", time.perf_counter() - t, window_ms_median=round(1e3 * float(np.median(per)), 2))


def two_threads(burn: bool):
    it = iter(wins)
    lock = threading.Lock()
    spent = [0.0, 0.0]

    def work(k, c):
        while True:
            with lock:
                w = next(it, None)
            if w is None:
                return
            t1 = time.perf_counter()
            hash_utils.get_upload_hashes_many(w, ctx=c)
            spent[k] += time.perf_counter() - t1

    ths = [threading.Thread(target=work, args=(0, ctx)), threading.Thread(target=work, args=(1, ctx2))]
    t = time.perf_counter()
    for th in ths:
        th.start()
    spins = 0
    if burn:
        while any(th.is_alive() for th in ths):
            for _ in range(1000):
                spins += 1
    for th in ths:
        th.join()
    return time.perf_counter() - t, spent


dt, spent = two_threads(False)
report("C two threads / two contexts", dt, thread_busy_s=[round(x, 3) for x in spent])
dt, spent = two_threads(True)
report("D same, main thread burning Python (GIL contention)", dt, thread_busy_s=[round(x, 3) for x in spent])
sys.setswitchinterval(0.0005)
dt, spent = two_threads(True)
report("E same as D with sys.setswitchinterval(0.5 ms)", dt, thread_busy_s=[round(x, 3) for x in spent])
sys.setswitchinterval(0.005)

import bench  # noqa: E402
from modal_client_b200 import blob_utils  # noqa: E402

blob_utils._upload_to_s3_url = bench._null_put
for _ in range(2):
    stub = bench.NullStub()
    dt, batches, _ = bench.run_map_pump(payloads, stub)
report("F the pump (InputPreprocessor -> InputPumper, null control plane)", dt, hash_batches=batches,
       stages=bench.run_map_pump.last_stats)
sys.setswitchinterval(0.0005)
stub = bench.NullStub()
dt, batches, _ = bench.run_map_pump(payloads, stub)
report("G the pump with sys.setswitchinterval(0.5 ms)", dt, hash_batches=batches, stages=bench.run_map_pump.last_stats)
