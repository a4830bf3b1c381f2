"""tools/dedupe_bench.py -- first-occurrence dedupe of a digest table: device-resident kernel time (CUDA events),
host entry point (H2D keys + D2H result inside), and the reference's way (a Python set walk over hex digests,
py/modal/mount.py:498,518-534) on the same rows.   usage: python tools/dedupe_bench.py [n ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from modal_client_b200 import _lib

ctx = _lib.Context(0)
ns = [int(a) for a in sys.argv[1:]] or [100_000, 1_048_576, 8_388_608]
for n in ns:
    rng = np.random.default_rng(n)
    pool = rng.integers(0, 256, (max(1, n * 3 // 4), 32), dtype=np.uint8)  # ~25 % duplicate rows
    keys = pool[rng.integers(0, len(pool), n)]
    d_keys = torch.from_numpy(keys).cuda()
    d_first = torch.empty(n, dtype=torch.int32, device="cuda")
    d_nd = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            ctx.dedupe_device(d_keys.data_ptr(), n, 32, d_first.data_ptr(), d_nd.data_ptr(), st.cuda_stream)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(st)
        for _ in range(reps):
            ctx.dedupe_device(d_keys.data_ptr(), n, 32, d_first.data_ptr(), d_nd.data_ptr(), st.cuda_stream)
        e1.record(st)
        st.synchronize()
    dev_ms = e0.elapsed_time(e1) / reps
    t = time.perf_counter()
    first, nd = ctx.dedupe(keys)
    host_ms = (time.perf_counter() - t) * 1e3
    assert nd == int(d_nd.item()) and np.array_equal(first, d_first.cpu().numpy().view(np.uint32))
    hexes = [k.tobytes().hex() for k in keys[: min(n, 1_048_576)]]
    t = time.perf_counter()
    seen = set()
    for h in hexes:
        if h not in seen:
            seen.add(h)
    cpu_ms = (time.perf_counter() - t) * 1e3 * (n / len(hexes))
    alg = n * 36 / 1e9  # 32 B key read + 4 B result written per row
    print(f"n={n}: distinct={nd}  device {dev_ms:.3f} ms ({alg / dev_ms * 1e3:.0f} GB/s algorithmic, "
          f"{n / dev_ms / 1e3:.0f} M rows/s) | host entry point {host_ms:.2f} ms | python set walk {cpu_ms:.0f} ms"
          f"{' (extrapolated)' if len(hexes) < n else ''}")
