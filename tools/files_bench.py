"""tools/files_bench.py -- Volume.batch_upload spec building over a real directory tree, end to end from files:
the reference's way (ThreadPoolExecutor over get_file_upload_spec_from_path == oracle/ref_port.file_spec_fields on
hashlib) vs modal_client_b200.blob_utils.get_file_upload_specs / file_upload_specs2 (native reader -> GPU).
usage: python tools/files_bench.py [n_files=20000] [total_GiB=2] [dir=/dev/shm|/tmp] [sigma=1.5]"""
import asyncio
import os
import shutil
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path, PurePosixPath

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modal_client_b200 import _backend, _lib, blob_utils
from modal_client_b200.synth import synth_array
from oracle import ref_port  # CPU baseline leg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
total = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
where = sys.argv[3] if len(sys.argv) > 3 else ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
root = Path(tempfile.mkdtemp(prefix="b200h_tree_", dir=where))
try:
    rng = np.random.default_rng(0)
    sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 1.5
    sizes = np.clip(rng.lognormal(np.log(102400) - sigma**2 / 2, sigma, n), 1, 1 << 28)
    sizes = np.maximum(1, (sizes * (total * 2**30 / sizes.sum())).astype(np.int64))
    blob = synth_array(9, int(sizes.max()) + (1 << 20))
    files = []
    t0 = time.perf_counter()
    for i, s in enumerate(sizes):
        d = root / f"d{i % 64}"
        if i < 64:
            d.mkdir()
        p = d / f"f{i}.bin"
        with open(p, "wb") as f:
            f.write(blob[i % 4096 : i % 4096 + int(s)].tobytes())
        files.append((p, PurePosixPath(f"/vol/d{i % 64}/f{i}.bin"), None))
    nbytes = int(sizes.sum())
    print(f"tree: {n} files, {nbytes / 2**30:.2f} GiB in {where} (written in {time.perf_counter() - t0:.1f}s), largest {sizes.max() / 2**20:.1f} MiB")

    ctx = _lib.Context(0)
    _backend.set_context(ctx)

    def report(name, dt):
        print(f"{name:66s} {dt:7.3f} s  {nbytes / 2**30 / dt:8.2f} GiB/s  {n / dt:10.0f} files/s", flush=True)

    def cpu_one(item):
        with open(item[0], "rb") as fp:
            return ref_port.file_spec_fields(fp)

    best = None
    for w in sorted({min(32, (os.cpu_count() or 1) + 4), os.cpu_count() or 1}):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=w) as ex:
            cpu = list(ex.map(cpu_one, files))
        dt = time.perf_counter() - t0
        report(f"reference way: ThreadPool({w}) x hashlib per file (v1 specs)", dt)
    blob_utils.get_file_upload_specs(files[:256])  # warm up (allocations)
    t0 = time.perf_counter()
    specs = blob_utils.get_file_upload_specs(files, cache_small_content=False)
    report("b200: get_file_upload_specs (stat + read + hash in the library, v1)", time.perf_counter() - t0)
    assert all(a.sha256_hex == b["sha256_hex"] and a.md5_hex == b["md5_hex"] for a, b in zip(specs, cpu)), "MISMATCH vs hashlib"
    t0 = time.perf_counter()
    specs2 = asyncio.run(blob_utils.file_upload_specs2(files))
    report("b200: file_upload_specs2 (v2: trimmed 8 MiB blocks, SHA-256)", time.perf_counter() - t0)
    t0 = time.perf_counter()
    sizes_l, _ = ctx.stat_files([str(f[0]) for f in files])
    sha, md5, _ = ctx.hash_files([str(f[0]) for f in files], sizes_l, 0, 3)
    report("b200: raw stat_files + hash_files (no Python spec objects)", time.perf_counter() - t0)
    print("digests match hashlib on every file: yes")
finally:
    shutil.rmtree(root, ignore_errors=True)
