"""tools/dropin_concurrency.py -- what the UNMODIFIED reference callers get when only hash_utils is swapped.

The reference hashes one file / payload per call: `Volume.batch_upload` v1 runs `get_file_upload_spec_from_path` in a
`ThreadPoolExecutor()` (py/modal/volume.py:1209-1216), the map pump calls `get_upload_hashes(payload)` serially on the
event-loop thread (py/modal/_utils/blob_utils.py:345).  With `modal._utils.hash_utils` replaced by
`modal_client_b200.hash_utils` those calls become one-message GPU batches; the library's combining queue merges the
concurrent ones and every BinaryIO digest runs on its own CUDA stream.  This tool measures, on one tree:

  ref+hashlib     the reference's blob_utils on its own hash_utils (CPU), default ThreadPoolExecutor
  ref+b200        the SAME unmodified blob_utils code on modal_client_b200.hash_utils (combining queue on / off)
  b200 batched    modal_client_b200.blob_utils.get_file_upload_specs(paths): the whole tree as one GPU batch
and the serial single-call latency of get_upload_hashes(bytes) against hashlib (the un-batched regression, stated).
Specs are compared field by field (sha256_hex, md5_hex, size, use_blob)."""
import json
import os
import shutil
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path, PurePosixPath

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from modal_client_b200 import _backend, _lib
from modal_client_b200 import blob_utils as our_blob_utils
from modal_client_b200 import hash_utils as our_hash_utils
from modal_client_b200.synth import synth_array
from oracle import ref_shim

GiB = float(1 << 30)
SHAPE = [(1500, 100 * 1024), (400, 1 << 20), (100, 8 << 20)]  # (files, bytes): cached / streamed / blob classes
if len(sys.argv) > 1:
    SHAPE = [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]

root = tempfile.mkdtemp(prefix="b200h_dropin_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
paths = []
k = 0
for n, size in SHAPE:
    blob = synth_array(100 + k, size + n)
    for i in range(n):
        p = os.path.join(root, f"c{k}_{i:05d}.bin")
        blob[i : i + size].tofile(p)  # distinct contents
        paths.append(Path(p))
    k += 1
total = sum(os.path.getsize(p) for p in paths)
print(json.dumps({"tree": SHAPE, "files": len(paths), "bytes": total}), flush=True)

ref_hash, ref_blob, _ = ref_shim.load()
ctx = _lib.Context(0)
_backend.set_context(ctx)
ref_blob_on_b200 = ref_shim.load_blob_utils_on(our_hash_utils)


def thread_pool_specs(blob_utils_module, workers=None):
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as ex:
        specs = list(ex.map(lambda p: blob_utils_module.get_file_upload_spec_from_path(p, PurePosixPath(p.name)), paths))
        used = ex._max_workers
    return time.perf_counter() - t0, specs, used


def key(specs):
    return [(s.sha256_hex, s.md5_hex, s.size, s.use_blob) for s in specs]


def row(name, dt, **kw):
    print(json.dumps({"case": name, "seconds": round(dt, 3), "files_per_s": round(len(paths) / dt), "GiBps": round(total / GiB / dt, 3), **kw}), flush=True)


dt, ref_specs, used = thread_pool_specs(ref_blob)
row(f"ref blob_utils + hashlib, ThreadPoolExecutor() default = {used} workers", dt)
dt, ref_specs, used = thread_pool_specs(ref_blob, os.cpu_count())
row(f"ref blob_utils + hashlib, ThreadPoolExecutor({used})", dt)
want = key(ref_specs)

thread_pool_specs(ref_blob_on_b200)  # warm-up (stream pool, wave buffers)
g0, r0 = ctx.combine_stats()
l0 = ctx.launch_count
dt, specs, used = thread_pool_specs(ref_blob_on_b200)
g1, r1 = ctx.combine_stats()
row(f"ref blob_utils + modal_client_b200.hash_utils (combining queue ON), ThreadPoolExecutor() = {used} workers", dt,
    one_message_calls=r1 - r0, gpu_batches_for_them=g1 - g0, calls_per_batch=round((r1 - r0) / max(g1 - g0, 1), 1),
    kernel_launches=ctx.launch_count - l0, specs_equal_reference=key(specs) == want)
dt, specs, used = thread_pool_specs(ref_blob_on_b200, 128)
g2, r2 = ctx.combine_stats()
row("same, ThreadPoolExecutor(128)", dt, one_message_calls=r2 - r1, gpu_batches_for_them=g2 - g1,
    calls_per_batch=round((r2 - r1) / max(g2 - g1, 1), 1), specs_equal_reference=key(specs) == want)

os.environ["B200H_COMBINE"] = "0"
ctx_off = _lib.Context(0)
os.environ.pop("B200H_COMBINE")
_backend.set_context(ctx_off)
thread_pool_specs(ref_blob_on_b200)
l0 = ctx_off.launch_count
dt, specs, used = thread_pool_specs(ref_blob_on_b200)
row(f"ref blob_utils + modal_client_b200.hash_utils (combining queue OFF), ThreadPoolExecutor() = {used} workers", dt,
    kernel_launches=ctx_off.launch_count - l0, specs_equal_reference=key(specs) == want)
_backend.set_context(ctx)
ctx_off.close()

our_blob_utils.get_file_upload_specs([(p, PurePosixPath(p.name), None) for p in paths[:8]])
l0 = ctx.launch_count
t0 = time.perf_counter()
specs = our_blob_utils.get_file_upload_specs([(p, PurePosixPath(p.name), None) for p in paths])
dt = time.perf_counter() - t0
row("modal_client_b200.blob_utils.get_file_upload_specs (the whole tree as ONE GPU batch, native reader)", dt,
    kernel_launches=ctx.launch_count - l0, specs_equal_reference=key(specs) == want)

# ---- serial callers: nothing to coalesce.  The map pump's get_upload_hashes(payload) on the loop thread, and the
# single-call latency table.
import hashlib

for size, reps in ((4096, 200), (256 * 1024, 100), (4 << 20, 20), (64 << 20, 3)):
    data = synth_array(7, size).tobytes()
    our_hash_utils.get_upload_hashes(data)
    t0 = time.perf_counter()
    for _ in range(reps):
        ours = our_hash_utils.get_upload_hashes(data)
    t_ours = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        theirs = ref_hash.get_upload_hashes(data)
    t_ref = (time.perf_counter() - t0) / reps
    print(json.dumps({"case": f"serial get_upload_hashes(bytes), {size} B", "b200_ms": round(1e3 * t_ours, 3),
                      "hashlib_ms": round(1e3 * t_ref, 3), "b200_MBps": round(size / 1e6 / t_ours, 1),
                      "hashlib_MBps": round(size / 1e6 / t_ref, 1), "equal": ours.sha256_hex() == theirs.sha256_hex()
                      and ours.md5_hex() == theirs.md5_hex()}), flush=True)
import io

for size in (4 << 20, 64 << 20):
    data = synth_array(8, size).tobytes()
    our_hash_utils.get_upload_hashes(io.BytesIO(data))
    t0 = time.perf_counter()
    ours = our_hash_utils.get_upload_hashes(io.BytesIO(data))
    t_ours = time.perf_counter() - t0
    t0 = time.perf_counter()
    theirs = ref_hash.get_upload_hashes(io.BytesIO(data))
    t_ref = time.perf_counter() - t0
    print(json.dumps({"case": f"serial get_upload_hashes(BinaryIO), {size} B (stream API)", "b200_ms": round(1e3 * t_ours, 2),
                      "hashlib_ms": round(1e3 * t_ref, 2), "b200_MBps": round(size / 1e6 / t_ours, 1),
                      "hashlib_MBps": round(size / 1e6 / t_ref, 1), "equal": ours.sha256_hex() == theirs.sha256_hex()
                      and ours.md5_hex() == theirs.md5_hex()}), flush=True)
shutil.rmtree(root, ignore_errors=True)
ctx.close()
