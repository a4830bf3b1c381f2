"""tools/sweep.py -- BASELINE.json configs other than the bench headline, HBM-resident, one GPU.
  C5: message size sweep 4 KiB .. 1 GiB, ~8 GiB per point, fused SHA-256+MD5
  C4: one 10 GiB stream as (i) a single chain [bounded sample], (ii) 1280 x 8 MiB SHA-256 blocks, (iii) 160 x 64 MiB MD5 parts
  C3: 12.5 GiB / 131072 files, log-normal sizes (one rank's share of the 100 GiB / 1M-file tree), v1 (SHA+MD5)
      and v2 (trimmed 8 MiB-block SHA-256)
Prints one JSON line per point."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from modal_client_b200 import _lib

ctx = _lib.Context(0)
dev = torch.device("cuda:0")
st = torch.cuda.Stream()
CAP = int(os.environ.get("B200H_SWEEP_BYTES", 8 << 30))
data = torch.empty(CAP + (1 << 20), dtype=torch.uint8, device=dev)
ctx.fill_synth_device(data.data_ptr(), CAP + (1 << 20), 5)
torch.cuda.synchronize()


def run(name, offs, lens, flags, reps=3, **extra):
    n = len(lens)
    off = torch.from_numpy(np.asarray(offs, np.int64)).to(dev)
    ln = torch.from_numpy(np.asarray(lens, np.int64)).to(dev)
    sha = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    md5 = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    tr = torch.empty(n, dtype=torch.int64, device=dev)
    with torch.cuda.stream(st):
        ctx.hash_batch_device(data.data_ptr(), off.data_ptr(), ln.data_ptr(), n, flags, sha.data_ptr(), md5.data_ptr(), tr.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            ctx.hash_batch_device(data.data_ptr(), off.data_ptr(), ln.data_ptr(), n, flags, sha.data_ptr(), md5.data_ptr(), tr.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    total = float(np.asarray(lens, np.float64).sum())
    print(json.dumps({"config": name, "n": n, "bytes": int(total), "ms": round(ms, 3), "GBps": round(total / ms / 1e6, 1),
                      "GiBps": round(total / ms / 1e6 / 1.073741824, 1), "flags": flags, **extra}), flush=True)


BOTH = _lib.SHA256 | _lib.MD5
which = sys.argv[1:] or ["c5", "c4", "c3"]
if "c5" in which:
    for k in range(0, int(os.environ.get("B200H_SWEEP_KMAX", 9)) + 1):
        size = 4096 * 4**k
        n = max(2, CAP // size)
        if size * n > CAP:
            n = CAP // size
        run(f"C5 size={size}", np.arange(n) * size, np.full(n, size), BOTH, reps=2 if size >= (1 << 24) else 3)
if "c4" in which:
    G = min(CAP, 10 << 30)
    nb = G // (8 << 20)
    run("C4(ii) 8MiB-block SHA-256 (+trim scan)", np.arange(nb) * (8 << 20), np.full(nb, 8 << 20), _lib.SHA256 | _lib.TRIM_ZEROS, reps=2, stream_bytes=G)
    npart = G // (64 << 20)
    run("C4(iii) 64MiB-part MD5", np.arange(npart) * (64 << 20), np.full(npart, 64 << 20), _lib.MD5, reps=1, stream_bytes=G)
    run("C4(i) single chain SHA-256+MD5 (256 MiB sample of the stream)", [0], [256 << 20], BOTH, reps=1)
if "c3" in which:
    rng = np.random.default_rng(0)
    n = 131072
    sizes = np.clip(rng.lognormal(np.log(102400) - 1.5**2 / 2, 1.5, n), 1, 1 << 30)
    sizes = np.maximum(1, (sizes * (min(CAP, 12.5 * 2**30) / sizes.sum())).astype(np.int64))
    offs = np.concatenate([[0], np.cumsum((sizes + 15) & ~15)])[:-1]
    run("C3-v1 log-normal files SHA-256+MD5 (one rank's 12.5 GiB share)", offs, sizes, BOTH, reps=2, max_file=int(sizes.max()))
    run("C3-v2 same files, trimmed-block SHA-256 (files < 8 MiB are one block)", offs, np.minimum(sizes, 8 << 20), _lib.SHA256 | _lib.TRIM_ZEROS, reps=2)
    nf = min(n, CAP // 102400)
    run(f"C3-fixed {nf} x 100 KiB SHA-256+MD5", np.arange(nf) * 102400, np.full(nf, 102400), BOTH, reps=3)
