"""tools/sharded_check.py -- NCCL-side parity of sharding.hash_table_sharded: every rank describes the same message set,
hashes its shard on its own GPU (digest table written into device memory by the hash call), the rows are all-gathered
over NCCL and un-permuted; rank 0 compares the whole table with the C oracle.  Covers pageable and page-locked sources,
zero trimming, and SHA-only / MD5-only tables.
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from modal_client_b200 import _backend, _lib, sharding
from modal_client_b200.synth import synth_array
from oracle import c_oracle

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = _lib.Context(local, pinned_bytes=64 << 20, device_bytes=512 << 20)
_backend.set_context(ctx)
rng = np.random.default_rng(17)
lens = np.concatenate([rng.integers(0, 200_000, 3000), [0, 1, 64, 3 << 20, (2 << 20) + 5, 70_000]]).astype(np.uint64)
offs = (np.concatenate([[0], np.cumsum(lens + np.uint64(9))])[:-1] + np.uint64(5)).astype(np.uint64)
buf = synth_array(99, int(offs[-1] + lens[-1]) + 16).copy()
for i in range(0, len(lens), 3):  # zero tails for the trim case
    z = min(int(lens[i]), int(rng.integers(0, 50_000)))
    buf[int(offs[i] + lens[i]) - z : int(offs[i] + lens[i])] = 0
ok = True
pinned = ctx.host_alloc(buf.size)
pinned[:] = buf
for name, src in (("pageable", buf), ("page-locked", pinned)):
    for kw in ({}, {"trim_zeros": True}, {"md5": False}, {"sha256": False}):
        tab = sharding.hash_table_sharded(src, offs, lens, **kw)
        if rank == 0:
            s, m, e = c_oracle.hash_batch(buf, offs, lens, trim=kw.get("trim_zeros", False))
            good = (tab.sha256 is None or np.array_equal(tab.sha256, s)) and (tab.md5 is None or np.array_equal(tab.md5, m)) \
                and np.array_equal(tab.hashed_len, e)
            good = good and (tab.sha256 is None) == (kw.get("sha256") is False) and (tab.md5 is None) == (kw.get("md5") is False)
            ok &= bool(good)
            print(f"sharded over {world} rank(s), {name}, {kw or 'sha256+md5'}: {'exact' if good else 'MISMATCH'}", flush=True)
ctx.host_free(pinned)
ctx.close()
if world > 1:
    dist.destroy_process_group()
if rank == 0:
    print("SHARDED CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)
