"""tools/c3_sharded.py -- BASELINE config C3: Volume.batch_upload synthetic tree, 100 GiB / 1 Mi files, sharded
over the ranks of one node (torchrun).  Sizes are log-normal (sigma 1.5, mean ~100 KiB, clipped to 1 GiB, seed 0),
rescaled to the requested total.  Every rank computes the same plan (sharding.shard_assignment), generates its own
shard's bytes in HBM, hashes v1 (whole-file SHA-256+MD5) and v2 (zero-trimmed <= 8 MiB blocks, SHA-256), and the
56-byte rows are all-gathered.  Prints one JSON line per mode from rank 0.
  torchrun --nproc-per-node 8 tools/c3_sharded.py [total_GiB=100] [files=1048576]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from modal_client_b200 import _lib, sharding

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
total_gib = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = _lib.Context(local)
rng = np.random.default_rng(0)
sizes = np.clip(rng.lognormal(np.log(102400) - 1.5**2 / 2, 1.5, nfiles), 1, 1 << 30)
sizes = np.maximum(1, (sizes * (total_gib * 2**30 / sizes.sum())).astype(np.int64))
mine = sharding.shard_assignment(sizes, world)[rank]
my_sizes = sizes[mine]
offs = np.concatenate([[0], np.cumsum((my_sizes + 15) & ~15)])
nbytes = int(offs[-1])
offs = offs[:-1]
data = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
ctx.fill_synth_device(data.data_ptr(), (nbytes + 64) & ~7, 0xC3 + rank)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)


def run(name, off, ln, flags):
    n = len(ln)
    off_t, len_t = torch.from_numpy(np.ascontiguousarray(off, np.int64)).to(dev), torch.from_numpy(np.ascontiguousarray(ln, np.int64)).to(dev)
    sha = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    md5 = torch.zeros((n, 16), dtype=torch.uint8, device=dev)
    tr = torch.empty(n, dtype=torch.int64, device=dev)
    counts = [0] * world
    if world > 1:
        c = torch.tensor([n], device=dev)
        allc = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(allc, c)
        counts = [int(x.item()) for x in allc]
        cap = max(counts)
        send = torch.zeros((cap, 56), dtype=torch.uint8, device=dev)
        recv = torch.empty((world * cap, 56), dtype=torch.uint8, device=dev)

    def step():
        ctx.hash_batch_device(data.data_ptr(), off_t.data_ptr(), len_t.data_ptr(), n, flags, sha.data_ptr(), md5.data_ptr(), tr.data_ptr(), st.cuda_stream)
        if world > 1:
            send[:n, :32] = sha
            send[:n, 32:48] = md5
            send[:n, 48:56] = tr.view(torch.uint8).view(n, 8)
            dist.all_gather_into_tensor(recv, send)

    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    step()
    e1.record(st)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    b = torch.tensor([float(np.asarray(ln, np.float64).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"config": name, "n_gpus": world, "files_total": nfiles, "bytes_total": int(b.item()), "ms_max_over_ranks": round(t.item(), 2),
                          "GiBps": round(b.item() / 2**30 / (t.item() / 1e3), 1), "largest_file": int(sizes.max()),
                          "msgs_on_rank0": n}), flush=True)


run("C3-v1 whole-file SHA-256+MD5", offs, my_sizes, _lib.SHA256 | _lib.MD5)
B = 8 << 20
nblk = (my_sizes + B - 1) // B
boff = np.repeat(offs, nblk) + (np.arange(nblk.sum()) - np.repeat(np.cumsum(nblk) - nblk, nblk)) * B
blen = np.minimum(B, np.repeat(my_sizes, nblk) - (boff - np.repeat(offs, nblk)))
run("C3-v2 zero-trimmed 8 MiB blocks SHA-256", boff, blen, _lib.SHA256 | _lib.TRIM_ZEROS)
if world > 1:
    dist.destroy_process_group()
