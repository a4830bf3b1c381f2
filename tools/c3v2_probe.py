"""tools/c3v2_probe.py -- where does C3-v2 (zero-trimmed <= 8 MiB blocks, SHA-256) spend its time on one rank's share
(12.5 GiB / 131 072 files)?  Times the batch (a) as shipped, (b) with the outlier path off (B200H_NO_OUTLIERS),
(c) only the messages the planner sends to the chain kernel, alone, (d) only the lane messages; prints the lane kernel's
own time for each (b200h_profile)."""
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
torch.cuda.set_stream(st)
rng = np.random.default_rng(0)
nfiles, world = 1 << 20, 8
sizes = np.clip(rng.lognormal(np.log(102400) - 1.5**2 / 2, 1.5, nfiles), 1, 1 << 30)
sizes = np.maximum(1, (sizes * (100 * 2**30 / sizes.sum())).astype(np.int64))
from modal_client_b200 import sharding

mine = sharding.shard_assignment(sizes, world)[0]
my = sizes[mine].astype(np.uint64)
offs = np.concatenate([[0], np.cumsum((my + np.uint64(15)) & ~np.uint64(15))]).astype(np.uint64)
nbytes = int(offs[-1])
offs = offs[:-1]
data = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
ctx.fill_synth_device(data.data_ptr(), (nbytes + 64) & ~7, 0xC3)
B = 8 << 20
nblk = ((my + np.uint64(B - 1)) // np.uint64(B)).astype(np.int64)
first = np.cumsum(nblk) - nblk
within = np.arange(int(nblk.sum())) - np.repeat(first, nblk)
boff = (np.repeat(offs, nblk) + (within * B).astype(np.uint64)).astype(np.uint64)
blen = np.minimum(np.uint64(B), np.repeat(my, nblk) - (within * B).astype(np.uint64)).astype(np.uint64)


def run(name, off, ln, flags, h_lengths=True):
    n = len(ln)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_len = torch.from_numpy(ln.astype(np.int64)).to(dev)
    sha = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    tr = torch.empty(n, dtype=torch.int64, device=dev)

    def step():
        ctx.hash_batch_device(data.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), n, flags, sha.data_ptr(), 0, tr.data_ptr(),
                              st.cuda_stream, h_lengths=ln if h_lengths else None)

    step()
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    step()
    e1.record(st)
    torch.cuda.synchronize()
    kms, kn = ctx.profile_read()
    ctx.profile_enable(False)
    print(json.dumps({"case": name, "messages": n, "GiB": round(float(ln.sum()) / 2**30, 2), "ms": round(e0.elapsed_time(e1), 2),
                      "lane_kernel_ms": round(kms / max(kn, 1), 2), "to_chain_kernel": ctx.last_outlier_count,
                      "longest": int(ln.max())}), flush=True)


S = _lib.SHA256
run("a. shipped: SHA-256 + TRIM (planner count read back)", boff, blen, S | _lib.TRIM_ZEROS, h_lengths=False)
run("a'. same without TRIM (host-side plan)", boff, blen, S)
run("b. outlier path off (everything on lanes)", boff, blen, S | _lib.NO_OUTLIERS)
big = blen >= np.uint64(4 << 20)
run("c. only the blocks >= 4 MiB (what the planner routes), alone", boff[big], blen[big], S)
run("c'. the same blocks forced onto lanes", boff[big], blen[big], S | _lib.NO_OUTLIERS)
run("d. only the blocks < 4 MiB", boff[~big], blen[~big], S)
mid = (blen >= np.uint64(1 << 20)) & ~big
run("e. only the blocks in [1 MiB, 4 MiB)", boff[mid], blen[mid], S)
ctx.close()
