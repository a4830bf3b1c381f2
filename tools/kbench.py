"""tools/kbench.py -- kernel-only timing of the HBM-resident hash path for one (n, size) point.
usage: python tools/kbench.py [n] [size_bytes] [reps] [flags]   (flags: 3=sha+md5, 1=sha, 2=md5)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from modal_client_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
size = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = _lib.Context(0)
dev = torch.device("cuda:0")
data = torch.empty(n * size, dtype=torch.uint8, device=dev)
ctx.fill_synth_device(data.data_ptr(), n * size, 1)
off = torch.arange(n, dtype=torch.int64, device=dev) * size
ln = torch.full((n,), size, dtype=torch.int64, device=dev)
sha = torch.empty((n, 32), dtype=torch.uint8, device=dev)
md5 = torch.empty((n, 16), dtype=torch.uint8, device=dev)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(2):
        ctx.hash_batch_device(data.data_ptr(), off.data_ptr(), ln.data_ptr(), n, flags, sha.data_ptr(), md5.data_ptr(), 0, st.cuda_stream)
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        ctx.hash_batch_device(data.data_ptr(), off.data_ptr(), ln.data_ptr(), n, flags, sha.data_ptr(), md5.data_ptr(), 0, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
kms, kn = ctx.profile_read()
gb = n * size / 1e9
print(f"lib={os.environ.get('B200H_LIB','default')} n={n} size={size} flags={flags}: step {ms:.3f} ms ({gb/ms*1e3:.1f} GB/s, {gb/ms*1e3/1.073741824:.1f} GiB/s) | lane kernel {kms/kn:.3f} ms ({gb/(kms/kn)*1e3:.1f} GB/s)")
