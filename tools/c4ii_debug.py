import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modal_client_b200 import _lib
ctx = _lib.Context(0)
dev = torch.device("cuda:0")
n, size = 1024, 8 << 20
data = torch.empty(n * size, dtype=torch.uint8, device=dev)
ctx.fill_synth_device(data.data_ptr(), n * size, 5)
off = torch.arange(n, dtype=torch.int64, device=dev) * size
ln = torch.full((n,), size, dtype=torch.int64, device=dev)
sha = torch.empty((n, 32), dtype=torch.uint8, device=dev); md5 = torch.empty((n, 16), dtype=torch.uint8, device=dev)
tr = torch.empty(n, dtype=torch.int64, device=dev)
st = torch.cuda.Stream()
ctx.profile_enable(True)
for flags in (5, 1, 5):
    with torch.cuda.stream(st):
        ctx.hash_batch_device(data.data_ptr(), off.data_ptr(), ln.data_ptr(), n, flags, sha.data_ptr(), md5.data_ptr(), tr.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ctx.hash_batch_device(data.data_ptr(), off.data_ptr(), ln.data_ptr(), n, flags, sha.data_ptr(), md5.data_ptr(), tr.data_ptr(), st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
    kms, kn = ctx.profile_read()
    print("lane kernel ms", round(kms / max(kn, 1), 2), "launches", kn, end=" | ")
    print("flags", flags, "ms", round(e0.elapsed_time(e1), 2), "outliers", ctx.last_outlier_count, "trim min/max", int(tr.min()), int(tr.max()), flush=True)
