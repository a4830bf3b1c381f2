"""Multi-GPU: shard independent messages over the ranks of one node, all-gather the digest table.

Files / map inputs / 8 MiB blocks / multipart parts carry no cross-message state, so the path shards with
no data-path collective; the only exchange is the fixed-width digest table (48 B per message:
SHA-256 || MD5), all-gathered once over NCCL (NVLink / NVSwitch) -- or gloo in the CPU tests.
One process per GPU (torchrun); every rank ends up with the full table in the caller's message order.
"""
from __future__ import annotations

import numpy as np

from ._backend import get_context
from ._lib import MD5, SHA256, TRIM_ZEROS
from .batch import DigestTable


def gpu_node_cpus(device: int) -> set[int] | None:
    """CPUs of the NUMA node CUDA device ``device`` hangs off (sysfs), or None when that cannot be told."""
    import os

    try:
        import torch

        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            text = f.read().strip()
    except Exception:
        return None
    cpus: set[int] = set()
    for part in text.split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    cpus &= os.sched_getaffinity(0)
    return cpus or None


def bind_process_to_gpu_node(device: int) -> bool:
    """One process per GPU: run this process (and every thread it starts later) on the CPUs next to its GPU, so that
    the payloads it allocates, the pinned staging ring and the packer threads all sit in the memory the GPU's DMA
    engine reads from.  torchrun does not do this; without it 8 ranks packing pageable payloads push most bytes
    across the socket interconnect (measured: map-pump e2e 8 GPUs 64 -> see profiles/).  Returns whether it bound."""
    import os

    cpus = gpu_node_cpus(device)
    if not cpus:
        return False
    try:
        os.sched_setaffinity(0, cpus)
        return True
    except OSError:
        return False


def shard_assignment(lengths, world: int) -> list[np.ndarray]:
    """Byte-balanced partition of message indices over ``world`` ranks: sort by length (longest first) and
    deal in boustrophedon order (0..w-1, w-1..0, ...), the vectorisable cousin of LPT list scheduling.
    Returns one ascending index array per rank.  Deterministic, so every rank computes the same plan."""
    lengths = np.asarray(lengths, dtype=np.uint64)
    n = lengths.size
    order = np.argsort(-lengths.astype(np.int64), kind="stable")
    pos = np.arange(n)
    lap, col = pos // world, pos % world
    owner_sorted = np.where(lap % 2 == 0, col, world - 1 - col)
    owner = np.empty(n, np.int64)
    owner[order] = owner_sorted
    return [np.flatnonzero(owner == r) for r in range(world)]


def all_gather_table(local: np.ndarray, counts: list[int], group=None, device=None) -> list[np.ndarray]:
    """All-gather per-rank uint8[count_r, W] tables (padded to the largest count) -> list of per-rank tables."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    width = local.shape[1]
    cap = max(max(counts), 1)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    send = torch.zeros((cap, width), dtype=torch.uint8, device=device)
    if local.shape[0]:
        send[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    recv = torch.empty((world * cap, width), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    host = recv.cpu().numpy().reshape(world, cap, width)
    return [host[r, : counts[r]] for r in range(world)]


def hash_table_sharded(base, offsets, lengths, *, sha256: bool = True, md5: bool = True, trim_zeros: bool = False,
                       group=None, ctx=None) -> DigestTable:
    """Every rank passes the SAME (offsets, lengths) description of the whole message set and a ``base`` in
    which at least its own shard is readable; each rank hashes only its shard on its own GPU, then the
    48-byte rows (+ 8-byte hashed length) are all-gathered and un-permuted into message order."""
    import torch.distributed as dist

    offsets = np.asarray(offsets, dtype=np.uint64)
    lengths = np.asarray(lengths, dtype=np.uint64)
    n = offsets.size
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        flags = (SHA256 if sha256 else 0) | (MD5 if md5 else 0) | (TRIM_ZEROS if trim_zeros else 0)
        s, m, t = (ctx or get_context()).hash_batch_host(base, offsets, lengths, flags)
        return DigestTable(s, m, t)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    plan = shard_assignment(lengths, world)
    mine = plan[rank]
    flags = (SHA256 if sha256 else 0) | (MD5 if md5 else 0) | (TRIM_ZEROS if trim_zeros else 0)
    s, m, t = (ctx or get_context()).hash_batch_host(base, offsets[mine], lengths[mine], flags)
    row = np.zeros((mine.size, 56), np.uint8)
    if s is not None:
        row[:, :32] = s
    if m is not None:
        row[:, 32:48] = m
    row[:, 48:56] = np.ascontiguousarray(t, dtype="<u8").view(np.uint8).reshape(-1, 8)
    tables = all_gather_table(row, [p.size for p in plan], group=group)
    full = np.zeros((n, 56), np.uint8)
    for r in range(world):
        full[plan[r]] = tables[r]
    return DigestTable(
        np.ascontiguousarray(full[:, :32]) if sha256 else None,
        np.ascontiguousarray(full[:, 32:48]) if md5 else None,
        np.ascontiguousarray(full[:, 48:56]).view("<u8").reshape(-1).astype(np.uint64),
    )
