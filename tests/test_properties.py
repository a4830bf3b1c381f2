"""Property tests (hypothesis, CPU) of host-side logic around the path: the directory walk, the shard plan, and the
dedupe oracle that the GPU kernel is checked against."""
import os
import shutil
import tempfile
from pathlib import Path, PurePosixPath

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from modal_client_b200 import sharding, volume
from oracle import ref_port

_names = st.text(alphabet="abcXYZ019 ._-é", min_size=1, max_size=8).filter(lambda s: s not in (".", "..") and "/" not in s)
_tree = st.recursive(
    st.one_of(st.just("file"), st.just("empty_dir"), st.just("link_file"), st.just("link_dir"), st.just("dangling")),
    lambda children: st.dictionaries(_names, children, max_size=4),
    max_leaves=12,
)


def _build(root: Path, node, outside: Path):
    for name, child in node.items():
        p = root / name
        if isinstance(child, dict):
            p.mkdir()
            _build(p, child, outside)
        elif child == "file":
            p.write_bytes(name.encode())
        elif child == "empty_dir":
            p.mkdir()
        elif child == "link_file":
            os.symlink(outside / "target.txt", p)
        elif child == "link_dir":
            os.symlink(outside, p)
        else:
            os.symlink(root / "does-not-exist", p)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(tree=st.dictionaries(_names, _tree, max_size=4), recursive=st.booleans())
def test_scandir_walk_equals_rglob_is_file(tree, recursive):
    """_walk_files selects exactly what the reference's rglob("*") + is_file() + relative_to() selects
    (py/modal/volume.py:1279-1284), for arbitrary nestings of files, directories and symlinks."""
    tmp = Path(tempfile.mkdtemp(prefix="b200h_walk_"))
    try:
        outside = tmp / "outside"
        outside.mkdir()
        (outside / "target.txt").write_bytes(b"t")
        root = tmp / "root"
        root.mkdir()
        _build(root, tree, outside)
        want = sorted((str(sub), (PurePosixPath("/r") / sub.relative_to(root)).as_posix())
                      for sub in (root.rglob("*") if recursive else root.glob("*")) if sub.is_file())
        got = sorted((p, f"/r/{rel}") for p, rel in volume._walk_files(str(root), recursive))
        assert got == want
    finally:
        shutil.rmtree(tmp)


@settings(max_examples=60, deadline=None)
@given(lengths=st.lists(st.integers(0, 1 << 34), max_size=300), world=st.integers(1, 8))
def test_shard_assignment_is_a_balanced_partition(lengths, world):
    plan = sharding.shard_assignment(lengths, world)
    assert len(plan) == world
    allidx = np.concatenate(plan) if plan else np.zeros(0, np.int64)
    assert np.array_equal(np.sort(allidx), np.arange(len(lengths)))          # a partition
    assert all(np.all(np.diff(p) > 0) for p in plan)                         # ascending per rank
    if lengths:
        sizes = [p.size for p in plan]
        assert max(sizes) - min(sizes) <= 1                                   # counts within one
        loads = [int(np.asarray(lengths, dtype=object)[p].sum()) if p.size else 0 for p in plan]
        assert max(loads) - min(loads) <= max(lengths)                        # bytes within one (longest) message
    plan2 = sharding.shard_assignment(lengths, world)
    assert all(np.array_equal(a, b) for a, b in zip(plan, plan2))            # deterministic: every rank agrees


@settings(max_examples=60, deadline=None)
@given(pool=st.lists(st.binary(min_size=32, max_size=32), min_size=1, max_size=20), picks=st.lists(st.integers(0, 19), max_size=200))
def test_first_occurrence_oracle_equals_numpy_unique(pool, picks):
    keys = [pool[i % len(pool)] for i in picks]
    first, nd = ref_port.first_occurrence(keys)
    if keys:
        arr = np.frombuffer(b"".join(keys), np.uint8).reshape(len(keys), 32)
        v = arr.view(np.dtype((np.void, 32))).ravel()
        _, idx, inv = np.unique(v, return_index=True, return_inverse=True)
        assert first == idx[inv].tolist() and nd == len(idx)
    else:
        assert (first, nd) == ([], 0)
    assert all(f <= i and keys[f] == keys[i] for i, f in enumerate(first))
    assert sorted(set(first)) == [i for i, f in enumerate(first) if f == i]


# ------------------------------------------------------------------ the pump's ordered, bounded upload stage

@settings(max_examples=60, deadline=None)
@given(
    n=st.integers(0, 60),
    concurrency=st.integers(1, 7),
    delays=st.lists(st.integers(0, 3), min_size=60, max_size=60),
    fail_at=st.one_of(st.none(), st.integers(0, 59)),
    sink_blocks=st.booleans(),
)
def test_bounded_each_ordered_against_its_sequential_meaning(n, concurrency, delays, fail_at, sink_blocks):
    """Whatever the completion order: results reach the sink in index order, exactly once, never more than
    ``concurrency`` calls run at once, and after a failure at index f the sink has seen a prefix 0..k-1 with k <= f."""
    import asyncio

    from modal_client_b200.async_utils import bounded_each_ordered

    async def run():
        in_flight = peak = 0
        out = []

        async def fn(i):
            nonlocal in_flight, peak
            in_flight += 1
            peak = max(peak, in_flight)
            for _ in range(delays[i]):
                await asyncio.sleep(0)
            in_flight -= 1
            if fail_at is not None and i == fail_at:
                raise KeyError(i)
            return i

        async def slow_append(r):
            await asyncio.sleep(0)
            out.append(r)

        sink = (lambda r: slow_append(r)) if sink_blocks else out.append
        failed = False
        try:
            await bounded_each_ordered(n, fn, concurrency, sink)
        except KeyError:
            failed = True
        assert peak <= concurrency
        if fail_at is not None and fail_at < n:
            assert failed and out == list(range(len(out))) and len(out) <= fail_at
        else:
            assert not failed and out == list(range(n))

    asyncio.run(run())
