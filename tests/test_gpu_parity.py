"""GPU: bit-exact parity of the CUDA path (through the C ABI) against the oracle and golden fixtures."""
import hashlib

import numpy as np
import pytest

from modal_client_b200 import _lib
from modal_client_b200.synth import materialize, synth_array, synth_bytes
from oracle import c_oracle, ref_port

pytestmark = pytest.mark.gpu

BOTH = _lib.SHA256 | _lib.MD5


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0, pinned_bytes=64 << 20, device_bytes=512 << 20)
    yield c
    c.close()


def _check(buf, off, ln, sha, md5):
    s, m, _ = c_oracle.hash_batch(buf, off, ln)
    assert np.array_equal(sha, s)
    assert np.array_equal(md5, m)


def test_every_length_0_to_300_unaligned_packed(ctx):
    lens = np.arange(0, 301, dtype=np.uint64)
    offs = np.concatenate([[3], 3 + np.cumsum(lens + 1)])[:-1].astype(np.uint64)  # odd, unaligned starts
    buf = synth_array(11, int(offs[-1] + lens[-1]) + 8)
    sha, md5, trimmed = ctx.hash_batch_host(buf, offs, lens, BOTH)
    assert np.array_equal(trimmed, lens)
    _check(buf, offs, lens, sha, md5)
    for i in (0, 1, 55, 56, 63, 64, 65, 119, 120, 127, 128, 300):
        msg = buf[int(offs[i]) : int(offs[i] + lens[i])].tobytes()
        assert sha[i].tobytes() == hashlib.sha256(msg).digest()
        assert md5[i].tobytes() == hashlib.md5(msg).digest()


def test_aligned_and_mixed_sizes(ctx):
    rng = np.random.default_rng(5)
    lens = np.concatenate([rng.integers(0, 70000, 700), [0, 64, 1 << 20, (1 << 20) + 1, 262144, 262144]]).astype(
        np.uint64
    )
    offs = np.zeros_like(lens)
    pos = 0
    for i, n in enumerate(lens):
        offs[i] = pos
        pos += (int(n) + 15) & ~15
    buf = synth_array(12, pos + 16)
    sha, md5, _ = ctx.hash_batch_host(buf, offs, lens, BOTH)
    _check(buf, offs, lens, sha, md5)


def test_single_flags(ctx):
    buf = synth_array(13, 100000)
    offs = np.array([0, 1000, 50000], np.uint64)
    lens = np.array([1000, 49000, 50000], np.uint64)
    s, m, _ = c_oracle.hash_batch(buf, offs, lens)
    sha, md5, _ = ctx.hash_batch_host(buf, offs, lens, _lib.SHA256)
    assert md5 is None and np.array_equal(sha, s)
    sha, md5, _ = ctx.hash_batch_host(buf, offs, lens, _lib.MD5)
    assert sha is None and np.array_equal(md5, m)


def test_separate_buffers_absolute_addresses(ctx):
    bufs = [synth_bytes(20 + i, n) for i, n in enumerate([0, 5, 64, 1000, 65536, 300001])]
    sha, md5, _ = ctx.hash_buffers(bufs)
    for b, s, m in zip(bufs, sha, md5):
        assert s.tobytes() == c_oracle.sha256(b) and m.tobytes() == c_oracle.md5(b)


def test_golden_hash_utils_vectors(ctx, golden):
    cases = golden("hash_utils.json")["bytes_cases"]
    bufs = [materialize(c["input"]) for c in cases]
    sha, md5, _ = ctx.hash_buffers(bufs)
    for c, s, m in zip(cases, sha, md5):
        assert s.tobytes().hex() == c["sha256_hex"]
        assert m.tobytes().hex() == c["md5_hex"]


def test_trim_zeros_blocks_golden(ctx, golden):
    for c in golden("blocks.json")["spec2"]:
        data = materialize(c["input"])
        bs = c["patch"].get("BLOCK_SIZE", 8 << 20)
        sha, _, trimmed, _ = ctx.hash_fixed_parts(data, bs, _lib.SHA256 | _lib.TRIM_ZEROS)
        got = [[i * bs, i * bs + int(t), s.tobytes().hex()] for i, (t, s) in enumerate(zip(trimmed, sha))]
        assert got == c["blocks"]


def test_trim_random_zero_runs(ctx):
    rng = np.random.default_rng(9)
    lens = rng.integers(0, 9000, 300).astype(np.uint64)
    gaps = rng.integers(0, 40, 300).astype(np.uint64)
    offs = (np.concatenate([[0], np.cumsum(lens + gaps)])[:-1] + 5).astype(np.uint64)
    buf = synth_array(14, int(offs[-1] + lens[-1]) + 64).copy()
    for i in range(0, 300, 2):
        z = min(int(lens[i]), int(rng.integers(0, 6000)))
        buf[int(offs[i] + lens[i]) - z : int(offs[i] + lens[i])] = 0
    sha, md5, trimmed = ctx.hash_batch_host(buf, offs, lens, BOTH | _lib.TRIM_ZEROS)
    s, m, e = c_oracle.hash_batch(buf, offs, lens, trim=True)
    assert np.array_equal(trimmed, e) and np.array_equal(sha, s) and np.array_equal(md5, m)


def test_multipart_golden(ctx, golden):
    for c in golden("multipart.json")["cases"]:
        data = materialize(c["input"])
        _, md5, _, etag = ctx.hash_fixed_parts(data, c["part_len"], _lib.MD5, want_etag=True)
        assert [m.tobytes().hex() for m in md5] == c["part_md5_hex"]
        assert f"{etag.hex()}-{len(md5)}" == c["etag"]


def test_stream_matches_one_shot(ctx):
    data = synth_bytes(15, 3 * (1 << 20) + 777)
    st = ctx.stream(BOTH)
    pos = 0
    for step in [1, 63, 64, 65, 4096, 65536, 1 << 20, 1 << 21]:
        st.update(data[pos : pos + step])
        pos += step
        s, m = st.digests()  # non-destructive
        assert s == hashlib.sha256(data[:pos]).digest() and m == hashlib.md5(data[:pos]).digest()
    st.update(data[pos:])
    s, m = st.digests()
    assert s == c_oracle.sha256(data) and m == c_oracle.md5(data)
    st.reset()
    st.update(b"abc")
    assert st.digests()[0].hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    st.close()


def test_device_resident_batch_matches_host_and_oracle(ctx):
    import torch

    n, size = 3000, 262144
    dev = torch.device("cuda:0")
    data = torch.empty(n * size, dtype=torch.uint8, device=dev)
    ctx.fill_synth_device(data.data_ptr(), n * size, seed=77)
    torch.cuda.synchronize()
    # the device generator and the numpy generator are the same stream
    assert np.array_equal(data[: 1 << 16].cpu().numpy(), synth_array(77, 1 << 16))
    off = torch.arange(n, dtype=torch.int64, device=dev) * size
    ln = torch.full((n,), size, dtype=torch.int64, device=dev)
    sha = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    md5 = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ctx.hash_batch_device(data.data_ptr(), off.data_ptr(), ln.data_ptr(), n, BOTH, sha.data_ptr(), md5.data_ptr(), 0, st)
    torch.cuda.synchronize()
    sha_h, md5_h = sha.cpu().numpy(), md5.cpu().numpy()
    for i in (0, 1, 31, 32, 1499, n - 1):
        msg = synth_bytes(77, size, start=i * size)
        up = ref_port.upload_hashes(msg)
        assert sha_h[i].tobytes().hex() == up.sha256_hex() and md5_h[i].tobytes().hex() == up.md5_hex()
    # size-independent property: the host path over the same bytes gives the same digest table
    host = data.cpu().numpy()
    sha2, md52, _ = ctx.hash_batch_host(host, off.cpu().numpy(), ln.cpu().numpy(), BOTH)
    assert np.array_equal(sha2, sha_h) and np.array_equal(md52, md5_h)


def test_one_large_message_and_64bit_lengths(ctx):
    n = (1 << 27) + 12345  # 128 MiB: crosses wave / pinned-slot boundaries
    data = synth_array(16, n)
    sha, md5, _ = ctx.hash_batch_host(data, [0], [n], BOTH)
    assert sha[0].tobytes() == hashlib.sha256(data).digest()
    assert md5[0].tobytes() == hashlib.md5(data).digest()


def test_chain_kernel_long_messages_all_tail_shapes(ctx):
    """Messages long enough to be routed to the warp-specialised chain kernel (>= 64 KiB and outliers of the
    batch): every padding shape, tile-boundary lengths, misaligned starts, single-digest modes, trim."""
    base_len = 1024 * 64  # 1024 blocks = 32 full tiles
    extras = [0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 2047, 2048, 2049, 4096 + 56, 100_003]
    lens = np.array([base_len + e for e in extras] + [300_000, 1 << 20], np.uint64)
    offs = np.zeros_like(lens)
    pos = 0
    for i, n in enumerate(lens):
        pos += i % 16  # every misalignment 0..15
        offs[i] = pos
        pos += int(n)
    buf = synth_array(17, pos + 64).copy()
    buf[int(offs[3] + lens[3]) - 70 : int(offs[3] + lens[3])] = 0  # trailing zeros on one message
    for flags in (BOTH, _lib.SHA256, _lib.MD5, BOTH | _lib.TRIM_ZEROS):
        sha, md5, trimmed = ctx.hash_batch_host(buf, offs, lens, flags)
        s, m, e = c_oracle.hash_batch(buf, offs, lens, sha=bool(flags & _lib.SHA256), md5=bool(flags & _lib.MD5),
                                      trim=bool(flags & _lib.TRIM_ZEROS))
        assert np.array_equal(trimmed, e)
        if s is not None:
            assert np.array_equal(sha, s)
        if m is not None:
            assert np.array_equal(md5, m)


def test_chain_and_lane_kernels_agree(ctx, monkeypatch):
    """The opt-in warp-specialised chain kernel (B200H_CHAIN=N, TMA tiles + mbarriers) gives the same table as
    the default lane kernel, for every padding shape / misalignment / digest mode."""
    base_len = 1024 * 64
    extras = [0, 1, 55, 56, 63, 64, 65, 120, 2047, 2048, 2049, 100_003]
    lens = np.array([base_len + e for e in extras] + [70_000, 1 << 20, 5, 123_457, 999_999], np.uint64)
    offs = np.zeros_like(lens)
    pos = 0
    for i, n in enumerate(lens):
        pos += i % 16
        offs[i] = pos
        pos += int(n)
    buf = synth_array(18, pos + 16)
    monkeypatch.setenv("B200H_CHAIN", "1184")
    c2 = _lib.Context(0, pinned_bytes=32 << 20, device_bytes=128 << 20)
    try:
        for flags in (BOTH, _lib.SHA256, _lib.MD5):
            a = ctx.hash_batch_host(buf, offs, lens, flags)
            b = c2.hash_batch_host(buf, offs, lens, flags)
            s, m, _ = c_oracle.hash_batch(buf, offs, lens, sha=bool(flags & _lib.SHA256), md5=bool(flags & _lib.MD5))
            for got in (a, b):
                assert s is None or np.array_equal(got[0], s)
                assert m is None or np.array_equal(got[1], m)
        st = c2.stream(BOTH)  # streamed continuation goes through the chain kernel's resume path
        data = synth_bytes(19, 3 * (1 << 20) + 5)
        st.update(data)
        sd, md = st.digests()
        assert sd == c_oracle.sha256(data) and md == c_oracle.md5(data)
        st.close()
    finally:
        c2.close()


def test_offsets_beyond_4gib_and_million_small_messages(ctx):
    """64-bit offsets (messages placed past the 4 GiB mark of one device buffer) and a 10^6-message batch of
    tiny ragged messages; checked through size-independent properties plus sampled oracle digests."""
    import torch

    dev = torch.device("cuda:0")
    total = (4 << 30) + (64 << 20)
    data = torch.empty(total, dtype=torch.uint8, device=dev)
    ctx.fill_synth_device(data.data_ptr(), total, seed=91)
    # (1) the same 3 MiB of bytes hashed at a low and at a > 4 GiB offset must agree with the oracle on both
    lo, hi = 1 << 20, (4 << 30) + (1 << 20) + 8
    sizes = np.array([0, 1, 63, 64, 65, 4097, 70001, 1 << 20], np.int64)
    offs = np.concatenate([lo + np.cumsum(sizes) - sizes, hi + np.cumsum(sizes) - sizes]).astype(np.int64)
    lens = np.concatenate([sizes, sizes]).astype(np.int64)
    n = len(lens)
    off_t, len_t = torch.from_numpy(offs).to(dev), torch.from_numpy(lens).to(dev)
    sha = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    md5 = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    ctx.hash_batch_device(data.data_ptr(), off_t.data_ptr(), len_t.data_ptr(), n, BOTH, sha.data_ptr(), md5.data_ptr())
    torch.cuda.synchronize()
    sha_h, md5_h = sha.cpu().numpy(), md5.cpu().numpy()
    for i in range(n):
        msg = synth_bytes(91, int(lens[i]), start=int(offs[i]))
        assert sha_h[i].tobytes() == c_oracle.sha256(msg) and md5_h[i].tobytes() == c_oracle.md5(msg), i
    # (2) 10^6 ragged tiny messages (0..255 bytes): every digest of an equal-content message must be equal,
    #     and a sample must match the oracle
    m = 1_000_000
    rng = np.random.default_rng(4)
    lens2 = rng.integers(0, 256, m).astype(np.int64)
    offs2 = (np.arange(m, dtype=np.int64) * 256) % (1 << 20)  # messages alias a 1 MiB window: many duplicates
    off_t, len_t = torch.from_numpy(offs2).to(dev), torch.from_numpy(lens2).to(dev)
    sha = torch.empty((m, 32), dtype=torch.uint8, device=dev)
    md5 = torch.empty((m, 16), dtype=torch.uint8, device=dev)
    ctx.hash_batch_device(data.data_ptr(), off_t.data_ptr(), len_t.data_ptr(), m, BOTH, sha.data_ptr(), md5.data_ptr())
    torch.cuda.synchronize()
    sha_h, md5_h = sha.cpu().numpy(), md5.cpu().numpy()
    key = offs2 * 256 + lens2
    order = np.argsort(key, kind="stable")
    same = key[order][1:] == key[order][:-1]
    assert same.sum() > 100_000
    assert np.array_equal(sha_h[order][1:][same], sha_h[order][:-1][same])
    assert np.array_equal(md5_h[order][1:][same], md5_h[order][:-1][same])
    for i in rng.integers(0, m, 300):
        msg = synth_bytes(91, int(lens2[i]), start=int(offs2[i]))
        assert sha_h[i].tobytes() == c_oracle.sha256(msg) and md5_h[i].tobytes() == c_oracle.md5(msg)


def test_native_file_reader_matches_memory_path_and_reports_io_errors(ctx, tmp_path):
    paths, blobs = [], []
    for i, n in enumerate([0, 1, 4095, 70_000, 300_001, (1 << 20) + 17]):
        data = synth_bytes(50 + i, n)
        p = tmp_path / f"f{i}"
        p.write_bytes(data)
        paths.append(str(p))
        blobs.append(data)
    sizes, modes = ctx.stat_files(paths)
    assert sizes.tolist() == [len(b) for b in blobs]
    sha, md5, trimmed = ctx.hash_files(paths, sizes, 0, BOTH)
    for b, s, m in zip(blobs, sha, md5):
        assert s.tobytes() == c_oracle.sha256(b) and m.tobytes() == c_oracle.md5(b)
    # fixed parts, file-major, trimmed
    sha, _, trimmed = ctx.hash_files(paths, sizes, 65536, _lib.SHA256 | _lib.TRIM_ZEROS)
    row = 0
    for b in blobs:
        for o in range(0, len(b), 65536):
            part = b[o : o + 65536].rstrip(b"\0")
            assert int(trimmed[row]) == len(part) and sha[row].tobytes() == c_oracle.sha256(part)
            row += 1
    assert row == len(trimmed)
    with pytest.raises(_lib.B200HashError, match="nope"):
        ctx.stat_files(paths + [str(tmp_path / "nope")])
    (tmp_path / "f3").write_bytes(b"short now")  # shrank after stat
    with pytest.raises(_lib.B200HashError, match="shorter"):
        ctx.hash_files(paths, sizes, 0, BOTH)
    s2, _, _ = ctx.hash_files(paths[:3], sizes[:3], 0, _lib.SHA256)  # the context is still usable
    assert s2[2].tobytes() == c_oracle.sha256(blobs[2])


def test_messages_larger_than_the_staging_wave_are_segmented():
    """A context with tiny staging (1 MiB waves) must hash multi-MiB messages by carrying the chaining state
    across waves -- from pageable memory, from page-locked memory and from files -- mixed with small messages."""
    c = _lib.Context(0, pinned_bytes=1 << 20, device_bytes=2 << 20)
    try:
        lens = np.array([100, 5 * (1 << 20) + 77, 0, 3 * (1 << 20), 64, (1 << 20) - 16, (1 << 20) - 15, 2_500_001], np.uint64)
        offs = (np.concatenate([[0], np.cumsum(lens + 5)])[:-1] + 3).astype(np.uint64)
        buf = synth_array(60, int(offs[-1] + lens[-1]) + 16)
        s, m, _ = c_oracle.hash_batch(buf, offs, lens)
        for flags in (BOTH, _lib.SHA256, _lib.MD5):
            sha, md5, trimmed = c.hash_batch_host(buf, offs, lens, flags)
            assert np.array_equal(trimmed, lens)
            assert sha is None or np.array_equal(sha, s)
            assert md5 is None or np.array_equal(md5, m)
        pinned = c.host_alloc(buf.size)
        pinned[:] = buf
        sha, md5, _ = c.hash_batch_host(pinned, offs, lens, BOTH)
        assert np.array_equal(sha, s) and np.array_equal(md5, m)
        c.host_free(pinned)
        with pytest.raises(_lib.B200HashError, match="TRIM_ZEROS"):
            c.hash_batch_host(buf, offs, lens, BOTH | _lib.TRIM_ZEROS)
        import tempfile, os

        with tempfile.TemporaryDirectory() as d:
            paths = []
            for i, (o, n) in enumerate(zip(offs, lens)):
                p = os.path.join(d, f"f{i}")
                buf[int(o) : int(o + n)].tofile(p)
                paths.append(p)
            sizes, _ = c.stat_files(paths)
            sha, md5, _ = c.hash_files(paths, sizes, 0, BOTH)
            assert np.array_equal(sha, s) and np.array_equal(md5, m)
    finally:
        c.close()


def test_bit_length_beyond_32_bits(ctx):
    """>= 512 MiB: the message bit length needs the high word (SHA-256 big-endian, MD5 little-endian)."""
    n = (512 << 20) + 4099
    data = synth_array(61, n)
    sha, md5, _ = ctx.hash_batch_host(data, [0], [n], BOTH)
    assert sha[0].tobytes() == hashlib.sha256(data).digest()
    assert md5[0].tobytes() == hashlib.md5(data).digest()


def test_hash_tensors_in_place(ctx):
    """Tensors that already live in HBM are hashed where they are (absolute device addresses, NULL base)."""
    import torch

    from modal_client_b200 import batch

    g = torch.Generator(device="cuda").manual_seed(3)
    tensors = [torch.randn(257, 129, device="cuda", generator=g), torch.arange(100_003, device="cuda", dtype=torch.int32),
               torch.zeros(0, device="cuda"), torch.randint(0, 255, (3, 70_001), device="cuda", dtype=torch.uint8, generator=g),
               torch.randn(1 << 20, device="cuda", generator=g).to(torch.bfloat16)]
    with torch.cuda.stream(torch.cuda.Stream()):
        sha, md5 = batch.hash_table_tensors(tensors, ctx=ctx)
        torch.cuda.current_stream().synchronize()
    for t, s, m in zip(tensors, sha.cpu().numpy(), md5.cpu().numpy()):
        raw = t.view(torch.uint8).cpu().numpy().tobytes() if t.numel() else b""
        assert s.tobytes() == c_oracle.sha256(raw) and m.tobytes() == c_oracle.md5(raw)
    sha2, _ = batch.hash_table_tensors(tensors[:2], md5=False, ctx=ctx)  # legacy default stream path
    assert sha2[1].cpu().numpy().tobytes() == c_oracle.sha256(tensors[1].cpu().numpy().tobytes())


def _np_first(keys):
    """numpy restatement used at sizes where the dict walk of the oracle is slow: np.unique over rows."""
    k = np.ascontiguousarray(keys)
    v = k.view(np.dtype((np.void, k.shape[1]))).ravel()
    _, idx, inv = np.unique(v, return_index=True, return_inverse=True)
    return idx[inv].astype(np.uint32), len(idx)


@pytest.mark.parametrize("width", [32, 16])
def test_dedupe_matches_the_mount_set_walk(ctx, width):
    """b200h_dedupe_host == the `accounted_hashes` walk of mount.py:498,518-534 (oracle.ref_port.first_occurrence)."""
    rng = np.random.default_rng(width)
    cases = [
        np.zeros((0, width), np.uint8),                                           # empty table
        rng.integers(0, 256, (1, width), dtype=np.uint8),                         # single row
        np.repeat(rng.integers(0, 256, (1, width), dtype=np.uint8), 1000, 0),     # every row the same content
        rng.integers(0, 256, (5000, width), dtype=np.uint8),                      # all distinct
        rng.integers(0, 256, (97, width), dtype=np.uint8)[rng.integers(0, 97, 20000)],  # heavy duplication
    ]
    # rows that share the first 8 bytes (same home slot, full-key comparison decides) and differ only in the last byte
    coll = np.repeat(rng.integers(0, 256, (1, width), dtype=np.uint8), 600, 0)
    coll[:, -1] = np.arange(600) % 7
    cases.append(coll)
    # digests of real (duplicated) contents
    blobs = [synth_bytes(i % 37, 100 + (i % 37)) for i in range(500)]
    dig = np.array([np.frombuffer(hashlib.sha256(b).digest() if width == 32 else hashlib.md5(b).digest(), np.uint8)
                    for b in blobs])
    cases.append(dig)
    for keys in cases:
        first, nd = ctx.dedupe(keys)
        want, want_nd = ref_port.first_occurrence(keys)
        assert first.tolist() == want and nd == want_nd


def test_dedupe_million_rows_and_device_entry_point(ctx):
    import torch

    rng = np.random.default_rng(5)
    pool = rng.integers(0, 256, (300_000, 32), dtype=np.uint8)
    keys = pool[rng.integers(0, len(pool), 1_000_000)]
    want, want_nd = _np_first(keys)
    first, nd = ctx.dedupe(keys)
    assert np.array_equal(first, want) and nd == want_nd
    # device-resident table (what the batch path hands over): same answer, nothing leaves HBM but the result
    d_keys = torch.from_numpy(keys).cuda()
    d_first = torch.empty(len(keys), dtype=torch.int32, device="cuda")
    d_nd = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(2):  # the table scratch is reused: a second run must give the same result
            ctx.dedupe_device(d_keys.data_ptr(), len(keys), 32, d_first.data_ptr(), d_nd.data_ptr(), st.cuda_stream)
        st.synchronize()
    assert np.array_equal(d_first.cpu().numpy().view(np.uint32), want) and int(d_nd.item()) == want_nd
    with pytest.raises(_lib.B200HashError):
        ctx.dedupe_device(d_keys.data_ptr(), 10, 24, d_first.data_ptr())  # unsupported key width


def test_outlier_routing_policy(ctx):
    """The planner sends true outliers -- and only those -- to the chain kernel (regression test: an overflow in
    the bucket bounds once made it route nothing at all).  Digests are checked in every case."""
    import torch

    def run(lens, flags=BOTH):
        lens = np.asarray(lens, np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens + np.uint64(7))])[:-1].astype(np.uint64)
        buf = synth_array(23, int(offs[-1] + lens[-1]) + 8)
        sha, md5, _ = ctx.hash_batch_host(buf, offs, lens, flags)
        routed = ctx.last_outlier_count
        s, m, _ = c_oracle.hash_batch(buf, offs, lens, sha=bool(flags & _lib.SHA256), md5=bool(flags & _lib.MD5))
        assert (s is None or np.array_equal(sha, s)) and (m is None or np.array_equal(md5, m))
        return routed

    assert run([3 << 20] + [5000] * 3000) == 1                     # one long file in a tree of small ones
    assert run([3 << 20, (3 << 20) - 4097, 1 << 20] + [70_000] * 500) == 2   # the 1 MiB one is < half the longest
    assert run([1 << 20] * 40, _lib.SHA256) == 40                  # few equally long messages: all of them fit
    assert run([100_000] * 2000) == 0                              # nothing stands out (one staging wave)
    assert run([80 * 1024] * 700) == 0                             # more equally long messages than chain CTAs
    assert run([2 << 20]) == 1 and run([100]) == 0 and run([0]) == 0
    torch.cuda.synchronize()
