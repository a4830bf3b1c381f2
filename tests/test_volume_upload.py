"""Volume.batch_upload v1 / v2 against in-process fakes that follow the reference's test doubles
(py/test/conftest.py:2224-2231 MountPutFile, :3024-3112 VolumePutFiles[2], :3399-3413 block PUT)."""
import asyncio
import hashlib
import io
import types

import pytest
from aiohttp import web

from modal_client_b200 import blob_utils, volume
from modal_client_b200.synth import synth_bytes
from tests.blob_server import FakeBlobStub, running_blob_server


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    return request.getfixturevalue("fake_backend" if request.param == "fake" else "gpu_backend")


class AlreadyExistsError(Exception):
    pass


class FakeVolumeStub(FakeBlobStub):
    def __init__(self, host, store, block_size):
        super().__init__(host, multipart_threshold=10_000_000)
        self.store, self.block_size = store, block_size
        self.files_sha2data, self.volume_files, self.blocks = {}, {}, {}
        self.mount_put_calls = 0

    async def MountPutFile(self, req):
        self.mount_put_calls += 1
        if req.WhichOneof("data_oneof") is not None:
            self.files_sha2data[req.sha256_hex] = {"data": req.data, "data_blob_id": req.data_blob_id}
            return types.SimpleNamespace(exists=True)
        return types.SimpleNamespace(exists=req.sha256_hex in self.files_sha2data)

    async def VolumePutFiles(self, req):
        for f in req.files:
            if f.filename in self.volume_files and req.disallow_overwrite_existing_files:
                raise AlreadyExistsError(f"{f.filename}: already exists")
            blob = self.files_sha2data[f.sha256_hex]
            data = blob["data"] if blob["data"] is not None else self.store.blobs[blob["data_blob_id"]]
            self.volume_files[f.filename] = (data, f.mode, f.sha256_hex)

    async def VolumePutFiles2(self, req):
        missing, created = [], {}
        for fi, f in enumerate(req.files):
            if f.path in self.volume_files and req.disallow_overwrite_existing_files:
                raise AlreadyExistsError(f"{f.path}: already exists")
            parts, file_missing = [], []
            for bi, b in enumerate(f.blocks):
                bid = b.contents_sha256.hex()
                ok = b.put_response == b"test-put-response:" + bid.encode() and bid in self.blocks
                if ok:
                    want = min(self.block_size, max(0, f.size - bi * self.block_size))
                    parts.append(self.blocks[bid].ljust(want, b"\0"))
                else:
                    file_missing.append(types.SimpleNamespace(file_index=fi, block_index=bi,
                                                              put_url=f"{self.host}/block/test-put-request"))
            if file_missing:
                missing.extend(file_missing)
            else:
                created[f.path] = (b"".join(parts), f.mode, None)
        if not missing:
            self.volume_files.update(created)
        return types.SimpleNamespace(missing_blocks=missing)


def _add_block_route(app_store, stub_holder, block_size):
    async def put_block(request: web.Request):
        body = await request.read()
        if len(body) > block_size:
            return web.Response(status=413, text="block too big")
        bid = hashlib.sha256(body).hexdigest()
        stub_holder["stub"].blocks[bid] = body
        stub_holder["stub"].block_puts = getattr(stub_holder["stub"], "block_puts", 0) + 1
        return web.Response(text=f"test-put-response:{bid}")

    return put_block


def _tree(tmp_path, sizes, seed0):
    files = {}
    root = tmp_path / "tree"
    (root / "sub" / "deep").mkdir(parents=True)
    for i, n in enumerate(sizes):
        data = bytearray(synth_bytes(seed0 + i, n))
        if n > 3000:
            data[n // 2 : n // 2 + 1500] = bytes(1500)
            data[-700:] = bytes(700)
        rel = ["a.bin", "sub/b.bin", "sub/deep/c.bin", "sub/d.bin", "e.bin", "sub/deep/f.bin"][i % 6] + str(i)
        (root / rel).write_bytes(bytes(data))
        files[rel] = bytes(data)
    return root, files


def test_batch_upload_v1_dedupes_and_uploads(backend, monkeypatch, tmp_path):
    monkeypatch.setattr(blob_utils, "LARGE_FILE_LIMIT", 20_000)  # make some files go through the blob path

    async def run():
        async with running_blob_server() as (host, store):
            stub = FakeVolumeStub(host, store, 8 << 20)
            client = types.SimpleNamespace(stub=stub)
            root, files = _tree(tmp_path, [0, 11, 5000, 19_999, 20_000, 70_001], 200)
            (root / "dup.bin").write_bytes(files["sub/b.bin1"])  # same content twice: second is deduped
            async with volume.VolumeUploadContextManager("vo-1", client) as batch:
                batch.put_directory(root, "/data")
                batch.put_file(io.BytesIO(b"hello world, this is a lot of text"), "/data/from_fileobj")
            for rel, data in files.items():
                got, mode, sha = stub.volume_files[f"/data/{rel}"]
                assert got == data and sha == hashlib.sha256(data).hexdigest()
            assert stub.volume_files["/data/dup.bin"][0] == files["sub/b.bin1"]
            assert stub.volume_files["/data/from_fileobj"][0] == b"hello world, this is a lot of text"
            assert len(store.blobs) == 2  # only the two files >= LARGE_FILE_LIMIT became blobs
            # dup.bin has the content of sub/b.bin1: one existence check + one data put for both paths
            n_files = len(files) + 2
            assert stub.mount_put_calls == 2 * (n_files - 1)
            # uploading again without force collides
            with pytest.raises(FileExistsError):
                async with volume.VolumeUploadContextManager("vo-1", client) as batch:
                    batch.put_file(root / "dup.bin", "/data/dup.bin")
            async with volume.VolumeUploadContextManager("vo-1", client, force=True) as batch:
                batch.put_file(root / "dup.bin", "/data/dup.bin")
            with pytest.raises(ValueError):  # like the reference, only a path that *normalises* to a directory ("/")
                batch.put_file(root / "dup.bin", "/")
            await blob_utils.ClientSessionRegistry.close_session()

    asyncio.run(run())


def test_batch_upload_v2_missing_block_loop(backend, monkeypatch, tmp_path):
    BS = 4096
    monkeypatch.setattr(blob_utils, "BLOCK_SIZE", BS)
    holder = {}

    async def run():
        async with running_blob_server() as (host, store):
            # extend the fake blob server with the block PUT route
            stub = FakeVolumeStub(host, store, BS)
            holder["stub"] = stub
            app = web.Application(client_max_size=1 << 26)
            app.add_routes([web.put("/block/{token}", _add_block_route(store, holder, BS))])
            runner = web.AppRunner(app)
            await runner.setup()
            site = web.TCPSite(runner, "127.0.0.1", 0)
            await site.start()
            stub.host = f"http://127.0.0.1:{site._server.sockets[0].getsockname()[1]}"
            try:
                client = types.SimpleNamespace(stub=stub)
                root, files = _tree(tmp_path, [0, 100, BS, BS + 1, 3 * BS + 4711, 40_000], 300)
                blank = (b"a" + bytes(BS - 1)) * 2 + b"cdef"  # the reference's blank-block fixture, scaled
                async with volume.VolumeUploadContextManager2("vo-2", client) as batch:
                    batch.put_directory(root, "/v")
                    batch.put_file(io.BytesIO(blank), "/v/blank", mode=0o600)
                for rel, data in files.items():
                    assert stub.volume_files[f"/v/{rel}"][0] == data
                assert stub.volume_files["/v/blank"][0] == blank and stub.volume_files["/v/blank"][1] == 0o600
                # trimmed blocks travelled trimmed: the blank blocks were sent as one byte
                assert stub.blocks[hashlib.sha256(b"a").hexdigest()] == b"a"
                # identical blocks (blank's two "a" blocks, ...) were PUT once each: one request per distinct content
                assert stub.block_puts == len(stub.blocks)
                # a second tree made of copies: every block is already known to the server after one round of PUTs
                copies = tmp_path / "copies"
                copies.mkdir()
                payload = synth_bytes(999, 2 * BS + 17)
                for k in range(5):
                    (copies / f"same{k}.bin").write_bytes(payload)
                before = stub.block_puts
                async with volume.VolumeUploadContextManager2("vo-2", client) as batch:
                    batch.put_directory(copies, "/copies")
                assert stub.block_puts - before == 3  # 3 distinct blocks, not 15
                assert all(stub.volume_files[f"/copies/same{k}.bin"][0] == payload for k in range(5))
                with pytest.raises(FileExistsError):
                    async with volume.VolumeUploadContextManager2("vo-2", client) as batch:
                        batch.put_file(io.BytesIO(b"x"), "/v/blank")
            finally:
                await runner.cleanup()
                await blob_utils.ClientSessionRegistry.close_session()

    asyncio.run(run())


def test_put_directory_walk_selects_what_rglob_is_file_selects(tmp_path):
    """put_directory's scandir walk == the reference's `rglob("*")` + `is_file()` + `relative_to` (volume.py:1279-1284):
    symlinks to files count, symlinked directories are not descended into, special files and directories are skipped."""
    import os
    from pathlib import Path, PurePosixPath

    root = tmp_path / "src"
    (root / "a" / "b").mkdir(parents=True)
    (root / "empty_dir").mkdir()
    for rel in ["top.txt", "a/x.bin", "a/b/deep.bin", "a/b/.hidden", "a/sp ace.txt"]:
        (root / rel).write_bytes(rel.encode())
    outside = tmp_path / "outside"
    outside.mkdir()
    (outside / "o.txt").write_bytes(b"o")
    os.symlink(outside / "o.txt", root / "link_to_file")
    os.symlink(outside, root / "link_to_dir")          # not descended into (rglob does not follow it either)
    os.symlink(root / "nope", root / "dangling")       # is_file() false
    os.mkfifo(root / "a" / "fifo")

    def reference_way(recursive):
        out = []
        for sub in (root.rglob("*") if recursive else root.glob("*")):
            if sub.is_file():
                out.append((str(sub), (PurePosixPath("/dst") / sub.relative_to(root)).as_posix()))
        return sorted(out)

    for recursive in (True, False):
        batch = volume._BatchBase("vo", client=None)
        batch.put_directory(root, "/dst", recursive=recursive)
        got = sorted((str(p), r) for p, r, _ in batch._paths)
        assert got == reference_way(recursive)
    batch = volume._BatchBase("vo", client=None)
    batch.put_directory(str(root), PurePosixPath("/dst/"), recursive=True)  # str source, trailing slash on the remote
    assert sorted(r for _, r, _ in batch._paths) == sorted(r for _, r in reference_way(True))


def test_resolve_picks_the_uploader_by_filesystem_version():
    """py/modal/volume.py:1156-1173: unspecified / v1 -> the volumefs1 uploader, v2 -> the volumefs2 uploader,
    anything else is an error; the reference's underscore names resolve to the same classes."""
    from modal_client_b200 import volume as v

    for version in (None, v.VOLUME_FS_VERSION_UNSPECIFIED, v.VOLUME_FS_VERSION_V1):
        m = v._AbstractVolumeUploadContextManager.resolve(version, "vo-1", client=None, force=True)
        assert type(m) is v._VolumeUploadContextManager and m._volume_id == "vo-1" and m._force
    m2 = v.AbstractVolumeUploadContextManager.resolve(v.VOLUME_FS_VERSION_V2, "vo-2", client=None)
    assert type(m2) is v._VolumeUploadContextManager2 and m2._put_concurrency == 128
    assert v._VolumeUploadContextManager2("vo-3", None, hash_concurrency=7, put_concurrency=3)._put_concurrency == 3
    with pytest.raises(RuntimeError, match="unsupported volume version"):
        v._AbstractVolumeUploadContextManager.resolve(99, "vo", client=None)
