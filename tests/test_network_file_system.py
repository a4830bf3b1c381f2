"""NetworkFileSystem uploads against an in-process fake following the reference's servicer
(py/test/conftest.py:2684-2697 SharedVolumePutFile / SharedVolumeGetFile) and mirroring the cases of
py/test/network_file_system_test.py:46-118 (single file, directory, big file through the blob path, read, write)."""
import asyncio
import hashlib
import io
import types

import pytest

from modal_client_b200 import blob_utils, network_file_system as nfs_mod
from modal_client_b200.synth import synth_bytes
from tests.blob_server import FakeBlobStub, running_blob_server


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    return request.getfixturevalue("fake_backend" if request.param == "fake" else "gpu_backend")


class NotFoundError(Exception):
    pass


class FakeNfsStub(FakeBlobStub):
    def __init__(self, host):
        super().__init__(host, multipart_threshold=10_000_000)
        self.nfs_files = {}
        self.puts = 0

    async def SharedVolumePutFile(self, req):
        self.puts += 1
        self.nfs_files.setdefault(req.shared_volume_id, {})[req.path] = req
        return types.SimpleNamespace(exists=True)

    async def SharedVolumeGetFile(self, req):
        put = self.nfs_files.get(req.shared_volume_id, {}).get(req.path)
        if not put:
            raise NotFoundError(f"No such file: {req.path}")
        if put.data_blob_id:
            return types.SimpleNamespace(WhichOneof=lambda _n: "data_blob_id", data_blob_id=put.data_blob_id)
        return types.SimpleNamespace(WhichOneof=lambda _n: "data", data=put.data)


def test_nfs_single_file_dir_big_file_read_write(backend, monkeypatch, tmp_path):
    async def run():
        async with running_blob_server() as (host, store):
            stub = FakeNfsStub(host)
            nfs = nfs_mod.NetworkFileSystemUploader("sv-1", types.SimpleNamespace(stub=stub))
            # single file, default and explicit destination (network_file_system_test.py:46-60)
            f = tmp_path / "some_file"
            f.write_text("hello world")
            assert await nfs.add_local_file(f) == 11
            await nfs.add_local_file(f.as_posix(), remote_path="/foo/other_destination")
            assert stub.nfs_files["sv-1"].keys() == {"/some_file", "/foo/other_destination"}
            assert stub.nfs_files["sv-1"]["/some_file"].data == b"hello world"
            # directory (…:63-82), plus content duplicated under two names and an empty file
            d = tmp_path / "some_dir"
            (d / "subdir").mkdir(parents=True)
            (d / "smol").write_text("###")
            (d / "subdir" / "other").write_text("####")
            (d / "subdir" / "copy").write_text("####")
            (d / "empty").write_bytes(b"")
            written = await nfs.add_local_dir(d)
            assert written == 3 + 4 + 4
            got = stub.nfs_files["sv-1"]
            assert {"/some_dir/smol", "/some_dir/subdir/other", "/some_dir/subdir/copy", "/some_dir/empty"} <= got.keys()
            assert got["/some_dir/smol"].data == b"###" and got["/some_dir/subdir/other"].data == b"####"
            assert got["/some_dir/empty"].data == b""
            # big file goes through the blob path with the GPU digests (…:85-100)
            monkeypatch.setattr(nfs_mod, "LARGE_FILE_LIMIT", 10)
            big = tmp_path / "bigfile"
            big.write_text("hello world, this is a lot of text")
            await nfs.add_local_file(big)
            req = stub.nfs_files["sv-1"]["/bigfile"]
            assert req.data == b"" and req.data_blob_id == "bl-1"
            assert req.sha256_hex == hashlib.sha256(big.read_bytes()).hexdigest()
            assert store.blobs["bl-1"] == b"hello world, this is a lot of text"
            # a directory whose files are all above the limit: one GPU batch, every file a blob with the right digests
            tree = tmp_path / "tree"
            (tree / "x").mkdir(parents=True)
            payloads = {"a.bin": synth_bytes(1, 5000), "x/b.bin": synth_bytes(2, 70_001), "x/c.bin": synth_bytes(3, 11)}
            for rel, data in payloads.items():
                (tree / rel).write_bytes(data)
            n_create = len(stub.requests)
            await nfs.add_local_dir(tree, "/t")
            for rel, data in payloads.items():
                r = stub.nfs_files["sv-1"][f"/t/{rel}"]
                assert store.blobs[r.data_blob_id] == data and r.sha256_hex == hashlib.sha256(data).hexdigest()
            creates = stub.requests[n_create:]
            assert sorted(c.content_length for c in creates) == sorted(len(v) for v in payloads.values())
            import base64
            assert {c.content_md5 for c in creates} == {base64.b64encode(hashlib.md5(v).digest()).decode() for v in payloads.values()}
            # read back: inline and blob-backed (…:103-118); missing file -> FileNotFoundError
            assert b"".join([c async for c in nfs.read_file("/some_file")]) == b"hello world"
            assert b"".join([c async for c in nfs.read_file("/t/x/b.bin")]) == payloads["x/b.bin"]
            with pytest.raises(FileNotFoundError):
                async for _ in nfs.read_file("idontexist.txt"):
                    pass
            # write_file from a file object, twice (overwrite through the provider)
            for _ in range(2):
                assert await nfs.write_file("remote_path.txt", io.BytesIO(b"0123456789" * 3)) == 30
            assert store.blobs[stub.nfs_files["sv-1"]["remote_path.txt"].data_blob_id] == b"0123456789" * 3
            await blob_utils.ClientSessionRegistry.close_session()

    asyncio.run(run())


def test_nfs_put_times_out(backend, monkeypatch, tmp_path):
    async def run():
        async with running_blob_server() as (host, _store):
            stub = FakeNfsStub(host)

            async def never(req):
                await asyncio.sleep(0.01)
                return types.SimpleNamespace(exists=False)

            stub.SharedVolumePutFile = never
            monkeypatch.setattr(nfs_mod, "NETWORK_FILE_SYSTEM_PUT_FILE_CLIENT_TIMEOUT", 0.05)
            nfs = nfs_mod.NetworkFileSystemUploader("sv-2", types.SimpleNamespace(stub=stub))
            with pytest.raises(TimeoutError, match="timed out"):
                await nfs.write_file("/x", io.BytesIO(b"abc"))

    asyncio.run(run())
