"""In-process fakes of the two remote ends of the blob path, following the contracts of the reference's own
test doubles (py/test/conftest.py:1337-1370 BlobCreate, :3356-3413 blob server): PUT returns ETag = md5 of the
received body; multipart completion returns md5(concat(md5(part)))-N; b"FAILURE" bodies get HTTP 500."""
from __future__ import annotations

import contextlib
import hashlib
import types
from collections import defaultdict

from aiohttp import web


class BlobStore:
    def __init__(self):
        self.blobs: dict[str, bytes] = {}
        self.parts: dict[str, dict[int, bytes]] = defaultdict(dict)
        self.puts = 0
        self.corrupt_etag = False
        self.content_md5_headers: list[str | None] = []

    async def upload(self, request: web.Request):
        blob_id = request.query["blob_id"]
        body = await request.read()
        self.puts += 1
        self.content_md5_headers.append(request.headers.get("Content-MD5"))
        if body == b"FAILURE":
            return web.Response(status=500)
        md5 = hashlib.md5(body).hexdigest()
        if self.corrupt_etag:
            md5 = "0" * 32
        if "part_number" in request.query:
            self.parts[blob_id][int(request.query["part_number"])] = body
        else:
            self.blobs[blob_id] = body
        return web.Response(text="ok", headers={"ETag": f'"{md5}"'})

    async def complete(self, request: web.Request):
        blob_id = request.query["blob_id"]
        parts = self.parts[blob_id]
        ordered = [parts[i] for i in range(min(parts), max(parts) + 1)]
        cat = hashlib.md5(b"".join(hashlib.md5(p).digest() for p in ordered)).hexdigest()
        self.blobs[blob_id] = b"".join(ordered)
        return web.Response(text=f'<etag>"{cat}-{len(ordered)}"</etag>')


    async def download(self, request: web.Request):  # py/test/conftest.py:3393-3397
        blob_id = request.query["blob_id"]
        if blob_id == "bl-failure":
            return web.Response(status=500)
        return web.Response(body=self.blobs[blob_id])


@contextlib.asynccontextmanager
async def running_blob_server():
    store = BlobStore()
    app = web.Application(client_max_size=1 << 30)
    app.add_routes([web.put("/upload", store.upload), web.post("/complete_multipart", store.complete),
                    web.get("/download", store.download)])
    runner = web.AppRunner(app)
    await runner.setup()
    site = web.TCPSite(runner, "127.0.0.1", 0)
    await site.start()
    port = site._server.sockets[0].getsockname()[1]
    try:
        yield f"http://127.0.0.1:{port}", store
    finally:
        await runner.cleanup()


class FakeBlobStub:
    """Duck-typed ``stub`` with BlobCreate: single URL below ``multipart_threshold``, multipart above."""

    def __init__(self, host: str, multipart_threshold: int = 10_000_000, providers: int = 2):
        self.host, self.threshold, self.providers = host, multipart_threshold, providers
        self.requests = []
        self.n = 0

    async def BlobGet(self, req):  # py/test/conftest.py:1382-1385
        return types.SimpleNamespace(download_url=f"{self.host}/download?blob_id={req.blob_id}")

    async def BlobCreate(self, req):
        self.requests.append(req)
        self.n += 1
        blob_id = f"bl-{self.n}"
        ids = [blob_id] * self.providers
        if req.content_length > self.threshold:
            nparts = -(-req.content_length // self.threshold)
            item = types.SimpleNamespace(
                part_length=self.threshold,
                upload_urls=[f"{self.host}/upload?blob_id={blob_id}&part_number={i + 1}" for i in range(nparts)],
                completion_url=f"{self.host}/complete_multipart?blob_id={blob_id}",
            )
            return types.SimpleNamespace(
                WhichOneof=lambda _n: "multiparts", blob_ids=ids,
                multiparts=types.SimpleNamespace(items=[item] * self.providers))
        url = f"{self.host}/upload?blob_id={blob_id}"
        return types.SimpleNamespace(
            WhichOneof=lambda _n: "upload_urls", blob_ids=ids,
            upload_urls=types.SimpleNamespace(items=[url] * self.providers))
