"""aiohttp payload that streams one segment of a file and knows that segment's MD5.

Counterpart of the reference's ``BytesIOSegmentPayload`` (py/modal/_utils/bytes_io_segment_payload.py:19-116).
The reference folds every chunk it sends into a hashlib.md5 on an executor thread -- i.e. it hashes every
uploaded byte a second time on the CPU.  Here the part MD5 normally arrives *precomputed* from the GPU
batch that hashed all parts at once (``blob_utils.perform_multipart_upload``), so sending is a pure
read+write loop; when no digest was supplied, the chunks are folded into a device-resident
``DigestStream`` instead (same update/hexdigest shape as the hashlib object the callers expect).
"""
from __future__ import annotations

import asyncio
from collections.abc import Callable
from contextlib import contextmanager
from typing import BinaryIO

from aiohttp import Payload
from aiohttp.abc import AbstractStreamWriter

from ._backend import get_context
from ._lib import MD5

DEFAULT_SEGMENT_CHUNK_SIZE = 2**24  # ~16 MiB reads


class _KnownMd5:
    """hashlib-shaped view of a digest that was already computed on the GPU."""

    def __init__(self, raw: bytes):
        self._raw = raw

    def update(self, _chunk) -> None:  # bytes are already accounted for
        return None

    def digest(self) -> bytes:
        return self._raw

    def hexdigest(self) -> str:
        return self._raw.hex()


class _GpuMd5:
    """hashlib-shaped incremental MD5 whose state lives on the device (b200h_stream_*)."""

    def __init__(self):
        self._s = get_context().stream(MD5)

    def update(self, chunk) -> None:
        self._s.update(chunk)

    def digest(self) -> bytes:
        return self._s.digests()[1]

    def hexdigest(self) -> str:
        return self.digest().hex()

    def close(self):
        self._s.close()


class BytesIOSegmentPayload(Payload):
    _value: BinaryIO

    def __init__(
        self,
        bytes_io: BinaryIO,  # one reader per payload: its position is not shared or locked
        segment_start: int,
        segment_length: int,
        chunk_size: int = DEFAULT_SEGMENT_CHUNK_SIZE,
        progress_report_cb: Callable | None = None,
        md5_digest: bytes | None = None,  # raw 16 bytes when the GPU batch already produced it
    ):
        super().__init__(bytes_io)
        self._size = segment_length
        self.initial_seek_pos = bytes_io.tell()
        self.segment_start = segment_start
        self.segment_length = segment_length
        self._value.seek(self.initial_seek_pos + segment_start)
        assert self.segment_length <= super().size
        self.chunk_size = chunk_size
        self.progress_report_cb = progress_report_cb or (lambda *_, **__: None)
        self._known_md5 = md5_digest
        self._md5_checksum = None
        self.reset_state()

    def decode(self, encoding: str = "utf-8", errors: str = "strict") -> str:
        self._value.seek(self.initial_seek_pos)
        return self._value.read().decode(encoding, errors)

    def reset_state(self):
        """Forget progress so that a retry re-sends (and, without a known digest, re-hashes) from scratch."""
        old = self._md5_checksum
        if isinstance(old, _GpuMd5):
            old.close()
        self._md5_checksum = _KnownMd5(self._known_md5) if self._known_md5 is not None else _GpuMd5()
        self.num_bytes_read = 0
        self._value.seek(self.initial_seek_pos)

    @contextmanager
    def reset_on_error(self, subtract_progress: bool = False):
        try:
            yield
        except Exception as exc:
            try:
                if subtract_progress:
                    self.progress_report_cb(advance=-self.num_bytes_read)
                else:
                    self.progress_report_cb(reset=True)
            except Exception as cb_exc:
                raise cb_exc from exc
            raise exc
        finally:
            self.reset_state()

    @property
    def size(self) -> int:
        return self.segment_length

    def md5_checksum(self):
        return self._md5_checksum

    def remaining_bytes(self) -> int:
        return self.segment_length - self.num_bytes_read

    async def write(self, writer: "AbstractStreamWriter"):
        await self.write_with_length(writer, None)

    async def write_with_length(self, writer: AbstractStreamWriter, content_length: int | None):
        loop = asyncio.get_event_loop()
        budget = self.segment_length if content_length is None else min(self.segment_length, content_length)
        while self.num_bytes_read < budget:
            self._value.seek(self.initial_seek_pos + self.segment_start + self.num_bytes_read)
            want = min(self.chunk_size, budget - self.num_bytes_read)
            chunk = await loop.run_in_executor(None, self._value.read, want)
            if not chunk:
                break
            if self._known_md5 is None:
                await loop.run_in_executor(None, self._md5_checksum.update, chunk)
            self.num_bytes_read += len(chunk)
            await writer.write(chunk)
            self.progress_report_cb(advance=len(chunk))
