"""Drop-in for the reference's ``modal/_utils/hash_utils.py`` with the arithmetic on the B200.

Same names, signatures, return types and stream semantics as the reference
(py/modal/_utils/hash_utils.py:14-101); the hashlib objects are replaced by libb200hash calls:

* ``bytes`` input  -> one message through ``b200h_hash_batch_host`` (both digests in one read);
* stream input     -> ``b200h_stream_*``: read from the current position to EOF in ``HASH_CHUNK_SIZE``
  pieces, restore the position afterwards (reference :20,:29), reject non-bytes chunks (:23-24).

``get_upload_hashes_many`` is the batched entry point the accelerated callers use (map input pump,
Volume.batch_upload): N payloads, one GPU batch, N ``UploadHashes``.
"""
from __future__ import annotations

import base64
import dataclasses
import time
from collections.abc import Callable, Sequence
from typing import BinaryIO

from ._backend import get_context
from ._lib import MD5, SHA256
from ._logging import logger

HASH_CHUNK_SIZE = 65536  # module global on purpose: the reference's tests patch it by name


def _update(hashers: Sequence[Callable[[bytes], None]], data: bytes | BinaryIO) -> None:
    """Feed ``data`` to every updater in ``hashers`` (reference :14-29)."""
    if isinstance(data, bytes):
        for push in hashers:
            push(data)
        return
    assert not isinstance(data, (bytearray, memoryview))
    start = data.tell()
    while True:
        chunk = data.read(HASH_CHUNK_SIZE)
        if not isinstance(chunk, bytes):
            raise ValueError(f"Only accepts bytes or byte buffer objects, not {type(chunk)} buffers")
        if len(chunk) == 0:
            break
        for push in hashers:
            push(chunk)
    data.seek(start)


def _digests(data: bytes | BinaryIO, flags: int) -> tuple[bytes | None, bytes | None]:
    """(sha256 raw | None, md5 raw | None) of one message, computed on the GPU."""
    ctx = get_context()
    if isinstance(data, bytes):
        sha, md5, _ = ctx.hash_buffers([data], flags)
        return (sha[0].tobytes() if sha is not None else None, md5[0].tobytes() if md5 is not None else None)
    stream = ctx.stream(flags)
    try:
        _update([stream.update], data)
        return stream.digests()
    finally:
        stream.close()


def get_sha256_hex(data: bytes | BinaryIO) -> str:
    t0 = time.monotonic()
    sha, _ = _digests(data, SHA256)
    logger.debug("get_sha256_hex took %.3fs", time.monotonic() - t0)
    return sha.hex()


def get_sha256_base64(data: bytes | BinaryIO) -> str:
    t0 = time.monotonic()
    sha, _ = _digests(data, SHA256)
    logger.debug("get_sha256_base64 took %.3fs", time.monotonic() - t0)
    return base64.b64encode(sha).decode("ascii")


def get_md5_base64(data: bytes | BinaryIO) -> str:
    t0 = time.monotonic()
    _, md5 = _digests(data, MD5)
    logger.debug("get_md5_base64 took %.3fs", time.monotonic() - t0)
    return base64.b64encode(md5).decode("utf-8")


@dataclasses.dataclass
class UploadHashes:
    md5_base64: str
    sha256_base64: str

    def md5_hex(self) -> str:
        return base64.b64decode(self.md5_base64).hex()

    def sha256_hex(self) -> str:
        return base64.b64decode(self.sha256_base64).hex()


def _b64(raw: bytes) -> str:
    return base64.b64encode(raw).decode("ascii")


def get_upload_hashes(
    data: bytes | BinaryIO, sha256_hex: str | None = None, md5_hex: str | None = None
) -> UploadHashes:
    """Both upload digests in a single pass over ``data``; a digest supplied by the caller is passed
    through (hex -> base64) and NOT recomputed (reference :74-93)."""
    t0 = time.monotonic()
    flags = (0 if sha256_hex else SHA256) | (0 if md5_hex else MD5)
    sha_raw = md5_raw = None
    if flags:
        sha_raw, md5_raw = _digests(data, flags)
    out = UploadHashes(
        md5_base64=_b64(bytes.fromhex(md5_hex) if md5_hex else md5_raw),
        sha256_base64=_b64(bytes.fromhex(sha256_hex) if sha256_hex else sha_raw),
    )
    logger.debug("get_upload_hashes took %.3fs (flags=%d)", time.monotonic() - t0, flags)
    return out


class _UploadHashesView(Sequence):
    """Read-only sequence of ``UploadHashes`` over a digest table; rows are formatted on access, so a batch of
    10^6 payloads does not pay 10^6 base64 round trips up front."""

    def __init__(self, sha, md5, n: int):
        self._sha, self._md5, self._n = sha, md5, n

    def __len__(self) -> int:
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return UploadHashes(md5_base64=_b64(self._md5[i].tobytes()) if self._md5 is not None else "",
                            sha256_base64=_b64(self._sha[i].tobytes()))


def get_upload_hashes_many(payloads: Sequence[bytes], *, want_md5: bool = True) -> Sequence[UploadHashes]:
    """N in-memory payloads -> N UploadHashes in ONE GPU batch.  This replaces the serial
    ``get_upload_hashes(payload)`` loop the map pump runs on its event-loop thread
    (py/modal/_utils/blob_utils.py:345 under parallel_map.py:139) and Go's per-call md5.Sum/sha256.Sum256
    (go/blob.go:51-52).  ``want_md5=False`` leaves md5_base64 empty (callers that hold a placeholder)."""
    if not payloads:
        return []
    flags = SHA256 | (MD5 if want_md5 else 0)
    sha, md5, _ = get_context().hash_buffers(payloads, flags)
    return _UploadHashesView(sha, md5, len(payloads))
