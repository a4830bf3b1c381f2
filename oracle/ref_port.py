"""oracle/ref_port.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference path's *control flow*, built on ``hashlib`` (the library the
reference itself calls).  Each function cites the reference lines it follows
(paths relative to /root/reference).  This is what ``bench.py`` times as the CPU baseline
(``cpu_baseline.kind == "port"``) and what the parity tests compare the CUDA path against.
Pinned by tests/test_oracle.py against tests/golden/ (outputs of the unmodified reference).
"""
from __future__ import annotations

import base64
import hashlib
import io
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import BinaryIO, Iterable

READ_CHUNK = 65536  # py/modal/_utils/hash_utils.py:11
BLOCK = 8 * 1024 * 1024  # py/modal/_utils/blob_utils.py:63
BIG_FILE = 4 * 1024 * 1024  # blob_utils.py:43
NO_MD5_ABOVE = 1024**3  # blob_utils.py:54
INLINE_BELOW = 256 * 1024  # blob_utils.py:468
MD5_PLACEHOLDER = "baadbaad" * 4  # blob_utils.py:461


def feed(sinks, data) -> None:
    """hash_utils.py:14-29 -- bytes: one update; stream: 64 KiB reads from the current
    position to EOF, position restored afterwards; non-bytes chunks are rejected."""
    if isinstance(data, bytes):
        for s in sinks:
            s.update(data)
        return
    here = data.tell()
    for piece in iter(lambda: data.read(READ_CHUNK), b""):
        if not isinstance(piece, bytes):
            raise ValueError(f"Only accepts bytes or byte buffer objects, not {type(piece)} buffers")
        for s in sinks:
            s.update(piece)
    data.seek(here)


@dataclass
class Hashes:
    """hash_utils.py:56-65."""

    md5_base64: str
    sha256_base64: str

    def md5_hex(self) -> str:
        return base64.b64decode(self.md5_base64).hex()

    def sha256_hex(self) -> str:
        return base64.b64decode(self.sha256_base64).hex()


def upload_hashes(data, sha256_hex: str | None = None, md5_hex: str | None = None) -> Hashes:
    """hash_utils.py:68-101 -- only the digests not supplied are computed, in one pass."""
    sha = None if sha256_hex else hashlib.sha256()
    md = None if md5_hex else hashlib.md5()
    sinks = [h for h in (sha, md) if h is not None]
    if sinks:
        feed(sinks, data)
    sha_raw = bytes.fromhex(sha256_hex) if sha256_hex else sha.digest()
    md_raw = bytes.fromhex(md5_hex) if md5_hex else md.digest()
    return Hashes(base64.b64encode(md_raw).decode("ascii"), base64.b64encode(sha_raw).decode("ascii"))


def sha256_hex(data) -> str:  # hash_utils.py:32-37
    h = hashlib.sha256()
    feed([h], data)
    return h.hexdigest()


def sha256_base64(data) -> str:  # hash_utils.py:40-45
    h = hashlib.sha256()
    feed([h], data)
    return base64.b64encode(h.digest()).decode("ascii")


def md5_base64(data) -> str:  # hash_utils.py:48-53
    h = hashlib.md5()
    feed([h], data)
    return base64.b64encode(h.digest()).decode("ascii")


def file_spec_fields(fp: BinaryIO) -> dict:
    """blob_utils.py:446-487 (the hashing/size-class part of ``_get_file_upload_spec``)."""
    fp.seek(0, os.SEEK_END)
    size = fp.tell()
    fp.seek(0)
    content = None
    if size >= BIG_FILE:
        h = upload_hashes(fp, md5_hex=MD5_PLACEHOLDER if size > NO_MD5_ABOVE else None)
        use_blob = True
    else:
        use_blob = False
        if size < INLINE_BELOW:
            content = fp.read()
            h = upload_hashes(content)
        else:
            h = upload_hashes(fp)
    return dict(use_blob=use_blob, sha256_hex=h.sha256_hex(), md5_hex=h.md5_hex(), size=size, content=content)


def block_end(buf: bytes, start: int, end: int) -> int:
    """blob_utils.py:667-705 -- absolute index just past the last non-zero byte in [start, end);
    ``start`` if the range is empty or all zero."""
    window = buf[start:end]
    kept = window.rstrip(b"\0")
    return start + len(kept)


def gather_blocks(buf: bytes, block_size: int = BLOCK) -> list[tuple[int, int, bytes]]:
    """blob_utils.py:622-664 -- [(start, trimmed_end, sha256 raw)] per ceil(size/block) block."""
    out = []
    for start in range(0, len(buf), block_size):
        end = block_end(buf, start, min(len(buf), start + block_size))
        out.append((start, end, hashlib.sha256(buf[start:end]).digest()))
    return out


def multipart_etag(buf, part_len: int) -> tuple[list[bytes], str]:
    """blob_utils.py:194-219 + bytes_io_segment_payload.py:58,102 -- per-part MD5 and
    ``md5(concat raw part digests).hexdigest() + "-<n>"``."""
    view = memoryview(buf)
    parts = [hashlib.md5(view[o : o + part_len]).digest() for o in range(0, len(view), part_len)]
    return parts, hashlib.md5(b"".join(parts)).hexdigest() + f"-{len(parts)}"


# ------------------------------------------------------------------ CPU baseline drivers


def first_occurrence(keys) -> tuple[list[int], int]:
    """In-batch dedupe as py/modal/mount.py:498,518-534 does it: walk the files in order, a content whose digest
    is already in `accounted_hashes` is skipped, otherwise it is added.  Returned per row: the index of the row
    that first carried the same digest (== own index for a first occurrence), and the number of distinct digests."""
    accounted: dict[bytes, int] = {}
    first = []
    for i, k in enumerate(keys):
        k = bytes(k)
        if k in accounted:  # mount.py:518
            first.append(accounted[k])
        else:
            accounted[k] = i  # mount.py:534
            first.append(i)
    return first, len(accounted)


def default_workers() -> int:
    """``ThreadPoolExecutor()`` default the reference relies on (volume.py:1211, mount.py:469)."""
    return min(32, (os.cpu_count() or 1) + 4)


def hash_payloads_serial(payloads: Iterable[bytes]) -> list[Hashes]:
    """What the map pump really does: ``get_upload_hashes(bytes)`` on the loop thread,
    one payload after another (blob_utils.py:345, parallel_map.py:139)."""
    return [upload_hashes(p) for p in payloads]


def hash_payloads_pool(payloads: list, workers: int) -> list[Hashes]:
    """All-cores variant (hashlib releases the GIL above 2 KiB): the strongest CPU arm."""
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(upload_hashes, payloads))


def hash_streams_pool(blobs: list[bytes], workers: int) -> list[Hashes]:
    """v1 file path: ``get_upload_hashes(BinaryIO)`` in a thread pool (volume.py:1209-1216)."""
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(lambda b: upload_hashes(io.BytesIO(b)), blobs))
