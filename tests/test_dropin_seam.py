"""The drop-in seam from the reference's side: the reference's UNMODIFIED blob_utils.py (spec builders, volumefs2 block
gatherer) is executed with `modal._utils.hash_utils` swapped for modal_client_b200.hash_utils, and must reproduce
the golden outputs its own hashlib-backed self produced (tests/golden/file_specs.json, blocks.json).  Needs a copy of
the reference (/root/reference here, baseline/_ref on the GPU box)."""
import asyncio
import io
import json
import os
from pathlib import PurePosixPath

import pytest

from modal_client_b200 import hash_utils as our_hash_utils
from modal_client_b200.synth import materialize
from oracle import ref_shim

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    return request.getfixturevalue("fake_backend" if request.param == "fake" else "gpu_backend")


@pytest.fixture
def ref_blob_on_ours(backend):
    if not ref_shim.available():
        pytest.skip("no copy of the reference here")
    try:
        mod = ref_shim.load_blob_utils_on(our_hash_utils)
    except Exception as exc:  # an environment that cannot load the reference at all is not a parity failure
        pytest.skip(f"reference blob_utils could not be loaded: {exc!r}")
    assert mod.get_upload_hashes is our_hash_utils.get_upload_hashes  # the reference code now calls OUR hashing
    return mod


def test_reference_spec_builder_on_our_hashes_reproduces_its_golden(ref_blob_on_ours, monkeypatch):
    doc = json.load(open(os.path.join(GOLDEN, "file_specs.json")))
    for c in doc["cases"]:
        with monkeypatch.context() as mp:
            for name, value in c["patch"].items():
                mp.setattr(ref_blob_on_ours, name, value)
            data = materialize(c["input"])
            spec = ref_blob_on_ours.get_file_upload_spec_from_fileobj(
                io.BytesIO(data), PurePosixPath(c["mount_filename"]), 0o100644 if not c["patch"] else 0o755)
            got = (spec.use_blob, spec.sha256_hex, spec.md5_hex, spec.mode, spec.size, spec.mount_filename, spec.content is not None)
            want = (c["use_blob"], c["sha256_hex"], c["md5_hex"], c["mode"], c["size"], c["mount_filename"], c["has_content"])
            assert got == want


def test_reference_block_gatherer_on_our_hashes_reproduces_its_golden(ref_blob_on_ours, monkeypatch, backend):
    # _hash_range_sha256 uses hashlib directly in the reference (blob_utils.py:648): the seam for volumefs2 is that
    # function, so it is swapped for ours the way INTEGRATION.md section 1 describes; the block planning, the
    # to_thread fan-out and FileUploadSpec2 assembly stay the reference's code
    from modal_client_b200 import blob_utils as ours

    doc = json.load(open(os.path.join(GOLDEN, "blocks.json")))
    for c in doc["spec2"]:
        data = materialize(c["input"])
        if len(data) > 20_000_000 and backend.device < 0:
            continue  # the oracle-backed stand-in is slow; the GPU variant runs every case
        with monkeypatch.context() as mp:
            for name, value in c["patch"].items():
                mp.setattr(ref_blob_on_ours, name, value)
                mp.setattr(ours, name, value)
            mp.setattr(ref_blob_on_ours, "_hash_range_sha256", ours._hash_range_sha256)
            spec = asyncio.run(ref_blob_on_ours.FileUploadSpec2.from_fileobj(
                io.BytesIO(data), PurePosixPath(c["path"]), asyncio.Semaphore(2), 0o644))
            assert [[b.start, b.end, b.contents_sha256.hex()] for b in spec.blocks] == c["blocks"]
            assert (spec.size, spec.mode, spec.path) == (c["size"], c["mode"], c["path"])
