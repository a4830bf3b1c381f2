"""oracle/ref_shim.py -- TEST INFRASTRUCTURE ONLY; works only where /root/reference exists.

Loads the reference's UNMODIFIED hash_utils.py / blob_utils.py / bytes_io_segment_payload.py
straight from /root/reference/py/modal/_utils without grpclib / synchronicity / generated
protos, by pre-seeding sys.modules so that modal/__init__.py never executes (SURVEY.md
Appendix A).  Used by oracle/gen_golden.py to produce tests/golden/*.json and by
tests/test_oracle.py (skipped when /root/reference is absent, e.g. on the GPU box).
"""
from __future__ import annotations

import asyncio
import contextlib
import importlib
import logging
import os
import sys
import types

REF_ROOT = os.environ.get("B200H_REFERENCE_ROOT", "/root/reference")
_PKG = os.path.join(REF_ROOT, "py", "modal")


def available() -> bool:
    return os.path.isfile(os.path.join(_PKG, "_utils", "hash_utils.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """-> (ref_hash_utils, ref_blob_utils, ref_segment_payload) modules of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    if "modal" in sys.modules and not getattr(sys.modules["modal"], "_b200h_shim", False):
        raise RuntimeError("a real `modal` package is already imported; shim needs a clean process")

    class ExecutionError(Exception):
        pass

    def retry(direct_fn=None, **_kw):
        passthrough = lambda fn: fn  # noqa: E731
        return passthrough(direct_fn) if direct_fn else passthrough

    @contextlib.asynccontextmanager
    async def asyncnullcontext(*_a, **_k):
        yield

    class TaskContext:
        @staticmethod
        async def gather(*coros):
            return await asyncio.gather(*coros)

    shell = _stub("modal", _b200h_shim=True)
    shell.__path__ = []
    _stub("modal.config", logger=logging.getLogger("modal-ref"), config={})
    _stub("modal.exception", ExecutionError=ExecutionError)
    utils = _stub("modal._utils")
    utils.__path__ = [os.path.join(_PKG, "_utils")]
    _stub(
        "modal._utils.async_utils",
        retry=retry,
        asyncnullcontext=asyncnullcontext,
        TaskContext=TaskContext,
        on_shutdown=lambda coro: None,
    )
    _stub("modal_proto").__path__ = []
    _stub("modal_proto.api_pb2")
    _stub("modal_proto.modal_api_grpc", ModalClientModal=object)
    h = importlib.import_module("modal._utils.hash_utils")
    b = importlib.import_module("modal._utils.blob_utils")
    s = importlib.import_module("modal._utils.bytes_io_segment_payload")
    _loaded = (h, b, s)
    return _loaded
