"""Where the host layer gets its GPU context from.

Everything above the C ABI (hash_utils, blob_utils, batch planners) calls ``get_context()``; by default
that is the process-wide ``_lib.default_context()`` (device = LOCAL_RANK).  ``set_context`` lets an
application pin a specific context.  There is no CPU implementation behind this seam: the unit tests
substitute an oracle-backed stand-in here to exercise host logic on machines without a GPU, which is
test scaffolding, not a fallback -- production code never does that.
"""
from __future__ import annotations

from . import _lib

_override = None


def set_context(ctx) -> None:
    global _override
    _override = ctx


def get_context():
    return _override if _override is not None else _lib.default_context()
