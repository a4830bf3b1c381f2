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


_pool: list = []


def context_pool(n: int) -> list:
    """``n`` contexts on this process's device for callers that keep several host batches in flight (the map pump
    hashes window k+1 while window k's tail is still on the GPU): the current context plus ``n-1`` more on the same
    device, created on first use.  A stand-in installed by the tests is returned ``n`` times."""
    first = get_context()
    if not isinstance(first, _lib.Context):
        return [first] * n
    while len(_pool) < n - 1 or any(c.device != first.device or not getattr(c, "_h", None) for c in _pool):
        _pool[:] = [c for c in _pool if c.device == first.device and getattr(c, "_h", None)]
        if len(_pool) >= n - 1:
            break
        _pool.append(_lib.Context(first.device))
    return [first] + _pool[: n - 1]
