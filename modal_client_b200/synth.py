"""Counter-based synthetic bytes shared by tests, fixtures and bench.

``synth_bytes(seed, n, start)`` is a pure function of (seed, absolute byte position), so the CPU
(numpy, here) and the GPU (``b200h_fill_synth`` in csrc/b200hash_kernels.cu) produce the same
stream without shipping data: 64-bit word ``i`` of stream ``seed`` is
``mix64(seed * 0xD1342543DE82EF95 + i)`` (splitmix64 finaliser), stored little-endian.

``materialize(recipe)`` expands the small JSON recipes stored in tests/golden/*.json.
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1
_SEED_MUL = 0xD1342543DE82EF95


def _mix64(z: np.ndarray) -> np.ndarray:
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_array(seed: int, nbytes: int, start: int = 0) -> np.ndarray:
    """uint8[nbytes]: bytes [start, start+nbytes) of stream ``seed``."""
    if nbytes <= 0:
        return np.zeros(0, np.uint8)
    w0 = start // 8
    w1 = (start + nbytes + 7) // 8
    base = np.uint64((seed * _SEED_MUL) & _M64)
    with np.errstate(over="ignore"):
        words = _mix64(np.arange(w0, w1, dtype=np.uint64) + base)
    raw = words.view(np.uint8)  # little-endian host
    lo = start - 8 * w0
    return raw[lo : lo + nbytes]


def synth_bytes(seed: int, nbytes: int, start: int = 0) -> bytes:
    return synth_array(seed, nbytes, start).tobytes()


def materialize(recipe: dict) -> bytes:
    """Expand a fixture recipe: synth | repeat | literal | concat | zeros."""
    kind = recipe["kind"]
    if kind == "synth":
        return synth_bytes(recipe["seed"], recipe["size"], recipe.get("start", 0))
    if kind == "repeat":
        return bytes.fromhex(recipe["unit"]) * recipe["count"]
    if kind == "zeros":
        return bytes(recipe["size"])
    if kind == "literal":
        return bytes.fromhex(recipe["hex"])
    if kind == "concat":
        return b"".join(materialize(p) for p in recipe["parts"])
    raise ValueError(f"unknown recipe kind {kind!r}")
