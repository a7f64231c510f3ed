"""Synthetic decoded surfaces (SURVEY.md 8d "Synthetic inputs").

Counter-based and numpy-version independent: byte i of frame f of stream s is byte (i & 7) of
splitmix64(key(s, f) + (i >> 3)), so the SHA-256 known-answer vectors under tests/golden/ are
reproducible anywhere.  Padding bytes [w, pitch) of every row are 0xCD; callers pre-fill outputs
with 0xA5 so that "bytes the reference leaves untouched" can be checked.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20240517
PAD_BYTE = 0xCD
OUT_FILL = 0xA5

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def frame_key(stream: int, frame: int, seed: int = BASE_SEED) -> int:
    return ((seed << 32) ^ (0x9E3779B9 * (stream * 100003 + frame) + 1)) & 0xFFFFFFFFFFFFFFFF


def random_bytes(n: int, key: int) -> np.ndarray:
    """n uniform bytes from the counter-based generator."""
    words = (n + 7) // 8
    with np.errstate(over="ignore"):
        ctr = np.arange(words, dtype=np.uint64) + np.uint64(key)
    return _splitmix64(ctr).view(np.uint8)[:n].copy()


def nv12_surface(w: int, h: int, pitch: int, stream: int = 0, frame: int = 0,
                 kind: str = "random", rows: int | None = None) -> np.ndarray:
    """A pitched NV12 surface of `rows` rows (default: the reference's pitch*h*3/2 bytes,
    nv_dec/nv_dec.cpp:453).  Y rows [0,h), UV rows from row h on."""
    nbytes = pitch * h * 3 // 2 if rows is None else pitch * rows
    nrows = -(-nbytes // pitch)
    s = np.full((nrows, pitch), PAD_BYTE, dtype=np.uint8)
    if kind == "random":
        s[:, :w] = random_bytes(nrows * w, frame_key(stream, frame)).reshape(nrows, w)
    elif kind == "gradient":            # catches U/V swaps and row/column slips
        x = np.arange(w, dtype=np.int64)[None, :]
        y = np.arange(nrows, dtype=np.int64)[:, None]
        s[:h, :w] = ((x + y[:h]) & 255).astype(np.uint8)
        cx = x >> 1
        uv = np.where((x & 1) == 0, cx & 255, 255 - (cx & 255)) + 0 * y[h:]
        s[h:, :w] = ((uv + (y[h:] - h)) & 255).astype(np.uint8)
    elif kind.startswith("const"):      # RGB clamp edges: const0, const16, const235 ...
        s[:, :w] = int(kind[5:])
    else:
        raise ValueError(kind)
    return s.reshape(-1)[:nbytes].copy()


def i420_frame(w: int, h: int, stream: int = 0, frame: int = 0) -> np.ndarray:
    """A tight planar frame of the reference's w*h*3/2 bytes."""
    return random_bytes(w * h * 3 // 2, frame_key(stream, frame) ^ 0x5A5A5A5A)
