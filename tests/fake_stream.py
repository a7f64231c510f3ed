"""Byte streams for the fake NVDEC library (tests/fake_nvcuvid): Annex-B framing with H.264-style
emulation prevention around a trivial payload (sequence header / tight NV12 pictures)."""
from __future__ import annotations

import numpy as np

START4 = np.array([0, 0, 0, 1], np.uint8)
START3 = np.array([0, 0, 1], np.uint8)


def escape(rbsp: np.ndarray) -> np.ndarray:
    """Insert 0x03 after every 00 00 that is followed by a byte <= 3 (encoder-side rule)."""
    b = np.ascontiguousarray(rbsp, dtype=np.uint8)
    if b.size < 3:
        return b
    cand = np.flatnonzero((b[:-2] == 0) & (b[1:-1] == 0) & (b[2:] <= 3)) + 2
    keep, last = [], -10
    for p in cand.tolist():
        if p - last >= 2:          # the zero run restarts at the byte we inserted before
            keep.append(p)
            last = p
    return np.insert(b, keep, 3) if keep else b


def nal(nal_type: int, rbsp: np.ndarray, long_start: bool = True) -> np.ndarray:
    return np.concatenate([START4 if long_start else START3, np.array([nal_type], np.uint8), escape(rbsp),
                           np.array([0x80], np.uint8)])


def sequence_header(w: int, h: int) -> np.ndarray:
    return nal(0x67, np.array([w, h], dtype="<u4").view(np.uint8))


def picture(tight_nv12: np.ndarray, long_start: bool = True) -> np.ndarray:
    return nal(0x65, tight_nv12, long_start)


def split_nals(stream: np.ndarray):
    """The NAL units of a stream, each with its start code (what test_nv_dec.cpp's find_nalu yields)."""
    b = stream
    hits = np.flatnonzero((b[:-2] == 0) & (b[1:-1] == 0) & (b[2:] == 1))
    starts = [int(p - 1) if p > 0 and b[p - 1] == 0 else int(p) for p in hits]
    return [b[s:e] for s, e in zip(starts, starts[1:] + [b.size])]
