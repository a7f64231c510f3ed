"""Annex-B NAL splitter (include/jmc_annexb.h) against the reference's own find_nalu_prefix / find_nalu
(test_nv_dec/test_nv_dec.cpp:30-86, compiled unmodified into oracle/_ref)."""
import ctypes as C

import numpy as np
import pytest

import jmcodec_b200 as J
import oracle

_u8p = C.POINTER(C.c_uint8)


def ours_prefix(L, buf, size):
    n = C.c_int(-9)
    off = L.jmc_annexb_find_prefix(buf.ctypes.data, size, C.byref(n))
    return off, n.value


def ours_nalu(L, buf, size):
    n = C.c_int(-9)
    p = L.jmc_annexb_find_nalu(buf.ctypes.data, size, C.byref(n))
    return (p - buf.ctypes.data if p else -1), n.value


def test_known_streams():
    L = J.load()
    b = np.array([9, 0, 0, 1, 0x67, 1, 2, 0, 0, 0, 1, 0x65, 7, 7, 0, 0, 1, 0x41, 0xFF], np.uint8)
    assert ours_prefix(L, b, b.size) == (1, 3)
    assert ours_nalu(L, b, b.size) == (1, 6)                 # 00 00 01 67 01 02 | next start code at the 4-byte form
    assert ours_nalu(L, b[7:].copy(), b.size - 7) == (0, 7)   # 00 00 00 01 65 07 07
    tail = b[14:].copy()
    assert ours_nalu(L, tail, tail.size) == (-1, 0)           # last NAL: no second start code -> caller refills / EOS
    assert ours_prefix(L, np.array([0, 0], np.uint8), 2) == (-1, 0)
    assert ours_prefix(L, np.array([0, 0, 0], np.uint8), 3) == (-1, 0)     # no over-read of buf[3]
    assert ours_prefix(L, np.array([0, 0, 0, 1], np.uint8), 4) == (0, 4)
    assert ours_prefix(L, None if False else np.zeros(0, np.uint8), 0) == (-1, 0)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref/libjmref.so not built")
def test_matches_reference_functions_on_random_buffers():
    L = J.load()
    R = C.CDLL(oracle.REF_SO)
    R.jmref_find_nalu_prefix.argtypes = [_u8p, C.c_int, C.POINTER(C.c_int)]
    R.jmref_find_nalu.argtypes = [_u8p, C.c_int, C.POINTER(C.c_int)]
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(3000):
        size = int(rng.integers(0, 48))
        buf = rng.choice(np.array([0, 0, 0, 1, 1, 2, 0x65], np.uint8), size=size + 8)
        buf[size:] = 0xFF                                   # what the reference's over-read (:47) would see
        n1 = C.c_int(-9)
        r_off = R.jmref_find_nalu_prefix(buf.ctypes.data_as(_u8p), size, C.byref(n1))
        assert ours_prefix(L, buf, size) == (r_off, n1.value), (buf[:size].tolist())
        if r_off >= 0:                                      # find_nalu presumes a leading start code (:72)
            n2 = C.c_int(-9)
            r_nal = R.jmref_find_nalu(buf.ctypes.data_as(_u8p), size, C.byref(n2))
            assert ours_nalu(L, buf, size) == (r_nal, n2.value), (buf[:size].tolist())
            checked += 1
    assert checked > 1000


def test_splitter_drives_a_whole_stream():
    """The refill loop of test_nv_dec.cpp:184-213 in miniature: every NAL comes out once, in order."""
    import fake_stream as FS
    L = J.load()
    frames = [np.full(40, v, np.uint8) for v in (5, 6, 7)]
    stream = np.concatenate([FS.sequence_header(16, 16)] + [FS.picture(f, long_start=(i != 1)) for i, f in enumerate(frames)])
    want = FS.split_nals(stream)
    got, pos = [], 0
    while True:
        off, n = ours_nalu(L, stream[pos:].copy(), stream.size - pos)
        if off < 0:
            got.append(stream[pos:])                       # end of stream: the remainder is the last NAL (:199-203)
            break
        got.append(stream[pos + off:pos + off + n])
        pos += off + n
    assert len(got) == len(want) == 4
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
