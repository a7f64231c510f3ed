"""NVDEC front-end of jm_nvdec_* (bitstream codecs through libnvcuvid), driven against the fake
library tests/fake_nvcuvid (the pool's boxes expose no NVDEC engine), and the reference's own test
program compiled against our header and library."""
import os
import re
import subprocess

import numpy as np
import pytest

import fake_stream as FS
import oracle
from jmcodec_b200 import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FAKE = os.path.join(HERE, "fake_nvcuvid", "libfake_nvcuvid.so")
REF_TEST = os.path.join(HERE, "ref_driver", "_ref", "test_nv_dec")


@pytest.fixture()
def fake_env(monkeypatch):
    if not os.path.exists(FAKE):
        pytest.skip("tests/fake_nvcuvid/libfake_nvcuvid.so not built (run __graft_entry__.build())")
    monkeypatch.setenv("JMC_NVCUVID_LIB", FAKE)
    monkeypatch.delenv("FAKE_NVCUVID_NO_ENGINE", raising=False)
    return monkeypatch


def _frames(w, h, n, stream=31):
    # tight NV12 pictures; a few zero runs so that emulation prevention is exercised
    out = []
    for f in range(n):
        t = synth.random_bytes(w * h * 3 // 2, synth.frame_key(stream, f))
        t[100:140] = 0
        t[1000:1003] = (0, 0, 1)
        out.append(t)
    return out


def _want(chk, tight, w, h, out_fmt):
    want = np.empty(w * h * 3 // 2, np.uint8)
    r, n = chk.nvdec_output_frame(tight, w, w, h, out_fmt, want, want.size)     # a tight NV12 frame is a surface with pitch == w
    assert r == n == want.size
    return want


@pytest.mark.parametrize("out_fmt", [0, 1])
@pytest.mark.parametrize("geom", [(1920, 1080), (320, 180), (66, 34)])
def test_bitstream_decode_through_cuvid_frontend(fake_env, out_fmt, geom):
    import jmcodec_b200 as J
    w, h = geom
    need = w * h * 3 // 2
    n = 7
    frames = _frames(w, h, n)
    chk = oracle.best()
    dec = J.NvDec(0)
    assert dec.init(0, out_fmt) == 0, J.last_error()                  # codec 0 = H.264 -> NVDEC front-end
    got_frames = []
    out = np.empty(need, np.uint8)

    def feed(buf, nbytes=None):
        r, got = dec.decode_frame(buf, nbytes)
        assert r == 0
        if got == 1:
            assert dec.output_frame(out, need) == (need, need)
            got_frames.append(out.copy())
        return got

    assert feed(FS.sequence_header(w, h)) == 0
    assert dec.stream_info() == (w, h)
    gots = [feed(FS.picture(f, long_start=(i % 2 == 0))) for i, f in enumerate(frames)]
    assert gots[:2] == [0, 0] and all(g == 1 for g in gots[2:])        # ulMaxDisplayDelay = 2 (nv_dec.cpp:346)
    # flush: EOS hands out the delayed pictures, one per call, then is_exit (test_nv_dec.cpp:232-246)
    calls = 0
    while not dec.is_exit():
        feed(None, 0)
        calls += 1
        assert calls < 10
    assert len(got_frames) == n
    for i in range(n):
        assert np.array_equal(got_frames[i], _want(chk, frames[i], w, h, out_fmt)), f"frame {i}"
    info = dec.show_dec_info()
    assert "Codec:\t\tH.264" in info and f"Frame Count:\t{n}" in info and f"Display:\t{w} x {h}" in info
    assert dec.deinit() == 0


def test_several_nals_in_one_packet_and_queueing(fake_env):
    """A packet holding many pictures fills the display queue; frames still come out one per call."""
    import jmcodec_b200 as J
    w, h, n = 128, 72, 6
    need = w * h * 3 // 2
    frames = _frames(w, h, n, stream=32)
    stream = np.concatenate([FS.sequence_header(w, h)] + [FS.picture(f) for f in frames])
    chk = oracle.best()
    dec = J.NvDec(0)
    assert dec.init(0, 1) == 0
    out = np.empty(need, np.uint8)
    got_frames = []
    r, got = dec.decode_frame(stream)
    while True:
        if got == 1:
            assert dec.output_frame(out, need) == (need, need)
            got_frames.append(out.copy())
        if dec.is_exit():
            break
        r, got = dec.decode_frame(None, 0)
    assert len(got_frames) == n
    for i in range(n):
        assert np.array_equal(got_frames[i], _want(chk, frames[i], w, h, 1))
    dec.deinit()


def test_no_engine_fails_loudly(fake_env):
    import jmcodec_b200 as J
    fake_env.setenv("FAKE_NVCUVID_NO_ENGINE", "1")
    dec = J.NvDec(0)
    assert dec.init(0, 1) == -4
    assert "NVDEC is not usable" in J.last_error()
    assert dec.decode_frame(np.zeros(32, np.uint8)) == (0, 0)
    assert dec.decode_frame(None, 0) == (0, 0)
    assert dec.is_exit()                                              # a drained dead stream still terminates the caller's loop
    dec.deinit()


def test_missing_library_fails_loudly(monkeypatch):
    import jmcodec_b200 as J
    monkeypatch.setenv("JMC_NVCUVID_LIB", "/nonexistent/libnvcuvid.so.1")
    dec = J.NvDec(0)
    assert dec.init(1, 1) == -4
    assert "cannot load the NVDEC library" in J.last_error()
    dec.deinit()


@pytest.mark.skipif(not os.path.exists(REF_TEST), reason="tests/ref_driver/_ref/test_nv_dec not built (needs /root/reference)")
def test_reference_test_program_runs_against_our_library(fake_env, tmp_path):
    """test_nv_dec/test_nv_dec.cpp, UNMODIFIED, compiled against include/jm_nv_dec.h + libjmcodec_b200.so:
    its own NAL splitter and decode loop (test_nv_dec.cpp:30-86,163-259) drive our API to completion."""
    w, h, n = 320, 180, 9
    frames = _frames(w, h, n, stream=33)
    stream = np.concatenate([FS.sequence_header(w, h)] + [FS.picture(f) for f in frames])
    path = tmp_path / "stream.264"
    stream.tofile(path)
    env = dict(os.environ, JM_TEST_INPUT=str(path), JMC_NVCUVID_LIB=FAKE)
    p = subprocess.run([REF_TEST], env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert re.search(rf"Frame Count:\s+{n}\b", p.stdout), p.stdout
    assert re.search(rf"Display:\s+{w} x {h}", p.stdout)
    assert "Pixel Format:\tYV12" in p.stdout                            # the program asks for out_fmt 1 (test_nv_dec.cpp:167)
    assert re.search(rf"nalu count = {n + 1 + 3}\b", p.stdout) or re.search(r"nalu count = \d+", p.stdout)


# --------------------------------------------------------------------------------------------
# batch drain: several displayed pictures mapped together and converted by one launch
# --------------------------------------------------------------------------------------------
def _fake_stats(reset=False):
    import ctypes as C
    lib = C.CDLL(FAKE)
    v = (C.c_int * 6)()
    lib.fake_nvcuvid_stats(v, 1 if reset else 0)
    return dict(zip(("maps", "unmaps", "cur_mapped", "max_mapped", "map_refused", "decoders"), list(v)))


def _run_stream(dec, packets, need):
    """Feed packets, then flush; returns the frames in output order and the decode_frame return codes."""
    out = np.empty(need, np.uint8)
    frames, rets = [], []

    def step(buf, nbytes=None):
        r, got = dec.decode_frame(buf, nbytes)
        rets.append(r)
        if got == 1:
            assert dec.output_frame(out, need) == (need, need)
            frames.append(out.copy())

    for p in packets:
        step(p)
    calls = 0
    while not dec.is_exit():
        step(None, 0)
        calls += 1
        assert calls < 200
    return frames, rets


@pytest.mark.parametrize("map_limit,per_packet", [(8, 8), (8, 3), (4, 8), (1, 8)])
def test_batch_drain_of_a_64_frame_stream(fake_env, map_limit, per_packet):
    """64 pictures, several per packet: every picture that becomes displayable in one call is mapped (up to the
    map limit) and converted by ONE launch; surfaces are unmapped only after the convert event, never per frame.
    The fake poisons a surface at unmap and overwrites decode surfaces at once, so any ordering bug shows up as a
    wrong frame."""
    import jmcodec_b200 as J
    w, h, n = 320, 180, 64
    need = w * h * 3 // 2
    frames = _frames(w, h, n, stream=34)
    chk = oracle.best()
    _fake_stats(reset=True)
    dec = J.NvDec(0)
    assert dec.set_option("map_limit", map_limit) == 0
    assert dec.init(0, 1) == 0, J.last_error()
    packets = [FS.sequence_header(w, h)]
    for i in range(0, n, per_packet):
        packets.append(np.concatenate([FS.picture(f) for f in frames[i:i + per_packet]]))
    got, rets = _run_stream(dec, packets, need)
    assert all(r == 0 for r in rets) and dec.dropped_frames == 0
    assert len(got) == n
    for i in range(n):
        assert np.array_equal(got[i], _want(chk, frames[i], w, h, 1)), f"frame {i}"
    launches = dec.launches
    st = _fake_stats()
    assert st["maps"] == n and st["map_refused"] == 0
    assert st["max_mapped"] <= map_limit
    if map_limit > 1 and per_packet > 1:
        assert st["max_mapped"] >= 2 and launches < n          # really batched
    else:
        assert launches == n
    assert dec.deinit() == 0
    st = _fake_stats()
    assert st["unmaps"] == n and st["cur_mapped"] == 0        # nothing left mapped


def test_many_pictures_in_one_packet_overflow_is_reported(fake_env):
    """More pictures in ONE packet than the handle can hold (64 converted frames): the surplus is dropped, the
    call returns -1 (the reference would have let the decoder overwrite queued surfaces silently), and every
    frame that does come out is a correct picture of the stream, in order."""
    import jmcodec_b200 as J
    w, h, n = 64, 36, 110
    need = w * h * 3 // 2
    frames = _frames(w, h, n, stream=35)
    chk = oracle.best()
    dec = J.NvDec(0)
    assert dec.init(0, 0) == 0
    stream = np.concatenate([FS.sequence_header(w, h)] + [FS.picture(f) for f in frames])
    got, rets = _run_stream(dec, [stream], need)
    dropped = dec.dropped_frames
    assert dropped > 0 and rets[0] == -1 and "dropped" in J.last_error()
    assert len(got) + dropped == n
    wants = [_want(chk, f, w, h, 0) for f in frames]
    pos = 0
    for g in got:                                             # a subsequence of the stream, order preserved
        while pos < n and not np.array_equal(g, wants[pos]):
            pos += 1
        assert pos < n
        pos += 1
    dec.deinit()


def test_format_change_mid_stream(fake_env):
    """A second sequence header: the old decoder's queued pictures are converted and unmapped before it is
    destroyed; frames of both geometries come out intact."""
    import jmcodec_b200 as J
    chk = oracle.best()
    (w1, h1), (w2, h2) = (320, 180), (192, 128)
    a, b = _frames(w1, h1, 5, stream=36), _frames(w2, h2, 6, stream=37)
    _fake_stats(reset=True)
    dec = J.NvDec(0)
    assert dec.init(0, 1) == 0
    stream = np.concatenate([FS.sequence_header(w1, h1)] + [FS.picture(f) for f in a] +
                            [FS.sequence_header(w2, h2)] + [FS.picture(f) for f in b])
    out = np.empty(w1 * h1 * 3 // 2, np.uint8)
    got = []
    r, g = dec.decode_frame(stream)
    while True:
        if g == 1:
            k = len(got)
            w, h = (w1, h1) if k < len(a) else (w2, h2)
            need = w * h * 3 // 2
            assert dec.output_frame(out, out.size) == (need, need)
            got.append(out[:need].copy())
        if dec.is_exit():
            break
        r, g = dec.decode_frame(None, 0)
    assert len(got) == len(a) + len(b)
    for i, f in enumerate(a):
        assert np.array_equal(got[i], _want(chk, f, w1, h1, 1)), f"first sequence, frame {i}"
    for i, f in enumerate(b):
        assert np.array_equal(got[len(a) + i], _want(chk, f, w2, h2, 1)), f"second sequence, frame {i}"
    assert dec.stream_info() == (w2, h2)
    dec.deinit()
    st = _fake_stats()
    assert st["decoders"] == 2 and st["maps"] == st["unmaps"] == len(a) + len(b)


def test_display_delay_on_top_of_the_parser(fake_env):
    """jm_nvdec_set_display_delay adds to the parser's own delay; every frame still comes out, in order."""
    import jmcodec_b200 as J
    w, h, n = 128, 72, 12
    need = w * h * 3 // 2
    frames = _frames(w, h, n, stream=38)
    chk = oracle.best()
    dec = J.NvDec(0)
    assert dec.set_display_delay(3) == 0
    assert dec.init(0, 1) == 0
    got, rets = _run_stream(dec, [FS.sequence_header(w, h)] + [FS.picture(f) for f in frames], need)
    assert len(got) == n
    for i in range(n):
        assert np.array_equal(got[i], _want(chk, frames[i], w, h, 1))
    dec.deinit()
