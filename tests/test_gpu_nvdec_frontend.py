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
