"""The checker itself: C restatement (oracle/jm_oracle.c) against (1) the committed golden vectors
generated from the unmodified reference, and (2) the compiled reference oracle/_ref when present."""
import json
import os

import numpy as np
import pytest

import cases as K
import oracle

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_sha256.json")))
ALL = K.all_cases()


def test_golden_covers_every_case():
    assert set(GOLD) == {K.case_id(c) for c in ALL}
    assert sum(1 for e in GOLD.values() if e["source"] == "reference") >= 100


@pytest.mark.parametrize("c", ALL, ids=K.case_id)
def test_port_matches_golden(c):
    g = GOLD[K.case_id(c)]
    r, n, out = K.run_case(oracle.port(), c)
    assert (int(r), int(n), int(out.size)) == (g["ret"], g["out_len"], g["nbytes"])
    assert K.sha(out) == g["sha256"]
    if "hex" in g:
        assert out.tobytes().hex() == g["hex"]


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref/libjmref.so not built")
@pytest.mark.parametrize("c", [c for c in ALL if c["op"] in K.REF_OPS], ids=K.case_id)
def test_port_matches_compiled_reference(c):
    r1, n1, o1 = K.run_case(oracle.port(), c)
    r2, n2, o2 = K.run_case(oracle.ref(), c)
    assert (r1, n1) == (r2, n2)
    assert np.array_equal(o1, o2)


@pytest.mark.parametrize("chk", ["port", "ref"])
def test_error_codes(chk):
    if chk == "ref" and not oracle.have_ref():
        pytest.skip("no compiled reference")
    k = oracle.port() if chk == "port" else oracle.ref()
    w, h, p = 16, 16, 32
    s = np.zeros(p * h * 3 // 2, np.uint8)
    out = np.full(w * h * 3 // 2, 0xA5, np.uint8)
    need = w * h * 3 // 2
    # nv_dec.cpp:757-758 no current frame; :768-771 NULL surface; :773-774 short buffer keeps *out_len
    assert k.nvdec_output_frame(s, p, w, h, 1, out, need, have_frame=False) == (-1, need)
    assert k.nvdec_output_frame(None, p, w, h, 1, out, need) == (-1, need)
    assert k.nvdec_output_frame(s, p, w, h, 1, out, need - 1) == (-2, need - 1)
    assert (out == 0xA5).all()
    assert k.nvdec_output_frame(s, p, w, h, 1, out, need) == (need, need)      # :827 returns the byte count
    # intel_dec.cpp:251-255 / :264-268 both zero *out_len
    assert k.inteldec_output_frame(s, p * h, p, (0, 0, w, h), 1, out, need, have_surface=False) == (-1, 0)
    assert k.inteldec_output_frame(s, p * h, p, (0, 0, w, h), 1, out, need - 1) == (-2, 0)
    assert k.inteldec_output_frame(s, p * h, p, (0, 0, w, h), 1, out, need) == (0, need)
    # intel_enc.cpp:254-259 no free surface
    yuv = np.zeros(need, np.uint8)
    surf = np.zeros(p * h * 3 // 2, np.uint8)
    assert k.intelenc_input(yuv, 1, surf, p * h, p, (w, h), (0, 0, w, h), surface_free=False) == -1


def test_i420_is_u_first():
    """out_fmt 1 is really I420 -- even bytes of the UV row go to the FIRST plane."""
    w, h, p = 8, 4, 16
    s = np.zeros(p * h * 3 // 2, np.uint8)
    uv = s[p * h:].reshape(-1, p)
    uv[:, 0:w:2] = 10   # U
    uv[:, 1:w:2] = 200  # V
    out = np.zeros(w * h * 3 // 2, np.uint8)
    oracle.port().nvdec_output_frame(s, p, w, h, 1, out, out.size)
    assert (out[w * h: w * h + 8] == 10).all() and (out[w * h + 8:] == 200).all()


def test_rgb_spec_known_points():
    """Builder-defined BT.601 spec (parity unpinned): black, white, clamps."""
    def px(y, u, v):
        s = np.zeros(2 * 2 * 3 // 2 + 2, np.uint8)
        s[0:4] = y
        s[4], s[5] = u, v
        o = np.zeros(12, np.uint8)
        assert oracle.nv12_to_rgb24(s[:6].copy(), 2, 2, 2, o, 6) == 0
        return tuple(int(x) for x in o[:3])
    assert px(16, 128, 128) == (0, 0, 0)
    assert px(235, 128, 128) == (255, 255, 255)
    assert px(0, 128, 128) == (0, 0, 0)          # clamps below
    assert px(255, 128, 128) == (255, 255, 255)  # clamps above
    assert px(81, 90, 240) == (255, 0, 0)        # BT.601 red
    assert px(145, 54, 34) == (0, 255, 1)        # BT.601 green: (386>>8)=1 by the integer formula
    assert px(41, 240, 110) == (0, 0, 255)       # BT.601 blue
    assert oracle.nv12_to_rgb24(np.zeros(4, np.uint8), 1, 1, 1, np.zeros(3, np.uint8), 3) == -1


def test_argb_is_bgra_bytes_with_opaque_alpha():
    """ARGB32 = NV_ENC_BUFFER_FORMAT_ARGB's memory order: a little-endian 0xAARRGGBB word, i.e. bytes B,G,R,A; the
    colour values are those of the RGB24 spec, alpha is 0xFF."""
    rng = np.random.default_rng(5)
    w, h, p = 10, 6, 16
    s = rng.integers(0, 256, p * h * 3 // 2, dtype=np.uint8)
    rgb, argb = np.zeros(3 * w * h, np.uint8), np.zeros(4 * w * h, np.uint8)
    assert oracle.nv12_to_rgb24(s, p, w, h, rgb, 3 * w) == 0
    assert oracle.nv12_to_argb32(s, p, w, h, argb, 4 * w) == 0
    a, r = argb.reshape(-1, 4), rgb.reshape(-1, 3)
    assert (a[:, 3] == 255).all()
    assert np.array_equal(a[:, 2], r[:, 0]) and np.array_equal(a[:, 1], r[:, 1]) and np.array_equal(a[:, 0], r[:, 2])


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref/libjmref.so not built")
@pytest.mark.parametrize("seed", range(4))
def test_port_matches_compiled_reference_on_random_geometries(seed):
    """Beyond the fixed case matrix: 60 random geometries per seed (odd sizes, crops, slack pitches) through every
    reference function the port restates, compared byte for byte including the bytes neither may touch."""
    rng = np.random.default_rng(7000 + seed)
    P, R = oracle.port(), oracle.ref()
    for _ in range(60):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 120))
        pitch = w + int(rng.integers(0, 40))
        surf = rng.integers(0, 256, pitch * (h * 3 // 2 + 2), dtype=np.uint8)
        cap = w * h * 3 // 2 + 32
        for fmt in (0, 1):                                                   # nv_dec.cpp:750-828
            a, b = np.full(cap, 0xA5, np.uint8), np.full(cap, 0xA5, np.uint8)
            assert P.nvdec_output_frame(surf, pitch, w, h, fmt, a, cap) == R.nvdec_output_frame(surf, pitch, w, h, fmt, b, cap)
            assert np.array_equal(a, b), ("nvdec", w, h, pitch, fmt)
        # intel_dec.cpp:244-332: a crop window inside a larger surface
        rows = h + int(rng.integers(0, 8))
        cw, ch = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
        cx, cy = int(rng.integers(0, w - cw + 1)), int(rng.integers(0, h - ch + 1))
        big = rng.integers(0, 256, pitch * (rows + rows // 2 + 2), dtype=np.uint8)
        ccap = cw * ch * 3 // 2 + 32
        for fmt in (0, 1):
            a, b = np.full(ccap, 0xA5, np.uint8), np.full(ccap, 0xA5, np.uint8)
            ra = P.inteldec_output_frame(big, pitch * rows, pitch, (cx, cy, cw, ch), fmt, a, ccap)
            rb = R.inteldec_output_frame(big, pitch * rows, pitch, (cx, cy, cw, ch), fmt, b, ccap)
            assert ra == rb and np.array_equal(a, b), ("inteldec", w, h, pitch, (cx, cy, cw, ch), fmt)
        # intel_enc.cpp:251-387: tight NV12 / I420 into the crop window of a surface
        yuv = rng.integers(0, 256, cw * ch * 3 // 2 + 8, dtype=np.uint8)
        for i420 in (0, 1):
            a = np.full(big.size, 0xCD, np.uint8)
            b = a.copy()
            ra = P.intelenc_input(yuv, i420, a, pitch * rows, pitch, (w, h), (cx, cy, cw, ch))
            rb = R.intelenc_input(yuv, i420, b, pitch * rows, pitch, (w, h), (cx, cy, cw, ch))
            assert ra == rb and np.array_equal(a, b), ("intelenc", w, h, pitch, (cx, cy, cw, ch), i420)
        # nv_enc.cpp:1023-1103 through the fake CUDA driver
        stride = ((w + 15) & ~15) + 16 * int(rng.integers(0, 3))
        tight = rng.integers(0, 256, w * h * 3 // 2 + 8, dtype=np.uint8)
        for fmt in (0x1, 0x10):
            a = np.full(stride * (h * 3 // 2 + 2), 0xCD, np.uint8)
            b = a.copy()
            ra = oracle.nvenc_upload(tight, fmt, w, h, a, stride)
            rb, _ = oracle.ref_nvenc_convert(tight, fmt, w, h, b, stride)
            assert ra == rb and np.array_equal(a, b), ("nvenc", w, h, stride, hex(fmt))


def test_forward_bt601_known_points():
    """Builder-defined forward transform (parity unpinned): the textbook limited-range values of the primaries, and
    chroma taken from the 2x2 block SUM (one red pixel in a black block moves U,V a quarter of the way)."""
    def blk(px):                         # px: four (r,g,b) of one 2x2 block, row-major
        rgb = np.array(px, np.uint8).reshape(2, 6)
        s = np.zeros(2 * 2 + 2, np.uint8)
        assert oracle.rgb24_to_nv12(rgb.reshape(-1), 6, 2, 2, s, 2) == 0
        return [int(x) for x in s]
    assert blk([(0, 0, 0)] * 4) == [16, 16, 16, 16, 128, 128]
    assert blk([(255, 255, 255)] * 4) == [235, 235, 235, 235, 128, 128]
    assert blk([(255, 0, 0)] * 4) == [82, 82, 82, 82, 90, 240]
    assert blk([(0, 255, 0)] * 4) == [144, 144, 144, 144, 54, 34]
    assert blk([(0, 0, 255)] * 4) == [41, 41, 41, 41, 240, 110]
    one_red = blk([(255, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0)])
    assert one_red[:4] == [82, 16, 16, 16] and one_red[4:] == [(-38 * 255 + 512 + 131072) >> 10, (112 * 255 + 512 + 131072) >> 10]
    # odd sizes: the last column / row has luma only, padding and the missing chroma stay untouched
    s = np.full(16 * 5, 0xCD, np.uint8)                    # 3 luma rows + 1 chroma row (h>>1) + one spare row
    assert oracle.rgb24_to_nv12(np.full(27, 255, np.uint8), 9, 3, 3, s, 16) == 0
    assert [int(x) for x in s[32:36]] == [235, 235, 235, 0xCD]           # third luma row: 3 pixels, then padding
    assert [int(x) for x in s[48:52]] == [128, 128, 0xCD, 0xCD]          # one chroma pair (w>>1), then padding
    assert (s[64:] == 0xCD).all()                                        # no second chroma row (h>>1 == 1)
    assert oracle.rgb24_to_nv12(np.zeros(3, np.uint8), 3, 0, 1, s, 16) == -1
