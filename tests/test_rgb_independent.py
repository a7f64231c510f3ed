"""Independent checks of the colour legs.

The reference has no YUV<->RGB code (SDL2 does it), so the integer formulas are builder-defined
and the oracle's restatement of them is "parity unpinned".  What can be pinned is that those integer formulas
ARE ITU-R BT.601 limited-range, not a shared misreading: here the oracle (which the CUDA kernels match bit for
bit, tests/test_gpu_parity.py) is compared with a float64 evaluation of the standard's equations written from the
standard, with known colour points, and through inverse-of-forward error bounds.  The GPU versions of the same
checks (marked gpu) run the kernels instead of the oracle.
"""
import numpy as np
import pytest

import oracle
from jmcodec_b200 import synth

# ITU-R BT.601-7: E'Y = 0.299 R + 0.587 G + 0.114 B; 8-bit limited range: Y = 16 + 219 E'Y, Cb = 128 + 224 (B-Y)/1.772,
# Cr = 128 + 224 (R-Y)/1.402 (E' in 0..1).  Everything below derives from these three numbers.
KR, KG, KB = 0.299, 0.587, 0.114


def float_yuv_to_rgb(y, cb, cr):
    """float64 BT.601 limited-range inverse, per pixel, unrounded (arrays of equal shape)."""
    ey = (y.astype(np.float64) - 16.0) / 219.0
    pb = (cb.astype(np.float64) - 128.0) / 224.0
    pr = (cr.astype(np.float64) - 128.0) / 224.0
    r = ey + (2 - 2 * KR) * pr
    b = ey + (2 - 2 * KB) * pb
    g = (ey - KR * r - KB * b) / KG
    return 255.0 * r, 255.0 * g, 255.0 * b


def float_rgb_to_yuv(r, g, b):
    er, eg, eb = r.astype(np.float64) / 255.0, g.astype(np.float64) / 255.0, b.astype(np.float64) / 255.0
    ey = KR * er + KG * eg + KB * eb
    return 16.0 + 219.0 * ey, 128.0 + 224.0 * (eb - ey) / (2 - 2 * KB), 128.0 + 224.0 * (er - ey) / (2 - 2 * KR)


def surface_from_planes(y, cb, cr, pitch):
    """Pitched NV12 surface from full-resolution luma and half-resolution chroma planes."""
    h, w = y.shape
    s = np.full((h * 3 // 2, pitch), synth.PAD_BYTE, np.uint8)
    s[:h, :w] = y
    s[h:, 0:w:2] = cb
    s[h:, 1:w:2] = cr
    return s.reshape(-1)


def oracle_rgb(surf, pitch, w, h):
    out = np.empty(3 * w * h, np.uint8)
    assert oracle.nv12_to_rgb24(surf, pitch, w, h, out, 3 * w) == 0
    return out.reshape(h, w, 3)


def gpu_rgb(ctx, surf, pitch, w, h):
    ds, dr = ctx.upload(surf), ctx.alloc(3 * w * h)
    j = ctx.job_rgb(w, h, pitch, 3 * w, False)
    j.n_frames, j.surf.base, j.rgb.base = 1, ds, dr
    ctx.convert(j)
    out = np.empty(3 * w * h, np.uint8)
    ctx.d2h(out, dr)
    ctx.free(ds), ctx.free(dr)
    return out.reshape(h, w, 3)


def oracle_nv12(rgb, w, h, pitch):
    surf = np.full(pitch * h * 3 // 2, synth.PAD_BYTE, np.uint8)
    assert oracle.rgb24_to_nv12(np.ascontiguousarray(rgb).reshape(-1), 3 * w, w, h, surf, pitch) == 0
    return surf


def gpu_nv12(ctx, rgb, w, h, pitch):
    dr, ds = ctx.upload(np.ascontiguousarray(rgb).reshape(-1)), ctx.alloc(pitch * h * 3 // 2)
    ctx.memset(ds, synth.PAD_BYTE, pitch * h * 3 // 2)
    j = ctx.job_rgb_to_nv12(w, h, 3 * w, pitch)
    j.n_frames, j.rgb.base, j.surf.base = 1, dr, ds
    ctx.convert(j)
    out = np.empty(pitch * h * 3 // 2, np.uint8)
    ctx.d2h(out, ds)
    ctx.free(dr), ctx.free(ds)
    return out


@pytest.fixture(scope="module")
def ctx():
    import jmcodec_b200 as J
    c = J.Ctx(0)
    yield c
    c.close()


def _impl(request, kind):
    """(nv12->rgb, rgb->nv12) for the oracle or, in the gpu-marked variants, the CUDA kernels."""
    if kind == "oracle":
        return oracle_rgb, oracle_nv12
    c = request.getfixturevalue("ctx")
    return (lambda s, p, w, h: gpu_rgb(c, s, p, w, h)), (lambda rgb, w, h, p: gpu_nv12(c, rgb, w, h, p))


KINDS = ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("kind", KINDS)
def test_inverse_matches_float_bt601_within_one_lsb(request, kind):
    """Every (Y, Cb, Cr) with Y in steps of 1 and chroma on a 5-step grid: |integer - round(float)| <= 1 on every
    channel, and the integer result equals round-half-up of the float result for >= 93 % of the samples (measured: 95 %)."""
    to_rgb, _ = _impl(request, kind)
    ys = np.arange(256, dtype=np.uint8)
    cs = np.arange(0, 256, 5, dtype=np.uint8)
    w, h = 2 * len(ys), 2 * len(cs) * len(cs)                       # one 2x2 block per (Y, Cb, Cr)
    y = np.repeat(np.repeat(ys[None, :], h, axis=0), 2, axis=1)
    cb = np.repeat(np.repeat(cs, len(cs))[:, None], len(ys), axis=1)
    cr = np.repeat(np.tile(cs, len(cs))[:, None], len(ys), axis=1)
    pitch = (w + 63) & ~63
    got = to_rgb(surface_from_planes(y, cb, cr, pitch), pitch, w, h).astype(np.int32)
    fr, fg, fb = float_yuv_to_rgb(y, np.repeat(np.repeat(cb, 2, 0), 2, 1), np.repeat(np.repeat(cr, 2, 0), 2, 1))
    want = np.stack([np.clip(np.floor(c + 0.5), 0, 255) for c in (fr, fg, fb)], axis=-1).astype(np.int32)
    diff = np.abs(got - want)
    assert diff.max() <= 1, f"max |integer - float| = {diff.max()}"
    assert (diff == 0).mean() >= 0.93


@pytest.mark.parametrize("kind", KINDS)
def test_known_colour_points(request, kind):
    """Black, white, mid grey and the 100 % colour bars of BT.601 (studio-swing codes from the standard's equations)."""
    to_rgb, to_nv12 = _impl(request, kind)
    bars = {  # name: (R, G, B)
        "black": (0, 0, 0), "white": (255, 255, 255), "grey": (128, 128, 128), "red": (255, 0, 0), "green": (0, 255, 0),
        "blue": (0, 0, 255), "yellow": (255, 255, 0), "cyan": (0, 255, 255), "magenta": (255, 0, 255),
    }
    # codes every textbook lists for 100 % bars: Y, Cb, Cr
    codes = {"black": (16, 128, 128), "white": (235, 128, 128), "red": (81, 90, 240), "green": (145, 54, 34),
             "blue": (41, 240, 110), "yellow": (210, 16, 146), "cyan": (170, 166, 16), "magenta": (106, 202, 222)}
    w = h = 16
    pitch = 64
    for name, (r, g, b) in bars.items():
        rgb = np.empty((h, w, 3), np.uint8)
        rgb[...] = (r, g, b)
        surf = to_nv12(rgb, w, h, pitch).reshape(-1, pitch)
        yv, cbv, crv = int(surf[0, 0]), int(surf[h, 0]), int(surf[h, 1])
        assert (surf[:h, :w] == yv).all() and (surf[h:, 0:w:2] == cbv).all() and (surf[h:, 1:w:2] == crv).all()
        fy, fcb, fcr = (float(v) for v in float_rgb_to_yuv(np.array(r), np.array(g), np.array(b)))
        assert abs(yv - fy) <= 1 and abs(cbv - fcb) <= 1 and abs(crv - fcr) <= 1, name
        if name in codes:
            assert max(abs(yv - codes[name][0]), abs(cbv - codes[name][1]), abs(crv - codes[name][2])) <= 1, (name, yv, cbv, crv)
            y = np.full((h, w), codes[name][0], np.uint8)
            cb = np.full((h // 2, w // 2), codes[name][1], np.uint8)
            cr = np.full((h // 2, w // 2), codes[name][2], np.uint8)
            back = to_rgb(surface_from_planes(y, cb, cr, pitch), pitch, w, h).astype(np.int32)
            assert np.abs(back - np.array([r, g, b])).max() <= 2, (name, back[0, 0])
    # exact anchors of the integer formulas: the legal-range end points and clamping beyond them
    for yv, want in ((16, 0), (235, 255), (0, 0), (255, 255), (126, 128)):
        y = np.full((h, w), yv, np.uint8)
        c = np.full((h // 2, w // 2), 128, np.uint8)
        assert (to_rgb(surface_from_planes(y, c, c, pitch), pitch, w, h) == want).all(), yv


@pytest.mark.parametrize("kind", KINDS)
def test_forward_matches_float_bt601_within_one_lsb(request, kind):
    """Random RGB: luma per pixel and chroma of the 2x2 block MEAN against the float equations."""
    _, to_nv12 = _impl(request, kind)
    w, h, pitch = 256, 128, 256
    rng = np.random.default_rng(601)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    surf = to_nv12(rgb, w, h, pitch).reshape(-1, pitch).astype(np.float64)
    fy, _, _ = float_rgb_to_yuv(rgb[..., 0], rgb[..., 1], rgb[..., 2])
    assert np.abs(surf[:h, :w] - fy).max() <= 1.0
    mean = rgb.reshape(h // 2, 2, w // 2, 2, 3).astype(np.float64).mean(axis=(1, 3))
    _, fcb, fcr = float_rgb_to_yuv(mean[..., 0], mean[..., 1], mean[..., 2])
    assert np.abs(surf[h:, 0:w:2] - fcb).max() <= 1.0 and np.abs(surf[h:, 1:w:2] - fcr).max() <= 1.0


@pytest.mark.parametrize("kind", KINDS)
def test_inverse_of_forward_error_bound(request, kind):
    """RGB -> NV12 -> RGB on frames whose 2x2 blocks are flat (so chroma subsampling loses nothing): the two 8-bit
    quantisations bound the error.  The float pipeline with the same roundings gives <= 2 on every channel except
    blue <= 3 (B carries 2.017 x the Cb rounding error); the integer pipeline must stay within that."""
    to_rgb, to_nv12 = _impl(request, kind)
    w, h, pitch = 512, 256, 512
    rng = np.random.default_rng(602)
    blocks = rng.integers(0, 256, (h // 2, w // 2, 3), dtype=np.uint8)
    rgb = np.repeat(np.repeat(blocks, 2, axis=0), 2, axis=1)
    back = to_rgb(to_nv12(rgb, w, h, pitch), pitch, w, h).astype(np.int32)
    err = np.abs(back - rgb.astype(np.int32))
    assert err[..., 0].max() <= 2 and err[..., 1].max() <= 2 and err[..., 2].max() <= 3, err.reshape(-1, 3).max(axis=0)
    assert err.mean() < 0.6
    # and the reverse order on legal-range grey: chroma exact, luma within 1 (the property the 4K device test uses)
    yv = rng.integers(16, 236, (h, w), dtype=np.uint8)
    c = np.full((h // 2, w // 2), 128, np.uint8)
    s2 = to_nv12(to_rgb(surface_from_planes(yv, c, c, pitch), pitch, w, h), w, h, pitch).reshape(-1, pitch)
    assert np.abs(s2[:h, :w].astype(np.int32) - yv).max() <= 1 and (s2[h:, :w] == 128).all()
