"""GPU parity: the CUDA path (through the C-ABI) against the oracle and the golden vectors.

Bit-exact bar: every output byte AND every sentinel byte (slack after the frame, surface padding)
must equal what the reference's CPU code produces on the same synthetic input.
"""
import json
import os

import numpy as np
import pytest

import cases as K
import gpu_runner as G
import oracle
from jmcodec_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_sha256.json")))
GPU_CASES = [c for c in K.all_cases() if not (c["op"] == "nvenc" and c["fmt"] in ("argb", "abgr"))]


@pytest.fixture(scope="module")
def ctx():
    import jmcodec_b200 as J
    c = J.Ctx(0)
    yield c
    c.close()


@pytest.mark.parametrize("c", GPU_CASES, ids=K.case_id)
def test_matches_golden_and_oracle(ctx, c):
    out = G.run_case_gpu(ctx, c)
    g = GOLD[K.case_id(c)]
    assert out.size == g["nbytes"]
    if K.sha(out) != g["sha256"]:
        _, _, ref = K.run_case(oracle.best() if c["op"] in K.REF_OPS else oracle.port(), c)
        bad = np.flatnonzero(out != ref)
        pytest.fail(f"{bad.size} byte(s) differ from the oracle, first at {bad[:8]}: got {out[bad[:8]]} want {ref[bad[:8]]}")


@pytest.mark.parametrize("c", [c for c in K.rgb_cases()], ids=K.case_id)
def test_fused_i420_rgb(ctx, c):
    """Fused op: RGB identical to the RGB-only op, I420 identical to the nv_dec I420 oracle."""
    rgb, tight = G.gpu_rgb(ctx, c, fused=True)
    assert K.sha(rgb) == GOLD[K.case_id(c)]["sha256"]
    want = np.full(tight.size, synth.OUT_FILL, np.uint8)
    oracle.best().nvdec_output_frame(K.rgb_input(c), c["pitch"], c["w"], c["h"], 1, want, want.size)
    assert np.array_equal(tight, want)


@pytest.mark.parametrize("mode", ["stride", "list", "list_aligned_flag", "list_misaligned"])
@pytest.mark.parametrize("op", ["i420", "nv12", "pack", "rgb"])
def test_batched_launch(ctx, op, mode):
    """One launch over a batch: frame strides and device pointer lists (aligned / deliberately odd)."""
    import jmcodec_b200 as J
    w, h, pitch, n = 96, 34, 128, 7
    surf_bytes = pitch * h * 3 // 2
    tight_bytes = w * h * 3 // 2
    rgb_bytes = 3 * w * h
    skew = 1 if mode == "list_misaligned" else 0          # odd byte offset defeats every vector path
    sstride, tstride, rstride = surf_bytes + 256 + skew, tight_bytes + 64 + skew, rgb_bytes + 32 + skew
    chk = oracle.best()
    surfs = [synth.nv12_surface(w, h, pitch, 9, f) for f in range(n)]
    tights = [synth.i420_frame(w, h, 9, f) for f in range(n)]
    if op == "pack":
        src = np.full(n * tstride, 0x11, np.uint8)
        for f in range(n):
            src[f * tstride:f * tstride + tight_bytes] = tights[f]
        dsrc = ctx.upload(src)
        ddst = ctx.alloc(n * sstride)
        ctx.memset(ddst, synth.PAD_BYTE, n * sstride)
        j = ctx.job_nvenc(w, h, pitch, 0x10)
        j.tight.base, j.tight.stride, j.surf.base, j.surf.stride = dsrc, tstride, ddst, sstride
        out_stride, out_total = sstride, n * sstride
    else:
        src = np.full(n * sstride, 0x11, np.uint8)
        for f in range(n):
            src[f * sstride:f * sstride + surf_bytes] = surfs[f]
        dsrc = ctx.upload(src)
        if op == "rgb":
            ddst = ctx.alloc(n * rstride)
            ctx.memset(ddst, synth.OUT_FILL, n * rstride)
            j = ctx.job_rgb(w, h, pitch, 3 * w, False)
            j.rgb.base, j.rgb.stride = ddst, rstride
            out_stride, out_total = rstride, n * rstride
        else:
            ddst = ctx.alloc(n * tstride)
            ctx.memset(ddst, synth.OUT_FILL, n * tstride)
            j = ctx.job_nvdec(w, h, pitch, 1 if op == "i420" else 0)
            j.tight.base, j.tight.stride = ddst, tstride
            out_stride, out_total = tstride, n * tstride
        j.surf.base, j.surf.stride = dsrc, sstride
    j.n_frames = n
    lists = []
    if mode != "stride":
        for fs in (j.surf, j.tight, j.rgb):
            if fs.base:
                d = G.device_ptr_array(ctx, [fs.base + f * fs.stride for f in range(n)])
                lists.append(d)
                fs.list, fs.base, fs.stride = d, None, 0
        if mode == "list_aligned_flag":
            j.flags = J.lib.JOB_ALIGNED16
    ctx.convert(j)
    got = np.empty(out_total, np.uint8)
    ctx.d2h(got, ddst)
    for f in range(n):
        g = got[f * out_stride:(f + 1) * out_stride]
        if op == "pack":
            want = np.full(out_stride, synth.PAD_BYTE, np.uint8)
            oracle.nvenc_upload(tights[f], 0x10, w, h, want, pitch)
        elif op == "rgb":
            want = np.full(out_stride, synth.OUT_FILL, np.uint8)
            oracle.nv12_to_rgb24(surfs[f], pitch, w, h, want, 3 * w)
        else:
            want = np.full(out_stride, synth.OUT_FILL, np.uint8)
            chk.nvdec_output_frame(surfs[f], pitch, w, h, 1 if op == "i420" else 0, want, want.size)
        assert np.array_equal(g, want), f"frame {f}"
    for d in [dsrc, ddst] + lists:
        ctx.free(d)


def test_bad_jobs_are_rejected(ctx):
    import jmcodec_b200 as J
    j = ctx.job_nvdec(64, 64, 64, 1)
    j.n_frames = 1
    with pytest.raises(J.JmcError):
        ctx.convert(j)                      # no frame sets
    with pytest.raises(J.JmcError):
        ctx.job_nvdec(64, 64, 32, 1)        # pitch < width
    with pytest.raises(J.JmcError):
        ctx.job_rgb(64, 64, 64, 100, False)  # rgb_pitch < 3w
    d = ctx.alloc(1 << 16)
    j = ctx.job_rgb(1, 8, 16, 3, False)
    j.n_frames, j.surf.base, j.rgb.base = 1, d, d
    with pytest.raises(J.JmcError):
        ctx.convert(j)                      # RGB needs w,h >= 2 (oracle returns -1)
    j = ctx.job_nvdec(0, 0, 0, 1)
    j.n_frames, j.surf.base, j.tight.base = 1, d, d
    ctx.convert(j)                          # empty frame: nothing to do, not an error
    ctx.free(d)


def _rand_geoms(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        w = int(rng.integers(1, 200))
        h = int(rng.integers(1, 120))
        pitch = w + int(rng.integers(0, 70))
        out.append((w, h, pitch, int(rng.integers(0, 16)), int(rng.integers(0, 16)), int(rng.integers(0, 16))))
    return out


@pytest.mark.parametrize("seed", range(6))
def test_random_geometries_against_oracle(ctx, seed):
    """Random sizes, pitches and buffer skews (every alignment class) for all four YUV ops: one frame each,
    sentinel-checked, against the oracle."""
    chk = oracle.best()
    for (w, h, pitch, ks, kt, kr) in _rand_geoms(40, 1000 + seed):
        surf = synth.nv12_surface(w, h, pitch, 21, w * 131 + h)
        tight_in = synth.i420_frame(w, h, 22, w * 17 + h)
        cap = w * h * 3 // 2 + K.SLACK
        nsurf = pitch * (h * 3 // 2 + 1)
        # device buffers at skewed (possibly odd) addresses
        dsurf = ctx.alloc(surf.size + 64)
        ctx.h2d(dsurf + ks, surf)
        for fmt in (0, 1):
            dout = ctx.alloc(cap + 64)
            ctx.memset(dout, synth.OUT_FILL, cap + 64)
            j = ctx.job_nvdec(w, h, pitch, fmt)
            j.n_frames, j.surf.base, j.tight.base = 1, dsurf + ks, dout + kt
            ctx.convert(j)
            got = np.empty(cap, np.uint8)
            ctx.d2h(got, dout + kt)
            want = np.full(cap, synth.OUT_FILL, np.uint8)
            chk.nvdec_output_frame(surf, pitch, w, h, fmt, want, cap)
            assert np.array_equal(got, want), (w, h, pitch, ks, kt, fmt)
            ctx.free(dout)
        dtin = ctx.alloc(tight_in.size + 64)
        ctx.h2d(dtin + kt, tight_in)
        for code in (0x1, 0x10):
            ds = ctx.alloc(nsurf + 64)
            ctx.memset(ds, synth.PAD_BYTE, nsurf + 64)
            j = ctx.job_nvenc(w, h, pitch, code)
            j.n_frames, j.surf.base, j.tight.base = 1, ds + ks, dtin + kt
            ctx.convert(j)
            got = np.empty(nsurf, np.uint8)
            ctx.d2h(got, ds + ks)
            want = np.full(nsurf, synth.PAD_BYTE, np.uint8)
            # tight_in is w*h*3/2 bytes; the yv12 path may read up to y_len*5/4 + (w/2)*(h/2) <= that
            oracle.nvenc_upload(tight_in, code, w, h, want, pitch)
            assert np.array_equal(got, want), (w, h, pitch, ks, kt, hex(code))
            ctx.free(ds)
        if w >= 2 and h >= 2:
            rp = 3 * w + kr
            rcap = rp * h + K.SLACK
            for fused in (False, True):
                drgb = ctx.alloc(rcap + 64)
                ctx.memset(drgb, synth.OUT_FILL, rcap + 64)
                dt = ctx.alloc(cap + 64)
                ctx.memset(dt, synth.OUT_FILL, cap + 64)
                j = ctx.job_rgb(w, h, pitch, rp, fused)
                j.n_frames, j.surf.base, j.rgb.base, j.tight.base = 1, dsurf + ks, drgb + kr, dt + kt
                ctx.convert(j)
                got = np.empty(rcap, np.uint8)
                ctx.d2h(got, drgb + kr)
                want = np.full(rcap, synth.OUT_FILL, np.uint8)
                oracle.nv12_to_rgb24(surf, pitch, w, h, want, rp)
                assert np.array_equal(got, want), (w, h, pitch, ks, kr, fused)
                gt = np.empty(cap, np.uint8)
                ctx.d2h(gt, dt + kt)
                wt = np.full(cap, synth.OUT_FILL, np.uint8)
                if fused:
                    chk.nvdec_output_frame(surf, pitch, w, h, 1, wt, cap)
                assert np.array_equal(gt, wt), (w, h, pitch, "fused tight", fused)
                ctx.free(drgb), ctx.free(dt)
        ctx.free(dsurf), ctx.free(dtin)


@pytest.mark.parametrize("seed", range(3))
def test_random_intel_crops_against_oracle(ctx, seed):
    """Random MFX-style surfaces with crop rectangles, both directions (intel_dec / intel_enc rules)."""
    chk = oracle.best()
    rng = np.random.default_rng(2000 + seed)
    for _ in range(40):
        pitch = int(rng.integers(8, 40)) * 8
        rows = int(rng.integers(4, 40)) * 2
        cw = int(rng.integers(1, pitch))
        ch = int(rng.integers(1, rows))
        cx = int(rng.integers(0, pitch - cw + 1))
        cy = int(rng.integers(0, rows - ch + 1))
        c = dict(pitch=pitch, rows=rows, cx=cx, cy=cy, cw=cw, ch=ch)
        for fmt in (0, 1):
            cc = dict(c, op="inteldec", fmt=fmt)
            _, _, want = K.run_inteldec(chk, cc)
            assert np.array_equal(G.gpu_inteldec(ctx, cc), want), cc
        for i420 in (0, 1):
            cc = dict(c, op="intelenc", i420=i420)
            _, _, want = K.run_intelenc(chk, cc)
            assert np.array_equal(G.gpu_intelenc(ctx, cc), want), cc


@pytest.mark.parametrize("c", [c for c in GPU_CASES if c["op"] in ("nvdec", "nvenc", "intelenc", "inteldec")
                               and (c.get("w", c.get("cw", 0)) % 32 == 0) and c.get("kind", "random") == "random"], ids=K.case_id)
def test_ldg_stg_kernel_still_matches(ctx, c, monkeypatch):
    """16-byte-aligned geometries normally take the bulk-copy (cp.async.bulk) kernel; JMC_NO_BULK=1
    routes them through the LDG/STG vector kernel, which must give the same bytes."""
    monkeypatch.setenv("JMC_NO_BULK", "1")
    out = G.run_case_gpu(ctx, c)
    assert K.sha(out) == GOLD[K.case_id(c)]["sha256"]


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("c", [c for c in K.rgb_cases() if c["w"] % 16 == 0 and c["kind"] == "random"], ids=K.case_id)
def test_rgb_ldg_stg_kernel_still_matches(ctx, c, fused, monkeypatch):
    """Same for the RGB kernels: JMC_NO_BULK=1 selects the warp-per-task LDG/STG kernel."""
    monkeypatch.setenv("JMC_NO_BULK", "1")
    if fused:
        rgb, tight = G.gpu_rgb(ctx, c, fused=True)
        want = np.full(tight.size, synth.OUT_FILL, np.uint8)
        oracle.best().nvdec_output_frame(K.rgb_input(c), c["pitch"], c["w"], c["h"], 1, want, want.size)
        assert np.array_equal(tight, want)
    else:
        rgb = G.gpu_rgb(ctx, c)
    assert K.sha(rgb) == GOLD[K.case_id(c)]["sha256"]


@pytest.mark.parametrize("c", [c for c in K.rgb_cases() if c["w"] % 2 == 0 and c["kind"] in ("random", "gradient")], ids=K.case_id)
def test_rgb_bulk_loaded_kernel_everywhere_it_can_run(ctx, c, monkeypatch):
    """The bulk-loaded RGB kernel is the default only where it is the fastest (fused op; RGB24 from 1664 pixels wide);
    JMC_RGB_BULK_ALWAYS=1 sends every even width through it - aligned rows (copy-engine stores) and rows at odd
    addresses (re-aligned stores) - and the bytes must not change."""
    monkeypatch.setenv("JMC_RGB_BULK_ALWAYS", "1")
    assert K.sha(G.gpu_rgb(ctx, c)) == GOLD[K.case_id(c)]["sha256"]


@pytest.mark.parametrize("flat", ["default", "0", "1"])
@pytest.mark.parametrize("always", [False, True])
@pytest.mark.parametrize("seed", range(2))
def test_random_widths_rgb_on_aligned_surfaces(ctx, seed, always, flat, monkeypatch):
    """RGB24, fused I420+RGB24 and ARGB32 on decoder-style surfaces with arbitrary widths and skewed output buffers;
    every kernel that can serve them: bulk-loaded (JMC_RGB_BULK_ALWAYS), flattened (JMC_RGB_FLAT=1, the default when
    the width is not a multiple of 512) and the warp-per-segment kernel (JMC_RGB_FLAT=0)."""
    if always:
        monkeypatch.setenv("JMC_RGB_BULK_ALWAYS", "1")
    if flat != "default":
        monkeypatch.setenv("JMC_RGB_FLAT", flat)
    _random_widths_rgb(ctx, 4000 + seed)


@pytest.mark.parametrize("pairs", ["0", "1", "2", "3", "4", "7"])
@pytest.mark.parametrize("seed", range(2))
def test_rgb_bulk_tiles_of_several_row_pairs(ctx, seed, pairs, monkeypatch):
    """JMC_RGB_BULK_PAIRS=n: the re-aligning bulk-loaded kernel takes n row pairs per CTA (frames whose rows are not
    16-byte multiples; 0 = the library's choice: two for the fused op, RGB24 alone on the warp-per-task kernel; 1 =
    the one-pair kernel).  Heights that are not a multiple of 2n, odd heights, widths up to the point where n pairs
    no longer fit and the launch falls back to fewer; RGB24 alone and fused."""
    monkeypatch.setenv("JMC_RGB_BULK_PAIRS", pairs)
    _random_widths_rgb(ctx, 4100 + seed, hmax=70)


def _random_widths_rgb(ctx, seed, hmax=40):
    chk = oracle.best()
    rng = np.random.default_rng(seed)
    for it in range(24):
        w = int(rng.integers(2, 2600))
        if it % 4:
            w &= ~1
        h = int(rng.integers(2, hmax))
        pitch = ((w + 15) & ~15) + 16 * int(rng.integers(0, 3))
        kt, kr = int(rng.integers(0, 16)), int(rng.integers(0, 16))
        surf = synth.nv12_surface(w, h, pitch, 29, w * 7 + h)
        dsurf = ctx.upload(surf)
        for op in ("rgb", "fused", "argb"):
            bpp = 4 if op == "argb" else 3
            rcap, tcap = bpp * w * h + K.SLACK, w * h * 3 // 2 + K.SLACK
            drgb, dt = ctx.alloc(rcap + 64), ctx.alloc(tcap + 64)
            ctx.memset(drgb, synth.OUT_FILL, rcap + 64), ctx.memset(dt, synth.OUT_FILL, tcap + 64)
            j = ctx.job_argb(w, h, pitch, 4 * w) if op == "argb" else ctx.job_rgb(w, h, pitch, 3 * w, op == "fused")
            j.n_frames, j.surf.base, j.rgb.base = 1, dsurf, drgb + kr
            if op == "fused":
                j.tight.base = dt + kt
            ctx.convert(j)
            got = np.empty(rcap, np.uint8)
            ctx.d2h(got, drgb + kr)
            want = np.full(rcap, synth.OUT_FILL, np.uint8)
            (oracle.nv12_to_argb32 if op == "argb" else oracle.nv12_to_rgb24)(surf, pitch, w, h, want, bpp * w)
            assert np.array_equal(got, want), (op, w, h, pitch, kr)
            if op == "fused":
                gt = np.empty(tcap, np.uint8)
                ctx.d2h(gt, dt + kt)
                wt = np.full(tcap, synth.OUT_FILL, np.uint8)
                chk.nvdec_output_frame(surf, pitch, w, h, 1, wt, tcap)
                assert np.array_equal(gt, wt), (op, w, h, pitch, kt)
            ctx.free(drgb), ctx.free(dt)
        ctx.free(dsurf)


@pytest.mark.parametrize("seed", range(3))
def test_rgb24_to_nv12_random_geometries(ctx, seed):
    """Forward transform: random sizes, RGB rows at any address / pitch, surfaces aligned or not, batched (stride and
    pointer-list modes), against the oracle including every byte that must stay untouched."""
    rng = np.random.default_rng(5000 + seed)
    for it in range(16):
        w, h = int(rng.integers(1, 1400)), int(rng.integers(1, 30))
        n = int(rng.integers(1, 4))
        rp = 3 * w + int(rng.integers(0, 9))
        pitch = (((w + 15) & ~15) + 16 * int(rng.integers(0, 3))) if it % 3 else w + int(rng.integers(0, 7))
        kr, ks = int(rng.integers(0, 16)), (0 if it % 3 else int(rng.integers(0, 16)))
        rgb_bytes, surf_bytes = rp * h + 7, pitch * (h + (h >> 1) + 1)
        rgbs = [rng.integers(0, 256, rgb_bytes, dtype=np.uint8) for _ in range(n)]
        drgb, dsurf = ctx.alloc(n * rgb_bytes + 64), ctx.alloc(n * surf_bytes + 64)
        ctx.h2d(drgb + kr, np.concatenate(rgbs))
        ctx.memset(dsurf, synth.PAD_BYTE, n * surf_bytes + 64)
        j = ctx.job_rgb_to_nv12(w, h, rp, pitch)
        j.n_frames = n
        dl = None
        if it % 2:
            j.rgb.base, j.rgb.stride, j.surf.base, j.surf.stride = drgb + kr, rgb_bytes, dsurf + ks, surf_bytes
        else:
            dl = (G.device_ptr_array(ctx, [drgb + kr + f * rgb_bytes for f in range(n)]),
                  G.device_ptr_array(ctx, [dsurf + ks + f * surf_bytes for f in range(n)]))
            j.rgb.list, j.surf.list = dl
        ctx.convert(j)
        got = np.empty(n * surf_bytes, np.uint8)
        ctx.d2h(got, dsurf + ks)
        for f in range(n):
            want = np.full(surf_bytes, synth.PAD_BYTE, np.uint8)
            assert oracle.rgb24_to_nv12(rgbs[f], rp, w, h, want, pitch) == 0
            assert np.array_equal(got[f * surf_bytes:(f + 1) * surf_bytes], want), (w, h, rp, pitch, kr, ks, f)
        ctx.free(drgb), ctx.free(dsurf)
        if dl:
            ctx.free(dl[0]), ctx.free(dl[1])


@pytest.mark.parametrize("kernels", ["bulk_rows", "ldg_rows", "ldg_rows_always"])
@pytest.mark.parametrize("seed", range(4))
def test_random_widths_on_aligned_surfaces(ctx, seed, kernels, monkeypatch):
    """The row kernels (bulk-loaded tiles by default; JMC_NO_BULK=1: the warp-per-row LDG kernel): decoder-style
    surfaces (16-byte aligned, pitch a multiple of 16) with arbitrary widths/heights (several tiles deep) and
    arbitrarily skewed tight buffers, all four YUV ops, vs the oracle."""
    if kernels != "bulk_rows":
        monkeypatch.setenv("JMC_NO_BULK", "1")
    if kernels == "ldg_rows_always":
        monkeypatch.setenv("JMC_ROWS_ALWAYS", "1")
    chk = oracle.best()
    rng = np.random.default_rng(3000 + seed)
    for it in range(30):
        w = int(rng.integers(1, 2600))
        h = int(rng.integers(1, 24)) if it % 3 else int(rng.integers(24, 150))
        pitch = ((w + 15) & ~15) + 16 * int(rng.integers(0, 5))
        kt = int(rng.integers(0, 16))
        surf = synth.nv12_surface(w, h, pitch, 23, w * 31 + h)
        tight_in = synth.i420_frame(w, h, 24, w * 13 + h)
        cap = w * h * 3 // 2 + K.SLACK
        nsurf = pitch * (h * 3 // 2 + 1)
        dsurf = ctx.upload(surf)
        for fmt in (0, 1):
            dout = ctx.alloc(cap + 64)
            ctx.memset(dout, synth.OUT_FILL, cap + 64)
            j = ctx.job_nvdec(w, h, pitch, fmt)
            j.n_frames, j.surf.base, j.tight.base = 1, dsurf, dout + kt
            ctx.convert(j)
            got = np.empty(cap, np.uint8)
            ctx.d2h(got, dout + kt)
            want = np.full(cap, synth.OUT_FILL, np.uint8)
            chk.nvdec_output_frame(surf, pitch, w, h, fmt, want, cap)
            assert np.array_equal(got, want), (w, h, pitch, kt, fmt)
            ctx.free(dout)
        dtin = ctx.alloc(tight_in.size + 64)
        ctx.h2d(dtin + kt, tight_in)
        for code in (0x1, 0x10):
            ds = ctx.alloc(nsurf)
            ctx.memset(ds, synth.PAD_BYTE, nsurf)
            j = ctx.job_nvenc(w, h, pitch, code)
            j.n_frames, j.surf.base, j.tight.base = 1, ds, dtin + kt
            ctx.convert(j)
            got = np.empty(nsurf, np.uint8)
            ctx.d2h(got, ds)
            want = np.full(nsurf, synth.PAD_BYTE, np.uint8)
            oracle.nvenc_upload(tight_in, code, w, h, want, pitch)
            assert np.array_equal(got, want), (w, h, pitch, kt, hex(code))
            ctx.free(ds)
        ctx.free(dsurf), ctx.free(dtin)


@pytest.mark.parametrize("c", [c for c in GPU_CASES if c["op"] in ("nvdec", "nvenc") and c.get("kind", "random") == "random"
                               and c["w"] % 16 != 0], ids=K.case_id)
def test_any_alignment_kernel_still_matches(ctx, c, monkeypatch):
    """Widths that are not multiples of 16 normally take the row-staged kernel when the surface is aligned;
    JMC_NO_ROWS=1 keeps them on the any-alignment vector kernel, which must give the same bytes."""
    monkeypatch.setenv("JMC_NO_ROWS", "1")
    out = G.run_case_gpu(ctx, c)
    assert K.sha(out) == GOLD[K.case_id(c)]["sha256"]


@pytest.mark.parametrize("geom", [(854, 480, 1024), (1366, 768, 1536), (1080, 1920, 1088), (427, 241, 512), (2562, 38, 2688), (255, 17, 256), (257, 16, 272)])
def test_pack_batches_of_odd_widths(ctx, geom):
    """Encode direction, widths that are not multiples of 16, several frames per launch in stride mode, tight frames at
    aligned and skewed addresses: intel_enc.cpp:291-307,366-380 / nv_enc.cpp:1029-1081 reproduced, padding left alone."""
    w, h, pitch = geom
    n = 5
    tight_bytes = w * h * 3 // 2
    tstride = (tight_bytes + 15 + 48) & ~15                 # frames a multiple of 16 bytes apart, with slack between them
    nsurf = pitch * (h * 3 // 2 + 1)
    for code in (0x1, 0x10):
        for skew in (0, 5):
            frames = [synth.i420_frame(w, h, 25, f * 7 + w) for f in range(n)]
            hin = np.full(n * tstride + 64, 0x77, np.uint8)
            for f in range(n):
                hin[skew + f * tstride: skew + f * tstride + tight_bytes] = frames[f]
            dt = ctx.upload(hin)
            ds = ctx.alloc(n * nsurf)
            ctx.memset(ds, synth.PAD_BYTE, n * nsurf)
            j = ctx.job_nvenc(w, h, pitch, code)
            j.n_frames, j.surf.base, j.surf.stride, j.tight.base, j.tight.stride = n, ds, nsurf, dt + skew, tstride
            ctx.convert(j)
            got = np.empty(n * nsurf, np.uint8)
            ctx.d2h(got, ds)
            for f in range(n):
                want = np.full(nsurf, synth.PAD_BYTE, np.uint8)
                oracle.nvenc_upload(frames[f], code, w, h, want, pitch)
                assert np.array_equal(got[f * nsurf:(f + 1) * nsurf], want), (geom, hex(code), skew, f)
            ctx.free(dt), ctx.free(ds)


def _zero_row_end_padding(want, w, h, pitch, nv12_chroma):
    """What JMC_JOB_PAD_ZERO allows: behind the last byte of every row the kernel wrote, zeros up to the next 16-byte boundary."""
    rows = want.reshape(-1, pitch)
    lw = w
    rows[:h, lw:(lw + 15) & ~15] = 0
    cwb = w if nv12_chroma else 2 * (w >> 1)                 # bytes per chroma row
    rows[h:h + (h >> 1), cwb:(cwb + 15) & ~15] = 0
    return want


@pytest.mark.parametrize("geom", [(854, 480, 1024), (1366, 768, 1536), (1080, 1920, 1088), (427, 241, 512), (255, 17, 256), (250, 16, 256), (1919, 1079, 2048)])
def test_pad_zero_flag(ctx, geom):
    """JMC_JOB_PAD_ZERO (encode ops): the active bytes are the reference's, the padding behind every row end is zero up to
    the next 16-byte boundary and untouched beyond -- for the I420 / NV12 pack and for RGB24 -> NV12."""
    import jmcodec_b200 as J
    w, h, pitch = geom
    nsurf = pitch * (h * 3 // 2 + 1)
    PAD_ZERO = 4
    tight = synth.i420_frame(w, h, 26, w + h)
    for code in (0x1, 0x10):
        dt, ds = ctx.upload(tight), ctx.alloc(nsurf)
        ctx.memset(ds, synth.PAD_BYTE, nsurf)
        j = ctx.job_nvenc(w, h, pitch, code)
        j.n_frames, j.surf.base, j.tight.base, j.flags = 1, ds, dt, PAD_ZERO
        ctx.convert(j)
        got = np.empty(nsurf, np.uint8)
        ctx.d2h(got, ds)
        want = np.full(nsurf, synth.PAD_BYTE, np.uint8)
        oracle.nvenc_upload(tight, code, w, h, want, pitch)
        if w % 16:                                           # widths that are multiples of 16 take the bulk kernel: nothing to zero
            _zero_row_end_padding(want, w, h, pitch, code == 0x1)
        assert np.array_equal(got, want), (geom, hex(code))
        ctx.free(dt), ctx.free(ds)
    if w >= 2 and h >= 2:
        rgb = synth.random_bytes(3 * w * h, synth.frame_key(27, w))
        dr, ds = ctx.upload(rgb), ctx.alloc(nsurf)
        ctx.memset(ds, synth.PAD_BYTE, nsurf)
        j = ctx.job_rgb_to_nv12(w, h, 3 * w, pitch)
        j.n_frames, j.rgb.base, j.surf.base, j.flags = 1, dr, ds, PAD_ZERO
        ctx.convert(j)
        got = np.empty(nsurf, np.uint8)
        ctx.d2h(got, ds)
        want = np.full(nsurf, synth.PAD_BYTE, np.uint8)
        oracle.rgb24_to_nv12(rgb, 3 * w, w, h, want, pitch)
        _zero_row_end_padding(want, w, h, pitch, False)
        assert np.array_equal(got, want), (geom, "rgb2nv12")
        ctx.free(dr), ctx.free(ds)
