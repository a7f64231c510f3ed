"""GPU tests of the drop-in API (jm_nvdec_*, jm_nvenc_*), the delivery pipeline, and full-size
size-independent properties.  Written to read like a port of test_nv_dec.cpp's call sequence."""
import ctypes as C

import numpy as np
import pytest

import oracle
from jmcodec_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import jmcodec_b200
    return jmcodec_b200


@pytest.fixture(scope="module")
def ctx(J):
    c = J.Ctx(0)
    yield c
    c.close()


# --------------------------------------------------------------------------------------------
# jm_nvdec_*: create / init / decode_frame / output_frame / deinit  (test_nv_dec.cpp:163-259)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("out_fmt", [0, 1])
@pytest.mark.parametrize("geom", [(1920, 1080, 2048), (1919, 1079, 2048), (64, 48, 64), (2, 2, 256)])
def test_nvdec_api_matches_reference_function(J, out_fmt, geom):
    w, h, pitch = geom
    chk = oracle.best()
    assert J.load().jm_nvdec_is_hw_support()
    dec = J.NvDec()
    assert dec.init(J.NvDec.CODEC_RAW_NV12, out_fmt) == 0
    need = w * h * 3 // 2
    cap = need + 64
    out = np.full(cap, synth.OUT_FILL, np.uint8)
    # nothing decoded yet: the reference returns -1 (nv_dec.cpp:757-758)
    assert dec.output_frame(out, cap) == (-1, cap)
    frames = 0
    for f in range(6):
        s = synth.nv12_surface(w, h, pitch, 11, f)
        pkt = J.NvDec.raw_packet(s, w, h, pitch)
        r, got = dec.decode_frame(pkt)
        assert (r, got) == (0, 1)
        assert dec.stream_info() == (w, h)
        # short buffer: -2 and *out_len untouched (nv_dec.cpp:773-774)
        assert dec.output_frame(out, need - 1) == (-2, need - 1)
        out[:] = synth.OUT_FILL
        want = out.copy()
        assert dec.output_frame(out, cap) == chk.nvdec_output_frame(s, pitch, w, h, out_fmt, want, cap) == (need, need)
        assert np.array_equal(out, want)
        frames += 1
    # flush: len 0 => end of stream; no frame left => is_exit + info block (nv_dec.cpp:389-392,460-466)
    assert not dec.is_exit()
    assert dec.decode_frame(None, 0) == (0, 0)
    assert dec.is_exit()
    info = dec.show_dec_info()
    assert f"Frame Count:\t{frames}" in info and f"Display:\t{w} x {h}" in info
    assert ("NV12" if out_fmt == 0 else "YV12") in info
    assert dec.deinit() == 0


def test_nvdec_api_queue_and_pinned_and_device_packets(J, ctx):
    """Several packets before fetching (display queue, one frame out per call), pinned out_buf (direct
    DMA) and device-pointer packets (what cuvidMapVideoFrame yields)."""
    w, h, pitch = 320, 180, 512
    need = w * h * 3 // 2
    chk = oracle.best()
    dec = J.NvDec(0)
    assert dec.init(J.NvDec.CODEC_RAW_NV12, 1) == 0
    pinned = dec.alloc_host(need)
    out = np.ctypeslib.as_array((C.c_uint8 * need).from_address(pinned))
    surfs = [synth.nv12_surface(w, h, pitch, 12, f) for f in range(4)]
    dptr = ctx.upload(surfs[3])
    assert dec.decode_frame(J.NvDec.raw_packet(surfs[0], w, h, pitch)) == (0, 1)       # frame 0 ready
    want = np.empty(need, np.uint8)
    for nxt, cur in ((1, 0), (2, 1)):
        # decode the next packet BEFORE fetching: the fetched frame is the newest announced one
        assert dec.decode_frame(J.NvDec.raw_packet(surfs[nxt], w, h, pitch)) == (0, 1)
        assert dec.output_frame(out, need) == (need, need)
        chk.nvdec_output_frame(surfs[nxt], pitch, w, h, 1, want, need)
        assert np.array_equal(out, want)
    assert dec.decode_frame(J.NvDec.raw_packet(None, w, h, pitch, device_ptr=dptr)) == (0, 1)
    assert dec.output_frame(out, need) == (need, need)
    chk.nvdec_output_frame(surfs[3], pitch, w, h, 1, want, need)
    assert np.array_equal(out, want)
    # garbage packet: consumed, no frame, still returns 0 like the reference's swallowed errors
    assert dec.decode_frame(np.zeros(64, np.uint8)) == (0, 0)
    dec.free_host(pinned)
    dec.deinit()
    ctx.free(dptr)


def test_nvdec_bitstream_codecs_fail_loudly(J):
    dec = J.NvDec()
    assert dec.init(0, 1) != 0           # H.264 needs the NVDEC parser front-end: no silent fallback
    assert dec.decode_frame(np.zeros(16, np.uint8)) == (0, 0)
    dec.deinit()


# --------------------------------------------------------------------------------------------
# jm_nvenc_*: encoder input path (surface-only on B200)
# --------------------------------------------------------------------------------------------
def test_nvenc_needs_surface_only_mode(J):
    enc = J.NvEnc()
    assert enc.init(64, 64, J.NvEnc.FMT_NV12, codec_id=0) == 1      # NV_ENC_ERR_NO_ENCODE_DEVICE
    enc.deinit()


@pytest.mark.parametrize("fmt", ["nv12", "yv12", "argb"])
@pytest.mark.parametrize("geom", [(1920, 1080), (1280, 720), (354, 290), (64, 64)])
def test_nvenc_api_upload(J, ctx, fmt, geom):
    w, h = geom
    code = {"nv12": J.NvEnc.FMT_NV12, "yv12": J.NvEnc.FMT_YV12, "argb": J.NvEnc.FMT_ARGB}[fmt]
    enc = J.NvEnc(0)
    assert enc.init(w, h, code) == 0
    assert enc.peek_surface()[0] == -1
    assert enc.enc_frame(None, 0) == (0, 0)                          # EOS before anything: fine
    n_in = w * h * 4 if fmt == "argb" else w * h * 3 // 2
    for f in range(3):
        src = synth.random_bytes(n_in, synth.frame_key(13, f))
        assert enc.enc_frame(src) == (0, 0)                          # uploaded + packed, no packet (no NVENC)
        assert enc.get_bitstream() == (-1, 0)
        r, dptr, pitch, rows = enc.peek_surface()
        assert r == 0 and pitch >= (w * 4 if fmt == "argb" else w)
        got = np.empty(pitch * rows, np.uint8)
        ctx.d2h(got, dptr)
        want = np.zeros(pitch * rows, np.uint8)                      # surfaces are zeroed at init
        if fmt == "argb":
            # the reference's flat copy ignores the pitch (nv_enc.cpp:1096); we honour it
            want.reshape(rows, pitch)[:, :w * 4] = src.reshape(h, w * 4)
        else:
            assert oracle.nvenc_upload(src, code, w, h, want, pitch) == 0
        assert np.array_equal(got, want)
    # surfaces stay locked until released, exactly 10 of them (MAX_NV_ENC_FRAME_NUM, nv_enc.cpp:916-927,90-93)
    src = synth.random_bytes(n_in, 1)
    for _ in range(7):
        assert enc.enc_frame(src)[0] == 0
    assert enc.enc_frame(src)[0] == -1
    assert enc.release_surface() == 0
    assert enc.enc_frame(src)[0] == 0
    enc.deinit()


# --------------------------------------------------------------------------------------------
# host-delivery pipeline
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op", ["i420", "nv12", "rgb", "fused", "pack", "argb", "rgb2nv12"])
def test_pipeline_batches(J, ctx, op):
    w, h, pitch, batch, nb = 640, 360, 768, 5, 7
    surf_bytes, tight_bytes, rgb_bytes = pitch * h * 3 // 2, w * h * 3 // 2, 3 * w * h
    chk = oracle.best()
    if op == "pack":
        shape = ctx.job_nvenc(w, h, pitch, 0x10)
        in_bytes, out_bytes = tight_bytes, surf_bytes
    elif op == "rgb2nv12":
        shape = ctx.job_rgb_to_nv12(w, h, 3 * w, pitch)
        in_bytes, out_bytes = rgb_bytes, surf_bytes
    elif op == "argb":
        shape = ctx.job_argb(w, h, pitch, 4 * w)
        in_bytes, out_bytes = surf_bytes, 4 * w * h
    elif op in ("rgb", "fused"):
        shape = ctx.job_rgb(w, h, pitch, 3 * w, op == "fused")
        in_bytes, out_bytes = surf_bytes, (rgb_bytes if op == "rgb" else tight_bytes)
    else:
        shape = ctx.job_nvdec(w, h, pitch, 1 if op == "i420" else 0)
        in_bytes, out_bytes = surf_bytes, tight_bytes
    shape.n_frames = batch
    pipe = J.Pipeline(ctx, shape, surf_bytes, depth=3)
    hin = ctx.alloc_host(nb * batch * in_bytes)
    hout = ctx.alloc_host(nb * batch * out_bytes)
    hout2 = ctx.alloc_host(nb * batch * rgb_bytes) if op == "fused" else None
    frames = []
    for i in range(nb * batch):
        if op == "rgb2nv12":
            f = synth.random_bytes(rgb_bytes, synth.frame_key(14, i) ^ 0x3C3C3C3C)
        else:
            f = synth.i420_frame(w, h, 14, i) if op == "pack" else synth.nv12_surface(w, h, pitch, 14, i)
        frames.append(f)
        hin.array[i * in_bytes:(i + 1) * in_bytes] = f
    for b in range(nb):
        n = batch if b != nb - 1 else batch - 2             # ragged last batch
        pipe.submit(hin.array[b * batch * in_bytes:], hout.array[b * batch * out_bytes:], n,
                    host_out2=hout2.array[b * batch * rgb_bytes:] if hout2 else None)
    pipe.drain()
    total = nb * batch - 2
    # decode-side uploads skip the pitch padding (2-D copy): only width of every pitch bytes cross PCIe
    assert pipe.h2d_bytes == (total * in_bytes if op in ("pack", "rgb2nv12") else total * (h * 3 // 2) * w)
    assert pipe.d2h_bytes == total * (out_bytes + (rgb_bytes if op == "fused" else 0))
    for i in range(total):
        got = hout.array[i * out_bytes:(i + 1) * out_bytes]
        if op == "pack":
            want = np.zeros(out_bytes, np.uint8)
            oracle.nvenc_upload(frames[i], 0x10, w, h, want, pitch)
        elif op == "rgb2nv12":
            want = np.zeros(out_bytes, np.uint8)                     # pipeline surfaces start zeroed; padding stays 0
            oracle.rgb24_to_nv12(frames[i], 3 * w, w, h, want, pitch)
        elif op == "rgb":
            want = np.empty(out_bytes, np.uint8)
            oracle.nv12_to_rgb24(frames[i], pitch, w, h, want, 3 * w)
        elif op == "argb":
            want = np.empty(out_bytes, np.uint8)
            oracle.nv12_to_argb32(frames[i], pitch, w, h, want, 4 * w)
        else:
            want = np.empty(out_bytes, np.uint8)
            chk.nvdec_output_frame(frames[i], pitch, w, h, 0 if op == "nv12" else 1, want, out_bytes)
        assert np.array_equal(got, want), f"frame {i}"
        if op == "fused":
            want2 = np.empty(rgb_bytes, np.uint8)
            oracle.nv12_to_rgb24(frames[i], pitch, w, h, want2, 3 * w)
            assert np.array_equal(hout2.array[i * rgb_bytes:(i + 1) * rgb_bytes], want2)
    pipe.close()
    for b in (hin, hout, hout2):
        if b:
            b.free()


# --------------------------------------------------------------------------------------------
# BASELINE.json sizes: size-independent properties + sampled oracle comparison
# --------------------------------------------------------------------------------------------
def _reference_outputs(chk, distinct, pitch, w, h, out_fmt):
    """The reference function on every distinct surface, on all host cores (one handle per thread)."""
    import os
    S = np.stack(distinct)
    out = np.zeros((len(distinct), w * h * 3 // 2), np.uint8)
    chk.nvdec_run(S, out, pitch, w, h, out_fmt, len(distinct), min(len(distinct), len(os.sched_getaffinity(0))))
    return out


@pytest.mark.parametrize("geom", [(1920, 1080, 2048, 300, 32), (3840, 2160, 4096, 64, 16)])
def test_full_size_every_frame_bit_exact_and_round_trip(ctx, geom):
    """BASELINE.json configs 2/3 sizes: EVERY frame of the batch equals the reference function's output for its
    surface ("bit-exact vs config 1 output for every frame"); then NV12 -(de-interleave)-> I420
    -(pack)-> NV12 must reproduce every active byte of all frames and leave the destination padding untouched."""
    w, h, pitch, n, nd = geom
    surf_bytes, tight_bytes = pitch * h * 3 // 2, w * h * 3 // 2
    chk = oracle.best()
    base = [synth.nv12_surface(w, h, pitch, 15, f) for f in range(nd)]
    want = _reference_outputs(chk, base, pitch, w, h, 1)
    host = np.concatenate([base[f % nd] for f in range(n)])
    dsurf = ctx.upload(host)
    dtight = ctx.alloc(n * tight_bytes)
    dback = ctx.alloc(n * surf_bytes)
    ctx.memset(dtight, synth.OUT_FILL, n * tight_bytes)
    ctx.memset(dback, synth.PAD_BYTE, n * surf_bytes)
    j = ctx.job_nvdec(w, h, pitch, 1)
    j.n_frames, j.surf.base, j.surf.stride, j.tight.base, j.tight.stride = n, dsurf, surf_bytes, dtight, tight_bytes
    ctx.convert(j)
    tight = np.empty(n * tight_bytes, np.uint8)
    ctx.d2h(tight, dtight)
    tight = tight.reshape(n, tight_bytes)
    for f in range(n):
        assert np.array_equal(tight[f], want[f % nd]), f"frame {f}"
    k = ctx.job_nvenc(w, h, pitch, 0x10)
    # nv_enc places V at y_len*5/4 == w*h + (w/2)*(h/2) for these even sizes, same layout as the I420 above
    k.n_frames, k.surf.base, k.surf.stride, k.tight.base, k.tight.stride = n, dback, surf_bytes, dtight, tight_bytes
    ctx.convert(k)
    back = np.empty(n * surf_bytes, np.uint8)
    ctx.d2h(back, dback)
    assert np.array_equal(back, host)            # synthetic padding is 0xCD on both sides, so whole surfaces match
    for d in (dsurf, dtight, dback):
        ctx.free(d)


@pytest.mark.parametrize("geom", [(1920, 1080, 2048, 300, 30, 32), (3840, 2160, 4096, 64, 16, 16)])
def test_full_size_pipeline_every_frame_bit_exact(J, ctx, geom):
    """The same sizes END TO END: pinned host surfaces -> H2D -> one launch per sub-batch -> pinned D2H
    (jmc_pipeline_*, what bench.py's e2e leg times); every delivered frame equals the reference function's output."""
    w, h, pitch, n, sub, nd = geom
    surf_bytes, tight_bytes = pitch * h * 3 // 2, w * h * 3 // 2
    chk = oracle.best()
    base = [synth.nv12_surface(w, h, pitch, 17, f) for f in range(nd)]
    want = _reference_outputs(chk, base, pitch, w, h, 1)
    hin, hout = ctx.alloc_host(n * surf_bytes), ctx.alloc_host(n * tight_bytes)
    for f in range(n):
        hin.array[f * surf_bytes:(f + 1) * surf_bytes] = base[f % nd]
    hout.array[:] = synth.OUT_FILL
    shape = ctx.job_nvdec(w, h, pitch, 1)
    shape.n_frames = sub
    pipe = J.Pipeline(ctx, shape, surf_bytes, depth=3)
    for b in range(0, n, sub):
        cnt = min(sub, n - b)
        pipe.submit(hin.array[b * surf_bytes:], hout.array[b * tight_bytes:], cnt)
    pipe.drain()
    got = hout.array.reshape(n, tight_bytes)
    for f in range(n):
        assert np.array_equal(got[f], want[f % nd]), f"frame {f}"
    pipe.close()
    hin.free(), hout.free()


# --------------------------------------------------------------------------------------------
# tools/jm_streams: the whole-box driver written in C++ against include/jmc_cuda.h only
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["e2e", "device"])
def test_cpp_streams_driver(mode):
    import json
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "jm_streams")
    if not os.path.exists(exe):
        pytest.skip("tools/jm_streams not built")
    p = subprocess.run([exe, "--gpus", "1", "--streams", "5", "--frames", "24", "--batch", "10", "--width", "640", "--height", "360",
                        "--pitch", "768", "--mode", mode], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    d = json.loads(p.stdout)
    assert d["mode"] == mode and d["n_gpus"] == 1 and d["frames"] == 5 * 24
    assert d["per_gpu"][0]["streams"] == 5 and d["frames_per_s"] > 0


def test_many_handles_many_threads(J):
    """One handle per host thread, several handles per process (SURVEY.md 8b "Threading"): four threads
    decode their own streams concurrently on one GPU; every frame must still be bit-exact."""
    import threading
    w, h, pitch, n = 352, 288, 512, 12
    need = w * h * 3 // 2
    chk = oracle.best()
    errors = []

    def worker(tid):
        try:
            dec = J.NvDec(0)
            assert dec.init(J.NvDec.CODEC_RAW_NV12, tid % 2) == 0
            out = np.empty(need, np.uint8)
            want = np.empty(need, np.uint8)
            for f in range(n):
                s = synth.nv12_surface(w, h, pitch, 40 + tid, f)
                assert dec.decode_frame(J.NvDec.raw_packet(s, w, h, pitch)) == (0, 1)
                assert dec.output_frame(out, need) == (need, need)
                chk.nvdec_output_frame(s, pitch, w, h, tid % 2, want, need)
                assert np.array_equal(out, want), (tid, f)
            dec.deinit()
        except Exception as e:      # noqa: BLE001
            errors.append((tid, repr(e)))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_8k_frames_take_the_bulk_path_with_smaller_tiles(ctx):
    """7680x4320: rows per tile shrink to fit shared memory (bulk planes) and 10*w bytes still fit (bulk RGB)."""
    w, h, pitch = 7680, 4320, 7680
    chk = oracle.best()
    s = synth.nv12_surface(w, h, pitch, 50, 0)
    need = w * h * 3 // 2
    ds = ctx.upload(s)
    dt = ctx.alloc(need)
    drgb = ctx.alloc(3 * w * h)
    j = ctx.job_rgb(w, h, pitch, 3 * w, True)
    j.n_frames, j.surf.base, j.tight.base, j.rgb.base = 1, ds, dt, drgb
    ctx.convert(j)
    got = np.empty(need, np.uint8)
    ctx.d2h(got, dt)
    want = np.empty(need, np.uint8)
    chk.nvdec_output_frame(s, pitch, w, h, 1, want, need)
    assert np.array_equal(got, want)
    rgb = np.empty(3 * w * h, np.uint8)
    ctx.d2h(rgb, drgb)
    wrgb = np.empty(3 * w * h, np.uint8)
    oracle.nv12_to_rgb24(s, pitch, w, h, wrgb, 3 * w)
    assert np.array_equal(rgb, wrgb)
    ctx.memset(dt, 0, need)
    j2 = ctx.job_nvdec(w, h, pitch, 1)
    j2.n_frames, j2.surf.base, j2.tight.base = 1, ds, dt
    ctx.convert(j2)
    ctx.d2h(got, dt)
    assert np.array_equal(got, want)
    # and back: pack the I420 frame into a fresh surface
    dback = ctx.alloc(s.size)
    k = ctx.job_nvenc(w, h, pitch, 0x10)
    k.n_frames, k.surf.base, k.tight.base = 1, dback, dt
    ctx.convert(k)
    back = np.empty(s.size, np.uint8)
    ctx.d2h(back, dback)
    assert np.array_equal(back, s)
    for d in (ds, dt, drgb, dback):
        ctx.free(d)


def test_handles_on_two_devices_interleaved(J):
    """Handles bound to different GPUs, driven alternately from ONE thread and then from two threads:
    every API entry must re-bind its own device (the reference pushes/pops its context per call)."""
    import threading
    if J.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w, h, pitch = 320, 200, 384
    need = w * h * 3 // 2
    chk = oracle.best()
    decs = [J.NvDec(d) for d in (0, 1)]
    for d in decs:
        assert d.init(J.NvDec.CODEC_RAW_NV12, 1) == 0
    enc = J.NvEnc(1)
    assert enc.init(w, h, J.NvEnc.FMT_YV12) == 0
    out, want = np.empty(need, np.uint8), np.empty(need, np.uint8)
    for f in range(6):
        dec = decs[f % 2]
        s = synth.nv12_surface(w, h, pitch, 60, f)
        assert dec.decode_frame(J.NvDec.raw_packet(s, w, h, pitch)) == (0, 1)
        assert enc.enc_frame(synth.i420_frame(w, h, 61, f))[0] == 0        # touches device 1 in between
        enc.release_surface()
        assert dec.output_frame(out, need) == (need, need)
        chk.nvdec_output_frame(s, pitch, w, h, 1, want, need)
        assert np.array_equal(out, want), f

    errors = []

    def worker(dev):
        try:
            o, wn = np.empty(need, np.uint8), np.empty(need, np.uint8)
            for f in range(8):
                s = synth.nv12_surface(w, h, pitch, 62 + dev, f)
                assert decs[dev].decode_frame(J.NvDec.raw_packet(s, w, h, pitch)) == (0, 1)
                assert decs[dev].output_frame(o, need) == (need, need)
                chk.nvdec_output_frame(s, pitch, w, h, 1, wn, need)
                assert np.array_equal(o, wn), (dev, f)
        except Exception as e:      # noqa: BLE001
            errors.append((dev, repr(e)))

    ts = [threading.Thread(target=worker, args=(d,)) for d in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for d in decs:
        d.deinit()
    enc.deinit()


def test_full_size_rgb_batch_every_frame(ctx):
    """64 x 4K through the display ops in one launch each: every frame of both kernels against the oracle
    (RGB24, fused RGB24 + I420)."""
    w, h, pitch, n = 3840, 2160, 4096, 64
    surf_bytes, tight_bytes, rgb_bytes = pitch * h * 3 // 2, w * h * 3 // 2, 3 * w * h
    base = [synth.nv12_surface(w, h, pitch, 16, f) for f in range(4)]
    dsurf = ctx.alloc(n * surf_bytes)
    for f in range(n):
        ctx.h2d(dsurf + f * surf_bytes, base[f % 4])
    drgb1, drgb2, dt = ctx.alloc(n * rgb_bytes), ctx.alloc(n * rgb_bytes), ctx.alloc(n * tight_bytes)
    j = ctx.job_rgb(w, h, pitch, 3 * w, False)
    j.n_frames, j.surf.base, j.surf.stride, j.rgb.base, j.rgb.stride = n, dsurf, surf_bytes, drgb1, rgb_bytes
    ctx.convert(j)
    k = ctx.job_rgb(w, h, pitch, 3 * w, True)
    k.n_frames, k.surf.base, k.surf.stride, k.rgb.base, k.rgb.stride = n, dsurf, surf_bytes, drgb2, rgb_bytes
    k.tight.base, k.tight.stride = dt, tight_bytes
    ctx.convert(k)
    a, b = np.empty(rgb_bytes, np.uint8), np.empty(rgb_bytes, np.uint8)
    wants = []
    for s4 in base:                                   # the oracle on every distinct surface
        wr, wt = np.empty(rgb_bytes, np.uint8), np.empty(tight_bytes, np.uint8)
        oracle.nv12_to_rgb24(s4, pitch, w, h, wr, 3 * w)
        oracle.best().nvdec_output_frame(s4, pitch, w, h, 1, wt, tight_bytes)
        wants.append((wr, wt))
    tight = np.empty(tight_bytes, np.uint8)
    for f in range(n):                                # EVERY frame of the batch, both kernels
        ctx.d2h(a, drgb1 + f * rgb_bytes)
        ctx.d2h(b, drgb2 + f * rgb_bytes)
        assert np.array_equal(a, wants[f % 4][0]), f"RGB24 frame {f}"
        assert np.array_equal(b, wants[f % 4][0]), f"fused RGB24 frame {f}"
        ctx.d2h(tight, dt + f * tight_bytes)
        assert np.array_equal(tight, wants[f % 4][1]), f"fused I420 frame {f}"
    for d in (dsurf, drgb1, drgb2, dt):
        ctx.free(d)


def test_full_size_transcode_loop_property(ctx):
    """16 x 4K: NV12 -> RGB24 -> NV12 on the device.  The chain is lossy in general, but for grey frames in the legal
    range (Y 16..235, U = V = 128) it must return chroma exactly and luma within 1 (a property of the two integer
    transforms, checked on the oracle for every Y); sampled frames of the second hop are compared with the oracle."""
    w, h, pitch, n = 3840, 2160, 4096, 16
    surf_bytes, rgb_bytes = pitch * h * 3 // 2, 3 * w * h
    rng = np.random.default_rng(77)
    base = []
    for f in range(2):
        s = np.full((h * 3 // 2, pitch), synth.PAD_BYTE, np.uint8)
        s[:h, :w] = rng.integers(16, 236, (h, w), dtype=np.uint8)
        s[h:, :w] = 128
        base.append(s.reshape(-1))
    dsurf, drgb, dback = ctx.alloc(n * surf_bytes), ctx.alloc(n * rgb_bytes), ctx.alloc(n * surf_bytes)
    for f in range(n):
        ctx.h2d(dsurf + f * surf_bytes, base[f % 2])
    ctx.memset(dback, synth.PAD_BYTE, n * surf_bytes)
    j = ctx.job_rgb(w, h, pitch, 3 * w, False)
    j.n_frames, j.surf.base, j.surf.stride, j.rgb.base, j.rgb.stride = n, dsurf, surf_bytes, drgb, rgb_bytes
    ctx.convert(j)
    k = ctx.job_rgb_to_nv12(w, h, 3 * w, pitch)
    k.n_frames, k.rgb.base, k.rgb.stride, k.surf.base, k.surf.stride = n, drgb, rgb_bytes, dback, surf_bytes
    ctx.convert(k)
    back, rgb = np.empty(surf_bytes, np.uint8), np.empty(rgb_bytes, np.uint8)
    want = np.empty(surf_bytes, np.uint8)
    for f in (0, 1, 7, 15):
        ctx.d2h(back, dback + f * surf_bytes)
        a, b = back.reshape(-1, pitch), base[f % 2].reshape(-1, pitch)
        assert np.abs(a[:h, :w].astype(np.int16) - b[:h, :w].astype(np.int16)).max() <= 1
        assert (a[h:, :w] == 128).all() and (a[:, w:] == synth.PAD_BYTE).all()
        ctx.d2h(rgb, drgb + f * rgb_bytes)
        want[:] = synth.PAD_BYTE
        assert oracle.rgb24_to_nv12(rgb, 3 * w, w, h, want, pitch) == 0
        assert np.array_equal(back, want)
    for d in (dsurf, drgb, dback):
        ctx.free(d)


def test_pipeline_argument_errors_and_depth_one(J, ctx):
    w, h, pitch, batch = 64, 32, 64, 4
    surf_bytes, tight_bytes = pitch * h * 3 // 2, w * h * 3 // 2
    shape = ctx.job_nvdec(w, h, pitch, 1)
    shape.n_frames = batch
    with pytest.raises(J.JmcError):
        J.Pipeline(ctx, shape, 0, depth=2)                  # surf_bytes == 0
    with pytest.raises(J.JmcError):
        J.Pipeline(ctx, shape, surf_bytes, depth=0)
    pipe = J.Pipeline(ctx, shape, surf_bytes, depth=1)      # a single slot: every submit waits for the previous one
    hin, hout = ctx.alloc_host(batch * surf_bytes), ctx.alloc_host(batch * tight_bytes)
    with pytest.raises(J.JmcError):
        pipe.submit(hin.array, hout.array, batch + 1)       # more frames than the slot holds
    with pytest.raises(J.JmcError):
        pipe.submit(None, hout.array, batch)                # neither host nor device input
    with pytest.raises(J.JmcError):
        pipe.wait(5)
    chk = oracle.best()
    want = np.empty(tight_bytes, np.uint8)
    for rep in range(5):
        frames = [synth.nv12_surface(w, h, pitch, 70 + rep, f) for f in range(batch)]
        hin.array[:] = np.concatenate(frames)
        slot = pipe.submit(hin.array, hout.array, batch)
        assert slot == 0
        pipe.wait(slot)
        for f in range(batch):
            chk.nvdec_output_frame(frames[f], pitch, w, h, 1, want, tight_bytes)
            assert np.array_equal(hout.array[f * tight_bytes:(f + 1) * tight_bytes], want)
    pipe.close()
    hin.free(), hout.free()


def test_raw_packet_fuzz_never_crashes(J):
    """Malformed packets are dropped (got_frame 0), never crash, and do not wedge the handle."""
    rng = np.random.default_rng(9)
    dec = J.NvDec(0)
    assert dec.init(J.NvDec.CODEC_RAW_NV12, 1) == 0
    w, h, pitch = 32, 16, 32
    good = J.NvDec.raw_packet(synth.nv12_surface(w, h, pitch, 80, 0), w, h, pitch)
    for i in range(200):
        bad = good.copy()
        k = int(rng.integers(0, 4))
        if k == 0:
            bad = bad[:int(rng.integers(0, bad.size))]                         # truncated
        elif k == 1:
            bad[int(rng.integers(0, 32))] ^= int(rng.integers(1, 256))         # corrupted header byte
        elif k == 2:
            bad = rng.integers(0, 256, size=int(rng.integers(1, 200)), dtype=np.uint8)
        else:
            bad[4:16] = np.array([-5, 7, 3], dtype="<i4").view(np.uint8)       # negative width, pitch < width
        r, got = dec.decode_frame(np.ascontiguousarray(bad))
        assert r == 0 and got in (0, 1)
    out = np.empty(w * h * 3 // 2, np.uint8)
    assert dec.decode_frame(good) == (0, 1)
    assert dec.output_frame(out, out.size) == (out.size, out.size)
    dec.deinit()


# --------------------------------------------------------------------------------------------
# delivery design behind jm_nvdec_decode_frame / jm_nvdec_output_frame
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("delay", [1, 2, 5])
@pytest.mark.parametrize("out_fmt", [0, 1])
def test_nvdec_display_delay(J, out_fmt, delay):
    """With a display delay of n the frame announced is the one decoded n calls earlier (like ulMaxDisplayDelay,
    nv_dec.cpp:346); end of stream hands out the rest one per call, then is_exit (test_nv_dec.cpp:232-246)."""
    w, h, pitch, n = 354, 290, 512, 11
    need = w * h * 3 // 2
    chk = oracle.best()
    dec = J.NvDec(0)
    assert dec.set_display_delay(delay) == 0
    assert dec.init(J.NvDec.CODEC_RAW_NV12, out_fmt) == 0
    surfs = [synth.nv12_surface(w, h, pitch, 90, f) for f in range(n)]
    out = np.full(need + 16, synth.OUT_FILL, np.uint8)
    got_frames, gots = [], []
    for s in surfs:
        r, got = dec.decode_frame(J.NvDec.raw_packet(s, w, h, pitch))
        assert r == 0
        gots.append(got)
        if got:
            assert dec.output_frame(out, need + 16) == (need, need)
            got_frames.append(out.copy())
    assert gots == [0] * delay + [1] * (n - delay)
    calls = 0
    while not dec.is_exit():
        r, got = dec.decode_frame(None, 0)
        if got:
            assert dec.output_frame(out, need + 16) == (need, need)
            got_frames.append(out.copy())
        calls += 1
        assert calls <= delay + 2
    assert len(got_frames) == n
    want = np.full(need + 16, synth.OUT_FILL, np.uint8)
    for i, s in enumerate(surfs):
        chk.nvdec_output_frame(s, pitch, w, h, out_fmt, want, need + 16)
        assert np.array_equal(got_frames[i], want), f"frame {i}"
    assert f"Frame Count:\t{n}" in dec.show_dec_info()
    dec.deinit()


@pytest.mark.parametrize("geom", [(1920, 1080, 2048), (1919, 1079, 2048), (66, 34, 128)])
@pytest.mark.parametrize("mode", ["pageable", "pageable_threads", "pinned", "registered", "lazy_pin", "device", "ref"])
def test_nvdec_output_destinations(J, ctx, mode, geom):
    """Every kind of out_buf receives exactly the bytes the reference writes (odd sizes: the tail stays untouched):
    pageable (copy out of the pinned delivery ring, optionally with helper threads), pinned / registered / lazily
    registered (direct DMA), device memory (device-to-device), and the zero-copy view of the ring."""
    w, h, pitch = geom
    need = w * h * 3 // 2
    cap = need + 32
    chk = oracle.best()
    dec = J.NvDec(0)
    if mode == "pageable_threads":
        assert dec.set_option("copy_threads", 3) == 0
    if mode == "lazy_pin":
        assert dec.set_option("lazy_pin", 1) == 0
    assert dec.init(J.NvDec.CODEC_RAW_NV12, 1) == 0
    pinned = dbuf = None
    if mode == "pinned":
        pinned = dec.alloc_host(cap)
        out = np.ctypeslib.as_array((C.c_uint8 * cap).from_address(pinned))
    else:
        out = np.empty(cap, np.uint8)
    registered = False
    if mode == "registered":
        # only whole pages inside the buffer are locked; a buffer with < 64 KB of them stays pageable
        registered = dec.register_host(out) == 0
        assert registered == (cap >= (1 << 17))
    if mode == "device":
        dbuf = ctx.alloc(cap)
    want = np.empty(cap, np.uint8)
    for f in range(5):
        s = synth.nv12_surface(w, h, pitch, 91, f)
        # alternate host payloads and device-pointer packets (SYNC: the surface is freed right after the call)
        if f % 2:
            d = ctx.upload(s)
            assert dec.decode_frame(J.NvDec.raw_packet(None, w, h, pitch, device_ptr=d, flags=J.RawPacket.SYNC)) == (0, 1)
            ctx.memset(d, 0x5A, s.size)
            ctx.free(d)
        else:
            assert dec.decode_frame(J.NvDec.raw_packet(s, w, h, pitch)) == (0, 1)
        want[:] = synth.OUT_FILL
        assert chk.nvdec_output_frame(s, pitch, w, h, 1, want, cap) == (need, need)
        if mode == "ref":
            r, view = dec.output_frame_ref()
            assert r == need and view.size == need
            total = w * h + 2 * (w >> 1) * (h >> 1)                        # the bytes the reference writes (nv_dec.cpp:801-818)
            assert np.array_equal(view[:total], want[:total])
            continue
        if mode == "device":
            ctx.memset(dbuf, synth.OUT_FILL, cap)
            assert dec.output_frame(dbuf, cap) == (need, need)
            ctx.d2h(out, dbuf)
        else:
            out[:] = synth.OUT_FILL
            assert dec.output_frame(out, cap) == (need, need)
            assert dec.output_frame(out, cap) == (need, need)              # fetching the same frame again is allowed
        assert np.array_equal(out, want), f"frame {f}"
    if registered:
        assert dec.unregister_host(out) == 0
    if mode == "registered":
        assert dec.unregister_host(out) == -1
    if pinned:
        dec.free_host(pinned)
    dec.deinit()
    if dbuf:
        ctx.free(dbuf)


def test_nvdec_wait_event_packet_and_reinit(J, ctx):
    """A device-pointer packet can carry the CUDA event that marks the surface complete; jm_nvdec_init on a live
    handle starts over without leaking."""
    w, h, pitch = 320, 200, 384
    need = w * h * 3 // 2
    chk = oracle.best()
    dec = J.NvDec(0)
    out, want = np.empty(need, np.uint8), np.empty(need, np.uint8)
    for rnd in range(3):
        assert dec.init(J.NvDec.CODEC_RAW_NV12, rnd % 2) == 0
        s = synth.nv12_surface(w, h, pitch, 92, rnd)
        d = ctx.upload(s)
        ev = J.Event(ctx)
        ev.record(0)
        pkt = J.NvDec.raw_packet(None, w, h, pitch, device_ptr=d, flags=J.RawPacket.WAIT_EVENT, ready_event=ev.h.value)
        assert dec.decode_frame(pkt) == (0, 1)
        assert dec.output_frame(out, need) == (need, need)
        chk.nvdec_output_frame(s, pitch, w, h, rnd % 2, want, need)
        assert np.array_equal(out, want)
        # truncated extended packet: refused, nothing announced
        assert dec.decode_frame(pkt[:pkt.size - 4]) == (0, 0)
        ev.close()
        ctx.free(d)
    dec.deinit()


def test_calls_restore_the_callers_device(J):
    """Every entry point switches to its handle's device and hands the caller's device back (the reference pushes /
    pops its context per call, nv_dec.cpp:378,398,423,471)."""
    if J.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    L = J.load()
    w, h, pitch = 64, 32, 64
    need = w * h * 3 // 2
    assert L.jmc_set_current_device(0) == 0
    dec = J.NvDec(1)
    assert dec.init(J.NvDec.CODEC_RAW_NV12, 1) == 0
    assert L.jmc_current_device() == 0
    out = np.empty(need, np.uint8)
    assert dec.decode_frame(J.NvDec.raw_packet(synth.nv12_surface(w, h, pitch, 93, 0), w, h, pitch)) == (0, 1)
    assert L.jmc_current_device() == 0
    assert dec.output_frame(out, need) == (need, need)
    assert L.jmc_current_device() == 0
    c1 = J.Ctx(1)
    d = c1.alloc(1024)
    c1.memset(d, 1, 1024)
    c1.free(d)
    assert L.jmc_current_device() == 0
    enc = J.NvEnc(1)
    assert enc.init(w, h, J.NvEnc.FMT_YV12) == 0
    assert enc.enc_frame(synth.i420_frame(w, h, 93, 1))[0] == 0
    assert L.jmc_current_device() == 0
    enc.deinit(), dec.deinit(), c1.close()
    assert L.jmc_current_device() == 0


def test_nvenc_short_buffer_and_reinit(J):
    w, h = 128, 64
    enc = J.NvEnc(0)
    for fmt, n_in in ((J.NvEnc.FMT_NV12, w * h * 3 // 2), (J.NvEnc.FMT_ARGB, w * h * 4)):
        assert enc.init(w, h, fmt) == 0                                # second round: init on a live handle
        src = synth.random_bytes(n_in, 5)
        assert enc.enc_frame(src, n_in - 1)[0] == 8                    # NV_ENC_ERR_INVALID_PARAM, nothing over-read
        assert "shorter" in J.last_error()
        assert enc.enc_frame(src)[0] == 0
    enc.deinit()


def test_job_with_host_side_pointer_list(J, ctx):
    """JMC_JOB_LIST_ON_HOST: up to 8 frame pointers travel as kernel arguments (how jm_nvdec_* batches mapped
    surfaces); alignment is judged from the pointers themselves, so skewed pointers take the any-alignment kernel."""
    w, h, pitch = 640, 360, 768
    surf_bytes, need = pitch * h * 3 // 2, w * h * 3 // 2
    chk = oracle.best()
    for skew in (0, 3):
        n = 5
        surfs = [synth.nv12_surface(w, h, pitch, 94, f) for f in range(n)]
        dsurf = [ctx.upload(s) for s in surfs]
        dout = [ctx.alloc(need + 16) for _ in range(n)]
        j = ctx.job_nvdec(w, h, pitch, 1)
        sl = (C.c_void_p * n)(*dsurf)
        tl = (C.c_void_p * n)(*[d + skew for d in dout])
        j.n_frames, j.flags = n, J.JOB_LIST_ON_HOST
        j.surf.list, j.tight.list = C.cast(sl, C.c_void_p), C.cast(tl, C.c_void_p)
        ctx.convert(j)
        got, want = np.empty(need, np.uint8), np.empty(need, np.uint8)
        for f in range(n):
            ctx.d2h(got, dout[f] + skew)
            chk.nvdec_output_frame(surfs[f], pitch, w, h, 1, want, need)
            assert np.array_equal(got, want), (skew, f)
        j.n_frames = 9
        with pytest.raises(J.JmcError):
            ctx.convert(j)                                             # more than JMC_INLINE_LIST_MAX frames
        for d in dsurf + dout:
            ctx.free(d)


def test_link_probe_and_tools(J, ctx):
    """jmc_link_probe (fixed number of copies and fixed window), tools/jm_link and tools/jm_dropin run and report sane numbers."""
    import json
    import os
    import subprocess
    up, down = ctx.link_probe(8 << 20, 3, 3)
    assert up > 0.5 and down > 0.5
    up, down = ctx.link_probe(8 << 20, -30, 2)              # 30 ms window, device -> host only
    assert up == 0 and down > 0.5
    up, down = ctx.link_probe(8 << 20, 2, 5)                # host -> device from write-combined memory
    assert up > 0.5 and down == 0
    with pytest.raises(J.JmcError):
        ctx.link_probe(0, 1, 3)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "jm_link")
    if os.path.exists(exe):
        d = json.loads(subprocess.run([exe, "--gpus", "1", "--mb", "8", "--ms", "30"], capture_output=True, text=True, timeout=120).stdout)
        assert d["n_gpus"] == 1 and d["bidirectional"]["box_h2d_gbs"] > 0.5 and d["d2h_only"]["box_d2h_gbs"] > 0.5
    exe = os.path.join(root, "tools", "jm_dropin")
    if os.path.exists(exe):
        # every variant of the drop-in loop (pageable / pinned / registered / lazy / zero-copy, delays, threads, 4 handles) on
        # small frames: each checks its first frame against a CPU restatement of nv_dec.cpp:798-820
        p = subprocess.run([exe, "--frames", "12", "--width", "640", "--height", "360", "--pitch", "768"], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        d = json.loads(p.stdout)
        assert len(d) >= 20 and all(isinstance(v, float) and v > 0 for v in d.values()), d
