"""Run the cases of tests/cases.py through the CUDA path (C-ABI, one launch per call).

Mirrors cases.run_case(): same inputs, same sentinel-prefilled outputs, so results can be compared
byte for byte with the oracle output and with the golden SHA-256 vectors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import cases as K
from jmcodec_b200 import synth


def _dev_filled(ctx, nbytes, byte):
    d = ctx.alloc(nbytes)
    ctx.memset(d, byte, nbytes)
    return d


def _download(ctx, d, nbytes):
    out = np.empty(nbytes, np.uint8)
    ctx.d2h(out, d)
    return out


def gpu_nvdec(ctx, c):
    s = K.nvdec_input(c)
    cap = c["w"] * c["h"] * 3 // 2 + K.SLACK
    ds, do = ctx.upload(s), _dev_filled(ctx, cap, synth.OUT_FILL)
    j = ctx.job_nvdec(c["w"], c["h"], c["pitch"], c["fmt"])
    j.n_frames, j.surf.base, j.tight.base = 1, ds, do
    ctx.convert(j)
    out = _download(ctx, do, cap)
    ctx.free(ds), ctx.free(do)
    return out


def gpu_inteldec(ctx, c):
    s = K.intel_surface(c)
    cap = c["cw"] * c["ch"] + (c["cw"] * c["ch"]) // 2 + K.SLACK
    ds, do = ctx.upload(s), _dev_filled(ctx, cap, synth.OUT_FILL)
    j = ctx.job_inteldec(c["pitch"], c["rows"], c["cx"], c["cy"], c["cw"], c["ch"], c["fmt"])
    j.n_frames, j.surf.base, j.tight.base = 1, ds, do
    ctx.convert(j)
    out = _download(ctx, do, cap)
    ctx.free(ds), ctx.free(do)
    return out


def gpu_intelenc(ctx, c):
    yuv = K.intelenc_input(c)
    w, h = K.intelenc_dims(c)          # CropW/H == 0 -> Info.Width/Height (intel_enc.cpp:271-278), host-side rule
    rows = c["rows"]
    nsurf = c["pitch"] * (rows + rows // 2)
    dt, dsurf = ctx.upload(yuv), _dev_filled(ctx, nsurf, synth.PAD_BYTE)
    j = ctx.job_intelenc(c["pitch"], rows, c["cx"], c["cy"], w, h, c["i420"])
    j.n_frames, j.surf.base, j.tight.base = 1, dsurf, dt
    ctx.convert(j)
    out = _download(ctx, dsurf, nsurf)
    ctx.free(dt), ctx.free(dsurf)
    return out


def gpu_nvenc(ctx, c):
    assert c["fmt"] in ("nv12", "yv12")
    yuv = K.nvenc_input(c)
    nsurf = K.nvenc_surface_bytes(c)
    dt, dsurf = ctx.upload(yuv), _dev_filled(ctx, nsurf, synth.PAD_BYTE)
    j = ctx.job_nvenc(c["w"], c["h"], c["stride"], K.NVENC_FMTS[c["fmt"]])
    j.n_frames, j.surf.base, j.tight.base = 1, dsurf, dt
    ctx.convert(j)
    out = _download(ctx, dsurf, nsurf)
    ctx.free(dt), ctx.free(dsurf)
    return out


def gpu_rgb(ctx, c, fused=False):
    s = K.rgb_input(c)
    w, h = c["w"], c["h"]
    rgb_pitch = 3 * w
    cap = rgb_pitch * h + K.SLACK
    ds, drgb = ctx.upload(s), _dev_filled(ctx, cap, synth.OUT_FILL)
    j = ctx.job_rgb(w, h, c["pitch"], rgb_pitch, fused)
    j.n_frames, j.surf.base, j.rgb.base = 1, ds, drgb
    dt, tcap = None, w * h * 3 // 2 + K.SLACK
    if fused:
        dt = _dev_filled(ctx, tcap, synth.OUT_FILL)
        j.tight.base = dt
    ctx.convert(j)
    out = _download(ctx, drgb, cap)
    tight = _download(ctx, dt, tcap) if fused else None
    ctx.free(ds), ctx.free(drgb)
    if dt:
        ctx.free(dt)
    return (out, tight) if fused else out


def gpu_argb(ctx, c):
    s = K.rgb_input(c)
    w, h = c["w"], c["h"]
    cap = 4 * w * h + K.SLACK
    ds, dout = ctx.upload(s), _dev_filled(ctx, cap, synth.OUT_FILL)
    j = ctx.job_argb(w, h, c["pitch"], 4 * w)
    j.n_frames, j.surf.base, j.rgb.base = 1, ds, dout
    ctx.convert(j)
    out = _download(ctx, dout, cap)
    ctx.free(ds), ctx.free(dout)
    return out


def gpu_rgb2nv12(ctx, c):
    rgb = K.rgb2nv12_input(c)
    w, h = c["w"], c["h"]
    nsurf = K.rgb2nv12_surface_bytes(c)
    drgb, dsurf = ctx.upload(rgb), _dev_filled(ctx, nsurf, synth.PAD_BYTE)
    j = ctx.job_rgb_to_nv12(w, h, 3 * w + c["skew"], c["pitch"])
    j.n_frames, j.rgb.base, j.surf.base = 1, drgb, dsurf
    ctx.convert(j)
    out = _download(ctx, dsurf, nsurf)
    ctx.free(drgb), ctx.free(dsurf)
    return out


GPU_RUNNERS = {"rgb2nv12": gpu_rgb2nv12, "argb32": gpu_argb, "nvdec": gpu_nvdec, "inteldec": gpu_inteldec, "intelenc": gpu_intelenc, "nvenc": gpu_nvenc, "rgb24": gpu_rgb}


def run_case_gpu(ctx, c):
    return GPU_RUNNERS[c["op"]](ctx, c)


def device_ptr_array(ctx, ptrs):
    """Upload a list of device pointers as a device array (jmc_frames.list)."""
    a = np.array(ptrs, dtype=np.uint64)
    return ctx.upload(a.view(np.uint8))
