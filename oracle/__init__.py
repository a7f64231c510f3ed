"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU checker.

Two libraries live here:

* ``libjmoracle.so``  -- plain-C restatement of the reference loops (``jm_oracle.c``), kind "port".
* ``_ref/libjmref.so`` -- the UNMODIFIED reference translation units compiled in place from
  ``/root/reference`` (``oracle/Makefile``), kind "reference".  Built in the dev container; the
  prebuilt file travels to the GPU box, ``/root/reference`` does not.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``jmcodec_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libjmoracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libjmref.so")
REF_O3_SO = os.path.join(_HERE, "_ref", "libjmref_o3avx2.so")      # same sources, gcc -O3 -mavx2 ("best-effort CPU")

_u8p = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)


def build(quiet: bool = True) -> None:
    """(Re)build the checker libraries; the _ref part only where /root/reference is mounted."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u8p)


class _Port:
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        L.jmo_nvdec_output_frame.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _ip]
        L.jmo_nvdec_output_frame.restype = C.c_int
        L.jmo_inteldec_output_frame.argtypes = [_u8p, _u8p] + [C.c_int] * 7 + [_u8p, _ip]
        L.jmo_inteldec_output_frame.restype = C.c_int
        L.jmo_intelenc_input.argtypes = [_u8p, C.c_int, C.c_int, _u8p, _u8p] + [C.c_int] * 8
        L.jmo_intelenc_input.restype = C.c_int
        L.jmo_nvenc_upload.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.jmo_nvenc_upload.restype = C.c_int
        L.jmo_nv12_to_rgb24.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.jmo_nv12_to_rgb24.restype = C.c_int
        L.jmo_nv12_to_argb32.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.jmo_nv12_to_argb32.restype = C.c_int
        L.jmo_rgb24_to_nv12.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.jmo_rgb24_to_nv12.restype = C.c_int
        L.jmo_nvdec_run.argtypes = [_u8p, C.c_size_t, C.c_int, _u8p, C.c_size_t, C.c_int] + [C.c_int] * 6
        L.jmo_nvdec_run.restype = C.c_double
        self.L = L
        self.kind = "port"
        self._nvdec = L.jmo_nvdec_output_frame
        self._inteldec = L.jmo_inteldec_output_frame
        self._intelenc = L.jmo_intelenc_input
        self._nvdec_run = L.jmo_nvdec_run


class _Ref:
    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = C.CDLL(REF_SO)
        L.jmref_nvdec_output_frame.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _ip]
        L.jmref_nvdec_output_frame.restype = C.c_int
        L.jmref_inteldec_output_frame.argtypes = [_u8p, _u8p] + [C.c_int] * 7 + [_u8p, _ip]
        L.jmref_inteldec_output_frame.restype = C.c_int
        L.jmref_intelenc_input.argtypes = [_u8p, C.c_int, C.c_int, _u8p, _u8p] + [C.c_int] * 8
        L.jmref_intelenc_input.restype = C.c_int
        L.jmref_nvdec_run.argtypes = [_u8p, C.c_size_t, C.c_int, _u8p, C.c_size_t, C.c_int] + [C.c_int] * 6
        L.jmref_nvdec_run.restype = C.c_double
        L.jmref_intelenc_run.argtypes = [_u8p, C.c_size_t, C.c_int, _u8p, C.c_size_t, C.c_int] + [C.c_int] * 5
        L.jmref_intelenc_run.restype = C.c_double
        L.jmref_nvenc_convert.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _ip]
        L.jmref_nvenc_convert.restype = C.c_int
        self.L = L
        self.kind = "reference"
        self._nvdec = L.jmref_nvdec_output_frame
        self._inteldec = L.jmref_inteldec_output_frame
        self._intelenc = L.jmref_intelenc_input
        self._nvdec_run = L.jmref_nvdec_run


class Checker:
    """Uniform numpy front-end over either library (same functions, same semantics)."""

    def __init__(self, impl):
        self.impl = impl
        self.kind = impl.kind

    # --- decode side -----------------------------------------------------------------------
    def nvdec_output_frame(self, surf, pitch, w, h, out_fmt, out, out_len, have_frame=True):
        """Returns (ret, out_len_after).  `out` is written in place."""
        n = C.c_int(int(out_len))
        r = self.impl._nvdec(_ptr(surf), pitch, w, h, out_fmt, 1 if have_frame else 0, _ptr(out), C.byref(n))
        return r, n.value

    def inteldec_output_frame(self, surf, uv_off, pitch, crop, out_fmt, out, out_len, have_surface=True):
        cx, cy, cw, ch = crop
        n = C.c_int(int(out_len))
        y = _ptr(surf)
        uv = C.cast(C.addressof(y.contents) + uv_off, _u8p)
        r = self.impl._inteldec(y, uv, pitch, cx, cy, cw, ch, out_fmt, 1 if have_surface else 0, _ptr(out), C.byref(n))
        return r, n.value

    # --- encode side -----------------------------------------------------------------------
    def intelenc_input(self, yuv, is_i420, surf, uv_off, pitch, info_wh, crop, surface_free=True):
        cx, cy, cw, ch = crop
        y = _ptr(surf)
        uv = C.cast(C.addressof(y.contents) + uv_off, _u8p)
        return self.impl._intelenc(_ptr(yuv), int(yuv.size), 1 if is_i420 else 0, y, uv, pitch,
                                   info_wh[0], info_wh[1], cx, cy, cw, ch, 1 if surface_free else 0)

    # --- timed loops -----------------------------------------------------------------------
    def nvdec_run(self, surfs, out, pitch, w, h, out_fmt, frames, nthreads):
        """surfs: (n_surf, surf_stride) uint8, out: (n_out, out_stride) uint8.  Returns seconds."""
        assert surfs.ndim == 2 and out.ndim == 2
        t = self.impl._nvdec_run(_ptr(surfs), surfs.shape[1], surfs.shape[0],
                                 _ptr(out), out.shape[1], out.shape[0],
                                 pitch, w, h, out_fmt, frames, nthreads)
        if t < 0:
            raise RuntimeError("CPU baseline loop reported a bad return code")
        return t


_port = None
_ref = None


def port() -> Checker:
    global _port
    if _port is None:
        _port = Checker(_Port())
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> Checker:
    global _ref
    if _ref is None:
        _ref = Checker(_Ref())
    return _ref


def ref_best_effort_run(surfs, out, pitch, w, h, out_fmt, frames, nthreads):
    """Timed loop of the reference function built with -O3 -mavx2 (BASELINE.md 3 "best-effort CPU").
    Returns seconds, or None if that build is absent or the CPU lacks AVX2."""
    if not os.path.exists(REF_O3_SO):
        return None
    try:
        if "avx2" not in open("/proc/cpuinfo").read():
            return None
    except OSError:
        return None
    L = C.CDLL(REF_O3_SO)
    L.jmref_nvdec_run.argtypes = [_u8p, C.c_size_t, C.c_int, _u8p, C.c_size_t, C.c_int] + [C.c_int] * 6
    L.jmref_nvdec_run.restype = C.c_double
    t = L.jmref_nvdec_run(_ptr(surfs), surfs.shape[1], surfs.shape[0], _ptr(out), out.shape[1], out.shape[0],
                          pitch, w, h, out_fmt, frames, nthreads)
    return t if t > 0 else None


def best() -> Checker:
    """The strongest checker available: the compiled reference if present, else the port."""
    return ref() if have_ref() else port()


def nvenc_upload(in_buf, fmt, w, h, surf, stride):
    """C restatement of nv_enc.cpp:1023-1103 (device side restated on the CPU)."""
    return port().impl.L.jmo_nvenc_upload(_ptr(in_buf), fmt, w, h, _ptr(surf), stride)


def ref_nvenc_convert(in_buf, fmt, w, h, surf, stride):
    """The reference's own nvenc_convert_yuv_data_to_nv12() run against a fake CUDA driver
    (oracle/ref_nvenc_driver.cpp).  Returns (ret, InterleaveUV launches issued)."""
    n = C.c_int(0)
    r = ref().impl.L.jmref_nvenc_convert(_ptr(in_buf), int(in_buf.size), fmt, w, h, _ptr(surf), stride, C.byref(n))
    return r, n.value


# exists only in the port: the reference has no YUV->RGB code at all

def nv12_to_rgb24(surf, pitch, w, h, rgb, rgb_pitch):
    return port().impl.L.jmo_nv12_to_rgb24(_ptr(surf), pitch, w, h, _ptr(rgb), rgb_pitch)


def nv12_to_argb32(surf, pitch, w, h, argb, argb_pitch):
    return port().impl.L.jmo_nv12_to_argb32(_ptr(surf), pitch, w, h, _ptr(argb), argb_pitch)


def rgb24_to_nv12(rgb, rgb_pitch, w, h, surf, pitch):
    return port().impl.L.jmo_rgb24_to_nv12(_ptr(rgb), rgb_pitch, w, h, _ptr(surf), pitch)
