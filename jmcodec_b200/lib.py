"""ctypes binding of libjmcodec_b200.so -- mirrors include/jmc_cuda.h, jm_nv_dec.h, jmnv_enc.h 1:1.

Nothing here computes: every method is one call into the C-ABI.  numpy arrays are only used as
host buffers (their .ctypes pointers are passed through).
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("JMCODEC_B200_LIB") or os.path.join(_HERE, "libjmcodec_b200.so")   # override: developer A/B builds
_lib = None


class JmcError(RuntimeError):
    pass


class JMC_OP(enum.IntEnum):
    NV12_TO_NV12 = 0
    NV12_TO_I420 = 1
    NV12_TO_SURF = 2
    I420_TO_SURF = 3
    NV12_TO_RGB24 = 4
    NV12_TO_I420_RGB24 = 5
    NV12_TO_ARGB32 = 6
    RGB24_TO_SURF = 7


JOB_ALIGNED16 = 1
JOB_LIST_ON_HOST = 2
INLINE_LIST_MAX = 8


class Frames(C.Structure):
    _fields_ = [("base", C.c_void_p), ("stride", C.c_size_t), ("list", C.c_void_p)]


class Job(C.Structure):
    _fields_ = [
        ("op", C.c_int32), ("n_frames", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("surf", Frames), ("pitch", C.c_int32), ("surf_y_off", C.c_int64), ("surf_uv_off", C.c_int64),
        ("tight", Frames), ("tight_u_off", C.c_int64), ("tight_v_off", C.c_int64),
        ("rgb", Frames), ("rgb_pitch", C.c_int32), ("flags", C.c_uint32),
    ]


class RawPacket(C.Structure):
    """struct jm_nvdec_raw_packet (include/jm_nv_dec.h)."""
    _fields_ = [("magic", C.c_uint32), ("width", C.c_int32), ("height", C.c_int32), ("pitch", C.c_int32),
                ("flags", C.c_uint32), ("reserved", C.c_uint32), ("device_ptr", C.c_uint64)]
    MAGIC = 0x53524D4A
    DEVICE_PTR, SYNC, WAIT_EVENT = 1, 2, 4


class NvEncParam(C.Structure):
    """nv_enc_param (include/jmnv_enc.h, reference nv_enc/jmnv_enc.h:23-53)."""
    _fields_ = [(n, C.c_int) for n in ("codec_id", "in_fmt", "preset", "src_width", "src_height", "dst_width",
                                       "dst_height", "fps", "bitrate_kb", "gop_len", "num_bframe",
                                       "is_external_alloc", "qp")]


def lib_path() -> str:
    return _SO


def build(verbose: bool = False) -> None:
    """Compile csrc/ for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_SIGS = {
    # name: (restype, argtypes)
    "jmc_last_error": (C.c_char_p, []),
    "jmc_version": (C.c_char_p, []),
    "jmc_reload_env": (None, []),
    "jmc_device_count": (C.c_int, []),
    "jmc_current_device": (C.c_int, []),
    "jmc_set_current_device": (C.c_int, [C.c_int]),
    "jmc_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "jmc_ctx_destroy": (C.c_int, [C.c_void_p]),
    "jmc_ctx_device": (C.c_int, [C.c_void_p]),
    "jmc_ctx_sm_count": (C.c_int, [C.c_void_p]),
    "jmc_ctx_stream": (C.c_void_p, [C.c_void_p, C.c_int]),
    "jmc_ctx_sync": (C.c_int, [C.c_void_p]),
    "jmc_ctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "jmc_alloc_device": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "jmc_alloc_pitched": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "jmc_free_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jmc_alloc_host": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "jmc_free_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jmc_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "jmc_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "jmc_memset_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "jmc_job_nvdec": (C.c_int, [C.POINTER(Job)] + [C.c_int] * 4),
    "jmc_job_inteldec": (C.c_int, [C.POINTER(Job)] + [C.c_int] * 7),
    "jmc_job_intelenc": (C.c_int, [C.POINTER(Job)] + [C.c_int] * 7),
    "jmc_job_nvenc": (C.c_int, [C.POINTER(Job)] + [C.c_int] * 4),
    "jmc_job_rgb": (C.c_int, [C.POINTER(Job)] + [C.c_int] * 5),
    "jmc_job_argb": (C.c_int, [C.POINTER(Job)] + [C.c_int] * 4),
    "jmc_job_rgb_to_nv12": (C.c_int, [C.POINTER(Job)] + [C.c_int] * 4),
    "jmc_tight_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "jmc_job_algorithmic_bytes": (C.c_int64, [C.POINTER(Job)]),
    "jmc_convert": (C.c_int, [C.c_void_p, C.POINTER(Job), C.c_void_p]),
    "jmc_convert_timed": (C.c_int, [C.c_void_p, C.POINTER(Job), C.c_int, C.POINTER(C.c_float)]),
    "jmc_link_probe": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double * 2)]),
    "jmc_event_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "jmc_event_destroy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jmc_event_record": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "jmc_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "jmc_pipeline_create": (C.c_int, [C.c_void_p, C.POINTER(Job), C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "jmc_pipeline_destroy": (C.c_int, [C.c_void_p]),
    "jmc_pipeline_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "jmc_pipeline_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "jmc_pipeline_drain": (C.c_int, [C.c_void_p]),
    "jmc_pipeline_h2d_bytes": (C.c_uint64, [C.c_void_p]),
    "jmc_pipeline_d2h_bytes": (C.c_uint64, [C.c_void_p]),
    # jmc_annexb.h
    "jmc_annexb_find_prefix": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "jmc_annexb_find_nalu": (C.c_void_p, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    # jm_nv_dec.h
    "jm_nvdec_create_handle": (C.c_void_p, []),
    "jm_nvdec_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_void_p]),
    "jm_nvdec_deinit": (C.c_int, [C.c_void_p]),
    "jm_nvdec_decode_frame": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "jm_nvdec_output_frame": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_void_p]),
    "jm_nvdec_stream_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "jm_nvdec_set_eof": (None, [C.c_bool, C.c_void_p]),
    "jm_nvdec_is_exit": (C.c_bool, [C.c_void_p]),
    "jm_nvdec_show_dec_info": (C.c_char_p, [C.c_void_p]),
    "jm_nvdec_is_hw_support": (C.c_bool, []),
    "jm_nvdec_set_device": (C.c_int, [C.c_int, C.c_void_p]),
    "jm_nvdec_memory_alloc_host": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "jm_nvdec_memory_release_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jm_nvdec_memory_register_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "jm_nvdec_memory_unregister_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jm_nvdec_set_display_delay": (C.c_int, [C.c_int, C.c_void_p]),
    "jm_nvdec_set_option": (C.c_int, [C.c_char_p, C.c_int, C.c_void_p]),
    "jm_nvdec_output_frame_ref": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_void_p]),
    "jm_nvdec_dropped_frames": (C.c_int, [C.c_void_p]),
    "jm_nvdec_launch_count": (C.c_longlong, [C.c_void_p]),
    "jm_nvdec_deliveries_in_flight": (C.c_int, [C.c_int]),
    # jmnv_enc.h
    "jm_nvenc_create_handle": (C.c_void_p, []),
    "jm_nvenc_init": (C.c_int, [C.POINTER(NvEncParam), C.c_void_p]),
    "jm_nvenc_deinit": (C.c_int, [C.c_void_p]),
    "jm_nvenc_enc_frame": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "jm_nvenc_get_bitstream": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "jm_nvenc_get_spspps_len": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "jm_nvenc_get_spspps": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jm_nvenc_memory_alloc_host": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "jm_nvenc_memory_release_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jm_nvenc_set_device": (C.c_int, [C.c_int, C.c_void_p]),
    "jm_nvenc_peek_surface": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "jm_nvenc_release_surface": (C.c_int, [C.c_void_p]),
}

EXPORTS = tuple(_SIGS)


def load():
    """dlopen the in-tree library.  Raises if it was not built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise JmcError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"(or make -C jmcodec_b200/csrc).  jmcodec_b200 has no CPU fallback.")
        L = C.CDLL(_SO)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)          # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return load().jmc_last_error().decode()


def version() -> str:
    return load().jmc_version().decode()


def reload_env() -> None:
    """Re-read the JMC_* kernel-variant switches (they are cached per process, not read per launch)."""
    load().jmc_reload_env()


def device_count() -> int:
    return load().jmc_device_count()


def _ck(r: int, what: str) -> int:
    if r < 0:
        raise JmcError(f"{what} failed ({r}): {last_error()}")
    return r


def _hptr(a):
    """numpy array / int / None -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)


class HostBuffer:
    """Pinned host memory from jmc_alloc_host, viewed as a numpy uint8 array."""

    def __init__(self, ctx: "Ctx", nbytes: int, write_combined: bool = False):
        self.ctx, self.nbytes = ctx, nbytes
        p = C.c_void_p()
        _ck(ctx.L.jmc_alloc_host(ctx.h, nbytes, 1 if write_combined else 0, C.byref(p)), "jmc_alloc_host")
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(self.ptr))[:nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            self.ctx.L.jmc_free_host(self.ctx.h, self.ptr)
            self.ptr = None


class Ctx:
    """jmc_ctx: one per (host thread, device)."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        _ck(self.L.jmc_ctx_create(device, C.byref(h)), "jmc_ctx_create")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.L.jmc_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def sm_count(self) -> int:
        return self.L.jmc_ctx_sm_count(self.h)

    @property
    def launches(self) -> int:
        return int(self.L.jmc_ctx_launch_count(self.h))

    def stream(self, which: int = 0) -> int:
        return self.L.jmc_ctx_stream(self.h, which) or 0

    def sync(self):
        _ck(self.L.jmc_ctx_sync(self.h), "jmc_ctx_sync")

    # memory
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _ck(self.L.jmc_alloc_device(self.h, nbytes, C.byref(p)), "jmc_alloc_device")
        return p.value

    def alloc_pitched(self, width_bytes: int, rows: int):
        p, pitch = C.c_void_p(), C.c_size_t()
        _ck(self.L.jmc_alloc_pitched(self.h, width_bytes, rows, C.byref(p), C.byref(pitch)), "jmc_alloc_pitched")
        return p.value, pitch.value

    def free(self, dptr: int):
        _ck(self.L.jmc_free_device(self.h, dptr), "jmc_free_device")

    def alloc_host(self, nbytes: int, write_combined: bool = False) -> HostBuffer:
        return HostBuffer(self, nbytes, write_combined)

    def h2d(self, dptr: int, host, nbytes: int | None = None):
        n = host.nbytes if nbytes is None else nbytes
        _ck(self.L.jmc_memcpy_h2d(self.h, dptr, _hptr(host), n), "jmc_memcpy_h2d")

    def d2h(self, host, dptr: int, nbytes: int | None = None):
        n = host.nbytes if nbytes is None else nbytes
        _ck(self.L.jmc_memcpy_d2h(self.h, _hptr(host), dptr, n), "jmc_memcpy_d2h")

    def memset(self, dptr: int, byte: int, nbytes: int):
        _ck(self.L.jmc_memset_device(self.h, dptr, byte, nbytes), "jmc_memset_device")

    def upload(self, host: np.ndarray) -> int:
        d = self.alloc(host.nbytes)
        self.h2d(d, host)
        return d

    # jobs
    def job_nvdec(self, w, h, pitch, out_fmt) -> Job:
        j = Job()
        _ck(self.L.jmc_job_nvdec(C.byref(j), w, h, pitch, out_fmt), "jmc_job_nvdec")
        return j

    def job_inteldec(self, pitch, rows, cx, cy, cw, ch, out_fmt) -> Job:
        j = Job()
        _ck(self.L.jmc_job_inteldec(C.byref(j), pitch, rows, cx, cy, cw, ch, out_fmt), "jmc_job_inteldec")
        return j

    def job_intelenc(self, pitch, rows, cx, cy, cw, ch, is_i420) -> Job:
        j = Job()
        _ck(self.L.jmc_job_intelenc(C.byref(j), pitch, rows, cx, cy, cw, ch, 1 if is_i420 else 0), "jmc_job_intelenc")
        return j

    def job_nvenc(self, w, h, stride, in_fmt) -> Job:
        j = Job()
        _ck(self.L.jmc_job_nvenc(C.byref(j), w, h, stride, in_fmt), "jmc_job_nvenc")
        return j

    def job_rgb(self, w, h, pitch, rgb_pitch, fused=False) -> Job:
        j = Job()
        _ck(self.L.jmc_job_rgb(C.byref(j), w, h, pitch, rgb_pitch, 1 if fused else 0), "jmc_job_rgb")
        return j

    def job_argb(self, w, h, pitch, argb_pitch) -> Job:
        j = Job()
        _ck(self.L.jmc_job_argb(C.byref(j), w, h, pitch, argb_pitch), "jmc_job_argb")
        return j

    def job_rgb_to_nv12(self, w, h, rgb_pitch, stride) -> Job:
        j = Job()
        _ck(self.L.jmc_job_rgb_to_nv12(C.byref(j), w, h, rgb_pitch, stride), "jmc_job_rgb_to_nv12")
        return j

    def link_probe(self, bytes_per_copy: int, copies: int, mode: int):
        """(h2d GB/s, d2h GB/s) of the PCIe link right now; mode 1 = H2D, 2 = D2H, 3 = both at once."""
        r = (C.c_double * 2)()
        _ck(self.L.jmc_link_probe(self.h, bytes_per_copy, copies, mode, C.byref(r)), "jmc_link_probe")
        return r[0], r[1]

    def algorithmic_bytes(self, job: Job) -> int:
        return int(self.L.jmc_job_algorithmic_bytes(C.byref(job)))

    def convert(self, job: Job, stream: int | None = None):
        _ck(self.L.jmc_convert(self.h, C.byref(job), stream), "jmc_convert")

    def convert_timed(self, job: Job, iters: int) -> float:
        ms = C.c_float()
        _ck(self.L.jmc_convert_timed(self.h, C.byref(job), iters, C.byref(ms)), "jmc_convert_timed")
        return ms.value


class Event:
    """jmc_event: a CUDA event on one of the context's streams."""

    def __init__(self, ctx: Ctx):
        self.ctx = ctx
        h = C.c_void_p()
        _ck(ctx.L.jmc_event_create(ctx.h, C.byref(h)), "jmc_event_create")
        self.h = h

    def record(self, which: int = 0):
        _ck(self.ctx.L.jmc_event_record(self.ctx.h, self.h, which), "jmc_event_record")

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        _ck(self.ctx.L.jmc_event_elapsed_ms(self.ctx.h, self.h, stop.h, C.byref(ms)), "jmc_event_elapsed_ms")
        return ms.value

    def close(self):
        if self.h:
            self.ctx.L.jmc_event_destroy(self.ctx.h, self.h)
            self.h = None


class Pipeline:
    """jmc_pipeline: upload -> convert -> pinned delivery ring."""

    def __init__(self, ctx: Ctx, shape: Job, surf_bytes: int, depth: int = 3):
        self.ctx, self.L = ctx, ctx.L
        h = C.c_void_p()
        _ck(self.L.jmc_pipeline_create(ctx.h, C.byref(shape), surf_bytes, depth, C.byref(h)), "jmc_pipeline_create")
        self.h = h

    def submit(self, host_in, host_out, n_frames: int, dev_in=None, host_out2=None) -> int:
        return _ck(self.L.jmc_pipeline_submit(self.h, _hptr(host_in), dev_in, _hptr(host_out), _hptr(host_out2), n_frames),
                   "jmc_pipeline_submit")

    def wait(self, slot: int):
        _ck(self.L.jmc_pipeline_wait(self.h, slot), "jmc_pipeline_wait")

    def drain(self):
        _ck(self.L.jmc_pipeline_drain(self.h), "jmc_pipeline_drain")

    @property
    def h2d_bytes(self) -> int:
        return int(self.L.jmc_pipeline_h2d_bytes(self.h))

    @property
    def d2h_bytes(self) -> int:
        return int(self.L.jmc_pipeline_d2h_bytes(self.h))

    def close(self):
        if self.h:
            self.L.jmc_pipeline_destroy(self.h)
            self.h = None


class NvDec:
    """The jm_nvdec_* API, called exactly as test_nv_dec.cpp:163-259 calls it."""

    CODEC_RAW_NV12 = 100

    def __init__(self, device: int | None = None):
        self.L = load()
        self.h = self.L.jm_nvdec_create_handle()
        if device is not None:
            self.L.jm_nvdec_set_device(device, self.h)

    def init(self, codec_type: int, out_fmt: int, extra: bytes | None = None) -> int:
        return self.L.jm_nvdec_init(codec_type, out_fmt, extra, len(extra) if extra else 0, self.h)

    def deinit(self) -> int:
        r = self.L.jm_nvdec_deinit(self.h)
        self.h = None
        return r

    def decode_frame(self, buf, nbytes: int | None = None):
        """Returns (ret, got_frame)."""
        got = C.c_int(-7)
        n = (buf.nbytes if isinstance(buf, np.ndarray) else 0) if nbytes is None else nbytes
        r = self.L.jm_nvdec_decode_frame(_hptr(buf), n, C.byref(got), self.h)
        return r, got.value

    def output_frame(self, out, capacity: int):
        """Returns (ret, out_len_after)."""
        n = C.c_int(capacity)
        r = self.L.jm_nvdec_output_frame(_hptr(out), C.byref(n), self.h)
        return r, n.value

    def stream_info(self):
        w, h = C.c_int(), C.c_int()
        self.L.jm_nvdec_stream_info(C.byref(w), C.byref(h), self.h)
        return w.value, h.value

    def set_eof(self, v: bool):
        self.L.jm_nvdec_set_eof(v, self.h)

    def is_exit(self) -> bool:
        return bool(self.L.jm_nvdec_is_exit(self.h))

    def show_dec_info(self) -> str:
        return self.L.jm_nvdec_show_dec_info(self.h).decode()

    def alloc_host(self, nbytes: int) -> int:
        p = C.c_void_p()
        if self.L.jm_nvdec_memory_alloc_host(C.byref(p), nbytes, self.h) != 0:
            raise JmcError("jm_nvdec_memory_alloc_host failed: " + last_error())
        return p.value

    def register_host(self, arr: np.ndarray) -> int:
        return self.L.jm_nvdec_memory_register_host(_hptr(arr), arr.nbytes, self.h)

    def unregister_host(self, arr: np.ndarray) -> int:
        return self.L.jm_nvdec_memory_unregister_host(_hptr(arr), self.h)

    def set_display_delay(self, n: int) -> int:
        return self.L.jm_nvdec_set_display_delay(n, self.h)

    def set_option(self, name: str, value: int) -> int:
        return self.L.jm_nvdec_set_option(name.encode(), value, self.h)

    def output_frame_ref(self):
        """Returns (ret, numpy view of the frame inside the pinned delivery ring or None)."""
        p, n = C.c_void_p(), C.c_int(0)
        r = self.L.jm_nvdec_output_frame_ref(C.byref(p), C.byref(n), self.h)
        if r <= 0:
            return r, None
        return r, np.ctypeslib.as_array((C.c_uint8 * n.value).from_address(p.value))

    @property
    def dropped_frames(self) -> int:
        return self.L.jm_nvdec_dropped_frames(self.h)

    @property
    def launches(self) -> int:
        return int(self.L.jm_nvdec_launch_count(self.h))

    def free_host(self, ptr: int):
        self.L.jm_nvdec_memory_release_host(ptr, self.h)

    @staticmethod
    def raw_packet(surface: np.ndarray | None, w: int, h: int, pitch: int, device_ptr: int = 0, flags: int = 0,
                   ready_event: int = 0) -> np.ndarray:
        """Build a JM_NVDEC_CODEC_RAW_NV12 packet: header + pitched surface bytes (or a device pointer)."""
        hdr = RawPacket(RawPacket.MAGIC, w, h, pitch, (RawPacket.DEVICE_PTR if device_ptr else 0) | flags, 0, device_ptr)
        hb = np.frombuffer(bytes(hdr), dtype=np.uint8)
        if device_ptr:
            if flags & RawPacket.WAIT_EVENT:
                return np.concatenate([hb, np.array([ready_event], dtype=np.uint64).view(np.uint8)])
            return hb.copy()
        return np.concatenate([hb, surface.reshape(-1)])


class NvEnc:
    """The jm_nvenc_* API (encoder input path; surface-only on B200)."""

    FMT_NV12, FMT_YV12, FMT_ARGB, FMT_ABGR = 0x1, 0x10, 0x01000000, 0x10000000
    CODEC_SURFACE_ONLY = -1

    def __init__(self, device: int | None = None):
        self.L = load()
        self.h = self.L.jm_nvenc_create_handle()
        if device is not None:
            self.L.jm_nvenc_set_device(device, self.h)

    def init(self, w: int, h: int, in_fmt: int, codec_id: int = -1) -> int:
        p = NvEncParam()
        p.codec_id, p.in_fmt, p.src_width, p.src_height, p.dst_width, p.dst_height = codec_id, in_fmt, w, h, w, h
        p.fps, p.bitrate_kb, p.gop_len, p.is_external_alloc = 30, 4000, 30, 1
        return self.L.jm_nvenc_init(C.byref(p), self.h)

    def deinit(self) -> int:
        r = self.L.jm_nvenc_deinit(self.h)
        self.h = None
        return r

    def enc_frame(self, yuv, nbytes: int | None = None):
        got = C.c_int(-7)
        n = (yuv.nbytes if isinstance(yuv, np.ndarray) else 0) if nbytes is None else nbytes
        r = self.L.jm_nvenc_enc_frame(_hptr(yuv), n, C.byref(got), self.h)
        return r, got.value

    def get_bitstream(self):
        n, key = C.c_int(-7), C.c_int(0)
        r = self.L.jm_nvenc_get_bitstream(None, C.byref(n), C.byref(key), self.h)
        return r, n.value

    def peek_surface(self):
        d, p, rows = C.c_void_p(), C.c_int(), C.c_int()
        r = self.L.jm_nvenc_peek_surface(C.byref(d), C.byref(p), C.byref(rows), self.h)
        return r, d.value, p.value, rows.value

    def release_surface(self) -> int:
        return self.L.jm_nvenc_release_surface(self.h)
