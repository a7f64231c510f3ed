"""CPU-side checks of the boundary: the shared library loads without a GPU, exports every symbol the
headers declare, its geometry fillers follow the reference's offset arithmetic, and it fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

import jmcodec_b200 as J
from jmcodec_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared(header):
    src = open(os.path.join(INC, header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"(?:JMC_API|JMDLL_FUNC)\s+[\w\s\*]+?\b(\w+)\s*\(", src))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(J.lib_path()):
        J.build()
    return J.load()


def test_every_declared_symbol_is_exported(lib):
    declared = _declared("jmc_cuda.h") | _declared("jm_nv_dec.h") | _declared("jmnv_enc.h") | _declared("jmc_annexb.h")
    assert len(declared) >= 55
    out = subprocess.run(["nm", "-D", "--defined-only", J.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert declared <= exported, sorted(declared - exported)
    assert declared == set(L.EXPORTS), sorted(declared ^ set(L.EXPORTS))      # the binding covers the whole ABI
    # nothing from the oracle or the reference leaks into the product library
    assert not [s for s in exported if s.startswith(("jmo_", "jmref_"))]


def test_reference_entry_points_present(lib):
    # nv_dec/jm_nv_dec.h:27-88 (11 functions) and nv_enc/jmnv_enc.h:55-67 (9 functions)
    for n in ("create_handle", "init", "deinit", "decode_frame", "output_frame", "stream_info", "set_eof",
              "is_exit", "show_dec_info", "is_hw_support"):
        assert hasattr(lib, "jm_nvdec_" + n)
    for n in ("create_handle", "init", "deinit", "enc_frame", "get_bitstream", "get_spspps_len", "get_spspps",
              "memory_alloc_host", "memory_release_host"):
        assert hasattr(lib, "jm_nvenc_" + n)


def _job(fn, *a):
    j = L.Job()
    r = getattr(J.load(), fn)(C.byref(j), *a)
    return r, j


@pytest.mark.parametrize("w,h,p", [(1920, 1080, 2048), (1919, 1079, 2048), (3, 5, 4), (1, 1, 256)])
def test_job_nvdec_offsets(lib, w, h, p):
    r, j = _job("jmc_job_nvdec", w, h, p, 1)
    assert r == 0 and j.op == J.JMC_OP.NV12_TO_I420
    assert (j.surf_y_off, j.surf_uv_off) == (0, p * h)                         # nv_dec.cpp:765
    assert (j.tight_u_off, j.tight_v_off) == (w * h, w * h + (w >> 1) * (h >> 1))   # :779,:810
    assert lib.jmc_tight_bytes(w, h) == w * h * 3 // 2                          # :773
    r, j = _job("jmc_job_nvdec", w, h, p, 0)
    assert j.op == J.JMC_OP.NV12_TO_NV12
    assert _job("jmc_job_nvdec", w, h, w - 1, 1)[0] == -1 if w > 0 else True


def test_job_intel_offsets(lib):
    p, rows, cx, cy, cw, ch = 1952, 1088, 5, 3, 1279, 719
    r, j = _job("jmc_job_inteldec", p, rows, cx, cy, cw, ch, 1)
    assert r == 0
    assert j.surf_y_off == cy * p + cx                                         # intel_dec.cpp:285
    assert j.surf_uv_off == p * rows + (cy // 2) * p + cx // 2                 # :292-293 (crop_x/2 BYTES)
    assert j.tight_v_off == cw * ch + (cw * ch // 2) // 2                      # :306-307, differs from (cw/2)*(ch/2)
    assert j.tight_v_off != cw * ch + (cw // 2) * (ch // 2)
    r, j = _job("jmc_job_intelenc", p, rows, cx, cy, cw, ch, 1)
    assert r == 0 and j.op == J.JMC_OP.I420_TO_SURF
    assert j.tight_v_off == cw * ch + (cw // 2) * (ch // 2)                    # intel_enc.cpp:370
    assert _job("jmc_job_intelenc", 70000, rows, 0, 0, 64, 64, 1)[0] == -1     # uint16_t pitch in the reference


def test_job_nvenc_offsets(lib):
    w, h, s = 1919, 1079, 2048
    r, j = _job("jmc_job_nvenc", w, h, s, 0x10)
    assert r == 0 and j.op == J.JMC_OP.I420_TO_SURF
    assert j.surf_uv_off == s * h                                              # nv_enc.cpp:1069
    assert (j.tight_u_off, j.tight_v_off) == (w * h, w * h * 5 // 4)           # :1055-1056
    assert _job("jmc_job_nvenc", w, h, s, 0x1)[1].op == J.JMC_OP.NV12_TO_SURF
    assert _job("jmc_job_nvenc", w, h, s, 0x01000000)[0] == -1                 # ARGB is a plain copy, not a surface job


def test_algorithmic_bytes(lib):
    # 3*w*h for the YUV ops, 4.5*w*h RGB, 6*w*h fused
    _, j = _job("jmc_job_nvdec", 1920, 1080, 2048, 1)
    assert lib.jmc_job_algorithmic_bytes(C.byref(j)) == 6220800
    _, j = _job("jmc_job_nvenc", 3840, 2160, 4096, 0x10)
    assert lib.jmc_job_algorithmic_bytes(C.byref(j)) == 24883200
    _, j = _job("jmc_job_rgb", 3840, 2160, 4096, 3 * 3840, 0)
    assert lib.jmc_job_algorithmic_bytes(C.byref(j)) == 37324800
    _, j = _job("jmc_job_rgb", 3840, 2160, 4096, 3 * 3840, 1)
    assert lib.jmc_job_algorithmic_bytes(C.byref(j)) == 49766400
    _, j = _job("jmc_job_argb", 3840, 2160, 4096, 4 * 3840)                   # 1.5*w*h read + 4*w*h written
    assert lib.jmc_job_algorithmic_bytes(C.byref(j)) == 45619200


def test_job_rgb_to_nv12_filler(lib):
    w, h, s_ = 1366, 768, 1536
    r, j = _job("jmc_job_rgb_to_nv12", w, h, 3 * w, s_)
    assert r == 0 and j.op == J.JMC_OP.RGB24_TO_SURF
    assert (j.width, j.height, j.pitch, j.rgb_pitch) == (w, h, s_, 3 * w)
    assert (j.surf_y_off, j.surf_uv_off) == (0, s_ * h)                        # nv_enc.cpp:1069
    assert lib.jmc_job_algorithmic_bytes(C.byref(j)) == 3 * w * h + w * h + 2 * (w >> 1) * (h >> 1)
    assert _job("jmc_job_rgb_to_nv12", w, h, 3 * w - 1, s_)[0] == -1
    assert _job("jmc_job_rgb_to_nv12", w, h, 3 * w, w - 1)[0] == -1


def test_job_argb_filler(lib):
    w, h, p = 1366, 768, 1536
    r, j = _job("jmc_job_argb", w, h, p, 4 * w)
    assert r == 0 and j.op == J.JMC_OP.NV12_TO_ARGB32
    assert (j.width, j.height, j.pitch, j.rgb_pitch) == (w, h, p, 4 * w)
    assert (j.surf_y_off, j.surf_uv_off) == (0, p * h)                         # the nv_dec surface layout, nv_dec.cpp:765
    assert _job("jmc_job_argb", w, h, p, 4 * w - 1)[0] == -1                   # rows must hold 4 bytes per pixel


def _no_gpu():
    return J.device_count() <= 0


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_fails_loudly_without_a_device(lib):
    """No CPU fallback: every entry that would compute refuses."""
    with pytest.raises(J.JmcError):
        J.Ctx(0)
    assert "CUDA" in J.last_error() or "device" in J.last_error()
    assert not lib.jm_nvdec_is_hw_support()
    d = J.NvDec()
    assert d.init(J.NvDec.CODEC_RAW_NV12, 1) == -2          # nvdec_cuda_init: no device (nv_dec.cpp:219-222)
    assert d.decode_frame(None, 0) == (0, 0)
    assert d.output_frame(None, 0)[0] == -1
    assert d.deinit() == 0
    e = J.NvEnc()
    assert e.init(64, 64, J.NvEnc.FMT_YV12) == 1            # NV_ENC_ERR_NO_ENCODE_DEVICE
    assert e.enc_frame(None, 0)[0] == -1
    e.deinit()


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "_SO", "/nonexistent/libjmcodec_b200.so")
    with pytest.raises(J.JmcError, match="no CPU fallback"):
        L.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jmcodec_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "jm_oracle" not in txt and "libjmref" not in txt, f


def test_headers_are_valid_c99_and_cxx(tmp_path):
    """The boundary is a C ABI: every public header must compile as strict C99 and as C++, and the C
    example must link against the library."""
    src = tmp_path / "inc.c"
    src.write_text('#include "jm_nv_dec.h"\n#include "jmnv_enc.h"\n#include "jmc_cuda.h"\n#include "jmc_annexb.h"\n'
                   "int main(void) { jmc_job j; nv_enc_param p; jm_nvdec_raw_packet k; (void)j; (void)p; (void)k; return 0; }\n")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", INC, "-fsyntax-only", str(src)], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-I", INC, "-x", "c++", "-fsyntax-only", str(src)], check=True)
    exe = tmp_path / "decode_raw"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", INC, os.path.join(ROOT, "examples", "decode_raw.c"),
                    "-L", os.path.dirname(J.lib_path()), "-ljmcodec_b200", "-Wl,-rpath," + os.path.dirname(J.lib_path()), "-o", str(exe)], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    if _no_gpu():
        assert p.returncode == 1 and "no CUDA device" in p.stderr       # fails loudly, no CPU fallback
    else:
        assert p.returncode == 0 and "decoded 8 of 8 frames, exit=1" in p.stdout


def test_nvdec_handle_options_without_a_device(lib):
    """Host-side logic of the jm_nvdec_* extensions that needs no CUDA: option names and ranges, defaults from the
    environment, calls on a handle that was never initialised."""
    h = lib.jm_nvdec_create_handle()
    assert h
    assert lib.jm_nvdec_set_display_delay(0, h) == 0 and lib.jm_nvdec_set_display_delay(20, h) == 0
    assert lib.jm_nvdec_set_display_delay(-1, h) == -1 and lib.jm_nvdec_set_display_delay(21, h) == -1
    for name, good, bad in ((b"display_delay", 3, 99), (b"copy_threads", 2, 17), (b"map_limit", 8, 9), (b"map_limit", 1, 0)):
        assert lib.jm_nvdec_set_option(name, good, h) == 0, name
        assert lib.jm_nvdec_set_option(name, bad, h) == -1, name
    assert lib.jm_nvdec_set_option(b"lazy_pin", 1, h) == 0
    assert lib.jm_nvdec_set_option(b"no_such_option", 1, h) == -1
    assert lib.jm_nvdec_set_option(None, 1, h) == -1
    got = C.c_int(5)
    assert lib.jm_nvdec_decode_frame(None, 0, C.byref(got), h) == 0 and got.value == 0      # never initialised: swallowed like nv_dec.cpp:491-493
    n = C.c_int(16)
    buf = (C.c_uint8 * 16)()
    assert lib.jm_nvdec_output_frame(buf, C.byref(n), h) == -1 and n.value == 16             # no frame: -1, *out_len untouched
    p = C.c_void_p()
    assert lib.jm_nvdec_output_frame_ref(C.byref(p), C.byref(n), h) == -1
    assert lib.jm_nvdec_memory_register_host(buf, 16, h) == -1                               # no context yet
    assert lib.jm_nvdec_dropped_frames(h) == 0 and lib.jm_nvdec_launch_count(h) == 0
    assert lib.jm_nvdec_deinit(h) == 0
    lib.jmc_reload_env()                                                                     # harmless without a device
