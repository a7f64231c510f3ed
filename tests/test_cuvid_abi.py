"""jmcodec_b200/csrc/cuvid_min.h against the NVDEC headers the reference vendors (compile-time only).

The decoder front-end binds libnvcuvid.so.1 at run time through our own minimal declaration of the CUVID C ABI;
no NVDEC engine is reachable on the build or GPU boxes, so the one thing that can be pinned is that every struct
size, field offset and enum value jm_nv_dec.cu relies on equals the reference's nv_sdk/inc/dynlink_nvcuvid.h /
dynlink_cuviddec.h.  tests/abi_check/cuvid_layout.cpp holds the static_asserts; the reference headers are included
in place and never copied."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/nv_sdk/inc"


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="/root/reference is not mounted here")
def test_cuvid_min_matches_the_reference_sdk_headers():
    cmd = ["g++", "-std=gnu++11", "-fpermissive", "-w", "-fsyntax-only", "-I" + os.path.join(ROOT, "oracle", "ref_shim"),
           "-I" + REF_INC, "-I" + os.path.join(ROOT, "jmcodec_b200", "csrc"), os.path.join(ROOT, "tests", "abi_check", "cuvid_layout.cpp")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="/root/reference is not mounted here")
def test_drop_in_headers_declare_the_reference_signatures():
    """include/jm_nv_dec.h and include/jmnv_enc.h against nv_dec/jm_nv_dec.h and nv_enc/jmnv_enc.h: every entry point
    has the reference's function type, nv_enc_param the reference's layout (tests/abi_check/api_signatures.cpp)."""
    cmd = ["g++", "-std=gnu++11", "-fsyntax-only", "-I/root/reference", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi_check", "api_signatures.cpp")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="/root/reference is not mounted here")
def test_the_layout_check_can_fail(tmp_path):
    """Guard against a vacuous check: a deliberately wrong assertion must stop the compile."""
    src = open(os.path.join(ROOT, "tests", "abi_check", "cuvid_layout.cpp")).read()
    bad = tmp_path / "neg.cpp"
    bad.write_text(src + '\nstatic_assert(sizeof(::CUVIDPROCPARAMS) == 1, "deliberate");\n')
    cmd = ["g++", "-std=gnu++11", "-fpermissive", "-w", "-fsyntax-only", "-I" + os.path.join(ROOT, "oracle", "ref_shim"),
           "-I" + REF_INC, "-I" + os.path.join(ROOT, "jmcodec_b200", "csrc"), str(bad)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "deliberate" in p.stderr
