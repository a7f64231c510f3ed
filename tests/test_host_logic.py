"""CPU logic of the delivery code (jm_nv_dec.cu) that otherwise only runs behind a GPU: the helper-thread copy pool and
the registered-range arithmetic.  The harness includes the translation unit whole, links the library's other objects
and runs without a device (tests/host_logic/copy_pool_test.cu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "jmcodec_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    subprocess.run(["make", "-C", CSRC, "-j4"], check=True, capture_output=True)        # the other objects of the library
    objs = [os.path.join(CSRC, "build", o) for o in ("jmc_runtime.o", "jmc_kernels.o", "jmc_annexb.o")]
    exe = tmp_path_factory.mktemp("host_logic") / "copy_pool_test"
    cmd = [NVCC, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           os.path.join(ROOT, "tests", "host_logic", "copy_pool_test.cu")] + objs + ["-o", str(exe), "-ldl", "-lpthread"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    return str(exe)


def test_copy_pool_and_registration_arithmetic(harness):
    p = subprocess.run([harness], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-3000:] + p.stderr[-1000:]


def test_copy_pool_under_thread_sanitizer(tmp_path):
    """The pool alone (no CUDA objects) under -fsanitize=thread: phases, slices, generations, shutdown."""
    src = tmp_path / "tsan.cpp"
    text = open(os.path.join(CSRC, "jm_nv_dec.cu")).read()
    a, b = text.index("struct copy_job {"), text.index("/* ---- state ---")
    src.write_text("#include <stdint.h>\n#include <string.h>\n#include <atomic>\n#include <chrono>\n#include <condition_variable>\n"
                   "#include <mutex>\n#include <thread>\n#include <vector>\n#include <stdio.h>\n" + text[a:b] + r'''
int main()
{
    copy_pool pool;
    pool.want_threads = 3;
    std::vector<uint8_t> src(3 << 20), dst(3 << 20);
    for (size_t i = 0; i < src.size(); i++) src[i] = (uint8_t)(i * 2654435761u >> 13);
    for (int it = 0; it < 40; it++) {
        memset(dst.data(), 0, dst.size());
        copy_job j = { dst.data(), 4096, src.data(), 4096, 4096, (3u << 20) / 4096, 100, 8 };
        pool.run(j, [](int) {});
        if (memcmp(src.data(), dst.data(), src.size())) { printf("mismatch\n"); return 1; }
        if (it == 20) pool.shutdown();
    }
    pool.shutdown();
    printf("OK\n");
    return 0;
}
''')
    exe = tmp_path / "tsan"
    p = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", str(src), "-o", str(exe), "-lpthread"], capture_output=True, text=True)
    if p.returncode != 0 and "sanitize" in p.stderr + p.stdout:
        pytest.skip("no thread sanitizer runtime in this toolchain")
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    if "FATAL: ThreadSanitizer" in p.stderr and "unexpected memory mapping" in p.stderr:
        pytest.skip("thread sanitizer cannot map its shadow memory in this container")
    assert p.returncode == 0 and "OK" in p.stdout and "WARNING: ThreadSanitizer" not in p.stderr, p.stderr[-3000:]
