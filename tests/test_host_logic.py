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


# ---- the whole host layer on a CUDA runtime simulator -----------------------------------------------------------------
HL = os.path.join(ROOT, "tests", "host_logic")
PRODUCT = ["jm_nv_dec.cu", "jmc_runtime.cu", "jmnv_enc.cu"]              # compiled unchanged, as C++, against fake_cuda/cuda_runtime.h
HARNESS = [os.path.join(HL, "fake_cuda", "fake_cuda.cpp"), os.path.join(HL, "fake_launch.cpp"), os.path.join(HL, "delivery_sim_test.cpp")]
INCLUDES = ["-I", os.path.join(HL, "fake_cuda"), "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _run(cmd, **kw):
    p = subprocess.run(cmd, capture_output=True, text=True, **kw)
    assert p.returncode == 0, " ".join(map(str, cmd)) + "\n" + p.stderr[-3000:]


def _build_sim(out_dir, flags, nvdec_source=None):
    """Objects + link of the simulation harness in out_dir; nvdec_source replaces jm_nv_dec.cu (mutation tests).
    The translation units are compiled side by side."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(out_dir, exist_ok=True)
    jobs = [(["gcc"] + flags + ["-c", os.path.join(ROOT, "oracle", "jm_oracle.c")], os.path.join(out_dir, "jm_oracle.o"))]
    for f in PRODUCT:
        src = nvdec_source if (nvdec_source and f == "jm_nv_dec.cu") else os.path.join(CSRC, f)
        jobs.append((["g++", "-std=c++17"] + flags + INCLUDES + ["-x", "c++", "-c", src], os.path.join(out_dir, f + ".o")))
    for f in HARNESS:
        jobs.append((["g++", "-std=c++17"] + flags + INCLUDES + ["-c", f], os.path.join(out_dir, os.path.basename(f) + ".o")))
    # the fake NVDEC library of the GPU tests, built against the simulator; its cuda* calls bind to the executable's
    lib = os.path.join(out_dir, "libfake_nvcuvid_sim.so")
    jobs.append((["g++", "-std=c++17"] + flags + ["-shared", "-fPIC", "-I", os.path.join(HL, "fake_cuda"), "-I", CSRC, "-x", "c++",
                  os.path.join(ROOT, "tests", "fake_nvcuvid", "fake_nvcuvid.cu")], lib))
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(lambda j: _run(j[0] + ["-o", j[1]]), jobs))
    exe = os.path.join(out_dir, "delivery_sim_test")
    _run(["g++"] + flags + [o for _, o in jobs if o.endswith(".o")] + ["-rdynamic", "-ldl", "-lpthread", "-o", exe])
    return exe, lib


@pytest.fixture(scope="module")
def sim_sanitized(tmp_path_factory):
    flags = ["-g", "-O1", "-fno-omit-frame-pointer", "-fsanitize=address,undefined"]
    probe = subprocess.run(["g++", "-fsanitize=address,undefined", "-x", "c++", "-", "-o", os.devnull], input="int main(){return 0;}", capture_output=True, text=True)
    if probe.returncode != 0:
        flags = ["-g", "-O1"]                                            # no sanitizer runtimes here: the checks still run
    return _build_sim(str(tmp_path_factory.mktemp("sim_san")), flags)


SCENARIOS = ["raw", "cuvid", "nvenc", "alloc-failure", "multi", "pipeline", "threads", "fuzz", "runtime", "call-failure"]


@pytest.fixture(scope="module")
def sim_runs(sim_sanitized):
    """All scenarios are started at once, one process each (they are independent); every test collects its own."""
    exe, lib = sim_sanitized
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="halt_on_error=1")
    procs = {sc: subprocess.Popen([exe, lib, sc], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for sc in SCENARIOS}
    yield procs
    for p in procs.values():
        if p.poll() is None:
            p.kill()


@pytest.mark.parametrize("scenario", SCENARIOS)
def test_host_layer_on_the_cuda_simulator(sim_runs, scenario):
    """jm_nv_dec.cu / jmnv_enc.cu / jmc_runtime.cu, unchanged, through the public C API on a CUDA runtime simulator whose
    streams run work as late as CUDA allows (only when waited for / at random moments / at once): every frame against
    the oracle for every input kind x out_buf kind x display delay, the NVDEC front-end against the fake library
    (batch drain, map limit, format change, overflow), the encoder-input API, an allocation failure at every
    allocation site, failing enqueue-type CUDA calls (copies, event records, stream waits, launches), five handles interleaved on one device, the jmc_pipeline_* batch pipeline, four threads with a
    handle each, malformed packets and calls on missing / uninitialised handles; no leak, no free under pending work, caller's device restored -- under ASan + UBSan."""
    p = sim_runs[scenario]
    out, err = p.communicate(timeout=900)
    assert p.returncode == 0 and out.strip().endswith("OK"), out[-3000:] + err[-3000:]


def test_handles_on_several_threads_under_thread_sanitizer(tmp_path):
    """What handles share across threads (per-device delivery count, handle count, environment switches, last-error
    string): four threads, one handle each, on the simulator in random-progress mode, built with -fsanitize=thread."""
    probe = subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", os.devnull], input="int main(){return 0;}", capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("no thread sanitizer runtime in this toolchain")
    exe, lib = _build_sim(str(tmp_path / "tsan"), ["-g", "-O1", "-fsanitize=thread"])
    p = subprocess.run([exe, lib, "threads"], capture_output=True, text=True, timeout=900, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    if "FATAL: ThreadSanitizer" in p.stderr and "unexpected memory mapping" in p.stderr:
        pytest.skip("thread sanitizer cannot map its shadow memory in this container")
    assert p.returncode == 0 and p.stdout.strip().endswith("OK") and "WARNING: ThreadSanitizer" not in p.stderr, p.stdout[-2000:] + p.stderr[-3000:]
    # tools/jm_dropin.cpp with four handles on four threads, copy helper threads included, in the same build
    out_dir = os.path.dirname(exe)
    objs = [os.path.join(out_dir, f) for f in os.listdir(out_dir) if f.endswith(".o") and f != "delivery_sim_test.cpp.o"]
    tool = str(tmp_path / "jm_dropin_tsan")
    _run(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=thread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "jm_dropin.cpp")] + objs +
         ["-rdynamic", "-ldl", "-lpthread", "-o", tool])
    for custom in ("device,pinned,0,-1,4", "host,pageable,2,2,4", "device,ref,2,0,3"):
        p = subprocess.run([tool, "--frames", "16", "--width", "1280", "--height", "720", "--pitch", "1280", "--custom", custom], capture_output=True, text=True,
                           timeout=900, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1", FAKE_CUDA_LAZINESS="1", FAKE_CUDA_SEED="4", FAKE_CUDA_DEVICES="2"))
        assert p.returncode == 0 and "WARNING: ThreadSanitizer" not in p.stderr, custom + "\n" + p.stdout[-2000:] + p.stderr[-3000:]


MUTATIONS = {
    # name: (text in jm_nv_dec.cu, replacement, scenario that must then fail)
    "staging buffer reused before its upload ran": (
        "if (c->stage_used[slot]) cudaEventSynchronize(c->stage_done[slot]);", ";", "raw"),
    "decoder surfaces unmapped before the launch that reads them ran": (
        "else if (cudaEventQuery(b.done) != cudaSuccess) { cudaGetLastError(); break; }", ";", "cuvid"),
    "ready_event of a device-pointer packet ignored": (
        "if (x.ready_event && cudaStreamWaitEvent(st, (cudaEvent_t)(uintptr_t)x.ready_event, 0) != cudaSuccess) { cudaGetLastError(); return -1; }", ";", "raw"),
    "delivery chunks copied out without waiting for them": (
        "auto wait_chunk = [&](int i) { if (cudaEventSynchronize(s.delivered[i]) != cudaSuccess) { cudaGetLastError(); failed = true; } };",
        "auto wait_chunk = [&](int i) { (void)i; (void)failed; };", "raw"),
    "upload surface written again while an unconverted frame still refers to it": (
        "if (it->pool_slot == slot) { it = c->pending.erase(it); c->dropped++; c->drop_flag = true; }", "if (false) { }", "alloc-failure"),
    "direct delivery started before the launch finished and never waited for": (
        "if (cudaEventRecord(s.direct, ds) != cudaSuccess || cudaEventSynchronize(s.direct) != cudaSuccess) {", "if (false) {", "raw"),
}


@pytest.fixture(scope="module")
def sim_plain(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("sim_plain"))
    _, lib = _build_sim(d, ["-g", "-O1"])
    return d, lib


@pytest.mark.parametrize("name", sorted(MUTATIONS))
def test_the_simulator_catches_injected_ordering_bugs(tmp_path, sim_plain, name):
    """The harness above is only worth something if it fails when the protocol is broken: each of these one-line
    removals of a synchronisation in jm_nv_dec.cu must make the named scenario fail."""
    old, new, scenario = MUTATIONS[name]
    text = open(os.path.join(CSRC, "jm_nv_dec.cu")).read()
    assert text.count(old) == 1, "the mutation no longer applies: update MUTATIONS"
    mutated = tmp_path / "jm_nv_dec_mutated.cu"
    mutated.write_text(text.replace(old, new))
    # everything but the mutated translation unit comes from the unsanitized build shared by these tests
    base_dir, lib = sim_plain
    o = str(tmp_path / "mutated.o")
    _run(["g++", "-std=c++17", "-g", "-O1"] + INCLUDES + ["-x", "c++", "-c", str(mutated), "-o", o])
    objs = [os.path.join(base_dir, f) for f in os.listdir(base_dir) if f.endswith(".o") and f != "jm_nv_dec.cu.o"] + [o]
    exe = str(tmp_path / "mutant")
    _run(["g++", "-g"] + objs + ["-rdynamic", "-ldl", "-lpthread", "-o", exe])
    p = subprocess.run([exe, lib, scenario], capture_output=True, text=True, timeout=900, env=dict(os.environ, SIM_FAIL_FAST="1"))
    assert p.returncode != 0 and "FAIL" in p.stdout, "mutation survived: " + name


REF_PROGRAM = "/root/reference/test_nv_dec/test_nv_dec.cpp"


@pytest.mark.skipif(not os.path.exists(REF_PROGRAM), reason="/root/reference is not mounted")
@pytest.mark.parametrize("laziness", [0, 1, 2])
def test_reference_test_program_on_the_simulator(sim_plain, tmp_path, laziness):
    """The reference's own test_nv_dec.cpp, unmodified and compiled in place, linked with the host layer on the CUDA
    simulator: its NAL splitter and decode / output loop (test_nv_dec.cpp:30-86,163-259) drive jm_nvdec_* to the end of
    a stream for the fake NVDEC library, whatever the timing of the streams underneath."""
    import re
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fake_stream as FS
    base_dir, lib = sim_plain
    o = str(tmp_path / "ref_program.o")
    shim = os.path.join(ROOT, "tests", "ref_driver", "shim")
    _run(["g++", "-std=gnu++11", "-fpermissive", "-w", "-O1", "-include", os.path.join(shim, "redirect.h"), "-I", shim,
          "-I", os.path.join(ROOT, "oracle", "ref_shim"), "-I", os.path.join(ROOT, "include"), "-c", REF_PROGRAM, "-o", o])
    objs = [os.path.join(base_dir, f) for f in os.listdir(base_dir) if f.endswith(".o") and f != "delivery_sim_test.cpp.o"] + [o]
    exe = str(tmp_path / "test_nv_dec_sim")
    _run(["g++"] + objs + ["-rdynamic", "-ldl", "-lpthread", "-o", exe])
    w, h, n = 320, 180, 23
    rng = np.random.default_rng(5)
    stream = np.concatenate([FS.sequence_header(w, h)] + [FS.picture(rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8)) for _ in range(n)])
    path = tmp_path / "stream.264"
    stream.tofile(path)
    env = dict(os.environ, JM_TEST_INPUT=str(path), JMC_NVCUVID_LIB=lib, FAKE_CUDA_LAZINESS=str(laziness), FAKE_CUDA_SEED="3")
    p = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert re.search(rf"Frame Count:\s+{n}\b", p.stdout) and re.search(rf"Display:\s+{w} x {h}", p.stdout), p.stdout
    assert "Pixel Format:\tYV12" in p.stdout


@pytest.mark.parametrize("laziness", [1, 2])
def test_the_measurement_tools_on_the_simulator(sim_plain, tmp_path, laziness):
    """tools/jm_dropin.cpp (every calling convention of the drop-in API, several handles on several threads; it compares
    the first frame of every handle with its own CPU loop) and tools/jm_streams.cpp (config 5: streams sharded over GPUs,
    one thread per GPU, static and dynamic assignment) use nothing but the public C API, so they run on the simulator
    too: functional coverage of what they otherwise only time."""
    import json
    base_dir, lib = sim_plain
    objs = [os.path.join(base_dir, f) for f in os.listdir(base_dir) if f.endswith(".o") and f != "delivery_sim_test.cpp.o"]
    env = dict(os.environ, FAKE_CUDA_LAZINESS=str(laziness), FAKE_CUDA_SEED="9", FAKE_CUDA_DEVICES="4")
    for tool in ("jm_dropin", "jm_streams"):
        exe = str(tmp_path / tool)
        _run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", tool + ".cpp")] + objs +
             ["-rdynamic", "-ldl", "-lpthread", "-o", exe])
    p = subprocess.run([str(tmp_path / "jm_dropin"), "--frames", "24", "--width", "322", "--height", "180", "--pitch", "384"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    d = json.loads(p.stdout[p.stdout.index("{"):])
    assert len(d) >= 20 and all(isinstance(v, (int, float)) and v > 0 for k, v in d.items() if k.endswith("_fps")), d
    for mode, assign in (("e2e", "static"), ("d2h", "dynamic"), ("device", "static")):
        p = subprocess.run([str(tmp_path / "jm_streams"), "--gpus", "4", "--streams", "8", "--frames", "12", "--batch", "4", "--width", "322",
                            "--height", "180", "--pitch", "384", "--mode", mode, "--assign", assign], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
        assert json.loads(line)["frames_per_s"] > 0 if "frames_per_s" in json.loads(line) else True, line
