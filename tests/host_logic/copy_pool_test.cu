/*
 * Host-only checks of the delivery code's CPU logic in jmcodec_b200/csrc/jm_nv_dec.cu -- the helper-thread
 * copy pool (phase-pipelined, sliced copies between pageable caller memory and the pinned rings) and the
 * arithmetic that decides which part of a caller buffer counts as registered.  No CUDA call is made: the
 * translation unit is included whole to reach its internal namespace, and only code that never touches the
 * device runs.  Built and run by tests/test_host_logic.py (no GPU needed).
 */
#include "../../jmcodec_b200/csrc/jm_nv_dec.cu"

#include <stdio.h>

#include <random>

static int g_fail = 0;
#define CHECK(cond, ...)                                                                     \
    do {                                                                                     \
        if (!(cond)) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); g_fail++; } \
    } while (0)

static void fill(std::vector<uint8_t> &v, uint32_t seed)
{
    std::mt19937 g(seed);
    for (auto &b : v) b = (uint8_t)g();
}

/* dst rows [0, rows) x width must equal src; everything else in dst must still be the sentinel */
static bool same(const copy_job &j, const std::vector<uint8_t> &dst, size_t dst_off, uint8_t sentinel)
{
    std::vector<uint8_t> want(dst.size(), sentinel);
    for (size_t r = 0; r < j.rows; r++) memcpy(&want[dst_off + r * j.dpitch], j.src + r * j.spitch, j.width);
    return want == dst;
}

static void test_copy_pool(int threads)
{
    copy_pool pool;
    pool.want_threads = threads;
    std::mt19937 g(1234 + threads);
    for (int it = 0; it < 60; it++) {
        /* sizes on both sides of the 512 KB / 8 rows threshold below which the caller copies alone */
        const size_t width = 1 + g() % 5000, rows = 1 + g() % (it % 3 == 0 ? 40 : 700);
        const size_t spitch = width + g() % 64, dpitch = (it & 1) ? width : width + g() % 64;
        const bool contiguous = it % 5 == 0;
        const size_t sp = contiguous ? width : spitch, dp = contiguous ? width : dpitch;
        std::vector<uint8_t> src(sp * rows + 7), dst(dp * rows + 64, 0xA5);
        fill(src, 99 + it);
        int n_phases = 1 + g() % 8;
        size_t phase_rows = (rows + n_phases - 1) / n_phases;
        if (it % 7 == 0) { n_phases = 0; phase_rows = 0; }                 /* "everything is there" */
        copy_job j = { dst.data() + 16, dp, src.data() + 3, sp, width, rows, phase_rows, n_phases };
        int waited = 0, last = -1;
        bool ordered = true;
        pool.run(j, [&](int ph) { ordered = ordered && ph == last + 1; last = ph; waited++; });
        CHECK(ordered, "phases released out of order (threads %d, it %d)", threads, it);
        CHECK(waited == (n_phases < 1 ? 1 : n_phases), "wait_phase called %d times for %d phases", waited, n_phases);
        CHECK(same(j, dst, 16, 0xA5), "copy mismatch: threads %d it %d width %zu rows %zu phases %d", threads, it, width, rows, n_phases);
    }
    /* a frame-sized job, as output_frame issues it: 4 KB units, chunked phases, last chunk partial */
    {
        const size_t total = 3110400, unit = 4096, rows = total / unit;
        std::vector<uint8_t> src(total), dst(total + 32, 0x5A);
        fill(src, 7);
        const int chunks = 5;
        const size_t cb = ((total + chunks - 1) / chunks + 4095) & ~(size_t)4095;
        copy_job j = { dst.data(), unit, src.data(), unit, unit, rows, cb / unit, chunks };
        pool.run(j, [](int) {});
        memcpy(dst.data() + rows * unit, src.data() + rows * unit, total % unit);
        CHECK(memcmp(dst.data(), src.data(), total) == 0, "frame-sized copy mismatch (threads %d)", threads);
        CHECK(dst[total] == 0x5A, "wrote past the frame");
    }
    CHECK((int)pool.workers.size() <= threads, "more workers (%zu) than asked for (%d)", pool.workers.size(), threads);
    if (threads > 0) CHECK(!pool.workers.empty(), "helper threads never started");
    /* stop, then use again: threads are restarted on demand */
    pool.shutdown();
    CHECK(pool.workers.empty(), "shutdown left workers");
    std::vector<uint8_t> src(2 << 20), dst(2 << 20, 0);
    fill(src, 5);
    copy_job j = { dst.data(), 4096, src.data(), 4096, 4096, (2u << 20) / 4096, 64, 8 };
    pool.copy(j);
    CHECK(src == dst, "copy after restart mismatch");
    pool.shutdown();
}

static void test_registered_interior()
{
    nvdec_b200 c;
    uint8_t *lo = nullptr, *hi = nullptr;
    const uintptr_t base = 0x7f0000001000ull;                              /* page aligned, never dereferenced */
    c.regs.push_back({ (void *)base, 1 << 20 });
    /* a buffer that starts up to a page before the registered range and ends up to a page after it */
    CHECK(registered_interior(&c, (void *)(base - 100), (1 << 20) + 200, &lo, &hi) && (uintptr_t)lo == base && (uintptr_t)hi == base + (1 << 20), "edges of 100 bytes");
    CHECK(registered_interior(&c, (void *)(base - 4096), (1 << 20) + 8192, &lo, &hi), "edges of exactly one page");
    CHECK(!registered_interior(&c, (void *)(base - 4097), (1 << 20) + 4097, &lo, &hi), "head edge of more than a page must not count");
    CHECK(!registered_interior(&c, (void *)base, (1 << 20) + 4097, &lo, &hi), "tail edge of more than a page must not count");
    /* a buffer inside the range: its own bounds come back */
    CHECK(registered_interior(&c, (void *)(base + 4096), 8192, &lo, &hi) && (uintptr_t)lo == base + 4096 && (uintptr_t)hi == base + 4096 + 8192, "buffer inside the range");
    /* disjoint */
    CHECK(!registered_interior(&c, (void *)(base + (2 << 20)), 65536, &lo, &hi), "disjoint buffer");
    c.regs.clear();
    CHECK(!registered_interior(&c, (void *)base, 65536, &lo, &hi), "nothing registered");
}

static void test_small_helpers()
{
    CHECK(written_bytes(0, 1920, 1080) == 3110400 && written_bytes(1, 1920, 1080) == 3110400, "even sizes write w*h*3/2");
    /* odd sizes: the reference's loops write fewer bytes than w*h*3/2 (nv_dec.cpp:792-796,807-818) */
    CHECK(written_bytes(0, 1919, 1079) == (size_t)1919 * 1079 + (size_t)539 * 1919, "NV12 odd size");
    CHECK(written_bytes(1, 1919, 1079) == (size_t)1919 * 1079 + 2 * (size_t)959 * 539, "I420 odd size");
    CHECK(written_bytes(1, 0, 0) == 0, "empty frame");
    setenv("JMC_TEST_ENV_INT", "7", 1);
    CHECK(env_int("JMC_TEST_ENV_INT", 1, 0, 10) == 7 && env_int("JMC_TEST_ENV_INT", 1, 0, 5) == 5 && env_int("JMC_TEST_ENV_INT", 1, 8, 10) == 8, "env_int clamps");
    unsetenv("JMC_TEST_ENV_INT");
    CHECK(env_int("JMC_TEST_ENV_INT", 3, 0, 10) == 3, "env_int default");
    CHECK(cuvid_codec_of(JM_NVDEC_CODEC_HEVC) == CUVID_CODEC_HEVC && cuvid_codec_of(12345) == CUVID_CODEC_H264, "codec map (nv_dec.cpp:295-333)");
}

int main()
{
    for (int t : { 0, 1, 3, 8 }) test_copy_pool(t);
    test_registered_interior();
    test_small_helpers();
    printf(g_fail ? "FAILED %d\n" : "OK\n", g_fail);
    return g_fail ? 1 : 0;
}
