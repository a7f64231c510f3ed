/*
 * The library's host layer (jm_nv_dec.cu, jmnv_enc.cu, jmc_runtime.cu -- compiled unchanged) driven through its public
 * C API on top of the CUDA runtime simulator (fake_cuda/) and an oracle-backed stand-in for the kernels
 * (fake_launch.cpp).  What this checks that the GPU tests cannot: the stream / event protocol of the delivery code
 * under the LEAST helpful legal timing (work that runs only when somebody waits for it, or at random moments), every
 * frame against the oracle, frees and unregistrations under pending work, leaks after deinit, allocation failures at
 * every allocation site, the caller's current device after every call.  Built with AddressSanitizer + UBSan by
 * tests/test_host_logic.py.  usage: delivery_sim_test <path of the fake nvcuvid .so> [scenario-filter]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "cuda_runtime.h"
#include "jm_nv_dec.h"
#include "jmc_cuda.h"
#include "jmnv_enc.h"
#include "../../oracle/jm_oracle.h"

extern std::atomic<int> g_fake_launches, g_fake_frames, g_fake_max_batch;
static int g_arm_kind = -1, g_arm_k = 0, g_arm_count = 1;      /* allocation failure(s) to arm right after the simulator reset */
static int g_arm_call = -1, g_arm_call_count = 1;             /* enqueue-call failure(s), same */

static std::atomic<int> g_fail{0}, g_checks{0};
static thread_local std::string g_ctx;
#define CHECK(cond, ...)                                                                                         \
    do {                                                                                                         \
        g_checks++;                                                                                              \
        if (!(cond)) { if (g_fail.load() < 40) { printf("FAIL [%s] line %d: ", g_ctx.c_str(), __LINE__); printf(__VA_ARGS__); printf("\n"); } g_fail++; \
                       if (getenv("SIM_FAIL_FAST")) { printf("FAILED (first failure)\n"); fflush(stdout); _exit(1); } }                     \
    } while (0)

struct geom { int w, h, pitch; };

static void random_bytes(uint8_t *p, size_t n, uint32_t seed)
{
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    for (size_t i = 0; i < n; i++) {
        if ((i & 7) == 0) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; }
        p[i] = (uint8_t)(x >> (8 * (i & 7)));
    }
}

static std::vector<uint8_t> surface(const geom &g, uint32_t seed)
{
    const size_t bytes = (size_t)g.pitch * g.h * 3 / 2;                   /* nv_dec.cpp:453 */
    std::vector<uint8_t> s(bytes, 0xCD);
    const size_t rows = g.pitch ? bytes / (size_t)g.pitch : 0;
    for (size_t y = 0; y < rows; y++) random_bytes(&s[y * g.pitch], (size_t)g.w, seed * 4099u + (uint32_t)y);
    return s;
}

static size_t written(int fmt, int w, int h) { return fmt == 0 ? (size_t)w * h + (size_t)(h >> 1) * w : (size_t)w * h + 2 * (size_t)(w >> 1) * (h >> 1); }

static std::vector<uint8_t> expected(const std::vector<uint8_t> &surf, const geom &g, int fmt)
{
    std::vector<uint8_t> out((size_t)g.w * g.h * 3 / 2 + 8, 0xA5);
    int len = g.w * g.h * 3 / 2;
    if (len > 0) jmo_nvdec_output_frame(surf.data(), g.pitch, g.w, g.h, fmt, 1, out.data(), &len);
    return out;
}

static void sim_clean(const fake_cuda_counts &base, bool streams_may_grow)
{
    for (auto &e : fake_cuda_take_errors()) CHECK(false, "simulator: %s", e.c_str());
    const fake_cuda_counts now = fake_cuda_live();
    CHECK(now.device == base.device, "device allocations leaked: %zu -> %zu", base.device, now.device);
    CHECK(now.pinned == base.pinned, "pinned allocations leaked: %zu -> %zu", base.pinned, now.pinned);
    CHECK(now.registered == base.registered, "host registrations leaked: %zu -> %zu", base.registered, now.registered);
    CHECK(now.events == base.events, "events leaked: %zu -> %zu", base.events, now.events);
    CHECK(now.streams == base.streams || (streams_may_grow && now.streams == base.streams + 1), "streams leaked: %zu -> %zu", base.streams, now.streams);
    for (int d = 0; d < 3; d++) CHECK(jm_nvdec_deliveries_in_flight(d) == 0, "device %d still counts %d deliveries in flight", d, jm_nvdec_deliveries_in_flight(d));
}

enum { IN_PAGEABLE, IN_PINNED, IN_REGISTERED, IN_DEVICE, IN_DEVICE_SYNC, IN_DEVICE_EVENT };
enum { OUT_PAGEABLE, OUT_PINNED, OUT_REGISTERED, OUT_LAZY_PIN, OUT_DEVICE, OUT_REF };

struct raw_cfg {
    geom g; int fmt, delay, in_kind, out_kind, threads, frames, device;
    bool tolerate_errors;           /* allocation-failure runs: frames may be missing, calls may fail */
};

/* One RAW-front-end session, the reference's calling loop (test_nv_dec.cpp:207-258): decode, fetch if a frame is
 * announced, flush at the end.  Returns the number of frames delivered. */
static int raw_session(const raw_cfg &c, unsigned seed);

static int run_raw(const raw_cfg &c, unsigned seed, int laziness)
{
    const int n_dev = c.device + 1 > 2 ? c.device + 1 : 2;
    fake_cuda_reset(seed, laziness, n_dev);
    cudaSetDevice(0);
    const fake_cuda_counts base = fake_cuda_live();
    if (g_arm_kind >= 0) fake_cuda_fail_alloc(g_arm_kind, g_arm_k, g_arm_count);
    if (g_arm_call >= 0) fake_cuda_fail_call(g_arm_call, g_arm_call_count);
    const int n = raw_session(c, seed);
    sim_clean(base, false);
    return n;
}

/* the session itself: usable from several threads at once (one handle each) */
static int raw_session(const raw_cfg &c, unsigned seed)
{
    cudaSetDevice(0);
    const geom &g = c.g;
    const size_t surf_bytes = (size_t)g.pitch * g.h * 3 / 2, tight = (size_t)g.w * g.h * 3 / 2, wr = written(c.fmt, g.w, g.h);
    handle_nvdec h = jm_nvdec_create_handle();
    CHECK(h != nullptr, "create_handle");
    if (!h) return 0;
    jm_nvdec_set_device(c.device, h);
    jm_nvdec_set_option("display_delay", c.delay, h);
    jm_nvdec_set_option("copy_threads", c.threads, h);
    if (c.out_kind == OUT_LAZY_PIN) jm_nvdec_set_option("lazy_pin", 1, h);
    int r = jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, c.fmt, nullptr, 0, h);
    if (r != 0) {
        CHECK(c.tolerate_errors, "init failed: %d (%s)", r, jmc_last_error());
        jm_nvdec_deinit(h);
        return 0;
    }
    /* caller-side buffers */
    const size_t pkt_bytes = sizeof(jm_nvdec_raw_packet_ex) + surf_bytes;
    uint8_t *pkt = nullptr, *pkt_base = nullptr;
    if (c.in_kind == IN_PINNED) { void *p = nullptr; jm_nvdec_memory_alloc_host(&p, (int)pkt_bytes, h); pkt = pkt_base = (uint8_t *)p; }
    else { pkt_base = (uint8_t *)malloc(pkt_bytes + 8192); pkt = pkt_base + 1234; }           /* unaligned inside the malloc block */
    if (c.in_kind == IN_REGISTERED) jm_nvdec_memory_register_host(pkt, (int)pkt_bytes, h);
    const size_t out_cap = tight + 8;
    uint8_t *out = nullptr, *out_base = nullptr;
    if (c.out_kind == OUT_PINNED) { void *p = nullptr; jm_nvdec_memory_alloc_host(&p, (int)out_cap, h); out = out_base = (uint8_t *)p; }
    else if (c.out_kind == OUT_DEVICE) { void *p = nullptr; cudaSetDevice(c.device); cudaMalloc(&p, out_cap); cudaSetDevice(0); out = out_base = (uint8_t *)p; }
    else { out_base = (uint8_t *)malloc(out_cap + 8192); out = out_base + 777; }
    /* buffers with less than 64 KB of whole pages inside are not worth a registration: the call says so (-1) */
    const bool out_registered = c.out_kind == OUT_REGISTERED && out && jm_nvdec_memory_register_host(out, (int)out_cap, h) == 0;
    if (c.out_kind == OUT_REGISTERED && !c.tolerate_errors) CHECK(out_registered == (out_cap >= (64u << 10) + 8192), "registration of a %zu-byte out_buf: %d", out_cap, (int)out_registered);
    CHECK((pkt && out) || c.tolerate_errors, "caller buffers");
    /* device-pointer inputs: a ring of surfaces, as a decoder has; a slot is rewritten only after the frame made from it was fetched */
    /* when frames may be lost the display pipeline can stall, and with it the moment a surface is free again: one surface per frame */
    const int n_dsurf = c.tolerate_errors ? c.frames : c.delay + 4;
    std::vector<uint8_t *> dsurf;
    cudaStream_t up = nullptr;
    cudaEvent_t up_ev = nullptr;
    uint8_t *up_stage = nullptr;
    if (c.in_kind >= IN_DEVICE) {
        cudaSetDevice(c.device);
        for (int i = 0; i < n_dsurf; i++) { void *p = nullptr; cudaMalloc(&p, surf_bytes ? surf_bytes : 1); dsurf.push_back((uint8_t *)p); }
        if (c.in_kind == IN_DEVICE_EVENT) {
            cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&up_ev, cudaEventDisableTiming);
            void *p = nullptr; cudaHostAlloc(&p, (surf_bytes ? surf_bytes : 1) * n_dsurf, cudaHostAllocDefault); up_stage = (uint8_t *)p;
        }
        cudaSetDevice(0);
    }
    bool caller_ok = pkt && out && (c.in_kind != IN_DEVICE_EVENT || up_stage);
    for (uint8_t *d : dsurf) caller_ok = caller_ok && d;
    std::vector<std::vector<uint8_t>> want;
    int delivered = 0, next_want = 0;
    bool order_ok = true;
    auto fetch = [&]() {
        int len = (int)tight;
        int rr;
        const uint8_t *got = nullptr;
        if (c.out_kind == OUT_REF) {
            const unsigned char *f = nullptr;
            rr = jm_nvdec_output_frame_ref(&f, &len, h);
            got = f;
        } else {
            if (c.out_kind != OUT_DEVICE) memset(out, 0xA5, out_cap);
            else cudaMemset(out, 0xA5, out_cap);
            rr = jm_nvdec_output_frame(out, &len, h);
            got = out;
        }
        int dev = -1; cudaGetDevice(&dev);
        CHECK(dev == 0, "output_frame left device %d current", dev);
        if (getenv("SIM_VERBOSE")) printf("  [%s] output_frame -> %d\n", g_ctx.c_str(), rr);
        if (rr < 0) { CHECK(c.tolerate_errors, "output_frame returned %d (%s)", rr, jmc_last_error()); return; }
        CHECK(rr == (int)tight && len == (int)tight, "output_frame returned %d, len %d, expected %zu", rr, len, tight);
        /* which frame is it?  In order, but frames may be missing when errors are tolerated */
        bool found = false;
        while (next_want < (int)want.size()) {
            const std::vector<uint8_t> &w = want[(size_t)next_want++];
            const size_t n = c.out_kind == OUT_REF ? wr : out_cap;           /* a ring slot only promises the written bytes */
            if (memcmp(got, w.data(), n) == 0) { found = true; break; }
            if (!c.tolerate_errors) break;
        }
        CHECK(found, "frame %d: bytes differ from the oracle (or frames out of order)", delivered);
        order_ok = order_ok && found;
        delivered++;
    };
    for (int f = 0; f < c.frames && caller_ok; f++) {
        const std::vector<uint8_t> s = surface(g, seed * 1000 + (unsigned)f);
        want.push_back(expected(s, g, c.fmt));
        jm_nvdec_raw_packet_ex x;
        memset(&x, 0, sizeof(x));
        x.base.magic = JM_NVDEC_RAW_MAGIC; x.base.width = g.w; x.base.height = g.h; x.base.pitch = g.pitch;
        int len;
        if (c.in_kind < IN_DEVICE) {
            memcpy(pkt, &x.base, sizeof(x.base));
            if (surf_bytes) memcpy(pkt + sizeof(x.base), s.data(), surf_bytes);
            len = (int)(sizeof(x.base) + surf_bytes);
        } else {
            uint8_t *d = dsurf[(size_t)f % dsurf.size()];
            x.base.flags = JM_NVDEC_RAW_DEVICE_PTR;
            x.base.device_ptr = (uint64_t)(uintptr_t)d;
            if (c.in_kind == IN_DEVICE_EVENT) {
                /* the surface is still being produced on another stream when the packet is handed over */
                uint8_t *st = up_stage + (size_t)(f % n_dsurf) * (surf_bytes ? surf_bytes : 1);
                if (surf_bytes) memcpy(st, s.data(), surf_bytes);
                cudaSetDevice(c.device);
                if (surf_bytes) cudaMemcpyAsync(d, st, surf_bytes, cudaMemcpyHostToDevice, up);
                cudaEventRecord(up_ev, up);
                cudaSetDevice(0);
                x.base.flags |= JM_NVDEC_RAW_WAIT_EVENT;
                x.ready_event = (uint64_t)(uintptr_t)up_ev;
            } else if (surf_bytes) {
                cudaMemcpy2D(d, surf_bytes, s.data(), surf_bytes, surf_bytes, 1, cudaMemcpyHostToDevice);      /* written at once */
            }
            if (c.in_kind == IN_DEVICE_SYNC) x.base.flags |= JM_NVDEC_RAW_SYNC;
            memcpy(pkt, &x, sizeof(x));
            len = (int)sizeof(x);
        }
        int got_frame = -1;
        r = jm_nvdec_decode_frame(pkt, len, &got_frame, h);
        int dev = -1; cudaGetDevice(&dev);
        CHECK(dev == 0, "decode_frame left device %d current", dev);
        CHECK(r == 0 || c.tolerate_errors, "decode_frame returned %d (%s)", r, jmc_last_error());
        if (getenv("SIM_VERBOSE")) printf("  [%s] frame %d: decode_frame -> %d, got_frame %d%s%s\n", g_ctx.c_str(), f, r, got_frame, r ? "  " : "", r ? jmc_last_error() : "");
        memset(pkt, 0x77, (size_t)len);                                    /* in_buf is the caller's again */
        if (c.in_kind == IN_DEVICE_SYNC && surf_bytes) cudaMemset(dsurf[(size_t)f % dsurf.size()], 0x33, surf_bytes);   /* ... and so is the surface */
        if (got_frame == 1) fetch();
    }
    for (int guard = 0; guard < c.frames + 50 && !jm_nvdec_is_exit(h); guard++) {
        int got_frame = 0;
        r = jm_nvdec_decode_frame(nullptr, 0, &got_frame, h);
        CHECK(r == 0 || c.tolerate_errors, "flush returned %d", r);
        if (got_frame == 1) fetch();
    }
    CHECK(jm_nvdec_is_exit(h), "the handle never reported the end of the stream");
    (void)order_ok;
    if (!c.tolerate_errors) {
        CHECK(delivered == c.frames, "%d of %d frames delivered", delivered, c.frames);
        CHECK(jm_nvdec_dropped_frames(h) == 0, "%d frames dropped", jm_nvdec_dropped_frames(h));
        CHECK(g_fake_max_batch >= 1, "no launch seen");
    }
    if (out_registered) CHECK(jm_nvdec_memory_unregister_host(out, h) == 0, "unregister out_buf");
    if (c.in_kind == IN_PINNED && pkt) jm_nvdec_memory_release_host(pkt, h);
    if (c.out_kind == OUT_PINNED && out) jm_nvdec_memory_release_host(out, h);
    jm_nvdec_deinit(h);                                                    /* releases what is still registered (IN_REGISTERED, lazy pin) */
    if (c.in_kind != IN_PINNED) free(pkt_base);
    if (c.out_kind == OUT_DEVICE) { if (out_base) cudaFree(out_base); }
    else if (c.out_kind != OUT_PINNED) free(out_base);
    for (uint8_t *d : dsurf) if (d) cudaFree(d);
    if (up) { cudaStreamDestroy(up); if (up_ev) cudaEventDestroy(up_ev); if (up_stage) cudaFreeHost(up_stage); }
    return delivered;
}

/* ---- NVDEC front-end against the fake library ------------------------------------------------------------------- */
static void put_nal(std::vector<uint8_t> &s, int type, const uint8_t *rbsp, size_t n, bool long_start)
{
    if (long_start) s.push_back(0);
    s.push_back(0); s.push_back(0); s.push_back(1);
    s.push_back((uint8_t)type);
    int zeros = 0;
    for (size_t i = 0; i < n; i++) {                                        /* H.264 emulation prevention */
        if (zeros >= 2 && rbsp[i] <= 3) { s.push_back(3); zeros = 0; }
        s.push_back(rbsp[i]);
        zeros = rbsp[i] == 0 ? zeros + 1 : 0;
    }
    s.push_back(0x80);
}

struct cuvid_cfg {
    std::vector<geom> formats;      /* one sequence per entry (pitch unused) */
    int pics_per_format, pics_per_packet, map_limit, parser_delay, delay, fmt, out_kind;
    bool expect_drops;
    int fault_what, fault_after, fault_count;       /* fake_nvcuvid_fail: 1 map, 2 decode, 3 create decoder */
    bool lossy;                                     /* pictures may vanish without being counted (decode / create failures, out of memory) */
};

#include <dlfcn.h>
typedef void (*fake_fail_fn)(int, int, int);
typedef void (*fake_stats_fn)(int *, int);

static void run_cuvid(const cuvid_cfg &c, unsigned seed, int laziness, const char *fake_lib)
{
    fake_cuda_reset(seed, laziness, 2);
    cudaSetDevice(0);
    const fake_cuda_counts base = fake_cuda_live();
    setenv("JMC_NVCUVID_LIB", fake_lib, 1);
    if (g_arm_kind >= 0) fake_cuda_fail_alloc(g_arm_kind, g_arm_k, g_arm_count);
    if (g_arm_call >= 0) fake_cuda_fail_call(g_arm_call, g_arm_call_count);
    void *fl = dlopen(fake_lib, RTLD_NOW | RTLD_LOCAL);                  /* the same library object the handle loads */
    fake_fail_fn fake_fail = fl ? (fake_fail_fn)dlsym(fl, "fake_nvcuvid_fail") : nullptr;
    fake_stats_fn fake_stats = fl ? (fake_stats_fn)dlsym(fl, "fake_nvcuvid_stats") : nullptr;
    CHECK(fake_fail && fake_stats, "the fake NVDEC library lacks its test hooks");
    if (!fake_fail || !fake_stats) return;
    fake_stats(nullptr, 1);
    fake_fail(c.fault_what, c.fault_after, c.fault_count);
    const bool lossy = c.lossy || c.fault_what != 0 || g_arm_kind >= 0 || g_arm_call >= 0;
    char num[16];
    snprintf(num, sizeof(num), "%d", c.parser_delay);
    setenv("JMC_NVDEC_PARSER_DELAY", num, 1);
    g_fake_launches = g_fake_frames = g_fake_max_batch = 0;
    handle_nvdec h = jm_nvdec_create_handle();
    jm_nvdec_set_option("map_limit", c.map_limit, h);
    jm_nvdec_set_option("display_delay", c.delay, h);
    int r = jm_nvdec_init(JM_NVDEC_CODEC_AVC, c.fmt, nullptr, 0, h);
    CHECK(r == 0 || g_arm_kind >= 0 || g_arm_call >= 0, "init with the fake NVDEC library failed: %d (%s)", r, jmc_last_error());
    if (r != 0) { jm_nvdec_deinit(h); fake_fail(0, 0, 0); dlclose(fl); sim_clean(base, true); return; }
    /* the stream and what must come out of it */
    std::vector<std::vector<uint8_t>> packets, want;
    std::vector<geom> want_geom;
    std::mt19937 rng(seed + 99);
    for (const geom &g : c.formats) {
        std::vector<uint8_t> cur;
        uint32_t wh[2] = { (uint32_t)g.w, (uint32_t)g.h };
        put_nal(cur, 0x67, (const uint8_t *)wh, 8, true);
        int in_packet = 0;
        for (int p = 0; p < c.pics_per_format; p++) {
            const geom tg = { g.w, g.h, g.w };
            std::vector<uint8_t> tight_nv12 = surface(tg, (uint32_t)rng());
            if (p % 5 == 0 && tight_nv12.size() > 64) memset(tight_nv12.data() + 20, 0, 40);      /* runs of zeros: escapes */
            want.push_back(expected(tight_nv12, tg, c.fmt));
            want_geom.push_back(tg);
            put_nal(cur, 0x65, tight_nv12.data(), tight_nv12.size(), (p & 1) == 0);
            if (++in_packet == c.pics_per_packet) { packets.push_back(cur); cur.clear(); in_packet = 0; }
        }
        if (!cur.empty()) packets.push_back(cur);
    }
    std::vector<uint8_t> out_buf_pageable;
    int delivered = 0, next_want = 0;
    uint8_t *pinned = nullptr;
    size_t pinned_cap = 0;
    for (const geom &g : c.formats) pinned_cap = std::max(pinned_cap, (size_t)g.w * g.h * 3 / 2 + 8);
    if (c.out_kind == OUT_PINNED) { void *p = nullptr; jm_nvdec_memory_alloc_host(&p, (int)pinned_cap, h); pinned = (uint8_t *)p; }
    out_buf_pageable.resize(pinned_cap);
    auto fetch = [&]() {
        int w = 0, hh = 0;
        jm_nvdec_stream_info(&w, &hh, h);
        uint8_t *out = pinned ? pinned : out_buf_pageable.data();
        memset(out, 0xA5, pinned_cap);
        int len = (int)pinned_cap;
        int rr = jm_nvdec_output_frame(out, &len, h);
        CHECK(rr > 0 || lossy, "output_frame returned %d (%s)", rr, jmc_last_error());
        if (rr <= 0) return;
        bool found = false;
        while (next_want < (int)want.size()) {
            const size_t i = (size_t)next_want++;
            const size_t n = (size_t)want_geom[i].w * want_geom[i].h * 3 / 2;
            if ((int)n == rr && memcmp(out, want[i].data(), n + 8 <= pinned_cap ? n + 8 : n) == 0) { found = true; break; }
            if (!c.expect_drops && !lossy) break;
        }
        CHECK(found, "picture %d: bytes differ from the oracle (or out of order)", delivered);
        delivered++;
    };
    bool saw_drop_report = false;
    for (auto &p : packets) {
        int got_frame = 0;
        std::vector<uint8_t> copy = p;                                      /* handed over, then scribbled on */
        r = jm_nvdec_decode_frame(copy.data(), (int)copy.size(), &got_frame, h);
        if (r != 0) saw_drop_report = true;
        CHECK(r == 0 || c.expect_drops || lossy, "decode_frame returned %d (%s)", r, jmc_last_error());
        memset(copy.data(), 0x77, copy.size());
        int dev = -1; cudaGetDevice(&dev);
        CHECK(dev == 0, "decode_frame left device %d current", dev);
        if (got_frame == 1) fetch();
    }
    for (int guard = 0; guard < (int)want.size() + 50 && !jm_nvdec_is_exit(h); guard++) {
        int got_frame = 0;
        if (jm_nvdec_decode_frame(nullptr, 0, &got_frame, h) != 0) saw_drop_report = true;
        if (got_frame == 1) fetch();
    }
    CHECK(jm_nvdec_is_exit(h), "the handle never reported the end of the stream");
    const int dropped = jm_nvdec_dropped_frames(h);
    if (lossy) {
        CHECK(delivered + dropped <= (int)want.size(), "delivered %d + dropped %d > %zu", delivered, dropped, want.size());
        if (c.fault_what == 1 && g_arm_kind < 0) CHECK(delivered + dropped == (int)want.size() && (dropped >= 1 || c.fault_count < 2), "map failures: delivered %d + dropped %d of %zu", delivered, dropped, want.size());
        CHECK(dropped == 0 || saw_drop_report, "pictures were dropped but no call reported it");
    } else if (!c.expect_drops) {
        CHECK(delivered == (int)want.size(), "%d of %zu pictures delivered", delivered, want.size());
        CHECK(dropped == 0, "%d pictures dropped", dropped);
    } else {
        CHECK(delivered + dropped == (int)want.size(), "delivered %d + dropped %d != %zu", delivered, dropped, want.size());
        CHECK(dropped == 0 || saw_drop_report, "pictures were dropped but no call reported it");
    }
    if (c.pics_per_packet >= 2 && c.map_limit >= 2 && !lossy) CHECK(g_fake_max_batch >= 2, "several pictures per packet were never converted by one launch (max batch %d)", g_fake_max_batch.load());
    CHECK(g_fake_max_batch <= c.map_limit, "a launch took %d pictures with a map limit of %d", g_fake_max_batch.load(), c.map_limit);
    if (pinned) jm_nvdec_memory_release_host(pinned, h);
    jm_nvdec_deinit(h);
    int st[6] = {};
    fake_stats(st, 0);
    CHECK(st[2] == 0 && st[0] == st[1], "decoder surfaces left mapped: %d maps, %d unmaps, %d mapped now", st[0], st[1], st[2]);
    fake_fail(0, 0, 0);
    dlclose(fl);
    sim_clean(base, true);
}

/* ---- encoder input API ------------------------------------------------------------------------------------------- */
static void run_nvenc(int fmt, const geom &g, unsigned seed, int laziness)
{
    fake_cuda_reset(seed, laziness, 2);
    cudaSetDevice(0);
    const fake_cuda_counts base = fake_cuda_live();
    handle_nvenc h = jm_nvenc_create_handle();
    jm_nvenc_set_device(1, h);
    nv_enc_param p;
    memset(&p, 0, sizeof(p));
    p.codec_id = JM_NVENC_CODEC_SURFACE_ONLY;
    p.src_width = g.w; p.src_height = g.h; p.in_fmt = fmt;
    int r = jm_nvenc_init(&p, h);
    CHECK(r == JM_NVENC_SUCCESS, "nvenc init %d (%s)", r, jmc_last_error());
    r = jm_nvenc_init(&p, h);                                               /* again on the live handle: starts over, leaks nothing */
    CHECK(r == JM_NVENC_SUCCESS, "nvenc re-init %d", r);
    const bool rgb = fmt == JM_NVENC_FMT_ARGB || fmt == JM_NVENC_FMT_ABGR;
    const size_t in_bytes = rgb ? (size_t)g.w * g.h * 4 : (size_t)g.w * g.h * 3 / 2;
    for (int f = 0; f < JM_NVENC_NUM_SURFACES + 2; f++) {
        std::vector<uint8_t> in(in_bytes);
        random_bytes(in.data(), in.size(), seed * 100 + (unsigned)f);
        const std::vector<uint8_t> in2 = in;                                /* enc_frame's buffer is scribbled on below */
        int got = -1;
        r = jm_nvenc_enc_frame(in.data(), (int)in.size(), &got, h);
        int dev = -1; cudaGetDevice(&dev);
        CHECK(dev == 0, "enc_frame left device %d current", dev);
        if (f >= JM_NVENC_NUM_SURFACES) {                                   /* every surface locked: the reference returns -1 (nv_enc.cpp:90-93) */
            CHECK(r == -1, "enc_frame with no free surface returned %d", r);
            CHECK(jm_nvenc_release_surface(h) == 0, "release_surface");
            r = jm_nvenc_enc_frame(in.data(), (int)in.size(), &got, h);     /* the oldest surface is free again */
        }
        CHECK(r == 0 && got == 0, "enc_frame returned %d got %d (%s)", r, got, jmc_last_error());
        memset(in.data(), 0x77, in.size());                                 /* consumed before return */
        void *d = nullptr; int pitch = 0, rows = 0;
        CHECK(jm_nvenc_peek_surface(&d, &pitch, &rows, h) == 0, "peek_surface");
        std::vector<uint8_t> want((size_t)pitch * rows, 0);
        if (rgb) { for (int y = 0; y < g.h; y++) memcpy(&want[(size_t)y * pitch], &in2[(size_t)y * g.w * 4], (size_t)g.w * 4); }   /* pitch honoured (documented fix of nv_enc.cpp:1096) */
        else jmo_nvenc_upload(in2.data(), fmt, g.w, g.h, want.data(), pitch);
        CHECK(memcmp(d, want.data(), want.size()) == 0, "surface %d differs from the oracle (fmt 0x%x)", f, fmt);
    }
    if (!rgb || true) {
        std::vector<uint8_t> small(in_bytes / 2 + 1, 1);
        int got = 0;
        r = jm_nvenc_enc_frame(small.data(), (int)small.size(), &got, h);
        if (fmt != JM_NVENC_FMT_YV12) CHECK(r == JM_NVENC_ERR_INVALID_PARAM, "short buffer accepted: %d", r);
    }
    jm_nvenc_deinit(h);
    sim_clean(base, false);
}

/* ---- several handles on one device, calls interleaved (the per-device cap on queued deliveries) ------------------- */
static void run_multi(unsigned seed, int laziness)
{
    fake_cuda_reset(seed, laziness, 2);
    cudaSetDevice(0);
    const fake_cuda_counts base = fake_cuda_live();
    const int N = 5, FRAMES = 22;
    const geom gs[N] = { { 640, 360, 640 }, { 64, 36, 64 }, { 1280, 720, 1280 }, { 199, 77, 256 }, { 640, 360, 768 } };
    const int delays[N] = { 0, 2, 1, 3, 0 }, outs[N] = { OUT_PAGEABLE, OUT_PINNED, OUT_REF, OUT_PAGEABLE, OUT_PINNED };
    handle_nvdec h[N];
    std::vector<std::vector<uint8_t>> want[N];
    size_t next[N] = {};
    uint8_t *out[N] = {};
    std::vector<uint8_t> pageable[N];
    for (int i = 0; i < N; i++) {
        h[i] = jm_nvdec_create_handle();
        jm_nvdec_set_option("display_delay", delays[i], h[i]);
        jm_nvdec_set_option("copy_threads", i == 0 ? 2 : 0, h[i]);
        CHECK(jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, i & 1, nullptr, 0, h[i]) == 0, "init handle %d", i);
        const size_t cap = (size_t)gs[i].w * gs[i].h * 3 / 2 + 8;
        if (outs[i] == OUT_PINNED) { void *p = nullptr; jm_nvdec_memory_alloc_host(&p, (int)cap, h[i]); out[i] = (uint8_t *)p; }
        else { pageable[i].resize(cap); out[i] = pageable[i].data(); }
    }
    auto fetch = [&](int i) {
        const size_t tight = (size_t)gs[i].w * gs[i].h * 3 / 2;
        int len = (int)tight + 8, rr;
        const uint8_t *got = out[i];
        if (outs[i] == OUT_REF) { const unsigned char *f = nullptr; rr = jm_nvdec_output_frame_ref(&f, &len, h[i]); got = f; }
        else { memset(out[i], 0xA5, tight + 8); rr = jm_nvdec_output_frame(out[i], &len, h[i]); }
        CHECK(rr == (int)tight, "handle %d: output_frame returned %d", i, rr);
        const size_t n = outs[i] == OUT_REF ? written(i & 1, gs[i].w, gs[i].h) : tight + 8;
        CHECK(next[i] < want[i].size() && rr > 0 && memcmp(got, want[i][next[i]].data(), n) == 0, "handle %d frame %zu differs", i, next[i]);
        next[i]++;
    };
    for (int f = 0; f < FRAMES; f++)
        for (int i = 0; i < N; i++) {
            const std::vector<uint8_t> s = surface(gs[i], seed * 77 + (unsigned)(f * N + i));
            want[i].push_back(expected(s, gs[i], i & 1));
            std::vector<uint8_t> pkt(sizeof(jm_nvdec_raw_packet) + s.size());
            jm_nvdec_raw_packet hd;
            memset(&hd, 0, sizeof(hd));
            hd.magic = JM_NVDEC_RAW_MAGIC; hd.width = gs[i].w; hd.height = gs[i].h; hd.pitch = gs[i].pitch;
            memcpy(pkt.data(), &hd, sizeof(hd));
            memcpy(pkt.data() + sizeof(hd), s.data(), s.size());
            int got = 0;
            CHECK(jm_nvdec_decode_frame(pkt.data(), (int)pkt.size(), &got, h[i]) == 0, "handle %d decode", i);
            if (got && (f + i) % 7 != 3) fetch(i);                          /* now and then a caller skips a frame: it is replaced, not leaked */
            else if (got) next[i]++;
        }
    for (int i = 0; i < N; i++) {
        for (int k = 0; k < FRAMES + 8 && !jm_nvdec_is_exit(h[i]); k++) { int got = 0; jm_nvdec_decode_frame(nullptr, 0, &got, h[i]); if (got) fetch(i); }
        CHECK(next[i] == want[i].size() && jm_nvdec_dropped_frames(h[i]) == 0, "handle %d: %zu of %zu frames, %d dropped", i, next[i], want[i].size(), jm_nvdec_dropped_frames(h[i]));
    }
    /* deinit in an order that leaves deliveries of other handles queued */
    for (int i : { 2, 0, 4, 1, 3 }) {
        if (outs[i] == OUT_PINNED) jm_nvdec_memory_release_host(out[i], h[i]);
        jm_nvdec_deinit(h[i]);
    }
    sim_clean(base, false);
}

/* ---- jmc_pipeline_*: the batch pipeline the bench's e2e leg uses --------------------------------------------------- */
static void run_pipeline(const geom &g, int fmt, int sub, int depth, int batches, bool device_input, unsigned seed, int laziness)
{
    fake_cuda_reset(seed, laziness, 2);
    cudaSetDevice(1);                                                       /* the caller sits on another device */
    const fake_cuda_counts base = fake_cuda_live();
    jmc_ctx *ctx = nullptr;
    CHECK(jmc_ctx_create(0, &ctx) == JMC_OK, "ctx");
    jmc_job shape;
    memset(&shape, 0, sizeof(shape));
    jmc_job_nvdec(&shape, g.w, g.h, g.pitch, fmt);
    shape.n_frames = sub;
    const size_t surf_bytes = (size_t)g.pitch * g.h * 3 / 2, tight = (size_t)g.w * g.h * 3 / 2;
    jmc_pipeline *p = nullptr;
    CHECK(jmc_pipeline_create(ctx, &shape, surf_bytes, depth, &p) == JMC_OK, "pipeline_create (%s)", jmc_last_error());
    const size_t n = (size_t)sub * batches;
    void *hin = nullptr, *hout = nullptr, *din = nullptr;
    jmc_alloc_host(ctx, surf_bytes * n, 0, &hin);
    jmc_alloc_host(ctx, tight * n + 8, 0, &hout);
    memset(hout, 0xA5, tight * n + 8);
    std::vector<std::vector<uint8_t>> want;
    for (size_t f = 0; f < n; f++) {
        const std::vector<uint8_t> s = surface(g, seed * 31 + (unsigned)f);
        memcpy((uint8_t *)hin + f * surf_bytes, s.data(), surf_bytes);
        want.push_back(expected(s, g, fmt));
    }
    if (device_input) { jmc_alloc_device(ctx, surf_bytes * n, &din); jmc_memcpy_h2d(ctx, din, hin, surf_bytes * n); memset(hin, 0x11, surf_bytes * n); }
    for (int b = 0; b < batches; b++) {
        const int nf = b == batches - 1 && sub > 1 ? sub - 1 : sub;       /* a short last batch */
        const int slot = jmc_pipeline_submit(p, device_input ? nullptr : (uint8_t *)hin + (size_t)b * sub * surf_bytes,
                                             device_input ? (uint8_t *)din + (size_t)b * sub * surf_bytes : nullptr,
                                             (uint8_t *)hout + (size_t)b * sub * tight, nullptr, nf);
        CHECK(slot >= 0 && slot < depth, "submit returned %d (%s)", slot, jmc_last_error());
        int dev = -1; cudaGetDevice(&dev);
        CHECK(dev == 1, "submit left device %d current", dev);
        if (b % 3 == 1) CHECK(jmc_pipeline_wait(p, slot) == JMC_OK, "wait");
    }
    CHECK(jmc_pipeline_drain(p) == JMC_OK, "drain");
    for (size_t f = 0; f < n; f++) {
        const bool skipped = sub > 1 && f == n - 1;                         /* the frame the short batch left out */
        if (skipped) { CHECK(((uint8_t *)hout)[f * tight] == 0xA5, "a frame beyond the short batch was written"); continue; }
        CHECK(memcmp((uint8_t *)hout + f * tight, want[f].data(), written(fmt, g.w, g.h)) == 0, "pipeline frame %zu differs", f);
    }
    CHECK(jmc_pipeline_d2h_bytes(p) > 0 && (device_input || jmc_pipeline_h2d_bytes(p) > 0), "byte counters");
    jmc_pipeline_destroy(p);
    if (din) jmc_free_device(ctx, din);
    jmc_free_host(ctx, hin);
    jmc_free_host(ctx, hout);
    jmc_ctx_destroy(ctx);
    sim_clean(base, false);
}

/* ---- malformed input ------------------------------------------------------------------------------------------------ */
static void run_fuzz(unsigned seed, int laziness)
{
    fake_cuda_reset(seed, laziness, 2);
    cudaSetDevice(0);
    const fake_cuda_counts base = fake_cuda_live();
    std::mt19937 rng(seed);
    handle_nvdec h = jm_nvdec_create_handle();
    CHECK(jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, 1, nullptr, 0, h) == 0, "init");
    const geom g = { 64, 36, 64 };
    std::vector<uint8_t> out((size_t)g.w * g.h * 3 / 2 + 8);
    std::vector<std::vector<uint8_t>> want;
    size_t next = 0;
    int good = 0;
    /* calls before any frame exists */
    int len = (int)out.size();
    CHECK(jm_nvdec_output_frame(out.data(), &len, h) == -1, "output_frame without a frame");
    CHECK(jm_nvdec_output_frame(nullptr, &len, h) == -1 && jm_nvdec_output_frame(out.data(), nullptr, h) == -1, "NULL arguments");
    int dummy = 0;
    CHECK(jm_nvdec_decode_frame(out.data(), 16, nullptr, h) == 0, "NULL got_frame is tolerated like the reference does");
    for (int it = 0; it < 400; it++) {
        const int kind = (int)(rng() % 8);
        std::vector<uint8_t> pkt;
        bool valid = false;
        jm_nvdec_raw_packet hd;
        memset(&hd, 0, sizeof(hd));
        hd.magic = JM_NVDEC_RAW_MAGIC; hd.width = g.w; hd.height = g.h; hd.pitch = g.pitch;
        const std::vector<uint8_t> s = surface(g, seed * 999 + (unsigned)it);
        switch (kind) {
        case 0: valid = true; break;
        case 1: hd.magic ^= 1u << (rng() % 32); break;                      /* wrong magic */
        case 2: hd.width = -(int)(rng() % 5000) - 1; break;                 /* negative size */
        case 3: hd.pitch = g.w - 1 - (int)(rng() % 60); break;              /* pitch below the width */
        case 4: hd.height = 0x7fffffff / 3; hd.pitch = 0x7fffffff; hd.width = 100; break;     /* byte count overflows an int */
        case 5: hd.flags = JM_NVDEC_RAW_DEVICE_PTR | JM_NVDEC_RAW_WAIT_EVENT; break;           /* claims an event, too short for one */
        case 6: hd.height = g.h * 50; break;                                /* payload much shorter than the header says */
        default: break;                                                     /* truncated below */
        }
        pkt.resize(sizeof(hd) + s.size());
        memcpy(pkt.data(), &hd, sizeof(hd));
        memcpy(pkt.data() + sizeof(hd), s.data(), s.size());
        int plen = (int)pkt.size();
        if (kind == 7) plen = 1 + (int)(rng() % (sizeof(hd) + 40));          /* cut inside the header or just after it (0 would mean end of stream) */
        if (kind == 5) plen = (int)sizeof(hd);
        if (valid) want.push_back(expected(s, g, 1));
        int got = -1;
        const int r = jm_nvdec_decode_frame(pkt.data(), plen, &got, h);
        CHECK(r == 0, "decode_frame returned %d for packet kind %d", r, kind);
        CHECK(valid ? got == 1 : got == 0, "packet kind %d: got_frame %d", kind, got);
        if (got == 1) {
            /* a buffer that is too small first: -2, *out_len untouched, the frame stays fetchable (nv_dec.cpp:773-776) */
            int small = (int)(rng() % ((size_t)g.w * g.h * 3 / 2));
            const int before = small;
            CHECK(jm_nvdec_output_frame(out.data(), &small, h) == -2 && small == before, "short out_buf");
            memset(out.data(), 0xA5, out.size());
            len = (int)out.size();
            const int rr = jm_nvdec_output_frame(out.data(), &len, h);
            CHECK(rr == g.w * g.h * 3 / 2 && next < want.size() && memcmp(out.data(), want[next].data(), out.size()) == 0, "valid frame %zu after %d malformed packets", next, it - good);
            next++; good++;
        }
    }
    (void)dummy;
    CHECK(next == want.size(), "%zu of %zu valid frames", next, want.size());
    jm_nvdec_deinit(h);
    /* calls on things that are not there */
    int got = 5;
    CHECK(jm_nvdec_decode_frame(out.data(), 10, &got, nullptr) == 0 && got == 0, "NULL handle");
    CHECK(jm_nvdec_is_exit(nullptr) == true && jm_nvdec_dropped_frames(nullptr) == -1 && jm_nvdec_deinit(nullptr) == -1, "NULL handle queries");
    handle_nvdec fresh = jm_nvdec_create_handle();                          /* never initialised */
    CHECK(jm_nvdec_decode_frame(out.data(), 10, &got, fresh) == 0 && got == 0, "uninitialised handle");
    len = 10;
    CHECK(jm_nvdec_output_frame(out.data(), &len, fresh) == -1, "uninitialised handle output");
    void *p = nullptr;
    CHECK(jm_nvdec_memory_alloc_host(&p, 100, fresh) == -1 && jm_nvdec_memory_register_host(out.data(), 100, fresh) == -1, "uninitialised handle memory calls");
    CHECK(jm_nvdec_set_option("no_such_option", 1, fresh) == -1 && jm_nvdec_set_option("display_delay", 99, fresh) == -1 && jm_nvdec_set_option("map_limit", 0, fresh) == -1, "bad options");
    CHECK(jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, 1, nullptr, 0, fresh) == 0, "init after all that");
    jm_nvdec_set_device(7, fresh);                                          /* refused on a live handle */
    jm_nvdec_deinit(fresh);
    handle_nvdec bad = jm_nvdec_create_handle();
    jm_nvdec_set_device(7, bad);
    CHECK(jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, 1, nullptr, 0, bad) == -3, "device id beyond the device count (nv_dec.cpp:227-231)");
    jm_nvdec_deinit(bad);
    sim_clean(base, false);
    /* no device at all (nv_dec.cpp:219-222) */
    fake_cuda_reset(seed, laziness, 0);
    handle_nvdec none = jm_nvdec_create_handle();
    CHECK(jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, 1, nullptr, 0, none) == -2, "no CUDA device");
    CHECK(jm_nvdec_is_hw_support() == false, "is_hw_support without a device");
    jm_nvdec_deinit(none);
    fake_cuda_reset(seed, laziness, 2);
}

/* ---- the jmc_* runtime layer: contexts, memory, events, timed launches, the link probe, error returns ---------------- */
static void run_runtime(unsigned seed, int laziness)
{
    fake_cuda_reset(seed, laziness, 3);
    cudaSetDevice(2);
    const fake_cuda_counts base = fake_cuda_live();
    jmc_ctx *ctx = nullptr;
    CHECK(jmc_device_count() == 3 && jmc_current_device() == 2, "device count / current device");
    CHECK(jmc_ctx_create(5, &ctx) == JMC_ERR_NO_DEVICE && ctx == nullptr, "bad device id");
    CHECK(jmc_ctx_create(0, nullptr) == JMC_ERR_INVALID, "NULL out");
    CHECK(jmc_ctx_create(1, &ctx) == JMC_OK && ctx, "ctx_create (%s)", jmc_last_error());
    CHECK(jmc_current_device() == 2, "ctx_create left device %d current", jmc_current_device());
    CHECK(jmc_ctx_device(ctx) == 1 && jmc_ctx_sm_count(ctx) == 148 && jmc_ctx_stream(ctx, 0) && jmc_ctx_stream(ctx, 2) && !jmc_ctx_stream(ctx, 3), "ctx queries");
    const geom g = { 322, 180, 384 };
    const size_t surf_bytes = (size_t)g.pitch * g.h * 3 / 2, tight = (size_t)g.w * g.h * 3 / 2;
    const int n = 4;
    void *d_in = nullptr, *d_out = nullptr, *h = nullptr, *dp = nullptr;
    size_t pitch = 0;
    CHECK(jmc_alloc_device(ctx, surf_bytes * n, &d_in) == JMC_OK && jmc_alloc_device(ctx, tight * n, &d_out) == JMC_OK, "alloc_device");
    CHECK(jmc_alloc_pitched(ctx, 100, 7, &dp, &pitch) == JMC_OK && pitch >= 100, "alloc_pitched");
    CHECK(jmc_alloc_host(ctx, tight * n, 1, &h) == JMC_OK, "alloc_host");
    CHECK(jmc_alloc_device(ctx, 16, nullptr) == JMC_ERR_INVALID && jmc_alloc_device(nullptr, 16, &dp) == JMC_ERR_INVALID, "NULL arguments");
    std::vector<std::vector<uint8_t>> want;
    for (int f = 0; f < n; f++) {
        const std::vector<uint8_t> s = surface(g, seed * 13 + (unsigned)f);
        want.push_back(expected(s, g, 1));
        CHECK(jmc_memcpy_h2d(ctx, (uint8_t *)d_in + (size_t)f * surf_bytes, s.data(), surf_bytes) == JMC_OK, "h2d");
    }
    CHECK(jmc_memset_device(ctx, d_out, 0xA5, tight * n) == JMC_OK, "memset");
    jmc_job j;
    memset(&j, 0, sizeof(j));
    CHECK(jmc_job_nvdec(&j, g.w, g.h, g.pitch, 1) == JMC_OK && jmc_job_nvdec(&j, g.w, g.h, g.w - 1, 1) == JMC_ERR_INVALID, "job filler");
    jmc_job_nvdec(&j, g.w, g.h, g.pitch, 1);
    j.n_frames = n; j.surf.base = d_in; j.surf.stride = surf_bytes; j.tight.base = d_out; j.tight.stride = tight;
    CHECK(jmc_job_algorithmic_bytes(&j) == (int64_t)(2 * tight), "algorithmic bytes of NV12->I420 = 3*w*h");
    jmc_event *e0 = nullptr, *e1 = nullptr;
    CHECK(jmc_event_create(ctx, &e0) == JMC_OK && jmc_event_create(ctx, &e1) == JMC_OK, "events");
    CHECK(jmc_event_record(ctx, e0, 0) == JMC_OK && jmc_event_record(ctx, e0, 3) == JMC_ERR_INVALID, "event_record");
    const uint64_t l0 = jmc_ctx_launch_count(ctx);
    CHECK(jmc_convert(ctx, &j, nullptr) == JMC_OK, "convert");
    CHECK(jmc_event_record(ctx, e1, 0) == JMC_OK, "event_record");
    float ms = -1.f;
    CHECK(jmc_event_elapsed_ms(ctx, e0, e1, &ms) == JMC_OK && ms >= 0.f, "elapsed");
    CHECK(jmc_memcpy_d2h(ctx, h, d_out, tight * n) == JMC_OK, "d2h");
    for (int f = 0; f < n; f++) CHECK(memcmp((uint8_t *)h + (size_t)f * tight, want[(size_t)f].data(), tight) == 0, "converted frame %d", f);
    CHECK(jmc_convert_timed(ctx, &j, 3, &ms) == JMC_OK && ms > 0.f && jmc_convert_timed(ctx, &j, 0, &ms) == JMC_ERR_INVALID, "convert_timed");
    CHECK(jmc_ctx_launch_count(ctx) == l0 + 4, "launch count");
    jmc_link_rates r;
    for (int mode : { 1, 2, 3, 7 })
        CHECK(jmc_link_probe(ctx, 1 << 16, 3, mode, &r) == JMC_OK && ((mode & 1) == 0 || r.h2d_gbs > 0) && ((mode & 2) == 0 || r.d2h_gbs > 0), "link probe mode %d", mode);
    CHECK(jmc_link_probe(ctx, 1 << 16, -3, 3, &r) == JMC_OK && r.h2d_gbs > 0 && r.d2h_gbs > 0, "link probe over a window");
    CHECK(jmc_link_probe(ctx, 0, 3, 3, &r) == JMC_ERR_INVALID && jmc_link_probe(ctx, 16, 3, 0, &r) == JMC_ERR_INVALID && jmc_link_probe(ctx, 16, 3, 3, nullptr) == JMC_ERR_INVALID, "link probe arguments");
    CHECK(jmc_current_device() == 2, "a jmc_* call left device %d current", jmc_current_device());
    jmc_event_destroy(ctx, e0); jmc_event_destroy(ctx, e1);
    jmc_free_device(ctx, d_in); jmc_free_device(ctx, d_out); jmc_free_device(ctx, dp); jmc_free_host(ctx, h);
    CHECK(jmc_ctx_destroy(ctx) == JMC_OK, "ctx_destroy");
    sim_clean(base, false);
    /* stream creation fails half way: the context leaves nothing behind */
    fake_cuda_reset(seed, laziness, 0);
    CHECK(jmc_device_count() == 0 && jmc_ctx_create(0, &ctx) == JMC_ERR_NO_DEVICE, "no device");
    fake_cuda_reset(seed, laziness, 2);
}

static const char *in_name[] = { "host-pageable", "host-pinned", "host-registered", "device", "device+sync", "device+event" };
static const char *out_name[] = { "pageable", "pinned", "registered", "lazy-pin", "device", "ref" };

int main(int argc, char **argv)
{
    if (argc < 2) { printf("usage: %s <fake nvcuvid .so> [filter]\n", argv[0]); return 2; }
    const char *fake_lib = argv[1];
    const char *filter = argc > 2 ? argv[2] : "";
    auto want_run = [&](const char *name) { return strstr(name, filter) != nullptr; };
    const geom geoms[] = { { 64, 36, 64 }, { 1366, 768, 1536 }, { 1280, 720, 1280 }, { 199, 77, 256 }, { 2, 2, 16 } };

    if (want_run("raw")) {
        for (int lazy = 0; lazy <= 2; lazy++)
            for (int in = 0; in <= IN_DEVICE_EVENT; in++)
                for (int out = 0; out <= OUT_REF; out++)
                    for (int delay : { 0, 2 }) {
                        const int gi = (in + out + delay) % 5;
                        const int fmt = (in + out) & 1;
                        char name[200];
                        snprintf(name, sizeof(name), "raw lazy=%d in=%s out=%s delay=%d %dx%d fmt=%d", lazy, in_name[in], out_name[out], delay, geoms[gi].w, geoms[gi].h, fmt);
                        g_ctx = name;
                        raw_cfg c = { geoms[gi], fmt, delay, in, out, (in + out) % 3 == 0 ? 3 : 0, 14, (in + out) & 1, false };
                        g_fake_max_batch = 0;
                        run_raw(c, 100u + (unsigned)(lazy * 1000 + in * 100 + out * 10 + delay), lazy);
                    }
        /* a display delay longer than the ring of upload surfaces / staging buffers (10): their reuse must wait for
         * uploads and launches that nobody has synchronised with yet */
        for (int lazy = 0; lazy <= 2; lazy++)
            for (int in : { (int)IN_PAGEABLE, (int)IN_PINNED, (int)IN_DEVICE }) {
                char name[100];
                snprintf(name, sizeof(name), "raw long delay lazy=%d in=%s", lazy, in_name[in]);
                g_ctx = name;
                raw_cfg c = { geoms[3], 1, 14, in, OUT_PAGEABLE, 0, 40, 0, false };
                run_raw(c, 900u + (unsigned)lazy, lazy);
            }
        /* full-size frames: payloads staged in pieces by the copy threads, deliveries in 512 KB chunks copied out behind
         * the DMA; both staging layouts */
        for (int lazy = 0; lazy <= 2; lazy++)
            for (int linear = 0; linear <= 1; linear++)
                for (int out : { (int)OUT_PAGEABLE, (int)OUT_REGISTERED }) {
                    char name[100];
                    snprintf(name, sizeof(name), "raw 1080p lazy=%d stage_linear=%d out=%s", lazy, linear, out_name[out]);
                    g_ctx = name;
                    setenv("JMC_NVDEC_STAGE_LINEAR", linear ? "1" : "0", 1);
                    raw_cfg c = { { 1920, 1080, 2048 }, 1, 0, linear ? IN_REGISTERED : IN_PAGEABLE, out, 3, 5, 0, false };
                    run_raw(c, 950u + (unsigned)lazy, lazy);
                    unsetenv("JMC_NVDEC_STAGE_LINEAR");
                }
        /* many seeds of the random-progress mode on the calling convention the reference uses */
        for (unsigned seed = 1; seed <= 40; seed++) {
            char name[100];
            snprintf(name, sizeof(name), "raw random-progress seed %u", seed);
            g_ctx = name;
            raw_cfg c = { geoms[seed % 5], (int)(seed & 1), (int)(seed % 4), (int)(seed % 6), (int)((seed / 2) % 6), (int)(seed % 3), 20, 1, false };
            run_raw(c, seed, 1);
        }
        /* a stream whose geometry grows: surfaces, staging buffers and ring slots are re-allocated under way */
        g_ctx = "raw geometry change";
        for (int lazy = 0; lazy <= 2; lazy++) {
            fake_cuda_reset(7, lazy, 2);
            const fake_cuda_counts base = fake_cuda_live();
            handle_nvdec h = jm_nvdec_create_handle();
            jm_nvdec_set_option("display_delay", 1, h);
            CHECK(jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, 1, nullptr, 0, h) == 0, "init");
            std::vector<std::vector<uint8_t>> want;
            std::vector<size_t> want_n;
            size_t next = 0;
            const geom seq[] = { { 64, 36, 64 }, { 64, 36, 64 }, { 640, 360, 640 }, { 640, 360, 768 }, { 64, 36, 128 }, { 1280, 720, 1280 }, { 1280, 720, 1280 }, { 64, 36, 64 } };
            std::vector<uint8_t> out(1280 * 720 * 3 / 2 + 8);
            auto fetch = [&]() {
                memset(out.data(), 0xA5, out.size());
                int len = (int)out.size();
                int rr = jm_nvdec_output_frame(out.data(), &len, h);
                CHECK(next < want.size() && rr == (int)want_n[next] && memcmp(out.data(), want[next].data(), want_n[next] + 8) == 0, "frame %zu after a geometry change", next);
                next++;
            };
            for (size_t f = 0; f < sizeof(seq) / sizeof(seq[0]); f++) {
                const std::vector<uint8_t> s = surface(seq[f], (uint32_t)f + 5);
                want.push_back(expected(s, seq[f], 1));
                want_n.push_back((size_t)seq[f].w * seq[f].h * 3 / 2);
                std::vector<uint8_t> pkt(sizeof(jm_nvdec_raw_packet) + s.size());
                jm_nvdec_raw_packet hd;
                memset(&hd, 0, sizeof(hd));
                hd.magic = JM_NVDEC_RAW_MAGIC; hd.width = seq[f].w; hd.height = seq[f].h; hd.pitch = seq[f].pitch;
                memcpy(pkt.data(), &hd, sizeof(hd));
                memcpy(pkt.data() + sizeof(hd), s.data(), s.size());
                int got = 0;
                CHECK(jm_nvdec_decode_frame(pkt.data(), (int)pkt.size(), &got, h) == 0, "decode");
                if (got) fetch();
            }
            for (int k = 0; k < 8 && !jm_nvdec_is_exit(h); k++) { int got = 0; jm_nvdec_decode_frame(nullptr, 0, &got, h); if (got) fetch(); }
            CHECK(next == want.size(), "%zu of %zu frames", next, want.size());
            /* init on the live handle, then straight to deinit with frames still inside */
            CHECK(jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, 0, nullptr, 0, h) == 0, "re-init");
            for (int f = 0; f < 3; f++) {
                const std::vector<uint8_t> s = surface(seq[2], 77);
                std::vector<uint8_t> pkt(sizeof(jm_nvdec_raw_packet) + s.size());
                jm_nvdec_raw_packet hd;
                memset(&hd, 0, sizeof(hd));
                hd.magic = JM_NVDEC_RAW_MAGIC; hd.width = seq[2].w; hd.height = seq[2].h; hd.pitch = seq[2].pitch;
                memcpy(pkt.data(), &hd, sizeof(hd));
                memcpy(pkt.data() + sizeof(hd), s.data(), s.size());
                int got = 0;
                jm_nvdec_decode_frame(pkt.data(), (int)pkt.size(), &got, h);
            }
            jm_nvdec_deinit(h);
            sim_clean(base, false);
        }
    }

    if (want_run("alloc-failure")) {
        /* every allocation site fails once: nothing crashes, nothing leaks, whatever is delivered is right */
        for (int kind = 0; kind < 3; kind++)
            for (int k = 0; k < 48; k++)
                for (int variant = 0; variant < 3; variant++) {
                    char name[100];
                    snprintf(name, sizeof(name), "alloc-failure kind %d at %d variant %d", kind, k, variant);
                    g_ctx = name;
                    raw_cfg c = { geoms[variant == 2 ? 2 : 0], 1, variant, variant == 1 ? IN_DEVICE : IN_PAGEABLE, variant == 1 ? OUT_PINNED : OUT_PAGEABLE, variant == 2 ? 2 : 0, 6, 0, true };
                    g_arm_kind = kind; g_arm_k = k; g_arm_count = 1;
                    run_raw(c, 5, 1);
                    g_arm_kind = -1;
                }
        /* the same for the set-up calls of the pipeline and of the encoder-input API: all or nothing */
        for (int kind : { 0, 2 })
            for (int k = 0; k < 16; k++) {
                char name[100];
                snprintf(name, sizeof(name), "alloc-failure in set-up calls, kind %d at %d", kind, k);
                g_ctx = name;
                fake_cuda_reset(9, 1, 2);
                cudaSetDevice(0);
                const fake_cuda_counts base = fake_cuda_live();
                fake_cuda_fail_alloc(kind, k, 1);
                jmc_ctx *ctx = nullptr;
                if (jmc_ctx_create(0, &ctx) == JMC_OK) {
                    jmc_job shape;
                    memset(&shape, 0, sizeof(shape));
                    jmc_job_nvdec(&shape, 64, 36, 64, 1);
                    shape.n_frames = 2;
                    jmc_pipeline *pl = nullptr;
                    const int r = jmc_pipeline_create(ctx, &shape, 64 * 36 * 3 / 2, 3, &pl);
                    CHECK((r == JMC_OK) == (pl != nullptr), "pipeline_create returned %d with pipeline %p", r, (void *)pl);
                    if (pl) jmc_pipeline_destroy(pl);
                    jmc_ctx_destroy(ctx);
                }
                handle_nvenc e = jm_nvenc_create_handle();
                nv_enc_param p;
                memset(&p, 0, sizeof(p));
                p.codec_id = JM_NVENC_CODEC_SURFACE_ONLY; p.src_width = 64; p.src_height = 36; p.in_fmt = JM_NVENC_FMT_YV12;
                const int r = jm_nvenc_init(&p, e);
                if (r != JM_NVENC_SUCCESS) {                                /* a failed init holds nothing; the handle can be initialised again */
                    const fake_cuda_counts mid = fake_cuda_live();
                    CHECK(mid.device == base.device && mid.streams == base.streams, "a failed jm_nvenc_init kept %zu device allocations / %zu streams", mid.device - base.device, mid.streams - base.streams);
                    CHECK(jm_nvenc_init(&p, e) == JM_NVENC_SUCCESS, "init after a failed init");
                }
                jm_nvenc_deinit(e);
                sim_clean(base, false);
            }
        /* an allocator that stays empty for a while: many failures in a row, then memory is back */
        for (int kind = 0; kind < 2; kind++)
            for (int k : { 3, 9, 14, 22 })
                for (int count : { 5, 40, 200 })
                    for (int lazy = 0; lazy <= 2; lazy++) {
                        char name[100];
                        snprintf(name, sizeof(name), "alloc-failure kind %d: %d in a row from %d lazy=%d", kind, count, k, lazy);
                        g_ctx = name;
                        raw_cfg c = { geoms[0], 1, (k & 1) * 2, IN_PAGEABLE, OUT_PAGEABLE, 0, 40, 0, true };
                        g_arm_kind = kind; g_arm_k = k; g_arm_count = count;
                        const int n = run_raw(c, 6, lazy);
                        g_arm_kind = -1; g_arm_count = 1;
                        if (count <= 5) CHECK(n >= 20, "only %d of 40 frames were delivered although just %d allocations failed", n, count);
                    }
    }

    if (want_run("debug-one")) {                                           /* SIM_VERBOSE=1 ./delivery_sim_test lib debug-one <k> <count> <variant> */
        const int k = argc > 3 ? atoi(argv[3]) : 0, count = argc > 4 ? atoi(argv[4]) : 1, variant = argc > 5 ? atoi(argv[5]) : 0;
        g_ctx = "debug-one";
        const int in = variant == 1 ? IN_DEVICE : (variant == 3 ? IN_PINNED : IN_PAGEABLE);
        const int out = variant == 1 ? OUT_PINNED : (variant == 2 ? OUT_REF : OUT_PAGEABLE);
        raw_cfg c = { geoms[variant == 2 ? 2 : 0], 1, variant == 0 ? 0 : 2, in, out, 0, 14, 0, true };
        g_arm_call = k; g_arm_call_count = count;
        run_raw(c, 8, (k + variant) % 3);
        g_arm_call = -1;
    }

    if (want_run("call-failure")) {
        /* an enqueue-type CUDA call (async copy, event record, stream wait, launch) fails once, or a few times in a row:
         * frames may be lost and calls may report it -- nothing crashes, leaks, stays mapped or comes out wrong */
        for (int k = 0; k < 120; k += (k < 40 ? 1 : 3))
            for (int count : { 1, 4 })
                for (int variant = 0; variant < 4; variant++) {
                    char name[100];
                    snprintf(name, sizeof(name), "call-failure at %d x%d variant %d", k, count, variant);
                    g_ctx = name;
                    const int in = variant == 1 ? IN_DEVICE : (variant == 3 ? IN_PINNED : IN_PAGEABLE);
                    const int out = variant == 1 ? OUT_PINNED : (variant == 2 ? OUT_REF : OUT_PAGEABLE);
                    raw_cfg c = { geoms[variant == 2 ? 2 : 0], 1, variant == 0 ? 0 : 2, in, out, 0, 14, 0, true };
                    g_arm_call = k; g_arm_call_count = count;
                    run_raw(c, 8, (k + variant) % 3);
                    g_arm_call = -1;
                }
        for (int k = 0; k < 200; k += (k < 40 ? 1 : 5))
            for (int count : { 1, 3 }) {
                char name[100];
                snprintf(name, sizeof(name), "cuvid call-failure at %d x%d", k, count);
                g_ctx = name;
                cuvid_cfg c = { { { 320, 180, 0 }, { 198, 102, 0 } }, 16, 4, 4, 2, 0, 1, (k & 1) ? OUT_PINNED : OUT_PAGEABLE, false, 0, 0, 0, true };
                g_arm_call = k; g_arm_call_count = count;
                run_cuvid(c, 800u + (unsigned)k, k % 3, fake_lib);
                g_arm_call = -1;
            }
    }

    if (want_run("call-failure")) {
        /* the same under the encoder-input API and the batch pipeline; the caller gives its buffers back right after an error */
        for (int k = 0; k < 40; k++)
            for (int fmt : { JM_NVENC_FMT_NV12, JM_NVENC_FMT_YV12, JM_NVENC_FMT_ARGB }) {
                char name[100];
                snprintf(name, sizeof(name), "nvenc call-failure at %d fmt 0x%x", k, fmt);
                g_ctx = name;
                fake_cuda_reset(3, k % 3, 2);
                cudaSetDevice(0);
                const fake_cuda_counts base = fake_cuda_live();
                handle_nvenc h = jm_nvenc_create_handle();
                nv_enc_param p;
                memset(&p, 0, sizeof(p));
                const geom g = { 322, 180, 0 };
                p.codec_id = JM_NVENC_CODEC_SURFACE_ONLY; p.src_width = g.w; p.src_height = g.h; p.in_fmt = fmt;
                if (jm_nvenc_init(&p, h) == JM_NVENC_SUCCESS) {
                    const bool rgb = fmt == JM_NVENC_FMT_ARGB;
                    const size_t in_bytes = rgb ? (size_t)g.w * g.h * 4 : (size_t)g.w * g.h * 3 / 2;
                    void *in = nullptr;
                    CHECK(jm_nvenc_memory_alloc_host(&in, (int)in_bytes, h) == 0, "pinned input buffer");
                    fake_cuda_fail_call(k, 2);
                    for (int f = 0; f < 5 && in; f++) {
                        std::vector<uint8_t> bytes(in_bytes);
                        random_bytes(bytes.data(), in_bytes, (uint32_t)(k * 10 + f));
                        memcpy(in, bytes.data(), in_bytes);
                        int got = 0;
                        const int r = jm_nvenc_enc_frame((const unsigned char *)in, (int)in_bytes, &got, h);
                        memset(in, 0x77, in_bytes);                          /* ours again, success or not */
                        if (r != 0) continue;
                        void *d = nullptr; int pitch = 0, rows = 0;
                        jm_nvenc_peek_surface(&d, &pitch, &rows, h);
                        std::vector<uint8_t> want((size_t)pitch * rows, 0);
                        if (rgb) { for (int y = 0; y < g.h; y++) memcpy(&want[(size_t)y * pitch], &bytes[(size_t)y * g.w * 4], (size_t)g.w * 4); }
                        else jmo_nvenc_upload(bytes.data(), fmt, g.w, g.h, want.data(), pitch);
                        CHECK(memcmp(d, want.data(), want.size()) == 0, "surface of frame %d differs although enc_frame succeeded", f);
                        jm_nvenc_release_surface(h);
                    }
                    if (in) jm_nvenc_memory_release_host(in, h);
                }
                jm_nvenc_deinit(h);
                sim_clean(base, false);
            }
        for (int k = 0; k < 60; k++) {
            char name[100];
            snprintf(name, sizeof(name), "pipeline call-failure at %d", k);
            g_ctx = name;
            fake_cuda_reset(4, k % 3, 2);
            cudaSetDevice(0);
            const fake_cuda_counts base = fake_cuda_live();
            jmc_ctx *ctx = nullptr;
            if (jmc_ctx_create(0, &ctx) != JMC_OK) continue;
            const geom g = { 322, 180, 384 };
            jmc_job shape;
            memset(&shape, 0, sizeof(shape));
            jmc_job_nvdec(&shape, g.w, g.h, g.pitch, 1);
            shape.n_frames = 2;
            const size_t surf_bytes = (size_t)g.pitch * g.h * 3 / 2, tight = (size_t)g.w * g.h * 3 / 2;
            jmc_pipeline *pl = nullptr;
            if (jmc_pipeline_create(ctx, &shape, surf_bytes, 2, &pl) == JMC_OK) {
                fake_cuda_fail_call(k, 1);
                for (int b = 0; b < 6; b++) {
                    void *hin = nullptr, *hout = nullptr;
                    jmc_alloc_host(ctx, surf_bytes * 2, 0, &hin);
                    jmc_alloc_host(ctx, tight * 2 + 8, 0, &hout);
                    if (!hin || !hout) { if (hin) jmc_free_host(ctx, hin); if (hout) jmc_free_host(ctx, hout); continue; }
                    std::vector<std::vector<uint8_t>> want;
                    for (int f = 0; f < 2; f++) {
                        const std::vector<uint8_t> s = surface(g, (uint32_t)(k * 100 + b * 2 + f));
                        memcpy((uint8_t *)hin + (size_t)f * surf_bytes, s.data(), surf_bytes);
                        want.push_back(expected(s, g, 1));
                    }
                    const int slot = jmc_pipeline_submit(pl, hin, nullptr, hout, nullptr, 2);
                    if (slot >= 0) {
                        CHECK(jmc_pipeline_wait(pl, slot) == JMC_OK, "wait");
                        for (int f = 0; f < 2; f++) CHECK(memcmp((uint8_t *)hout + (size_t)f * tight, want[(size_t)f].data(), tight) == 0, "batch %d frame %d differs although submit succeeded", b, f);
                    }
                    jmc_free_host(ctx, hin);                                /* given back at once, also after a failed submit */
                    jmc_free_host(ctx, hout);
                }
                jmc_pipeline_destroy(pl);
            }
            jmc_ctx_destroy(ctx);
            sim_clean(base, false);
        }
    }

    if (want_run("cuvid")) {
        struct { int ppp, map_limit, pdelay, delay; } v[] = { { 8, 8, 2, 0 }, { 8, 3, 2, 0 }, { 4, 8, 0, 2 }, { 1, 8, 2, 0 }, { 8, 1, 1, 1 }, { 3, 2, 4, 0 } };
        for (int lazy = 0; lazy <= 2; lazy++)
            for (size_t i = 0; i < sizeof(v) / sizeof(v[0]); i++) {
                char name[160];
                snprintf(name, sizeof(name), "cuvid lazy=%d pics/packet=%d map_limit=%d parser_delay=%d delay=%d", lazy, v[i].ppp, v[i].map_limit, v[i].pdelay, v[i].delay);
                g_ctx = name;
                cuvid_cfg c = { { { 320, 180, 0 } }, 48, v[i].ppp, v[i].map_limit, v[i].pdelay, v[i].delay, (int)(i & 1), i % 3 == 0 ? OUT_PINNED : OUT_PAGEABLE, false };
                run_cuvid(c, 300u + (unsigned)i, lazy, fake_lib);
                snprintf(name, sizeof(name), "cuvid format change lazy=%d pics/packet=%d map_limit=%d", lazy, v[i].ppp, v[i].map_limit);
                g_ctx = name;
                cuvid_cfg c2 = { { { 320, 180, 0 }, { 198, 102, 0 }, { 640, 360, 0 } }, 14, v[i].ppp, v[i].map_limit, v[i].pdelay, v[i].delay, 1, OUT_PAGEABLE, false };
                run_cuvid(c2, 400u + (unsigned)i, lazy, fake_lib);
            }
        for (int lazy = 0; lazy <= 2; lazy++) {
            g_ctx = "cuvid overflow";
            cuvid_cfg c = { { { 128, 72, 0 } }, 110, 8, 8, 2, 0, 1, OUT_PAGEABLE, true };
            run_cuvid(c, 500, lazy, fake_lib);
        }
        /* the decoder library itself fails: maps, decodes, decoder creation (first sequence / after a format change) */
        for (int lazy = 0; lazy <= 2; lazy++)
            for (int what = 1; what <= 3; what++)
                for (int after : { 0, 1, 5, 17 })
                    for (int count : { 1, 3 }) {
                        if (what == 3 && (after > 1 || count > 1)) continue;
                        char name[120];
                        snprintf(name, sizeof(name), "cuvid fault what=%d after=%d count=%d lazy=%d", what, after, count, lazy);
                        g_ctx = name;
                        cuvid_cfg c = { { { 320, 180, 0 }, { 198, 102, 0 } }, 20, 1 + (after % 4) * 2, 1 + (after + count) % 8, 2, after & 1, 1, OUT_PAGEABLE, false, what, after, count, true };
                        run_cuvid(c, 600u + (unsigned)(what * 100 + after * 10 + count), lazy, fake_lib);
                    }
        /* device / pinned / event allocations fail under the NVDEC front-end */
        for (int kind = 0; kind < 3; kind++)
            for (int k = 0; k < 60; k += (k < 24 ? 1 : 4))
                for (int count : { 1, 6 }) {
                    char name[120];
                    snprintf(name, sizeof(name), "cuvid alloc-failure kind %d at %d x%d", kind, k, count);
                    g_ctx = name;
                    cuvid_cfg c = { { { 320, 180, 0 }, { 198, 102, 0 } }, 16, 4, 4, 2, 0, 1, (k & 1) ? OUT_PINNED : OUT_PAGEABLE, false, 0, 0, 0, true };
                    g_arm_kind = kind; g_arm_k = k; g_arm_count = count;
                    run_cuvid(c, 700u + (unsigned)k, 1, fake_lib);
                    g_arm_kind = -1; g_arm_count = 1;
                }
        for (unsigned seed = 1; seed <= 25; seed++) {
            char name[100];
            snprintf(name, sizeof(name), "cuvid random-progress seed %u", seed);
            g_ctx = name;
            cuvid_cfg c = { { { 200, 120, 0 }, { 64, 48, 0 } }, 20, 1 + (int)(seed % 8), 1 + (int)(seed % 8), (int)(seed % 5), (int)(seed % 3), (int)(seed & 1), OUT_PAGEABLE, false };
            run_cuvid(c, seed, 1, fake_lib);
        }
    }

    if (want_run("runtime")) {
        for (int lazy = 0; lazy <= 2; lazy++) {
            char name[100];
            snprintf(name, sizeof(name), "runtime lazy=%d", lazy);
            g_ctx = name;
            run_runtime(21, lazy);
        }
    }

    if (want_run("fuzz")) {
        for (int lazy = 0; lazy <= 2; lazy++)
            for (unsigned seed = 1; seed <= 4; seed++) {
                char name[100];
                snprintf(name, sizeof(name), "fuzz lazy=%d seed %u", lazy, seed);
                g_ctx = name;
                run_fuzz(seed, lazy);
            }
    }

    if (want_run("multi")) {
        for (int lazy = 0; lazy <= 2; lazy++)
            for (unsigned seed = 1; seed <= (lazy == 1 ? 12u : 2u); seed++) {
                char name[100];
                snprintf(name, sizeof(name), "multi-handle lazy=%d seed %u", lazy, seed);
                g_ctx = name;
                run_multi(seed, lazy);
            }
    }

    if (want_run("pipeline")) {
        for (int lazy = 0; lazy <= 2; lazy++)
            for (int depth = 1; depth <= 3; depth++)
                for (int dev_in = 0; dev_in <= 1; dev_in++) {
                    char name[100];
                    snprintf(name, sizeof(name), "pipeline lazy=%d depth=%d device_input=%d", lazy, depth, dev_in);
                    g_ctx = name;
                    run_pipeline(geoms[(depth + dev_in) % 2 == 0 ? 1 : 3], (depth + lazy) & 1, 3, depth, 7, dev_in != 0, 40u + (unsigned)depth, lazy);
                    run_pipeline(geoms[0], 1, 1, depth, 9, dev_in != 0, 50u + (unsigned)depth, lazy);
                }
    }

    if (want_run("threads")) {
        /* one handle per thread, four threads at once on one device (random-progress mode): built with ThreadSanitizer
         * by the test suite, this is the check of what handles share -- the per-device delivery count, the handle
         * count, the environment switches, the last-error string */
        g_ctx = "threads";
        fake_cuda_reset(11, 1, 2);
        cudaSetDevice(0);
        const fake_cuda_counts base = fake_cuda_live();
        std::vector<std::thread> ts;
        for (int t = 0; t < 4; t++)
            ts.emplace_back([t, &geoms] {
                for (int k = 0; k < 6; k++) {
                    char name[64];
                    snprintf(name, sizeof(name), "threads t=%d k=%d", t, k);
                    g_ctx = name;
                    raw_cfg c = { geoms[(t + k) % 5], (t + k) & 1, (t * 3 + k) % 4, (t + 2 * k) % 6, (2 * t + k) % 6, (t + k) % 3, 10, 0, false };
                    raw_session(c, 1000u + (unsigned)(t * 10 + k));
                }
            });
        for (auto &t : ts) t.join();
        g_ctx = "threads";
        sim_clean(base, false);
    }

    if (want_run("nvenc")) {
        for (int lazy = 0; lazy <= 2; lazy++)
            for (int fmt : { JM_NVENC_FMT_NV12, JM_NVENC_FMT_YV12, JM_NVENC_FMT_ARGB })
                for (const geom &g : { geom{ 64, 36, 0 }, geom{ 322, 180, 0 }, geom{ 1366, 768, 0 } }) {
                    char name[100];
                    snprintf(name, sizeof(name), "nvenc lazy=%d fmt=0x%x %dx%d", lazy, fmt, g.w, g.h);
                    g_ctx = name;
                    run_nvenc(fmt, g, 9, lazy);
                }
    }
    printf("%d checks, %d failed\n", g_checks.load(), g_fail.load());
    printf(g_fail.load() ? "FAILED\n" : "OK\n");
    return g_fail.load() ? 1 : 0;
}
