/*
 * tools/jm_dropin.cpp -- frames/s of the jm_nvdec_* drop-in API, called the way the reference's own program calls
 * it (test_nv_dec/test_nv_dec.cpp:215-218: jm_nvdec_decode_frame, then jm_nvdec_output_frame if got_frame, one
 * frame at a time, one handle per thread), in plain C++ on include/jm_nv_dec.h + include/jmc_cuda.h only.
 *
 *   tools/jm_dropin [--device D] [--frames N] [--width W --height H --pitch P] [--only NAME]
 *
 * Input side   host   : pageable packet = header + pitched NV12 surface bytes (the caller's malloc memory)
 *              device : device-resident surface (what cuvidMapVideoFrame yields), packet = header only
 * Output side  pageable (malloc, what test_nv_dec.cpp:207 passes) | pinned (jm_nvdec_memory_alloc_host) |
 *              registered (malloc + jm_nvdec_memory_register_host) | lazy (malloc + option lazy_pin) | ref (zero copy)
 * Options      delay = display delay, threads = copy helper threads, handles = concurrent handles (one thread each)
 *
 * Prints one JSON object: variant name -> frames/s (wall clock around the timed loop, all handles together).
 * The first frame of every handle is checked against a CPU restatement of nv_dec.cpp:798-820 written here.
 */
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jm_nv_dec.h"
#include "jmc_cuda.h"

struct Variant {
    const char *name;
    bool device_in;
    const char *out;        /* pageable | pinned | registered | lazy | ref */
    int delay, threads, handles;
};

static const Variant VARIANTS[] = {
    /* the reference's calling convention: pageable packet in, pageable frame out.  threads -1 = the library's
     * default (min(4, cores/4) copy helper threads, started on the first pageable copy) */
    { "host_packet_pageable_out_fps", false, "pageable", 0, -1, 1 },
    { "host_packet_pageable_out_1thread_fps", false, "pageable", 0, 0, 1 },
    { "host_packet_pageable_out_delay2_fps", false, "pageable", 2, -1, 1 },
    { "host_packet_pageable_out_delay2_1thread_fps", false, "pageable", 2, 0, 1 },
    { "host_packet_registered_fps", false, "registered", 0, -1, 1 },      /* packet buffers and out_buf registered explicitly */
    { "host_packet_registered_delay2_fps", false, "registered", 2, -1, 1 },
    /* behind NVDEC: device-resident surface in */
    { "device_surface_pageable_out_fps", true, "pageable", 0, -1, 1 },
    { "device_surface_pageable_out_1thread_fps", true, "pageable", 0, 0, 1 },
    { "device_surface_pageable_out_delay2_fps", true, "pageable", 2, -1, 1 },
    { "device_surface_pageable_out_delay2_1thread_fps", true, "pageable", 2, 0, 1 },
    { "device_surface_pageable_out_delay2_8threads_fps", true, "pageable", 2, 8, 1 },
    { "device_surface_lazy_pin_fps", true, "lazy", 0, -1, 1 },
    { "device_surface_lazy_pin_delay2_fps", true, "lazy", 2, -1, 1 },
    { "device_surface_registered_out_fps", true, "registered", 0, -1, 1 },
    { "device_surface_pinned_out_fps", true, "pinned", 0, -1, 1 },
    { "device_surface_pinned_out_delay2_fps", true, "pinned", 2, -1, 1 },
    { "device_surface_ref_fps", true, "ref", 0, -1, 1 },
    { "device_surface_ref_delay2_fps", true, "ref", 2, -1, 1 },
    { "device_surface_pinned_out_4_handles_fps", true, "pinned", 0, -1, 4 },
    { "device_surface_pinned_out_delay2_4_handles_fps", true, "pinned", 2, -1, 4 },
    { "device_surface_pageable_out_delay2_4_handles_fps", true, "pageable", 2, -1, 4 },
    { "device_surface_ref_delay2_4_handles_fps", true, "ref", 2, -1, 4 },
};

struct Options { int device = 0, frames = 400, width = 1920, height = 1080, pitch = 2048; std::string only, custom; bool verbose = false; };

static void fill(unsigned char *p, size_t n, unsigned seed)
{
    unsigned x = seed * 2654435761u + 12345u;
    for (size_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; p[i] = (unsigned char)(x >> 24); }
}

/* NV12 -> I420 as nv_dec.cpp:798-820 does it (check of the first frame only) */
static void cpu_i420(const unsigned char *s, int pitch, int w, int h, unsigned char *out)
{
    for (int y = 0; y < h; y++) memcpy(out + (size_t)y * w, s + (size_t)y * pitch, (size_t)w);
    const unsigned char *uv = s + (size_t)pitch * h;
    const int w2 = w >> 1, h2 = h >> 1;
    unsigned char *u = out + (size_t)w * h, *v = u + (size_t)w2 * h2;
    for (int y = 0; y < h2; y++)
        for (int x = 0; x < w2; x++) { u[y * w2 + x] = uv[(size_t)y * pitch + 2 * x]; v[y * w2 + x] = uv[(size_t)y * pitch + 2 * x + 1]; }
}

struct Result { long long frames = 0; double seconds = 0; std::string error; };

static void run_handle(const Options &o, const Variant &v, int tid, int warm, int timed, std::atomic<int> *go, Result *res)
{
    const int w = o.width, h = o.height, pitch = o.pitch;
    const size_t surf = (size_t)pitch * h * 3 / 2, need = (size_t)w * h * 3 / 2;
    const int NS = 4;
    jmc_ctx *ctx = nullptr;
    if (jmc_ctx_create(o.device, &ctx) != JMC_OK) { res->error = jmc_last_error(); return; }
    handle_nvdec dec = jm_nvdec_create_handle();
    jm_nvdec_set_device(o.device, dec);
    jm_nvdec_set_display_delay(v.delay, dec);
    if (v.threads >= 0) jm_nvdec_set_option("copy_threads", v.threads, dec);
    jm_nvdec_set_option("lazy_pin", !strcmp(v.out, "lazy") ? 1 : 0, dec);
    if (jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, 1, nullptr, 0, dec) != 0) { res->error = std::string("init: ") + jmc_last_error(); return; }

    std::vector<std::vector<unsigned char>> pkts(NS);
    std::vector<unsigned char> host_surf(surf);
    void *dsurf[NS] = {};
    std::vector<unsigned char> want(need);
    for (int i = 0; i < NS; i++) {
        fill(host_surf.data(), surf, 1000u * (unsigned)tid + (unsigned)i);
        if (i == 0) cpu_i420(host_surf.data(), pitch, w, h, want.data());
        jm_nvdec_raw_packet hd;
        memset(&hd, 0, sizeof(hd));
        hd.magic = JM_NVDEC_RAW_MAGIC; hd.width = w; hd.height = h; hd.pitch = pitch;
        if (v.device_in) {
            jmc_alloc_device(ctx, surf, &dsurf[i]);
            jmc_memcpy_h2d(ctx, dsurf[i], host_surf.data(), surf);
            hd.flags = JM_NVDEC_RAW_DEVICE_PTR;
            hd.device_ptr = (uint64_t)(uintptr_t)dsurf[i];
            pkts[i].resize(sizeof(hd));
            memcpy(pkts[i].data(), &hd, sizeof(hd));
        } else {
            pkts[i].resize(sizeof(hd) + surf);
            memcpy(pkts[i].data(), &hd, sizeof(hd));
            memcpy(pkts[i].data() + sizeof(hd), host_surf.data(), surf);
            if (!strcmp(v.out, "registered")) jm_nvdec_memory_register_host(pkts[i].data(), (int)pkts[i].size(), dec);
        }
    }
    unsigned char *out = nullptr;
    void *pinned = nullptr;
    if (!strcmp(v.out, "pinned")) { jm_nvdec_memory_alloc_host(&pinned, (int)need, dec); out = (unsigned char *)pinned; }
    else {
        out = (unsigned char *)malloc(need);                              /* test_nv_dec.cpp:207 */
        memset(out, 0, need);
        if (!strcmp(v.out, "registered") && jm_nvdec_memory_register_host(out, (int)need, dec) != 0) { res->error = "register failed"; return; }
    }
    const bool ref = !strcmp(v.out, "ref");
    long long fetched = 0;
    bool checked = false;
    auto step = [&](unsigned char *buf, int len) {
        int got = 0;
        jm_nvdec_decode_frame(buf, len, &got, dec);
        if (got == 1) {
            int n = (int)need;
            const unsigned char *p = out;
            int r = ref ? jm_nvdec_output_frame_ref(&p, &n, dec) : jm_nvdec_output_frame(out, &n, dec);
            if (r != (int)need) res->error = "output_frame returned " + std::to_string(r);
            if (!checked) { checked = true; if (memcmp(p, want.data(), need) != 0) res->error = "first frame differs from the CPU loop"; }
            fetched++;
        }
    };
    for (int k = 0; k < warm; k++) step(pkts[k % NS].data(), (int)pkts[k % NS].size());
    go->fetch_add(1);
    while (go->load() < v.handles) { }                                           /* all handles start together */
    const long long f0 = fetched;
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < timed; k++) step(pkts[(k + warm) % NS].data(), (int)pkts[(k + warm) % NS].size());
    auto t1 = std::chrono::steady_clock::now();
    res->frames = fetched - f0;
    res->seconds = std::chrono::duration<double>(t1 - t0).count();
    while (!jm_nvdec_is_exit(dec)) step(nullptr, 0);                      /* flush, test_nv_dec.cpp:232-246 */
    if (!strcmp(v.out, "registered")) {
        jm_nvdec_memory_unregister_host(out, dec);
        if (!v.device_in) for (int i = 0; i < NS; i++) jm_nvdec_memory_unregister_host(pkts[i].data(), dec);
    }
    if (pinned) jm_nvdec_memory_release_host(pinned, dec);
    jm_nvdec_deinit(dec);
    if (!pinned) free(out);
    for (int i = 0; i < NS; i++) if (dsurf[i]) jmc_free_device(ctx, dsurf[i]);
    jmc_ctx_destroy(ctx);
}

int main(int argc, char **argv)
{
    Options o;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { return i + 1 < argc ? argv[++i] : (char *)"0"; };
        if (a == "--device") o.device = atoi(val());
        else if (a == "--frames") o.frames = atoi(val());
        else if (a == "--width") o.width = atoi(val());
        else if (a == "--height") o.height = atoi(val());
        else if (a == "--pitch") o.pitch = atoi(val());
        else if (a == "--only") o.only = val();
        else if (a == "--custom") o.custom = val();        /* in,out,delay,threads,handles  e.g. device,ref,2,-1,4 */
        else if (a == "--verbose") o.verbose = true;
    }
    std::vector<Variant> variants(VARIANTS, VARIANTS + sizeof(VARIANTS) / sizeof(VARIANTS[0]));
    static char c_in[32], c_out[32], c_name[128];
    if (!o.custom.empty()) {
        int d = 0, t = -1, hn = 1;
        if (sscanf(o.custom.c_str(), "%31[^,],%31[^,],%d,%d,%d", c_in, c_out, &d, &t, &hn) != 5) { fprintf(stderr, "bad --custom\n"); return 2; }
        snprintf(c_name, sizeof(c_name), "custom_%s_%s_delay%d_threads%d_handles%d", c_in, c_out, d, t, hn);
        variants.clear();
        variants.push_back(Variant{ c_name, !strcmp(c_in, "device"), c_out, d, t, hn });
        o.only.clear();
    }
    printf("{");
    bool first = true;
    for (const Variant &v : variants) {
        if (!o.only.empty() && o.only != v.name) continue;
        std::vector<Result> res((size_t)v.handles);
        std::vector<std::thread> th;
        std::atomic<int> go{0};
        const int timed = v.device_in ? o.frames * 4 : o.frames;
        for (int t = 0; t < v.handles; t++) th.emplace_back(run_handle, std::cref(o), std::cref(v), t, 20, timed, &go, &res[(size_t)t]);
        for (auto &t : th) t.join();
        long long frames = 0;
        double secs = 0;
        std::string err;
        for (auto &r : res) {
            frames += r.frames; if (r.seconds > secs) secs = r.seconds; if (!r.error.empty()) err = r.error;
            if (o.verbose) fprintf(stderr, "%s: handle %ld: %lld frames in %.4f s = %.0f fps\n", v.name, (long)(&r - &res[0]), r.frames, r.seconds, r.seconds > 0 ? r.frames / r.seconds : 0.0);
        }
        printf("%s\"%s\": ", first ? "" : ", ", v.name);
        if (!err.empty()) printf("\"error: %s\"", err.c_str());
        else printf("%.1f", secs > 0 ? (double)frames / secs : 0.0);
        first = false;
        fflush(stdout);
    }
    printf("}\n");
    return 0;
}
