/*
 * tools/jm_streams.cpp -- whole-box driver in the reference's own language (C++), on the C-ABI only.
 *
 * BASELINE.json config 5: S independent synthetic 1080p streams of F frames, stream s handled by GPU
 * s % N, ONE host thread + ONE pinned ring per GPU, no collective.  It includes
 * nothing but include/jmc_cuda.h and links nothing but libjmcodec_b200.so: no CUDA headers, no torch.
 *
 *   g++ -O2 -std=c++17 -Iinclude tools/jm_streams.cpp -Ljmcodec_b200 -ljmcodec_b200 \
 *       -Wl,-rpath,'$ORIGIN/../jmcodec_b200' -lpthread -o tools/jm_streams
 *   tools/jm_streams [--gpus N] [--streams 32] [--frames 300] [--batch 30] [--mode e2e|device]
 *
 * mode e2e   : every frame travels pinned host -> HBM -> NV12->I420 kernel -> pinned host
 *              (jmc_pipeline_*, what a decoder-less caller of the library pays)
 * mode device: the surfaces of a stream are resident in HBM (what NVDEC would have produced);
 *              one launch per batch, tight frames stay on the device
 * mode d2h   : surfaces resident in HBM, tight frames delivered to pinned host memory (the reference's real data flow)
 * Every GPU thread enters its timed region through a barrier, so whole-box rates are concurrent rates.
 * Prints one JSON line with whole-box frames/s (wall clock over all GPU threads) and per-GPU rates.
 */
#include <pthread.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jmc_cuda.h"

struct Options {
    int gpus = 0, streams = 32, frames = 300, batch = 30, width = 1920, height = 1080, pitch = 2048;
    bool e2e = true;          /* host buffers both ways */
    bool d2h = false;         /* surfaces resident in HBM (as after NVDEC), tight frames delivered to pinned host memory */
    bool dynamic = false;     /* --assign dynamic: a GPU takes the next unprocessed stream when it is free, instead of s % nGPU */
};
static std::atomic<int> g_next_stream{0};
static pthread_barrier_t g_start;   /* every GPU enters its timed region together: whole-box rates are then concurrent rates */

struct GpuResult {
    int device = 0, streams = 0;
    long long frames = 0;
    double seconds = 0, device_ms = 0;
    std::string error;
};

static void fill_surface(unsigned char *p, size_t n, unsigned seed)
{
    unsigned x = seed * 2654435761u + 1;
    for (size_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; p[i] = (unsigned char)(x >> 24); }
}

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        int r_ = (call);                                                                             \
        if (r_ < 0) { res->error = std::string(#call) + ": " + jmc_last_error(); return; }          \
    } while (0)

static void gpu_worker_body(const Options &o, int device, int n_gpus, GpuResult *res, bool *at_barrier)
{
    res->device = device;
    jmc_ctx *ctx = nullptr;
    CHECK(jmc_ctx_create(device, &ctx));
    const size_t surf = (size_t)o.pitch * o.height * 3 / 2, tight = (size_t)jmc_tight_bytes(o.width, o.height);
    jmc_job shape;
    memset(&shape, 0, sizeof(shape));
    CHECK(jmc_job_nvdec(&shape, o.width, o.height, o.pitch, /*out_fmt=*/1));     /* NV12 -> I420, nv_dec.cpp:798-820 */
    shape.n_frames = o.batch;

    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    jmc_pipeline *pipe = nullptr;
    jmc_event *ev0 = nullptr, *ev1 = nullptr;
    CHECK(jmc_event_create(ctx, &ev0));
    CHECK(jmc_event_create(ctx, &ev1));
    /* one stream's worth of frames; every stream of this GPU reuses the buffers (content is irrelevant to the rate) */
    CHECK(jmc_alloc_host(ctx, surf * o.frames, 0, &h_in));
    fill_surface((unsigned char *)h_in, surf * (size_t)(o.frames < 8 ? o.frames : 8), (unsigned)device);
    for (int f = 8; f < o.frames; f++) memcpy((unsigned char *)h_in + f * surf, (unsigned char *)h_in + (f % 8) * surf, surf);
    if (o.e2e || o.d2h) {
        CHECK(jmc_alloc_host(ctx, tight * o.frames, 0, &h_out));
        CHECK(jmc_pipeline_create(ctx, &shape, surf, 3, &pipe));
    }
    if (o.d2h) {
        CHECK(jmc_alloc_device(ctx, surf * o.frames, &d_in));
        CHECK(jmc_memcpy_h2d(ctx, d_in, h_in, surf * o.frames));
    } else if (!o.e2e) {
        CHECK(jmc_alloc_device(ctx, surf * o.frames, &d_in));
        CHECK(jmc_alloc_device(ctx, tight * o.frames, &d_out));
        CHECK(jmc_memcpy_h2d(ctx, d_in, h_in, surf * o.frames));
    }

    auto run_stream = [&](bool timed) {
        for (int f0 = 0; f0 < o.frames; f0 += o.batch) {
            const int n = o.frames - f0 < o.batch ? o.frames - f0 : o.batch;
            if (o.e2e || o.d2h) {
                int slot = jmc_pipeline_submit(pipe, o.d2h ? nullptr : (unsigned char *)h_in + f0 * surf, o.d2h ? (unsigned char *)d_in + f0 * surf : nullptr,
                                               (unsigned char *)h_out + f0 * tight, nullptr, n);
                if (slot < 0) { res->error = jmc_last_error(); return; }
            } else {
                jmc_job j = shape;
                j.n_frames = n;
                j.surf.base = (unsigned char *)d_in + f0 * surf; j.surf.stride = surf;
                j.tight.base = (unsigned char *)d_out + f0 * tight; j.tight.stride = tight;
                if (jmc_convert(ctx, &j, nullptr) < 0) { res->error = jmc_last_error(); return; }
            }
            if (timed) res->frames += n;
        }
    };
    run_stream(false);                                    /* warm-up: one whole stream */
    if (pipe) CHECK(jmc_pipeline_drain(pipe));
    CHECK(jmc_ctx_sync(ctx));

    *at_barrier = true;
    pthread_barrier_wait(&g_start);
    const auto t0 = std::chrono::steady_clock::now();
    CHECK(jmc_event_record(ctx, ev0, o.e2e ? 1 : 0));
    if (o.dynamic) {
        /* the GPUs of one box do not all reach host memory at the same rate (this pool: GPUs 4-7 deliver 1.5x what GPUs 0-3
         * do, profiles/r2_host_link_8gpu.txt); with s % nGPU the box waits for its slowest GPU, with a shared queue of
         * streams every GPU works until the queue is empty */
        for (int s; (s = g_next_stream.fetch_add(1)) < o.streams && res->error.empty();) {
            run_stream(true);
            res->streams++;
        }
    } else {
        for (int s = device; s < o.streams && res->error.empty(); s += n_gpus) {      /* stream s lives on GPU s % N */
            run_stream(true);
            res->streams++;
        }
    }
    CHECK(jmc_event_record(ctx, ev1, (o.e2e || o.d2h) ? 2 : 0));
    float ms = 0;
    CHECK(jmc_event_elapsed_ms(ctx, ev0, ev1, &ms));
    if (pipe) CHECK(jmc_pipeline_drain(pipe));
    CHECK(jmc_ctx_sync(ctx));
    res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    res->device_ms = ms;

    if (pipe) jmc_pipeline_destroy(pipe);
    if (h_in) jmc_free_host(ctx, h_in);
    if (h_out) jmc_free_host(ctx, h_out);
    if (d_in) jmc_free_device(ctx, d_in);
    if (d_out) jmc_free_device(ctx, d_out);
    jmc_event_destroy(ctx, ev0);
    jmc_event_destroy(ctx, ev1);
    jmc_ctx_destroy(ctx);
}

/* a GPU whose set-up failed must still show up at the start barrier, or the others would wait for ever */
static void gpu_worker(const Options &o, int device, int n_gpus, GpuResult *res)
{
    bool at_barrier = false;
    gpu_worker_body(o, device, n_gpus, res, &at_barrier);
    if (!at_barrier) pthread_barrier_wait(&g_start);
}

int main(int argc, char **argv)
{
    Options o;
    for (int i = 1; i < argc; i++) {
        auto val = [&](int &dst) { if (i + 1 < argc) dst = atoi(argv[++i]); };
        if (!strcmp(argv[i], "--gpus")) val(o.gpus);
        else if (!strcmp(argv[i], "--streams")) val(o.streams);
        else if (!strcmp(argv[i], "--frames")) val(o.frames);
        else if (!strcmp(argv[i], "--batch")) val(o.batch);
        else if (!strcmp(argv[i], "--width")) val(o.width);
        else if (!strcmp(argv[i], "--height")) val(o.height);
        else if (!strcmp(argv[i], "--pitch")) val(o.pitch);
        else if (!strcmp(argv[i], "--assign") && i + 1 < argc) o.dynamic = !strcmp(argv[++i], "dynamic");
        else if (!strcmp(argv[i], "--mode") && i + 1 < argc) { const char *m = argv[++i]; o.e2e = !strcmp(m, "e2e"); o.d2h = !strcmp(m, "d2h"); }
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    const int have = jmc_device_count();
    if (have <= 0) { fprintf(stderr, "jm_streams: no CUDA device (%s); there is no CPU fallback\n", jmc_last_error()); return 1; }
    if (o.gpus <= 0 || o.gpus > have) o.gpus = have;
    if (o.batch < 1 || o.frames < 1 || o.streams < 1 || o.pitch < o.width) { fprintf(stderr, "bad geometry\n"); return 2; }

    pthread_barrier_init(&g_start, nullptr, (unsigned)o.gpus);
    std::vector<GpuResult> res(o.gpus);
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; g < o.gpus; g++) th.emplace_back(gpu_worker, std::cref(o), g, o.gpus, &res[g]);
    for (auto &t : th) t.join();
    (void)t0;
    long long frames = 0;
    double slowest = 0;
    for (auto &r : res) {
        if (!r.error.empty()) { fprintf(stderr, "GPU %d: %s\n", r.device, r.error.c_str()); return 1; }
        frames += r.frames;
        if (r.seconds > slowest) slowest = r.seconds;
    }
    printf("{\"tool\": \"jm_streams\", \"mode\": \"%s\", \"n_gpus\": %d, \"streams\": %d, \"frames_per_stream\": %d, \"batch\": %d, "
           "\"assign\": \"%s\", \"width\": %d, \"height\": %d, \"pitch\": %d, \"frames\": %lld, \"seconds_slowest_gpu\": %.6f, \"frames_per_s\": %.1f, \"per_gpu\": [",
           o.e2e ? "e2e" : (o.d2h ? "d2h" : "device"), o.gpus, o.streams, o.frames, o.batch, o.dynamic ? "dynamic" : "stream % n_gpus", o.width, o.height, o.pitch, frames, slowest, frames / slowest);
    for (size_t i = 0; i < res.size(); i++)
        printf("%s{\"device\": %d, \"streams\": %d, \"frames\": %lld, \"wall_s\": %.6f, \"device_ms\": %.3f, \"frames_per_s\": %.1f}", i ? ", " : "",
               res[i].device, res[i].streams, res[i].frames, res[i].seconds, res[i].device_ms, res[i].frames / (res[i].device_ms * 1e-3));
    printf("]}\n");
    return 0;
}
