/*
 * jm_nv_dec.cu -- the jm_nvdec_* drop-in API (include/jm_nv_dec.h) over the jmc_* layer.
 *
 * Same calls, same return conventions as the reference's nv_dec/nv_dec.cpp; the surface-format path
 * behind them is batched and overlapped instead of synchronous:
 *
 *   reference                                         here
 *   ------------------------------------------------  ----------------------------------------------------
 *   decode_frame: parse -> display queue (:368-403)   decode_frame: front-end -> pending surfaces
 *     pop 1 frame, cuvidMapVideoFrame (:439)            ALL pending surfaces (up to the map limit) are mapped
 *     cuMemcpyDtoH pitch*h*3/2, SYNC (:452)             and converted by ONE launch (NV12 -> tight NV12 / I420 in
 *     ONE pinned buffer (MAX_OUTPUT_FRAMES 1)           HBM, pointer list as kernel arguments) into a RING of
 *                                                       tight frames; the D2H of every tight frame is enqueued at
 *                                                       once behind its kernel into a pinned ring slot (prefetch);
 *                                                       surfaces are unmapped when the launch's event has fired --
 *                                                       nothing waits per frame.  Deliveries run on their own
 *                                                       stream and are enqueued only when the launch that made
 *                                                       the frame has finished: no cross-stream semaphores
 *   output_frame: CPU strip/de-interleave (:782-820)  output_frame: waits for THAT frame's delivery events only;
 *                                                       pinned / registered out_buf: direct DMA of the tight frame;
 *                                                       pageable out_buf: chunk-pipelined copy out of the pinned ring
 *
 * An optional display delay (jm_nvdec_set_display_delay, like the parser's ulMaxDisplayDelay nv_dec.cpp:346)
 * lets frame k's delivery overlap the upload / decode / conversion of frames k+1.. .
 *
 * Front-ends:
 *   JM_NVDEC_CODEC_RAW_NV12  decoded surfaces as packets (host bytes or device pointers);
 *   bitstream codecs         NVDEC through libnvcuvid.so.1, bound at run time (dlopen): the same
 *                            parser/decoder callback structure as nv_dec.cpp:23-52,278-403,496-540.
 *                            If the library or an NVDEC engine is not available jm_nvdec_init fails (-4);
 *                            nothing falls back to a CPU decoder.
 */
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "cuvid_min.h"
#include "jm_nv_dec.h"
#include "jmc_internal.h"

#define NVDEC_MAX_FRAMES 10         /* decode surfaces / upload surfaces, nv_dec/nv_dec.h:32 */
#define MAX_LEN_DEC_INFO 1024       /* nv_dec/nv_dec.h:33 */
#define RING_MAX 64                 /* converted frames a handle holds at most (announced + not yet announced) */
#define MAX_CHUNKS 8                /* delivery events per frame (chunk-pipelined copy-out) */
#define MAP_LIMIT_MAX JMC_INLINE_LIST_MAX   /* decoder surfaces mapped at a time = frames per conversion launch */
#define MAX_DECODE_SURFACES 64
#define MAX_LAZY_REGS 8
#define COPY_THREADS_MAX 16
#define DELAY_MAX 20

namespace {

/* Deliveries (device-to-host copies of tight frames) queued on the copy engine, per device, over ALL handles of the
 * process, and the handles alive.  Measured on B200 (profiles/r2_dropin_multi_handle.txt): with more than about four
 * delivery copies queued at once -- e.g. four handles that each keep two in flight -- single handles fall, erratically,
 * into a state where every frame takes ~1.4 ms; with at most four queued the same handles run at the link rate,
 * stably.  So a handle enqueues its next delivery only while fewer than DEVICE_INFLIGHT_MAX are queued on its device
 * (or when the caller is waiting for that very frame). */
std::atomic<int> g_dev_inflight[64];
std::atomic<int> g_live_handles{0};
std::atomic<int> g_dev_inflight_max{-1};

int env_int(const char *name, int dflt, int lo, int hi)
{
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}

/* ---- host copies: calling thread + helper threads -------------------------------------------------- */
struct copy_job {
    uint8_t *dst; size_t dpitch;
    const uint8_t *src; size_t spitch;
    size_t width, rows;
    size_t phase_rows;              /* rows per phase; phase p may be copied once `ready` > p */
    int n_phases;
};

void copy_rows(const copy_job &j, size_t r0, size_t r1)
{
    if (r1 <= r0) return;
    if (j.width == j.spitch && j.width == j.dpitch) { memcpy(j.dst + r0 * j.width, j.src + r0 * j.width, (r1 - r0) * j.width); return; }
    for (size_t r = r0; r < r1; r++) memcpy(j.dst + r * j.dpitch, j.src + r * j.spitch, j.width);
}

/* A handle's helper threads for host copies between pageable caller memory and the pinned rings (the reference's
 * callers pass malloc'ed buffers, test_nv_dec.cpp:207).  A job is cut into PHASES (the chunks of a frame that is
 * still arriving by DMA) and every phase into one slice per thread: the threads are woken once per frame and
 * spin on `ready` while the next chunk is in flight, so the copy-out runs at several cores' memcpy rate right
 * behind the DMA.  Threads are started on first use; with none, everything runs on the calling thread. */
struct copy_pool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    copy_job job;
    int want_threads = 0;
    int n_parts = 0, next_part = 0, done_parts = 0;
    std::atomic<uint64_t> gen{0};
    std::atomic<bool> stop{false};
    std::atomic<int> ready{0};

    void run_part(const copy_job &j, int part, int np)
    {
        for (int ph = 0; ph < j.n_phases; ph++) {
            while (ready.load(std::memory_order_acquire) <= ph) __builtin_ia32_pause();
            const size_t r0 = (size_t)ph * j.phase_rows;
            const size_t r1 = r0 + j.phase_rows < j.rows ? r0 + j.phase_rows : j.rows;
            if (r1 > r0) copy_rows(j, r0 + (r1 - r0) * part / np, r0 + (r1 - r0) * (part + 1) / np);
        }
    }

    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            /* a streaming caller hands over the next copy within tens of microseconds: spin that long before
             * going to sleep, a condition-variable wake-up costs more than the wait */
            const auto t0 = std::chrono::steady_clock::now();
            while (gen.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed) &&
                   std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(150)) __builtin_ia32_pause();
            std::unique_lock<std::mutex> lk(m);
            cv_work.wait(lk, [&] { return stop.load() || (gen.load() != seen && next_part < n_parts); });
            if (stop.load()) return;
            seen = gen.load();
            if (next_part < n_parts) {
                const int part = next_part++;
                const copy_job j = job;
                const int np = n_parts;
                lk.unlock();
                run_part(j, part, np);
                lk.lock();
                if (++done_parts == n_parts) cv_done.notify_all();
            }
        }
    }

    void ensure_started()
    {
        try {
            for (int i = (int)workers.size(); i < want_threads; i++) workers.emplace_back([this] { worker(); });
        } catch (...) {
            want_threads = (int)workers.size();            /* no more threads to be had: copy with the ones we have */
        }
    }

    void shutdown()
    {
        { std::lock_guard<std::mutex> lk(m); stop.store(true); }
        cv_work.notify_all();
        for (auto &t : workers) t.join();
        workers.clear();
        stop.store(false);
    }

    /* wait_phase(p): called on the calling thread before phase p is released (NULL: everything is there). */
    template <class Wait> void run(copy_job j, Wait wait_phase)
    {
        if (j.n_phases < 1) { j.n_phases = 1; j.phase_rows = j.rows; }
        const size_t bytes = j.width * j.rows;
        if (want_threads > 0 && bytes >= (512u << 10) && j.rows >= 8) ensure_started();
        if (workers.empty() || bytes < (512u << 10) || j.rows < 8) {
            for (int ph = 0; ph < j.n_phases; ph++) {
                wait_phase(ph);
                const size_t r0 = (size_t)ph * j.phase_rows;
                copy_rows(j, r0, r0 + j.phase_rows < j.rows ? r0 + j.phase_rows : j.rows);
            }
            return;
        }
        std::unique_lock<std::mutex> lk(m);
        job = j;
        n_parts = (int)workers.size() + 1;
        next_part = 1;                                     /* part 0 is the caller's */
        done_parts = 0;
        ready.store(0, std::memory_order_release);
        gen.fetch_add(1, std::memory_order_release);
        lk.unlock();
        cv_work.notify_all();
        const int np = n_parts;
        for (int ph = 0; ph < j.n_phases; ph++) {
            wait_phase(ph);
            ready.store(ph + 1, std::memory_order_release);
            const size_t r0 = (size_t)ph * j.phase_rows;
            const size_t r1 = r0 + j.phase_rows < j.rows ? r0 + j.phase_rows : j.rows;
            if (r1 > r0) copy_rows(j, r0, r0 + (r1 - r0) / np);     /* slice 0 */
        }
        lk.lock();
        ++done_parts;
        cv_done.wait(lk, [&] { return done_parts == n_parts; });
    }

    void copy(const copy_job &j) { run(j, [](int) {}); }
};

/* ---- state --------------------------------------------------------------------------------------- */
struct decoded_surface {            /* what cuvidMapVideoFrame yields: device pointer + pitch */
    uint8_t *dptr;
    int pitch, width, height;
    int pool_slot;                  /* >= 0: one of our upload surfaces; -1: caller-owned device memory;
                                       -2: a CUVID picture still to be mapped (disp valid) */
    bool sync_consume;              /* JM_NVDEC_RAW_SYNC: do not return before the kernel has read the surface */
    CUVIDPARSERDISPINFO disp;       /* copied, not pointed to (the reference queues the parser's pointer, nv_dec.cpp:156) */
};

/* One converted frame: tight NV12 / I420 in device memory, optionally on its way into pinned host memory. */
struct ring_slot {
    uint8_t *d_tight = nullptr; size_t d_bytes = 0;
    uint8_t *h_tight = nullptr; size_t h_bytes = 0;
    cudaEvent_t converted = nullptr;        /* the launch that fills d_tight (recorded on the last slot of a batch) */
    int conv_slot = -1;                     /* ring slot whose `converted` event covers this slot's launch */
    cudaEvent_t direct = nullptr;           /* a direct D2H / D2D into the caller's buffer has finished */
    cudaEvent_t delivered[MAX_CHUNKS] = {};
    int n_chunks = 0;
    size_t chunk_bytes = 0, total = 0;      /* total: the bytes the reference writes for this frame */
    bool prefetched = false, busy = false;
    int w = 0, h = 0;
};

struct unmap_batch {
    cudaEvent_t done;
    CUvideodecoder decoder;
    int n;
    unsigned long long ptr[MAP_LIMIT_MAX];
    int pic[MAP_LIMIT_MAX];
};

struct cuvid_api {
    void *lib;
    tcuvidCreateVideoParser create_parser;
    tcuvidParseVideoData parse;
    tcuvidDestroyVideoParser destroy_parser;
    tcuvidCreateDecoder create_decoder;
    tcuvidDestroyDecoder destroy_decoder;
    tcuvidDecodePicture decode_picture;
    tcuvidMapVideoFrame64 map_frame;
    tcuvidUnmapVideoFrame64 unmap_frame;
};

struct lazy_reg { void *base; size_t len; };

struct nvdec_b200 {
    int device = 0;
    int codec_type = 0;
    int out_fmt = 0;                /* 0: NV12, else "YV12" = I420 (nv_dec.h:94) */
    bool inited = false, is_eof = false, is_exit = false;
    jmc_ctx *ctx = nullptr;

    /* decoded surfaces waiting for conversion (the display queue of nv_dec.h:88-91) */
    std::deque<decoded_surface> pending;

    /* upload surfaces standing in for the decoder's surfaces (RAW front-end, host payloads), used round-robin;
     * h_stage: pinned staging for pageable payloads (compact rows), stage_done: its H2D has been read */
    uint8_t *pool[NVDEC_MAX_FRAMES] = {};
    uint8_t *h_stage[NVDEC_MAX_FRAMES] = {};
    cudaEvent_t stage_done[NVDEC_MAX_FRAMES] = {};
    bool stage_used[NVDEC_MAX_FRAMES] = {};
    size_t pool_bytes = 0, stage_bytes = 0;
    int pool_next = 0;

    /* converted frames: ring + FIFO of announced-later frames + the current output frame (nv_dec.h:119-123) */
    std::vector<ring_slot> ring;
    std::deque<int> ready;
    std::deque<int> to_prefetch;    /* converted, delivery not enqueued yet (see flush_prefetch) */
    std::deque<int> inflight;       /* delivery enqueued, not yet seen complete */
    int max_inflight = 2;
    int cur = -1;
    int slot_rr = 0;
    int delay = 0;                  /* frames kept back before they are announced (display delay) */
    bool staged = true;             /* prefetch tight frames into the pinned ring (pageable out_buf callers) */
    int lazy_pin = 0;               /* cudaHostRegister a pageable out_buf / in_buf seen twice (opt-in) */
    const void *cand[MAX_LAZY_REGS] = {};    /* pageable buffers seen once (callers rotate over a few) */
    int cand_next = 0;
    std::vector<lazy_reg> regs;     /* registered by us: lazily or through jm_nvdec_memory_register_host */
    uint8_t *h_edge = nullptr;      /* pinned bounce buffer for the unregistered < 4 KB edges of such buffers: 2 x 4 KB out, 2 x 4 KB in */
    copy_pool copier;
    int copy_threads = 0;
    bool stage_linear = false;      /* pageable payloads: compact staging rows + 2-D H2D (JMC_NVDEC_STAGE_LINEAR=1: pitched staging + one linear H2D) */

    int disp_w = 0, disp_h = 0;     /* dec_create_info.ulTargetWidth/Height */
    uint32_t num_frames = 0, dropped = 0;
    bool drop_flag = false;
    struct timespec t_start = {};
    bool started = false;
    char dec_info[MAX_LEN_DEC_INFO] = "";

    /* NVDEC front-end (nv_dec.h:80-86) */
    cuvid_api nv = {};
    CUvideoparser parser = nullptr;
    CUvideodecoder decoder = nullptr;
    CUVIDEOFORMATEX parse_ext = {};
    int cuvid_codec = 0;
    bool decoder_failed = false;
    int map_limit = MAP_LIMIT_MAX, n_mapped = 0;
    int n_decode_surfaces = 0;
    int in_use[MAX_DECODE_SURFACES] = {};   /* displayed, not yet converted + unmapped (nv_dec.h:90 is_frame_in_use) */
    std::deque<unmap_batch> unmaps;
    std::vector<cudaEvent_t> free_events;
};

/* Two streams of the handle's context: uploads of host payloads, the decoder's post-processing and the conversion
 * launches are ordered on the CONVERT stream; every device-to-host copy runs on the DELIVERY stream and is enqueued by
 * the CPU only when the launch that produced the frame is known to have finished (flush_prefetch) -- so the kernel of
 * frame k+1 overlaps the delivery of frame k without a single cross-stream event wait on the device. */
cudaStream_t convert_stream(nvdec_b200 *c) { return (cudaStream_t)jmc_ctx_stream(c->ctx, 0); }
cudaStream_t delivery_stream(nvdec_b200 *c) { return (cudaStream_t)jmc_ctx_stream(c->ctx, 2); }

const char *codec_name(int t)
{
    switch (t) {                    /* nv_dec.cpp:629-661 */
    case JM_NVDEC_CODEC_AVC: return "H.264";
    case JM_NVDEC_CODEC_HEVC: return "H.265";
    case JM_NVDEC_CODEC_MJPEG: return "JPEG";
    case JM_NVDEC_CODEC_MPEG4: return "MPEG4";
    case JM_NVDEC_CODEC_MPEG2: return "MPEG2";
    case JM_NVDEC_CODEC_VP8: return "VP8";
    case JM_NVDEC_CODEC_VP9: return "VP9";
    case JM_NVDEC_CODEC_VC1: return "VC1";
    case JM_NVDEC_CODEC_RAW_NV12: return "RAW-NV12";
    default: return "UNKNOW";
    }
}

void show_info(nvdec_b200 *c)       /* nv_dec.cpp:663-683 */
{
    struct timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
    double ms = c->started ? (now.tv_sec - c->t_start.tv_sec) * 1e3 + (now.tv_nsec - c->t_start.tv_nsec) * 1e-6 : 0.0;
    snprintf(c->dec_info, MAX_LEN_DEC_INFO,
             "==========================================\n"
             "Codec:\t\t%s\n"
             "Display:\t%d x %d\n"
             "Pixel Format:\t%s\n"
             "Frame Count:\t%d\n"
             "Elapsed Time:\t%d ms\n"
             "Decode FPS:\t%f fps\n"
             "==========================================\n",
             codec_name(c->codec_type), c->disp_w, c->disp_h, c->out_fmt == 0 ? "NV12" : "YV12",
             (int)c->num_frames, (int)ms, ms > 0 ? c->num_frames * 1e3 / ms : 0.0);
}

void mark_started(nvdec_b200 *c)
{
    if (!c->started) { clock_gettime(CLOCK_MONOTONIC, &c->t_start); c->started = true; }   /* nv_dec.cpp:537 */
}

/* ---- events, ring slots -------------------------------------------------------------------------- */
cudaEvent_t get_event(nvdec_b200 *c)
{
    if (!c->free_events.empty()) { cudaEvent_t e = c->free_events.back(); c->free_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return e;
}

/* bytes of a tight frame the reference actually writes: for odd sizes fewer than w*h*3/2, the rest of
 * out_buf stays untouched (h>>1 chroma rows, w>>1 samples: nv_dec.cpp:792-796,807-818) */
size_t written_bytes(int out_fmt, int w, int h)
{
    const size_t luma = (size_t)w * h;
    return out_fmt == 0 ? luma + (size_t)(h >> 1) * w : luma + 2 * (size_t)(w >> 1) * (h >> 1);
}

/* A free ring slot able to hold a w x h frame, or -1 (ring full / out of memory).  Slots are taken round-robin
 * so that a slot whose delivery may still be in flight is reused as late as possible; the convert stream waits
 * for that delivery before the slot's device frame is overwritten. */
int acquire_slot(nvdec_b200 *c, int w, int h)
{
    size_t need = (size_t)jmc_tight_bytes(w, h);
    if (need == 0) need = 1;
    int idx = -1;
    const int n = (int)c->ring.size();
    for (int k = 0; k < n; k++) {
        const int i = (c->slot_rr + k) % n;
        if (!c->ring[i].busy) { idx = i; break; }
    }
    if (idx < 0) {
        if (n >= RING_MAX) return -1;
        c->ring.emplace_back();
        idx = n;
    }
    ring_slot &s = c->ring[idx];
    /* each event is created once; a slot whose events could not all be created is retried on its next use */
    auto have = [](cudaEvent_t &e) {
        if (e) return true;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess) return true;
        cudaGetLastError();
        e = nullptr;
        return false;
    };
    if (!have(s.converted) || !have(s.direct)) return -1;
    for (int i = 0; i < MAX_CHUNKS; i++)
        if (!have(s.delivered[i])) return -1;
    if (s.d_bytes < need) {                                   /* replaces cuMemAllocHost(pitch*h*3/2), nv_dec.cpp:569 */
        if (s.d_tight) { cudaFree(s.d_tight); s.d_tight = nullptr; s.d_bytes = 0; }   /* cudaFree waits for the device */
        if (cudaMalloc((void **)&s.d_tight, need) != cudaSuccess) { cudaGetLastError(); jmc_set_error("out of device memory for a %dx%d frame", w, h); return -1; }
        s.d_bytes = need;
        s.prefetched = false;
    }
    /* a frame nobody fetched: its delivery may still be reading d_tight (rare; fetched frames were waited for) */
    if (s.prefetched && s.n_chunks > 0 && cudaEventQuery(s.delivered[s.n_chunks - 1]) != cudaSuccess) {
        cudaGetLastError();
        cudaStreamWaitEvent(convert_stream(c), s.delivered[s.n_chunks - 1], 0);
    }
    s.prefetched = false;
    s.n_chunks = 0;
    s.w = w; s.h = h;
    s.total = written_bytes(c->out_fmt, w, h);
    s.busy = true;
    c->slot_rr = (idx + 1) % (int)c->ring.size();
    return idx;
}

void release_slot(nvdec_b200 *c, int idx)
{
    if (idx < 0) return;
    c->ring[idx].busy = false;
    for (auto it = c->inflight.begin(); it != c->inflight.end(); ++it)
        if (*it == idx) { c->inflight.erase(it); g_dev_inflight[c->device & 63].fetch_sub(1, std::memory_order_relaxed); break; }
    for (auto it = c->to_prefetch.begin(); it != c->to_prefetch.end(); ++it)
        if (*it == idx) { c->to_prefetch.erase(it); break; }              /* replaced before anyone fetched it */
}

/* Enqueue the D2H of a converted frame (its launch has finished) into the slot's pinned buffer on the delivery stream, in chunks with an
 * event each, so that output_frame can copy chunk i out while chunk i+1 is still crossing PCIe. */
bool prefetch_slot(nvdec_b200 *c, ring_slot &s)
{
    if (s.prefetched) return true;
    if (s.h_bytes < s.d_bytes) {
        if (s.h_tight) { cudaFreeHost(s.h_tight); s.h_tight = nullptr; s.h_bytes = 0; }
        if (cudaHostAlloc((void **)&s.h_tight, s.d_bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            jmc_set_error("out of pinned host memory for the delivery ring");
            return false;
        }
        s.h_bytes = s.d_bytes;
    }
    cudaStream_t ds = delivery_stream(c);
    /* no display delay: output_frame is waiting for this very frame, so pipeline its copy-out against the DMA;
     * with a delay the frame lands long before it is fetched and one copy is cheapest for the link */
    int chunks = c->delay > 0 ? 1 : (int)(s.total / (512u << 10));
    chunks = chunks < 1 ? 1 : (chunks > MAX_CHUNKS ? MAX_CHUNKS : chunks);
    const size_t cb = ((s.total + chunks - 1) / chunks + 4095) & ~(size_t)4095;
    s.chunk_bytes = cb ? cb : 4096;
    s.n_chunks = 0;
    for (size_t off = 0; off < s.total || s.n_chunks == 0; off += s.chunk_bytes) {
        const size_t n = s.total - off < s.chunk_bytes ? s.total - off : s.chunk_bytes;
        if (n && cudaMemcpyAsync(s.h_tight + off, s.d_tight + off, n, cudaMemcpyDeviceToHost, ds) != cudaSuccess) return false;
        if (cudaEventRecord(s.delivered[s.n_chunks], ds) != cudaSuccess) return false;
        s.n_chunks++;
        if (s.total == 0) break;
    }
    s.prefetched = true;
    return true;
}

/* Deliveries are enqueued by the CPU, on the delivery stream, only once the launch that produced the frame HAS
 * FINISHED (a non-blocking event query on the next API call; the call that wants the frame itself waits for the
 * event) -- never parked behind an unfinished kernel -- and only while few enough are queued: max_inflight (2) per
 * handle, DEVICE_INFLIGHT (4) per device over all handles of the process.  Both rules come from measurements
 * (profiles/r2_dropin_multi_handle.txt): with copies queued behind kernels through cross-stream events, or with
 * more than ~4 delivery copies queued on the device at once, four handles on one GPU ran anywhere between 3 k and
 * 17 k frames/s from run to run -- single handles dropping to ~700 frames/s, ~1.4 ms per frame, for the rest of the
 * run; with the two rules the same four handles deliver 16.5-16.9 k frames/s every time and one handle 17.2 k.
 * The kernel takes ~5 us and the next call comes tens of microseconds later, so the query is almost always positive
 * and frame k's copy runs while frame k+1 is converted.
 * wanted >= 0: that slot's delivery must be enqueued when this returns. */
bool flush_prefetch(nvdec_b200 *c, int wanted)
{
    std::atomic<int> &dev = g_dev_inflight[c->device & 63];
    for (;;) {
        /* forget the deliveries that have landed */
        while (!c->inflight.empty()) {
            ring_slot &f = c->ring[c->inflight.front()];
            if (f.prefetched && f.n_chunks > 0 && cudaEventQuery(f.delivered[f.n_chunks - 1]) == cudaErrorNotReady) { cudaGetLastError(); break; }
            c->inflight.pop_front();
            dev.fetch_sub(1, std::memory_order_relaxed);
        }
        if (c->to_prefetch.empty()) break;
        const int idx = c->to_prefetch.front();
        ring_slot &s = c->ring[idx];
        /* at most max_inflight deliveries queued per handle, DEVICE_INFLIGHT per device */
        if ((int)c->inflight.size() >= c->max_inflight || (!c->inflight.empty() && dev.load(std::memory_order_relaxed) >= g_dev_inflight_max.load(std::memory_order_relaxed))) {
            if (wanted < 0) return true;                                  /* later: the link is busy enough and nobody waits for this frame */
            ring_slot &f = c->ring[c->inflight.front()];
            if (cudaEventSynchronize(f.delivered[f.n_chunks - 1]) != cudaSuccess) { cudaGetLastError(); return false; }
            c->inflight.pop_front();
            dev.fetch_sub(1, std::memory_order_relaxed);
        } else if (c->inflight.empty() && dev.load(std::memory_order_relaxed) >= g_dev_inflight_max.load(std::memory_order_relaxed) && wanted < 0 && c->delay > 0) {
            return true;                                                  /* other handles fill the link; a display delay leaves slack */
        }
        cudaEvent_t ev = c->ring[s.conv_slot >= 0 ? s.conv_slot : idx].converted;
        cudaError_t q = cudaEventQuery(ev);
        if (q == cudaErrorNotReady) {
            cudaGetLastError();
            if (wanted < 0) return true;                                  /* later: nobody is waiting for it yet */
            q = cudaEventSynchronize(ev);
        }
        if (q != cudaSuccess) { cudaGetLastError(); return false; }
        c->to_prefetch.pop_front();
        if (!prefetch_slot(c, s)) return false;
        c->inflight.push_back(idx);
        dev.fetch_add(1, std::memory_order_relaxed);
        if (idx == wanted) wanted = -1;
    }
    return true;
}

/* ---- host memory classification / registration ----------------------------------------------------- */
enum { MEM_PAGEABLE = 0, MEM_PINNED = 1, MEM_DEVICE = 2 };

/* What kind of memory is [p, p+len)?  Both ends must agree; anything the copy engine cannot reach directly
 * (plain malloc memory, managed memory) counts as pageable. */
int host_kind_of(const void *p, size_t len)
{
    if (!p) return MEM_PAGEABLE;
    int kind = -1;
    cudaPointerAttributes a;
    for (int end = 0; end < 2; end++) {
        const void *q = end ? (const uint8_t *)p + (len ? len - 1 : 0) : p;
        if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return MEM_PAGEABLE; }
        const int k = a.type == cudaMemoryTypeHost ? MEM_PINNED : (a.type == cudaMemoryTypeDevice ? MEM_DEVICE : MEM_PAGEABLE);
        if (kind >= 0 && k != kind) return MEM_PAGEABLE;
        kind = k;
    }
    return kind;
}

bool is_device_accessible_host(const void *p, size_t len) { return host_kind_of(p, len) == MEM_PINNED; }

/* Page-lock the WHOLE PAGES INSIDE [p, p+len) -- never the partial pages at its ends: those may hold other
 * allocations of the caller, and a buffer that is only partly registered makes every CUDA copy from / to it fail
 * ("invalid argument"), also copies that have nothing to do with this library.  The < 4 KB edges of such a buffer
 * go through a small pinned bounce buffer (split_*). */
bool register_range(nvdec_b200 *c, const void *p, size_t len)
{
    const uintptr_t page = 4096, lo = ((uintptr_t)p + page - 1) & ~(page - 1), hi = ((uintptr_t)p + len) & ~(page - 1);
    if (hi <= lo || hi - lo < (64u << 10)) return false;                  /* not worth a registration */
    if ((int)c->regs.size() >= MAX_LAZY_REGS * 4) return false;
    if (cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return false; }
    c->regs.push_back({ (void *)lo, hi - lo });
    return true;
}

/* The part of [p, p+len) that this handle has registered, if it leaves at most 4 KB at either end. */
bool registered_interior(nvdec_b200 *c, const void *p, size_t len, uint8_t **lo, uint8_t **hi)
{
    const uintptr_t a = (uintptr_t)p, b = a + len;
    for (const lazy_reg &r : c->regs) {
        const uintptr_t l = a > (uintptr_t)r.base ? a : (uintptr_t)r.base;
        const uintptr_t h = b < (uintptr_t)r.base + r.len ? b : (uintptr_t)r.base + r.len;
        if (h > l && l - a <= 4096 && b - h <= 4096) { *lo = (uint8_t *)l; *hi = (uint8_t *)h; return true; }
    }
    return false;
}

bool ensure_edges(nvdec_b200 *c)
{
    if (c->h_edge) return true;
    if (cudaHostAlloc((void **)&c->h_edge, 4 * 4096, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); c->h_edge = nullptr; return false; }
    return true;
}

/* Opt-in (JMC_NVDEC_LAZY_PIN=1 / jm_nvdec_set_option): a pageable out_buf seen for the second time is page-locked in
 * place so that it receives frames by direct DMA (test_nv_dec.cpp:207 mallocs one buffer for the whole run).  Off by
 * default, and never applied to packet buffers: the registration outlives a free() of the buffer by the caller, which
 * the library cannot observe -- the virtual range then maps new pages while the DMA still targets the old ones. */
bool maybe_lazy_pin(nvdec_b200 *c, const void *p, size_t len)
{
    if (!c->lazy_pin) return false;
    for (int i = 0; i < MAX_LAZY_REGS; i++)
        if (c->cand[i] == p) { c->cand[i] = nullptr; return register_range(c, p, len); }
    c->cand[c->cand_next] = p;
    c->cand_next = (c->cand_next + 1) % MAX_LAZY_REGS;
    return false;
}

/* ---- NVDEC front-end ------------------------------------------------------------------------ */
int cuvid_codec_of(int t)           /* nv_dec.cpp:295-333 */
{
    switch (t) {
    case JM_NVDEC_CODEC_AVC: return CUVID_CODEC_H264;
    case JM_NVDEC_CODEC_HEVC: return CUVID_CODEC_HEVC;
    case JM_NVDEC_CODEC_MJPEG: return CUVID_CODEC_JPEG;
    case JM_NVDEC_CODEC_MPEG4: return CUVID_CODEC_MPEG4;
    case JM_NVDEC_CODEC_MPEG2: return CUVID_CODEC_MPEG2;
    case JM_NVDEC_CODEC_VP8: return CUVID_CODEC_VP8;
    case JM_NVDEC_CODEC_VP9: return CUVID_CODEC_VP9;
    case JM_NVDEC_CODEC_VC1: return CUVID_CODEC_VC1;
    default: return CUVID_CODEC_H264;                                    /* :330-332 */
    }
}

bool cuvid_load(nvdec_b200 *c)
{
    const char *name = getenv("JMC_NVCUVID_LIB");
    c->nv.lib = dlopen(name && *name ? name : "libnvcuvid.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!c->nv.lib) { jmc_set_error("cannot load the NVDEC library: %s", dlerror()); return false; }
    void *l = c->nv.lib;
    c->nv.create_parser = (tcuvidCreateVideoParser)dlsym(l, "cuvidCreateVideoParser");
    c->nv.parse = (tcuvidParseVideoData)dlsym(l, "cuvidParseVideoData");
    c->nv.destroy_parser = (tcuvidDestroyVideoParser)dlsym(l, "cuvidDestroyVideoParser");
    c->nv.create_decoder = (tcuvidCreateDecoder)dlsym(l, "cuvidCreateDecoder");
    c->nv.destroy_decoder = (tcuvidDestroyDecoder)dlsym(l, "cuvidDestroyDecoder");
    c->nv.decode_picture = (tcuvidDecodePicture)dlsym(l, "cuvidDecodePicture");
    c->nv.map_frame = (tcuvidMapVideoFrame64)dlsym(l, "cuvidMapVideoFrame64");
    c->nv.unmap_frame = (tcuvidUnmapVideoFrame64)dlsym(l, "cuvidUnmapVideoFrame64");
    if (!c->nv.create_parser || !c->nv.parse || !c->nv.destroy_parser || !c->nv.create_decoder || !c->nv.destroy_decoder ||
        !c->nv.decode_picture || !c->nv.map_frame || !c->nv.unmap_frame) {
        jmc_set_error("the NVDEC library lacks a required cuvid* entry point");
        return false;
    }
    /* Is there an engine behind it?  (cuvidGetDecoderCaps exists since Video Codec SDK 8; layout: cuvid_min.h) */
    tcuvidGetDecoderCaps caps = (tcuvidGetDecoderCaps)dlsym(l, "cuvidGetDecoderCaps");
    if (caps) {
        CUVIDDECODECAPS q;
        memset(&q, 0, sizeof(q));
        q.eCodecType = c->cuvid_codec; q.eChromaFormat = CUVID_CHROMA_420; q.nBitDepthMinus8 = 0;
        int r = caps(&q);
        if (r != 0 || !q.bIsSupported) {
            jmc_set_error("NVDEC is not usable here: cuvidGetDecoderCaps returned %d, supported=%d", r, (int)q.bIsSupported);
            return false;
        }
    }
    return true;
}

int convert_stage(nvdec_b200 *c, bool must_progress);
void reap_unmaps(nvdec_b200 *c, bool wait_all);

/* Everything that still refers to the current decoder -- displayed pictures not yet converted, mapped
 * surfaces not yet unmapped -- is converted into ring slots and released.  Frames the ring cannot take are
 * dropped (counted, reported by jm_nvdec_decode_frame). */
void cuvid_drain(nvdec_b200 *c)
{
    while (!c->pending.empty()) {
        if (convert_stage(c, true) <= 0) break;
    }
    while (!c->pending.empty()) {                                         /* ring full and nothing fetchable: give up on them */
        if (c->pending.front().pool_slot == -2) {
            const int idx = c->pending.front().disp.picture_index;
            if (idx >= 0 && idx < MAX_DECODE_SURFACES && c->in_use[idx] > 0) c->in_use[idx]--;
        }
        c->pending.pop_front();
        c->dropped++; c->drop_flag = true;
    }
    reap_unmaps(c, true);
}

/* nvdec_create_decoder, nv_dec.cpp:496-540, called from the parser on the caller's thread */
int cuvid_on_sequence(void *user, CUVIDEOFORMAT *f)
{
    nvdec_b200 *c = (nvdec_b200 *)user;
    if (c->decoder) {
        /* format change: pictures of the old decoder are still queued -- convert and unmap them first, the
         * indices mean nothing to the new decoder */
        cuvid_drain(c);
        c->nv.destroy_decoder(c->decoder);
        c->decoder = nullptr;
    }
    unsigned surfaces = NVDEC_MAX_FRAMES;                                 /* :526 */
    if (f->min_num_decode_surfaces > surfaces) surfaces = f->min_num_decode_surfaces;
    if (surfaces > MAX_DECODE_SURFACES) surfaces = MAX_DECODE_SURFACES;
    CUVIDDECODECREATEINFO ci;
    memset(&ci, 0, sizeof(ci));
    ci.CodecType = f->codec;
    ci.ChromaFormat = f->chroma_format;
    ci.OutputFormat = CUVID_SURFACE_NV12;                                 /* :507 */
    ci.DeinterlaceMode = f->progressive_sequence ? CUVID_DEINTERLACE_WEAVE : CUVID_DEINTERLACE_ADAPTIVE;   /* :508 */
    ci.bitDepthMinus8 = f->bit_depth_luma_minus8;
    ci.ulWidth = f->coded_width;                                          /* :510-511 */
    ci.ulHeight = f->coded_height;
    ci.ulTargetWidth = (unsigned long)(f->display_area.right - f->display_area.left);    /* :513-514 */
    ci.ulTargetHeight = (unsigned long)(f->display_area.bottom - f->display_area.top);
    ci.display_area.left = (short)f->display_area.left;
    ci.display_area.top = (short)f->display_area.top;
    ci.display_area.right = (short)f->display_area.right;
    ci.display_area.bottom = (short)f->display_area.bottom;
    ci.ulNumDecodeSurfaces = surfaces;
    /* the reference maps one surface at a time (:527); here up to map_limit displayed pictures are mapped
     * together and converted by one launch */
    ci.ulNumOutputSurfaces = (unsigned long)c->map_limit;
    ci.ulCreationFlags = CUVID_CREATE_PREFER_CUVID;                       /* :528 */
    ci.vidLock = nullptr;                                                 /* as the reference: never created (nv_dec.h:96) */
    int r = c->nv.create_decoder(&c->decoder, &ci);
    if (r != 0 || f->bit_depth_luma_minus8 != 0 || f->chroma_format != CUVID_CHROMA_420) {
        jmc_set_error("cuvidCreateDecoder failed (%d) or unsupported format (bit depth %d, chroma %d)", r,
                      8 + f->bit_depth_luma_minus8, f->chroma_format);
        if (c->decoder) { c->nv.destroy_decoder(c->decoder); c->decoder = nullptr; }
        c->decoder_failed = true;
        return 0;                                                         /* stop the parser */
    }
    c->decoder_failed = false;
    c->n_decode_surfaces = (int)surfaces;
    memset(c->in_use, 0, sizeof(c->in_use));
    c->n_mapped = 0;
    c->disp_w = (int)ci.ulTargetWidth;
    c->disp_h = (int)ci.ulTargetHeight;
    mark_started(c);                                                      /* :537 */
    return (int)surfaces;                                                 /* > 1: tells newer parsers the surface count */
}

int cuvid_on_decode(void *user, void *pic)                                /* nv_dec.cpp:33-41 */
{
    nvdec_b200 *c = (nvdec_b200 *)user;
    if (!c->decoder) return 0;
    /* The target surface may still hold a displayed picture that has not been converted yet (a packet with
     * many pictures): get it out of the decoder before it is overwritten.  The reference sets is_frame_in_use
     * (nv_dec.cpp:155) and never looks at it. */
    const int idx = ((const CUVIDPICPARAMS_HEAD *)pic)->CurrPicIdx;
    if (idx >= 0 && idx < MAX_DECODE_SURFACES && c->in_use[idx] > 0) {
        cuvid_drain(c);
        c->in_use[idx] = 0;
    }
    return c->nv.decode_picture(c->decoder, pic) == 0 ? 1 : 0;
}

int cuvid_on_display(void *user, CUVIDPARSERDISPINFO *d)                  /* nv_dec.cpp:44-52,151-161 */
{
    nvdec_b200 *c = (nvdec_b200 *)user;
    if (!d) return 1;                                                     /* newer parsers signal EOS with NULL */
    c->num_frames += 1;
    decoded_surface s;
    memset(&s, 0, sizeof(s));
    s.pool_slot = -2;
    s.disp = *d;
    s.width = c->disp_w; s.height = c->disp_h;                            /* ulTargetWidth/Height, :440-442 */
    if (d->picture_index >= 0 && d->picture_index < MAX_DECODE_SURFACES) c->in_use[d->picture_index]++;   /* :155 */
    c->pending.push_back(s);
    return 1;
}

int cuvid_open(nvdec_b200 *c, const char *extra, int len)                 /* nvdec_create_parser, nv_dec.cpp:278-366 */
{
    c->cuvid_codec = cuvid_codec_of(c->codec_type);
    if (!cuvid_load(c)) return -4;
    CUVIDPARSERPARAMS pp;
    memset(&pp, 0, sizeof(pp));
    memset(&c->parse_ext, 0, sizeof(c->parse_ext));
    pp.CodecType = c->cuvid_codec;
    pp.pExtVideoInfo = &c->parse_ext;
    c->parse_ext.format.chroma_format = CUVID_CHROMA_420;                 /* :335-336 */
    c->parse_ext.format.progressive_sequence = 1;
    if (extra && len > 0) {                                               /* :339-342 */
        const int n = len < (int)sizeof(c->parse_ext.raw_seqhdr_data) ? len : (int)sizeof(c->parse_ext.raw_seqhdr_data);
        c->parse_ext.format.seqhdr_data_length = (unsigned)n;
        memcpy(c->parse_ext.raw_seqhdr_data, extra, (size_t)n);
    }
    pp.ulMaxNumDecodeSurfaces = NVDEC_MAX_FRAMES;                         /* :345 */
    pp.ulMaxDisplayDelay = (unsigned)env_int("JMC_NVDEC_PARSER_DELAY", 2, 0, 8);   /* :346 */
    pp.pUserData = c;
    pp.pfnSequenceCallback = cuvid_on_sequence;
    pp.pfnDecodePicture = cuvid_on_decode;
    pp.pfnDisplayPicture = cuvid_on_display;
    int r = c->nv.create_parser(&c->parser, &pp);
    if (r != 0 || !c->parser) { jmc_set_error("cuvidCreateVideoParser failed (%d)", r); return -4; }
    if (c->parse_ext.format.seqhdr_data_length > 0) {                     /* :357-362: prime the parser with the sequence header */
        CUVIDSOURCEDATAPACKET pkt;
        memset(&pkt, 0, sizeof(pkt));
        pkt.payload = c->parse_ext.raw_seqhdr_data;
        pkt.payload_size = c->parse_ext.format.seqhdr_data_length;
        c->nv.parse(c->parser, &pkt);
    }
    return 0;
}

void cuvid_packet(nvdec_b200 *c, const unsigned char *buf, int len)       /* nvdec_decode_packet, nv_dec.cpp:368-403 */
{
    CUVIDSOURCEDATAPACKET pkt;
    memset(&pkt, 0, sizeof(pkt));
    if (buf && len > 0) {
        pkt.payload = buf;
        pkt.payload_size = (unsigned long)len;
        pkt.flags = CUVID_PKT_TIMESTAMP;                                  /* :386-387 */
    } else {
        pkt.flags = CUVID_PKT_ENDOFSTREAM;                                /* :390 */
    }
    c->nv.parse(c->parser, &pkt);
}

/* Surfaces go back to the decoder once the launch that read them has finished (nv_dec.cpp:469) -- checked
 * without blocking on every call; waited for only when the map limit is reached or the decoder goes away. */
void reap_unmaps(nvdec_b200 *c, bool wait_all)
{
    while (!c->unmaps.empty()) {
        unmap_batch &b = c->unmaps.front();
        if (wait_all) cudaEventSynchronize(b.done);
        else if (cudaEventQuery(b.done) != cudaSuccess) { cudaGetLastError(); break; }
        for (int i = 0; i < b.n; i++) {
            if (b.decoder == c->decoder && c->decoder) c->nv.unmap_frame(c->decoder, b.ptr[i]);
            if (b.pic[i] >= 0 && b.pic[i] < MAX_DECODE_SURFACES && c->in_use[b.pic[i]] > 0) c->in_use[b.pic[i]]--;   /* nvdec_frame_item_release, :458 */
            c->n_mapped--;
        }
        c->free_events.push_back(b.done);
        c->unmaps.pop_front();
    }
}

void cuvid_close(nvdec_b200 *c)
{
    if (c->decoder) reap_unmaps(c, true);
    if (c->parser) { c->nv.destroy_parser(c->parser); c->parser = nullptr; }     /* nv_dec.cpp:95-101 */
    if (c->decoder) { c->nv.destroy_decoder(c->decoder); c->decoder = nullptr; }
    if (c->nv.lib) { dlclose(c->nv.lib); c->nv.lib = nullptr; }
}

/* ---- RAW front-end: one packet = one decoded surface -> pending ------------------------------------ */
int raw_packet(nvdec_b200 *c, const unsigned char *buf, int len)
{
    if (len < (int)sizeof(jm_nvdec_raw_packet)) return -1;
    jm_nvdec_raw_packet h;
    memcpy(&h, buf, sizeof(h));
    if (h.magic != JM_NVDEC_RAW_MAGIC || h.width < 0 || h.height < 0 || h.pitch < h.width) return -1;
    if ((int64_t)h.pitch * h.height * 3 / 2 > 0x7fffffffll) return -1;     /* the API counts frame bytes in int (nv_dec.cpp:773) */
    decoded_surface s;
    memset(&s, 0, sizeof(s));
    s.width = h.width; s.height = h.height; s.pitch = h.pitch;
    s.sync_consume = (h.flags & JM_NVDEC_RAW_SYNC) != 0;
    cudaStream_t st = convert_stream(c);
    if (h.flags & JM_NVDEC_RAW_DEVICE_PTR) {
        s.dptr = (uint8_t *)(uintptr_t)h.device_ptr;
        s.pool_slot = -1;
        if (h.flags & JM_NVDEC_RAW_WAIT_EVENT) {
            /* the surface is being produced on another stream: order the conversion after the caller's event */
            if (len < (int)sizeof(jm_nvdec_raw_packet_ex)) return -1;
            jm_nvdec_raw_packet_ex x;
            memcpy(&x, buf, sizeof(x));
            if (x.ready_event && cudaStreamWaitEvent(st, (cudaEvent_t)(uintptr_t)x.ready_event, 0) != cudaSuccess) { cudaGetLastError(); return -1; }
        }
    } else {
        const size_t bytes = (size_t)h.pitch * h.height * 3 / 2;         /* nv_dec.cpp:453 */
        if ((size_t)len < sizeof(h) + bytes) return -1;
        if (bytes > c->pool_bytes) {                                     /* geometry grew: new surfaces */
            jmc_ctx_sync(c->ctx);
            for (int i = 0; i < NVDEC_MAX_FRAMES; i++) if (c->pool[i]) { cudaFree(c->pool[i]); c->pool[i] = nullptr; }
            c->pool_bytes = bytes;
        }
        const int slot = c->pool_next;
        c->pool_next = (c->pool_next + 1) % NVDEC_MAX_FRAMES;
        /* a surface that could not be converted for ten packets (no ring slot to be had: out of memory) is given up before
         * its upload surface is written again -- never converted from somebody else's bytes */
        for (auto it = c->pending.begin(); it != c->pending.end();) {
            if (it->pool_slot == slot) { it = c->pending.erase(it); c->dropped++; c->drop_flag = true; }
            else ++it;
        }
        /* uploads and kernels share the convert stream: the next upload into this surface is ordered behind the kernel that read it */
        if (!c->pool[slot]) {
            if (cudaMalloc((void **)&c->pool[slot], c->pool_bytes ? c->pool_bytes : 1) != cudaSuccess) { cudaGetLastError(); return -1; }
        }
        /* the "decode": the surface lands in HBM; only the active `width` bytes of every row are moved.
         * in_buf is consumed before we return. */
        const uint8_t *src = buf + sizeof(h);
        const size_t rows = (size_t)h.height * 3 / 2, wbytes = (size_t)h.width;
        if (bytes > 0 && wbytes > 0 && rows > 0) {
            const bool pinned = is_device_accessible_host(src, bytes);
            uint8_t *blo = nullptr, *bhi = nullptr;
            bool split = false;
            if (!pinned) {
                /* only buffers the caller registered explicitly (jm_nvdec_memory_register_host): packet buffers are
                 * typically transient, and a registration that outlives its buffer makes the DMA read stale pages */
                split = registered_interior(c, src, bytes, &blo, &bhi);
                if (split && !ensure_edges(c)) split = false;
            }
            if (split) {
                /* registered interior of a malloc'ed packet: the whole pitched payload moves by DMA out of the caller's
                 * buffer (one linear copy, like the reference's own cuMemcpyDtoH of pitch*h*3/2, nv_dec.cpp:452-453, in the
                 * other direction); its two < 4 KB edges take the pinned bounce buffer */
                const size_t head = (size_t)(blo - src), body = (size_t)(bhi - blo), tail = bytes - head - body;
                uint8_t *e = c->h_edge + 2 * 4096;
                cudaError_t er = cudaSuccess;
                if (head) { memcpy(e, src, head); er = cudaMemcpyAsync(c->pool[slot], e, head, cudaMemcpyHostToDevice, st); }
                if (er == cudaSuccess) er = cudaMemcpyAsync(c->pool[slot] + head, blo, body, cudaMemcpyHostToDevice, st);
                if (er == cudaSuccess && tail) { memcpy(e + 4096, bhi, tail); er = cudaMemcpyAsync(c->pool[slot] + head + body, e + 4096, tail, cudaMemcpyHostToDevice, st); }
                if (er == cudaSuccess) er = cudaStreamSynchronize(st);
                if (er != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(st); return -1; }    /* nothing of ours may still be reading in_buf */
            } else if (pinned) {
                /* pinned / registered payload: DMA straight out of the caller's buffer, complete before returning */
                if (cudaMemcpy2DAsync(c->pool[slot], (size_t)h.pitch, src, (size_t)h.pitch, wbytes, rows, cudaMemcpyHostToDevice, st) != cudaSuccess) { cudaGetLastError(); return -1; }
                if (cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); return -1; }
            } else {
                /* pageable payload: compact the rows into this surface's pinned staging buffer (the call returns when
                 * in_buf has been read) and let the H2D run behind us */
                /* default: compact rows in the staging buffer + a 2-D H2D that adds the pitch (6 % fewer bytes on the link).
                 * JMC_NVDEC_STAGE_LINEAR=1: staging in the surface's pitch + one linear H2D -- measured the same within the
                 * run-to-run noise (profiles/r2_stage_linear_ab.txt): enqueueing the 2-D copy is not what this path waits for */
                const bool linear = c->stage_linear;
                const size_t spitch = linear ? (size_t)h.pitch : wbytes;
                const size_t need = spitch * rows;
                if (need > c->stage_bytes) {
                    jmc_ctx_sync(c->ctx);
                    for (int i = 0; i < NVDEC_MAX_FRAMES; i++) if (c->h_stage[i]) { cudaFreeHost(c->h_stage[i]); c->h_stage[i] = nullptr; }
                    c->stage_bytes = need;
                }
                if (!c->h_stage[slot] && cudaHostAlloc((void **)&c->h_stage[slot], c->stage_bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return -1; }
                if (!c->stage_done[slot] && cudaEventCreateWithFlags(&c->stage_done[slot], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return -1; }
                if (c->stage_used[slot]) cudaEventSynchronize(c->stage_done[slot]);        /* ten uploads ago: long done */
                /* rows are copied into the staging buffer by the copy threads in a few pieces; each piece's H2D starts while the
                 * next is being copied */
                const size_t pieces = (need >= (2u << 20) && c->delay == 0) ? 4 : 1, prow = (rows + pieces - 1) / pieces;
                for (size_t r0 = 0; r0 < rows; r0 += prow) {
                    const size_t nr = rows - r0 < prow ? rows - r0 : prow;
                    copy_job cj = { c->h_stage[slot] + r0 * spitch, spitch, src + r0 * (size_t)h.pitch, (size_t)h.pitch, wbytes, nr, nr, 1 };
                    c->copier.copy(cj);
                    cudaError_t e2 = linear
                        ? cudaMemcpyAsync(c->pool[slot] + r0 * (size_t)h.pitch, c->h_stage[slot] + r0 * spitch, (nr - 1) * spitch + wbytes, cudaMemcpyHostToDevice, st)
                        : cudaMemcpy2DAsync(c->pool[slot] + r0 * (size_t)h.pitch, (size_t)h.pitch, c->h_stage[slot] + r0 * wbytes, wbytes, wbytes, nr, cudaMemcpyHostToDevice, st);
                    if (e2 != cudaSuccess) { cudaGetLastError(); return -1; }
                }
                cudaEventRecord(c->stage_done[slot], st);
                c->stage_used[slot] = true;
            }
        }
        s.dptr = c->pool[slot];
        s.pool_slot = slot;
    }
    mark_started(c);
    c->disp_w = h.width; c->disp_h = h.height;
    c->num_frames += 1;                                                   /* nv_dec.cpp:48 */
    c->pending.push_back(s);
    return 0;
}

/* ---- conversion: nvdec_decode_output_frame (nv_dec.cpp:406-478) for a whole batch ------------------- */
/* Converts the longest run of pending surfaces that share one geometry (at most MAP_LIMIT_MAX, at most the free
 * ring slots, at most the free map slots) with ONE launch.  Returns the number of frames converted, 0 if
 * nothing could be done now (ring full), < 0 on error.  must_progress: wait for outstanding unmaps instead of
 * giving up when the map limit is reached. */
int convert_stage(nvdec_b200 *c, bool must_progress)
{
    if (c->pending.empty()) return 0;
    const bool cuvid = c->pending.front().pool_slot == -2;
    if (cuvid && !c->decoder) {                                           /* :414-417 */
        while (!c->pending.empty() && c->pending.front().pool_slot == -2) c->pending.pop_front();
        return -1;
    }
    int limit = MAP_LIMIT_MAX;
    if (cuvid) {
        if (c->n_mapped >= c->map_limit) reap_unmaps(c, false);
        if (c->n_mapped >= c->map_limit) {
            if (!must_progress && c->map_limit > 1) return 0;
            reap_unmaps(c, true);
        }
        limit = c->map_limit - c->n_mapped;
        if (limit > MAP_LIMIT_MAX) limit = MAP_LIMIT_MAX;
        if (limit < 1) return 0;
    }
    const decoded_surface first = c->pending.front();
    int slots[MAP_LIMIT_MAX];
    decoded_surface surf[MAP_LIMIT_MAX];
    unsigned long long mapped[MAP_LIMIT_MAX];
    int k = 0;
    int pitch = first.pitch;
    cudaStream_t st = convert_stream(c);
    while (k < limit && !c->pending.empty()) {
        decoded_surface s = c->pending.front();
        if ((s.pool_slot == -2) != cuvid || s.width != first.width || s.height != first.height) break;
        if (!cuvid && s.pitch != pitch) break;
        const int slot = acquire_slot(c, s.width, s.height);
        if (slot < 0) break;
        mapped[k] = 0;
        if (cuvid) {
            /* cuvidMapVideoFrame (nv_dec.cpp:427-442): post-processed NV12 surface, produced on OUR convert
             * stream so that the kernel below is ordered after it without a host sync */
            CUVIDPROCPARAMS pp;
            memset(&pp, 0, sizeof(pp));
            pp.progressive_frame = s.disp.progressive_frame;
            pp.top_field_first = s.disp.top_field_first;
            pp.unpaired_field = s.disp.repeat_first_field < 0;
            pp.output_stream = st;
            unsigned int mp = 0;
            if (c->nv.map_frame(c->decoder, s.disp.picture_index, &mapped[k], &mp, &pp) != 0 || !mapped[k]) {
                release_slot(c, slot);
                if (k == 0) {                                             /* cannot be mapped at all: skip it */
                    const int idx = s.disp.picture_index;
                    if (idx >= 0 && idx < MAX_DECODE_SURFACES && c->in_use[idx] > 0) c->in_use[idx]--;
                    c->pending.pop_front();
                    c->dropped++; c->drop_flag = true;
                    return -1;
                }
                break;
            }
            if (k > 0 && (int)mp != pitch) {                              /* never seen; a new launch takes it */
                c->nv.unmap_frame(c->decoder, mapped[k]);
                release_slot(c, slot);
                break;
            }
            c->n_mapped++;
            s.dptr = (uint8_t *)(uintptr_t)mapped[k];
            s.pitch = pitch = (int)mp;
        }
        slots[k] = slot;
        surf[k] = s;
        c->pending.pop_front();
        k++;
    }
    if (k == 0) return 0;

    jmc_job j;
    memset(&j, 0, sizeof(j));
    jmc_job_nvdec(&j, first.width, first.height, pitch, c->out_fmt);
    j.n_frames = k;
    void *slist[MAP_LIMIT_MAX], *tlist[MAP_LIMIT_MAX];
    if (k == 1) {
        j.surf.base = surf[0].dptr;
        j.tight.base = c->ring[slots[0]].d_tight;
    } else {
        for (int i = 0; i < k; i++) { slist[i] = surf[i].dptr; tlist[i] = c->ring[slots[i]].d_tight; }
        j.surf.list = slist;
        j.tight.list = tlist;
        j.flags = JMC_JOB_LIST_ON_HOST;                                   /* the pointers ride in the kernel arguments */
    }
    const int r = jmc_launch_job(c->ctx, &j, st);
    bool ok = r == JMC_OK;
    bool need_sync = false;
    for (int i = 0; i < k; i++) need_sync = need_sync || surf[i].sync_consume;
    if (ok) ok = cudaEventRecord(c->ring[slots[k - 1]].converted, st) == cudaSuccess;
    for (int i = 0; i < k && ok; i++) {
        c->ring[slots[i]].conv_slot = slots[k - 1];
        if (c->staged) c->to_prefetch.push_back(slots[i]);               /* delivery is enqueued by flush_prefetch, once the launch is done */
    }
    if (cuvid) {
        unmap_batch b;
        b.done = get_event(c);
        b.decoder = c->decoder;
        b.n = k;
        for (int i = 0; i < k; i++) { b.ptr[i] = mapped[i]; b.pic[i] = surf[i].disp.picture_index; }
        if (b.done && cudaEventRecord(b.done, st) == cudaSuccess) c->unmaps.push_back(b);
        else {                                                            /* no event: fall back to waiting */
            cudaStreamSynchronize(st);
            for (int i = 0; i < k; i++) {
                c->nv.unmap_frame(c->decoder, mapped[i]);
                if (b.pic[i] >= 0 && b.pic[i] < MAX_DECODE_SURFACES && c->in_use[b.pic[i]] > 0) c->in_use[b.pic[i]]--;
                c->n_mapped--;
            }
            if (b.done) c->free_events.push_back(b.done);
        }
    }
    if (!ok) {
        cudaGetLastError();
        for (int i = 0; i < k; i++) release_slot(c, slots[i]);
        c->dropped += (uint32_t)k; c->drop_flag = true;
        return -1;
    }
    if (need_sync) cudaEventSynchronize(c->ring[slots[k - 1]].converted);
    for (int i = 0; i < k; i++) c->ready.push_back(slots[i]);
    return k;
}

void free_everything(nvdec_b200 *c)
{
    for (auto &s : c->ring) {
        if (s.d_tight) cudaFree(s.d_tight);
        if (s.h_tight) cudaFreeHost(s.h_tight);
        if (s.converted) cudaEventDestroy(s.converted);
        if (s.direct) cudaEventDestroy(s.direct);
        for (int i = 0; i < MAX_CHUNKS; i++) if (s.delivered[i]) cudaEventDestroy(s.delivered[i]);
    }
    c->ring.clear();
    c->ready.clear();
    c->to_prefetch.clear();
    g_dev_inflight[c->device & 63].fetch_sub((int)c->inflight.size(), std::memory_order_relaxed);
    c->inflight.clear();
    c->cur = -1;
    for (int i = 0; i < NVDEC_MAX_FRAMES; i++) {
        if (c->pool[i]) { cudaFree(c->pool[i]); c->pool[i] = nullptr; }
        if (c->h_stage[i]) { cudaFreeHost(c->h_stage[i]); c->h_stage[i] = nullptr; }
        if (c->stage_done[i]) { cudaEventDestroy(c->stage_done[i]); c->stage_done[i] = nullptr; }
        c->stage_used[i] = false;
    }
    c->pool_bytes = c->stage_bytes = 0;
    for (cudaEvent_t e : c->free_events) cudaEventDestroy(e);
    c->free_events.clear();
    for (auto &r : c->regs) if (cudaHostUnregister(r.base) != cudaSuccess) cudaGetLastError();
    c->regs.clear();
    if (c->h_edge) { cudaFreeHost(c->h_edge); c->h_edge = nullptr; }
    c->pending.clear();
}

void teardown(nvdec_b200 *c)
{
    if (!c->ctx) return;
    {
        jmc_device_guard g(c->ctx);
        jmc_ctx_sync(c->ctx);
        cuvid_close(c);
        free_everything(c);
    }
    jmc_ctx_destroy(c->ctx);
    c->ctx = nullptr;
    c->inited = false;
}

} /* namespace */

extern "C" {

handle_nvdec jm_nvdec_create_handle(void)
{
    nvdec_b200 *c = nullptr;
    try { c = new nvdec_b200(); } catch (...) { return nullptr; }    /* new + memset, nv_dec.cpp:54-60; nothing may throw across the C ABI */
    c->device = env_int("JMC_DEVICE", 0, 0, 1023);
    c->delay = env_int("JMC_NVDEC_DISPLAY_DELAY", 0, 0, DELAY_MAX);
    /* helper threads for copies from / to PAGEABLE caller buffers (started only when such a copy happens):
     * a quarter of the host's cores, at most 4, unless JMC_NVDEC_COPY_THREADS / option "copy_threads" says otherwise */
    c->copy_threads = env_int("JMC_NVDEC_COPY_THREADS", -1, -1, COPY_THREADS_MAX);      /* -1: decided at jm_nvdec_init */
    c->lazy_pin = env_int("JMC_NVDEC_LAZY_PIN", 0, 0, 1);
    c->map_limit = env_int("JMC_NVDEC_MAP_LIMIT", MAP_LIMIT_MAX, 1, MAP_LIMIT_MAX);
    c->max_inflight = env_int("JMC_NVDEC_MAX_INFLIGHT", 2, 1, 16);
    c->stage_linear = env_int("JMC_NVDEC_STAGE_LINEAR", 0, 0, 1) != 0;
    if (g_dev_inflight_max.load() < 0) g_dev_inflight_max.store(env_int("JMC_NVDEC_DEVICE_INFLIGHT", 4, 1, 64));   /* racing handles store the same value */
    g_live_handles.fetch_add(1);
    return c;
}

int jm_nvdec_set_device(int device, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || c->inited) return -1;
    c->device = device;
    return 0;
}

int jm_nvdec_set_display_delay(int frames, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || frames < 0 || frames > DELAY_MAX) return -1;
    c->delay = frames;
    return 0;
}

int jm_nvdec_set_option(const char *name, int value, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || !name) return -1;
    if (!strcmp(name, "display_delay")) return jm_nvdec_set_display_delay(value, handle);
    if (!strcmp(name, "lazy_pin")) { c->lazy_pin = value != 0; return 0; }
    if (!strcmp(name, "copy_threads")) {
        if (value < 0 || value > COPY_THREADS_MAX) return -1;
        if (value < (int)c->copier.workers.size()) c->copier.shutdown();
        c->copy_threads = c->copier.want_threads = value;
        return 0;
    }
    if (!strcmp(name, "map_limit")) {                                    /* takes effect when the next decoder is created */
        if (value < 1 || value > MAP_LIMIT_MAX) return -1;
        c->map_limit = value;
        return 0;
    }
    return -1;
}

int jm_nvdec_init(int codec_type, int out_fmt, char *extra_data, int len, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c) return -1;
    if (c->ctx) teardown(c);                                          /* init on a live handle: start over, leak nothing */
    c->out_fmt = out_fmt;
    c->codec_type = codec_type;
    c->is_eof = c->is_exit = false;
    c->num_frames = c->dropped = 0;
    c->started = false;
    c->staged = true;
    int r = jmc_ctx_create(c->device, &c->ctx);
    if (r == JMC_ERR_NO_DEVICE) return jmc_device_count() <= 0 ? -2 : -3;   /* nvdec_cuda_init, nv_dec.cpp:219-231 */
    if (r) return -1;
    c->inited = true;
    /* helper threads for copies from / to PAGEABLE caller buffers, started on the first such copy: unless told
     * otherwise a quarter of the host's cores shared by the handles alive now, at most 4 per handle */
    if (c->copy_threads < 0) {
        const int hw = (int)std::thread::hardware_concurrency(), live = g_live_handles.load() > 0 ? g_live_handles.load() : 1;
        const int share = hw / (4 * live);
        c->copier.want_threads = share > 4 ? 4 : share;
    } else {
        c->copier.want_threads = c->copy_threads;
    }
    if (codec_type != JM_NVDEC_CODEC_RAW_NV12) {
        /* bitstream codecs: NVDEC parser + decoder (nvdec_create_parser, nv_dec.cpp:278-366) */
        jmc_device_guard g(c->ctx);
        int r2 = g.err ? -1 : cuvid_open(c, extra_data, len);
        if (r2) { cuvid_close(c); return r2; }
    }
    return 0;
}

int jm_nvdec_deinit(handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c) return -1;
    teardown(c);
    c->copier.shutdown();
    g_live_handles.fetch_sub(1);
    delete c;
    return 0;
}

static int decode_frame_impl(unsigned char *in_buf, int in_data_len, int *got_frame, handle_nvdec handle);
static int output_frame_impl(unsigned char *out_buf, int *out_len, handle_nvdec handle);

int jm_nvdec_decode_frame(unsigned char *in_buf, int in_data_len, int *got_frame, handle_nvdec handle)
{
    try { return decode_frame_impl(in_buf, in_data_len, got_frame, handle); }
    catch (...) { jmc_set_error("jm_nvdec_decode_frame: out of host memory"); if (got_frame) *got_frame = 0; return -1; }
}

int jm_nvdec_output_frame(unsigned char *out_buf, int *out_len, handle_nvdec handle)
{
    try { return output_frame_impl(out_buf, out_len, handle); }
    catch (...) { jmc_set_error("jm_nvdec_output_frame: out of host memory"); return -1; }
}

static int decode_frame_impl(unsigned char *in_buf, int in_data_len, int *got_frame, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (got_frame) *got_frame = 0;
    if (!c || !got_frame) return 0;
    if (!c->inited || !c->ctx) return 0;                       /* decoder never created: the reference swallows -1 (nv_dec.cpp:414-417,491-493) */
    jmc_device_guard guard(c->ctx);                            /* the reference pushes / pops its context around every call (nv_dec.cpp:378,398,423,471) */
    if (guard.err) return 0;
    const bool cuvid = c->codec_type != JM_NVDEC_CODEC_RAW_NV12;
    c->drop_flag = false;
    if (cuvid && !c->unmaps.empty()) reap_unmaps(c, false);
    flush_prefetch(c, -1);                                     /* deliveries of the frames converted by the previous call */
    if (!c->is_eof) {                                          /* nv_dec.cpp:486-488 */
        if (in_buf && in_data_len > 0) {
            if (!cuvid) raw_packet(c, in_buf, in_data_len);    /* malformed packet: consumed, no frame (errors swallowed like :491-493) */
            else if (c->parser) cuvid_packet(c, in_buf, in_data_len);      /* no parser: init failed, packet dropped */
        } else {
            if (cuvid && c->parser) cuvid_packet(c, nullptr, 0);           /* flush: the parser hands out its delayed pictures */
            c->is_eof = true;                                  /* CUVID_PKT_ENDOFSTREAM, nv_dec.cpp:389-392 */
        }
    }
    /* nvdec_decode_output_frame (nv_dec.cpp:406-478), for everything that is pending */
    while (!c->pending.empty()) {
        if (convert_stage(c, false) <= 0) break;
    }
    /* nothing to announce while pictures wait for a map slot (every mapped surface still belongs to a launch in
     * flight): wait for those launches rather than hand the caller an empty call it can only repeat */
    if (c->ready.empty() && !c->pending.empty()) convert_stage(c, true);
    flush_prefetch(c, -1);                                     /* whatever has finished converting meanwhile (uploads take longer than kernels) */
    /* announce at most one frame per call (:455); frames beyond the display delay wait in `ready` */
    if (!c->ready.empty() && ((int)c->ready.size() > c->delay || c->is_eof)) {
        release_slot(c, c->cur);                               /* single current frame: fetch it before the next one is announced (nv_dec.h:119-123) */
        c->cur = c->ready.front();
        c->ready.pop_front();
        *got_frame = 1;
    } else if (c->ready.empty() && c->pending.empty() && c->is_eof) {
        c->is_exit = true;                                     /* :460-466 */
        show_info(c);
    }
    if (c->drop_flag) {
        jmc_set_error("jm_nvdec_decode_frame: %u decoded frame(s) dropped so far (the handle holds at most %d converted frames; fetch them with jm_nvdec_output_frame)",
                      c->dropped, RING_MAX);
        return -1;                                             /* deviation: the reference cannot lose frames silently here, its queue is unbounded */
    }
    return 0;
}

static int output_frame_impl(unsigned char *out_buf, int *out_len, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || c->cur < 0) return -1;                           /* nv_dec.cpp:757-758 */
    if (!out_buf || !out_len) return -1;                       /* :768-771 */
    jmc_device_guard guard(c->ctx);
    if (guard.err) return -1;
    ring_slot &s = c->ring[c->cur];
    const int need = s.w * s.h * 3 / 2;
    if (*out_len < need) return -2;                            /* :773-774 */
    *out_len = 0;                                              /* :776 */
    if (s.total > 0) {
        const int kind = host_kind_of(out_buf, s.total);
        uint8_t *blo = nullptr, *bhi = nullptr;
        bool split = false;
        if (kind == MEM_PAGEABLE) {
            split = registered_interior(c, out_buf, s.total, &blo, &bhi);
            if (!split && maybe_lazy_pin(c, out_buf, (size_t)need)) split = registered_interior(c, out_buf, s.total, &blo, &bhi);
            if (split && !ensure_edges(c)) split = false;
        }
        if (split) {
            /* malloc'ed out_buf whose interior pages are registered: the body arrives by direct DMA, the two < 4 KB
             * edges through the pinned bounce buffer */
            c->staged = false;
            c->to_prefetch.clear();                                       /* this caller takes its frames by direct DMA */
            const size_t head = (size_t)(blo - out_buf), body = (size_t)(bhi - blo), tail = s.total - head - body;
            cudaStream_t ds = delivery_stream(c);
            cudaError_t er = cudaEventSynchronize(c->ring[s.conv_slot >= 0 ? s.conv_slot : c->cur].converted);   /* launch done: the copies never wait on the device */
            if (er == cudaSuccess && head) er = cudaMemcpyAsync(c->h_edge, s.d_tight, head, cudaMemcpyDeviceToHost, ds);
            if (er == cudaSuccess) er = cudaMemcpyAsync(blo, s.d_tight + head, body, cudaMemcpyDeviceToHost, ds);
            if (er == cudaSuccess && tail) er = cudaMemcpyAsync(c->h_edge + 4096, s.d_tight + head + body, tail, cudaMemcpyDeviceToHost, ds);
            if (er == cudaSuccess) er = cudaEventRecord(s.direct, ds);
            if (er == cudaSuccess) er = cudaEventSynchronize(s.direct);
            if (er != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(ds); return -1; }       /* no copy into out_buf outlives the call */
            if (head) memcpy(out_buf, c->h_edge, head);
            if (tail) memcpy(bhi, c->h_edge + 4096, tail);
        } else if (kind != MEM_PAGEABLE) {
            /* pinned / registered out_buf: only the tight frame crosses PCIe, by DMA straight into the caller's
             * buffer (a device out_buf gets a device-to-device copy: the frame never leaves HBM) */
            c->staged = false;
            c->to_prefetch.clear();                                       /* this caller takes its frames by direct DMA */
            cudaStream_t ds = delivery_stream(c);
            if (cudaEventSynchronize(c->ring[s.conv_slot >= 0 ? s.conv_slot : c->cur].converted) != cudaSuccess) { cudaGetLastError(); return -1; }   /* launch done: the copy never waits on the device */
            if (cudaMemcpyAsync(out_buf, s.d_tight, s.total, kind == MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ds) != cudaSuccess) { cudaGetLastError(); return -1; }
            if (cudaEventRecord(s.direct, ds) != cudaSuccess || cudaEventSynchronize(s.direct) != cudaSuccess) {
                cudaGetLastError();
                cudaStreamSynchronize(ds);                                    /* no copy into out_buf outlives the call */
                return -1;
            }
        } else {
            /* pageable out_buf (what the reference's callers pass): the frame is (being) prefetched into the pinned
             * ring; copy chunk i out while chunk i+1 is still in flight */
            c->staged = true;
            if (!s.prefetched) {
                bool queued = false;
                for (int q : c->to_prefetch) queued = queued || q == c->cur;
                if (!queued) c->to_prefetch.push_back(c->cur);
                if (!flush_prefetch(c, c->cur)) return -1;
            }
            const size_t unit = 4096, rows = s.total / unit;
            bool failed = false;
            auto wait_chunk = [&](int i) { if (cudaEventSynchronize(s.delivered[i]) != cudaSuccess) { cudaGetLastError(); failed = true; } };
            if (rows) {
                copy_job cj = { out_buf, unit, s.h_tight, unit, unit, rows, s.chunk_bytes / unit, s.n_chunks };
                c->copier.run(cj, wait_chunk);
            } else {
                wait_chunk(0);
            }
            if (s.total % unit) {
                wait_chunk(s.n_chunks - 1);
                memcpy(out_buf + rows * unit, s.h_tight + rows * unit, s.total % unit);
            }
            if (failed) return -1;
            flush_prefetch(c, -1);                             /* this frame's delivery is done: let the next one start */
        }
    }
    *out_len = need;                                           /* :824 */
    return need;                                               /* :827 */
}

int jm_nvdec_output_frame_ref(const unsigned char **frame, int *frame_len, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || c->cur < 0 || !frame) return -1;
    jmc_device_guard guard(c->ctx);
    if (guard.err) return -1;
    ring_slot &s = c->ring[c->cur];
    c->staged = true;
    if (!s.prefetched) {
        bool queued = false;
        for (int q : c->to_prefetch) queued = queued || q == c->cur;
        if (!queued) c->to_prefetch.push_back(c->cur);
        if (!flush_prefetch(c, c->cur)) return -1;
    }
    for (int i = 0; i < s.n_chunks; i++)
        if (cudaEventSynchronize(s.delivered[i]) != cudaSuccess) { cudaGetLastError(); return -1; }
    *frame = s.h_tight;
    const int need = s.w * s.h * 3 / 2;
    if (frame_len) *frame_len = need;
    return need;
}

int jm_nvdec_stream_info(int *disp_width, int *disp_height, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || !disp_width || !disp_height) return -1;
    *disp_width = c->disp_w;
    *disp_height = c->disp_h;
    return 0;
}

void jm_nvdec_set_eof(bool is_eof, handle_nvdec handle)
{
    if (handle) ((nvdec_b200 *)handle)->is_eof = is_eof;       /* nv_dec.cpp:619-622 */
}

bool jm_nvdec_is_exit(handle_nvdec handle)
{
    return handle ? ((nvdec_b200 *)handle)->is_exit : true;
}

char *jm_nvdec_show_dec_info(handle_nvdec handle)
{
    static char empty[1] = "";
    return handle ? ((nvdec_b200 *)handle)->dec_info : empty;
}

bool jm_nvdec_is_hw_support(void)
{
    return jmc_device_count() > 0;                             /* nv_dec.cpp:188-200 */
}

int jm_nvdec_memory_alloc_host(void **buf, int buf_len, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || !c->ctx || !buf || buf_len < 0) return -1;
    return jmc_alloc_host(c->ctx, (size_t)buf_len, 0, buf) == JMC_OK ? 0 : -1;
}

int jm_nvdec_memory_release_host(void *buf, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || !c->ctx) return -1;
    return jmc_free_host(c->ctx, buf) == JMC_OK ? 0 : -1;
}

int jm_nvdec_memory_register_host(void *buf, int buf_len, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || !c->ctx || !buf || buf_len <= 0) return -1;
    jmc_device_guard guard(c->ctx);
    if (guard.err) return -1;
    if (is_device_accessible_host(buf, (size_t)buf_len)) return 0;
    uint8_t *lo, *hi;
    if (registered_interior(c, buf, (size_t)buf_len, &lo, &hi)) return 0;
    return register_range(c, buf, (size_t)buf_len) ? 0 : -1;
}

int jm_nvdec_memory_unregister_host(void *buf, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || !c->ctx || !buf) return -1;
    jmc_device_guard guard(c->ctx);
    if (guard.err) return -1;
    for (size_t i = 0; i < c->regs.size(); i++) {
        const uintptr_t lo = (uintptr_t)c->regs[i].base, hi = lo + c->regs[i].len;
        if ((uintptr_t)buf + 4096 > lo && (uintptr_t)buf < hi) {          /* buf starts at most one page before its registered interior */
            jmc_ctx_sync(c->ctx);
            if (cudaHostUnregister(c->regs[i].base) != cudaSuccess) cudaGetLastError();
            c->regs.erase(c->regs.begin() + (long)i);
            return 0;
        }
    }
    return -1;
}

int jm_nvdec_dropped_frames(handle_nvdec handle)
{
    return handle ? (int)((nvdec_b200 *)handle)->dropped : -1;
}

long long jm_nvdec_launch_count(handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    return c && c->ctx ? (long long)jmc_ctx_launch_count(c->ctx) : 0;
}

int jm_nvdec_deliveries_in_flight(int device)
{
    return device >= 0 && device < 64 ? g_dev_inflight[device].load(std::memory_order_relaxed) : -1;
}

} /* extern "C" */
