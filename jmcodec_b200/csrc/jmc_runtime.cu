/*
 * jmc_runtime.cu -- context, memory, geometry fillers and the host-delivery pipeline of the
 * C-ABI declared in include/jmc_cuda.h.  Host-side C++ over the CUDA runtime; no torch types.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <mutex>
#include <new>

#include "jmc_internal.h"

/* ---- errors -------------------------------------------------------------------------------- */
static thread_local char g_err[512] = "";

void jmc_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int jmc_cuda_fail(cudaError_t e, const char *what)
{
    jmc_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();             /* reported: do not let it surface again at the next launch's error check */
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) ? JMC_ERR_NO_DEVICE
         : (e == cudaErrorMemoryAllocation) ? JMC_ERR_NOMEM : JMC_ERR_CUDA;
}

extern "C" {

const char *jmc_last_error(void) { return g_err; }
const char *jmc_version(void) { return "jmcodec_b200 0.1 (sm_100a)"; }

/* ---- device + context --------------------------------------------------------------------- */
int jmc_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { jmc_cuda_fail(e, "cudaGetDeviceCount"); cudaGetLastError(); return 0; }
    return n;
}

int jmc_current_device(void)
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}

int jmc_set_current_device(int device)
{
    JMC_CUDA(cudaSetDevice(device));
    return JMC_OK;
}

int jmc_ctx_create(int device, jmc_ctx **out)
{
    if (!out) { jmc_set_error("jmc_ctx_create: NULL out"); return JMC_ERR_INVALID; }
    *out = nullptr;
    int n = jmc_device_count();
    if (n <= 0) { if (!g_err[0]) jmc_set_error("no CUDA device"); return JMC_ERR_NO_DEVICE; }       /* nv_dec.cpp:219-222 */
    if (device < 0 || device >= n) { jmc_set_error("invalid device id %d (have %d)", device, n); return JMC_ERR_NO_DEVICE; }   /* :227-231 */
    jmc_device_guard guard_(device);
    if (guard_.err) return guard_.err;
    jmc_ctx *c = new (std::nothrow) jmc_ctx();
    if (!c) return JMC_ERR_NOMEM;
    c->device = device;
    c->launches = 0;
    for (int i = 0; i < 3; i++) c->stream[i] = nullptr;
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) { delete c; return jmc_cuda_fail(e, "cudaDeviceGetAttribute"); }
    c->sm_count = sms;
    for (int i = 0; i < 3; i++) {
        e = cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            for (int k = 0; k < i; k++) cudaStreamDestroy(c->stream[k]);       /* do not leak the streams already created */
            delete c;
            return jmc_cuda_fail(e, "cudaStreamCreateWithFlags");
        }
    }
    *out = c;
    return JMC_OK;
}

} /* extern "C" */

/* ---- device guard + environment switches ------------------------------------------------------ */
void jmc_device_guard::enter(int device)
{
    prev = -1; err = 0; switched = false;
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (prev == device) return;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { err = jmc_cuda_fail(e, "cudaSetDevice"); return; }
    switched = prev >= 0;
}
jmc_device_guard::jmc_device_guard(int device) { enter(device); }
jmc_device_guard::jmc_device_guard(const jmc_ctx *c)
{
    if (!c) { prev = -1; switched = false; jmc_set_error("NULL jmc_ctx"); err = JMC_ERR_INVALID; return; }
    enter(c->device);
}
jmc_device_guard::~jmc_device_guard()
{
    if (switched) cudaSetDevice(prev);                  /* hand the caller's device back (nv_dec.cpp:398,471 cuCtxPopCurrent) */
}

static jmc_env_flags g_env;
static std::atomic<bool> g_env_loaded{false};
static std::mutex g_env_mutex;
static bool env_on(const char *name) { const char *e = getenv(name); return e && atoi(e) != 0; }
static int env_tri(const char *name) { const char *e = getenv(name); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }
static void env_load_locked()
{
    jmc_env_flags f;
    f.no_bulk = env_on("JMC_NO_BULK");
    f.no_rows = env_on("JMC_NO_ROWS");
    f.rows_single = env_on("JMC_ROWS_SINGLE");
    f.rows_always = env_on("JMC_ROWS_ALWAYS");
    f.rgb_bulk_always = env_on("JMC_RGB_BULK_ALWAYS");
    f.pipeline_h2d_2d = env_tri("JMC_PIPELINE_H2D_2D") != 0;
    f.rgb_flat = env_tri("JMC_RGB_FLAT");
    f.rgb2_flat = env_tri("JMC_RGB2_FLAT");
    f.pad_zero = env_on("JMC_PAD_ZERO");
    f.rgb_bulk_pairs = getenv("JMC_RGB_BULK_PAIRS") ? atoi(getenv("JMC_RGB_BULK_PAIRS")) : 0;
    f.brows_rows = getenv("JMC_BROWS_ROWS") ? atoi(getenv("JMC_BROWS_ROWS")) : 0;
    g_env = f;
    g_env_loaded.store(true, std::memory_order_release);
}
const jmc_env_flags &jmc_env()
{
    if (!g_env_loaded.load(std::memory_order_acquire)) {       /* first use, possibly from several threads at once */
        std::lock_guard<std::mutex> lk(g_env_mutex);
        if (!g_env_loaded.load(std::memory_order_relaxed)) env_load_locked();
    }
    return g_env;
}

extern "C" {

void jmc_reload_env(void) { std::lock_guard<std::mutex> lk(g_env_mutex); env_load_locked(); }   /* tests; not while launches run on other threads */

int jmc_ctx_destroy(jmc_ctx *c)
{
    JMC_BIND(c);
    for (int i = 0; i < 3; i++) { cudaStreamSynchronize(c->stream[i]); cudaStreamDestroy(c->stream[i]); }
    delete c;
    return JMC_OK;
}

int jmc_ctx_device(const jmc_ctx *c) { return c ? c->device : JMC_ERR_INVALID; }
int jmc_ctx_sm_count(const jmc_ctx *c) { return c ? c->sm_count : JMC_ERR_INVALID; }
void *jmc_ctx_stream(jmc_ctx *c, int which) { return (c && which >= 0 && which < 3) ? (void *)c->stream[which] : nullptr; }
uint64_t jmc_ctx_launch_count(const jmc_ctx *c) { return c ? c->launches : 0; }

int jmc_ctx_sync(jmc_ctx *c)
{
    JMC_BIND(c);
    for (int i = 0; i < 3; i++) JMC_CUDA(cudaStreamSynchronize(c->stream[i]));
    return JMC_OK;
}

/* ---- memory ---------------------------------------------------------------------------------- */
int jmc_alloc_device(jmc_ctx *c, size_t bytes, void **dptr)
{
    JMC_BIND(c);
    if (!dptr) return JMC_ERR_INVALID;
    JMC_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return JMC_OK;
}

int jmc_alloc_pitched(jmc_ctx *c, size_t width_bytes, size_t rows, void **dptr, size_t *pitch)
{
    JMC_BIND(c);
    if (!dptr || !pitch) return JMC_ERR_INVALID;
    JMC_CUDA(cudaMallocPitch(dptr, pitch, width_bytes ? width_bytes : 1, rows ? rows : 1));
    return JMC_OK;
}

int jmc_free_device(jmc_ctx *c, void *dptr)
{
    JMC_BIND(c);
    JMC_CUDA(cudaFree(dptr));
    return JMC_OK;
}

int jmc_alloc_host(jmc_ctx *c, size_t bytes, int write_combined, void **hptr)
{
    JMC_BIND(c);
    if (!hptr) return JMC_ERR_INVALID;
    JMC_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    return JMC_OK;
}

int jmc_free_host(jmc_ctx *c, void *hptr)
{
    JMC_BIND(c);
    JMC_CUDA(cudaFreeHost(hptr));
    return JMC_OK;
}

int jmc_memcpy_h2d(jmc_ctx *c, void *dptr, const void *hptr, size_t bytes)
{
    JMC_BIND(c);
    JMC_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, c->stream[0]));
    JMC_CUDA(cudaStreamSynchronize(c->stream[0]));
    return JMC_OK;
}

int jmc_memcpy_d2h(jmc_ctx *c, void *hptr, const void *dptr, size_t bytes)
{
    JMC_BIND(c);
    JMC_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream[0]));
    JMC_CUDA(cudaStreamSynchronize(c->stream[0]));
    return JMC_OK;
}

int jmc_memset_device(jmc_ctx *c, void *dptr, int byte, size_t bytes)
{
    JMC_BIND(c);
    JMC_CUDA(cudaMemsetAsync(dptr, byte, bytes, c->stream[0]));
    JMC_CUDA(cudaStreamSynchronize(c->stream[0]));
    return JMC_OK;
}

/* ---- geometry fillers: the reference's own offset arithmetic ------------------------------- */
int64_t jmc_tight_bytes(int width, int height) { return (int64_t)width * height * 3 / 2; }

int jmc_job_nvdec(jmc_job *j, int width, int height, int pitch, int out_fmt)
{
    if (!j || width < 0 || height < 0 || pitch < width) { jmc_set_error("jmc_job_nvdec: bad geometry"); return JMC_ERR_INVALID; }
    j->op = out_fmt == 0 ? JMC_OP_NV12_TO_NV12 : JMC_OP_NV12_TO_I420;      /* nv_dec.cpp:782,798 */
    j->width = width; j->height = height; j->pitch = pitch;
    j->surf_y_off = 0;
    j->surf_uv_off = (int64_t)pitch * height;                              /* :765 */
    j->tight_u_off = (int64_t)width * height;                              /* :779 xy_offset */
    j->tight_v_off = j->tight_u_off + (int64_t)(width >> 1) * (height >> 1);   /* :810 uv_offset = w2*h2 */
    return JMC_OK;
}

int jmc_job_inteldec(jmc_job *j, int pitch, int surf_rows, int crop_x, int crop_y, int crop_w, int crop_h, int out_fmt)
{
    if (!j || crop_w < 0 || crop_h < 0 || crop_x < 0 || crop_y < 0 || pitch < crop_x + crop_w || surf_rows < crop_y + crop_h) {
        jmc_set_error("jmc_job_inteldec: bad geometry");
        return JMC_ERR_INVALID;
    }
    j->op = out_fmt == 0 ? JMC_OP_NV12_TO_NV12 : JMC_OP_NV12_TO_I420;      /* intel_dec.cpp:289,301 */
    j->width = crop_w; j->height = crop_h; j->pitch = pitch;
    j->surf_y_off = (int64_t)crop_y * pitch + crop_x;                      /* :285 */
    /* UV plane follows the allocated luma rows; origin (crop_y/2) rows and crop_x/2 BYTES in (:292-293,303-304) */
    j->surf_uv_off = (int64_t)pitch * surf_rows + (int64_t)(crop_y / 2) * pitch + (crop_x / 2);
    const int y_len = crop_w * crop_h, uv_len = y_len / 2;                 /* :261-262 */
    j->tight_u_off = y_len;                                                /* :306 */
    j->tight_v_off = (int64_t)y_len + uv_len / 2;                          /* :307 */
    return JMC_OK;
}

int jmc_job_intelenc(jmc_job *j, int pitch, int surf_rows, int crop_x, int crop_y, int crop_w, int crop_h, int is_i420)
{
    if (!j || crop_w < 0 || crop_h < 0 || crop_x < 0 || crop_y < 0 || pitch < crop_x + crop_w || surf_rows < crop_y + crop_h ||
        crop_w > 65535 || crop_h > 65535 || pitch > 65535) {               /* the reference computes in uint16_t (:265) */
        jmc_set_error("jmc_job_intelenc: bad geometry");
        return JMC_ERR_INVALID;
    }
    j->op = is_i420 ? JMC_OP_I420_TO_SURF : JMC_OP_NV12_TO_SURF;
    j->width = crop_w; j->height = crop_h; j->pitch = pitch;
    j->surf_y_off = (int64_t)crop_y * pitch + crop_x;                      /* intel_enc.cpp:292 */
    j->surf_uv_off = (int64_t)pitch * surf_rows + (int64_t)(crop_y / 2) * pitch + (crop_x / 2);   /* :298-299,371 */
    j->tight_u_off = (int64_t)crop_w * crop_h;                             /* :368-369 */
    j->tight_v_off = j->tight_u_off + (int64_t)(crop_w / 2) * (crop_h / 2);    /* :370 */
    return JMC_OK;
}

int jmc_job_nvenc(jmc_job *j, int width, int height, int stride, int in_fmt)
{
    if (!j || width < 0 || height < 0 || stride < width) { jmc_set_error("jmc_job_nvenc: bad geometry"); return JMC_ERR_INVALID; }
    if (in_fmt == 0x1) j->op = JMC_OP_NV12_TO_SURF;                        /* NV_ENC_BUFFER_FORMAT_NV12, nv_enc.cpp:1029 */
    else if (in_fmt == 0x10) j->op = JMC_OP_I420_TO_SURF;                  /* NV_ENC_BUFFER_FORMAT_YV12 read as I420, :1041 */
    else { jmc_set_error("jmc_job_nvenc: format 0x%x has no NV12 surface conversion", in_fmt); return JMC_ERR_INVALID; }
    j->width = width; j->height = height; j->pitch = stride;
    j->surf_y_off = 0;
    j->surf_uv_off = (int64_t)stride * height;                             /* :1069 */
    const int64_t y_len = (int64_t)width * height;                         /* :1054 */
    j->tight_u_off = y_len;                                                /* :1055 */
    j->tight_v_off = y_len * 5 / 4;                                        /* :1056 */
    return JMC_OK;
}

int jmc_job_rgb(jmc_job *j, int width, int height, int pitch, int rgb_pitch, int fused)
{
    int r = jmc_job_nvdec(j, width, height, pitch, 1);
    if (r) return r;
    if (rgb_pitch < 3 * width) { jmc_set_error("jmc_job_rgb: rgb_pitch < 3*width"); return JMC_ERR_INVALID; }
    j->op = fused ? JMC_OP_NV12_TO_I420_RGB24 : JMC_OP_NV12_TO_RGB24;
    j->rgb_pitch = rgb_pitch;
    return JMC_OK;
}

int jmc_job_argb(jmc_job *j, int width, int height, int pitch, int argb_pitch)
{
    int r = jmc_job_nvdec(j, width, height, pitch, 1);
    if (r) return r;
    if (argb_pitch < 4 * width) { jmc_set_error("jmc_job_argb: argb_pitch < 4*width"); return JMC_ERR_INVALID; }
    j->op = JMC_OP_NV12_TO_ARGB32;
    j->rgb_pitch = argb_pitch;
    return JMC_OK;
}

int jmc_job_rgb_to_nv12(jmc_job *j, int width, int height, int rgb_pitch, int stride)
{
    int r = jmc_job_nvenc(j, width, height, stride, 0x1);        /* surface geometry of nv_enc.cpp:1029-1069 */
    if (r) return r;
    if (rgb_pitch < 3 * width) { jmc_set_error("jmc_job_rgb_to_nv12: rgb_pitch < 3*width"); return JMC_ERR_INVALID; }
    j->op = JMC_OP_RGB24_TO_SURF;
    j->rgb_pitch = rgb_pitch;
    return JMC_OK;
}

int64_t jmc_job_algorithmic_bytes(const jmc_job *j)
{
    if (!j) return 0;
    const int64_t w = j->width, h = j->height;
    const int64_t luma = w * h;
    int64_t yuv;                                   /* bytes of the frame actually read (= written for YUV->YUV) */
    switch (j->op) {
    case JMC_OP_NV12_TO_NV12: case JMC_OP_NV12_TO_SURF: yuv = luma + w * (h >> 1); break;
    default: yuv = luma + 2 * (w >> 1) * (h >> 1); break;
    }
    switch (j->op) {
    case JMC_OP_NV12_TO_RGB24: return luma + w * ((h + 1) >> 1) + 3 * luma;          /* reads every chroma row it uses */
    case JMC_OP_NV12_TO_ARGB32: return luma + w * ((h + 1) >> 1) + 4 * luma;
    case JMC_OP_RGB24_TO_SURF: return 3 * luma + luma + 2 * (w >> 1) * (h >> 1);
    case JMC_OP_NV12_TO_I420_RGB24: return luma + w * ((h + 1) >> 1) + yuv + 3 * luma;
    default: return 2 * yuv;
    }
}

/* ---- conversion ------------------------------------------------------------------------------ */
int jmc_convert(jmc_ctx *c, const jmc_job *job, void *stream)
{
    JMC_BIND(c);
    return jmc_launch_job(c, job, stream ? (cudaStream_t)stream : c->stream[0]);
}

int jmc_convert_timed(jmc_ctx *c, const jmc_job *job, int iters, float *ms_per_launch)
{
    JMC_BIND(c);
    if (iters < 1 || !ms_per_launch) return JMC_ERR_INVALID;
    cudaEvent_t a, b;
    JMC_CUDA(cudaEventCreate(&a));
    JMC_CUDA(cudaEventCreate(&b));
    cudaStream_t s = c->stream[0];
    JMC_CUDA(cudaStreamSynchronize(s));
    JMC_CUDA(cudaEventRecord(a, s));
    for (int i = 0; i < iters; i++) {
        int r = jmc_launch_job(c, job, s);
        if (r) { cudaEventDestroy(a); cudaEventDestroy(b); return r; }
    }
    JMC_CUDA(cudaEventRecord(b, s));
    JMC_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    JMC_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *ms_per_launch = ms / (float)iters;
    return JMC_OK;
}

/* ---- host link probe ------------------------------------------------------------------------- */
int jmc_link_probe(jmc_ctx *c, size_t bytes_per_copy, int copies, int mode, jmc_link_rates *out)
{
    JMC_BIND(c);
    const bool wc = (mode & 4) != 0;                       /* write-combined (uncached, unsnooped) host source for the upload */
    mode &= 3;
    if (!out || bytes_per_copy == 0 || copies == 0 || mode < 1) { jmc_set_error("jmc_link_probe: bad arguments"); return JMC_ERR_INVALID; }
    out->h2d_gbs = out->d2h_gbs = 0.0;
    void *h_up = nullptr, *h_down = nullptr, *d_up = nullptr, *d_down = nullptr;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    cudaError_t e = cudaSuccess;
    if (mode & 1) { e = cudaHostAlloc(&h_up, bytes_per_copy, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault); if (e == cudaSuccess) e = cudaMalloc(&d_up, bytes_per_copy); }
    if (e == cudaSuccess && (mode & 2)) { e = cudaHostAlloc(&h_down, bytes_per_copy, cudaHostAllocDefault); if (e == cudaSuccess) e = cudaMalloc(&d_down, bytes_per_copy); }
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&ev[i]);
    if (e == cudaSuccess && h_up) memset(h_up, 0x5A, bytes_per_copy);
    if (e == cudaSuccess && d_down) e = cudaMemset(d_down, 0xA5, bytes_per_copy);
    /* one warm-up copy each way, then the timed ones; the two directions run on the upload and the delivery stream.
     * copies > 0: that many copies per direction.  copies < 0: keep copying for -copies milliseconds of host time (at most
     * two copies queued per direction), so that several GPUs probed together are measured over the SAME window -- with a
     * fixed number of copies the faster GPUs finish early and the slower ones then run alone, which overstates what the
     * box delivers concurrently. */
    const bool timed = copies < 0;
    const double window_s = timed ? -copies * 1e-3 : 0.0;
    long long done_up = 0, done_down = 0;
    cudaEvent_t mark[2][2] = { { nullptr, nullptr }, { nullptr, nullptr } };
    for (int d = 0; d < 2 && e == cudaSuccess; d++) for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&mark[d][k], cudaEventDisableTiming);
    for (int rep = 0; rep < 2 && e == cudaSuccess; rep++) {
        const int n = rep == 0 ? 1 : (timed ? 0x7fffffff : copies);
        if (mode & 1) e = cudaEventRecord(ev[0], c->stream[1]);
        if (e == cudaSuccess && (mode & 2)) e = cudaEventRecord(ev[2], c->stream[2]);
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < n && e == cudaSuccess; i++) {
            if (timed && rep == 1) {
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) >= window_s) break;
                if (i >= 2) {                                              /* at most two copies queued per direction */
                    if (mode & 1) e = cudaEventSynchronize(mark[0][i & 1]);
                    if (e == cudaSuccess && (mode & 2)) e = cudaEventSynchronize(mark[1][i & 1]);
                }
            }
            if (e == cudaSuccess && (mode & 1)) { e = cudaMemcpyAsync(d_up, h_up, bytes_per_copy, cudaMemcpyHostToDevice, c->stream[1]); if (e == cudaSuccess) e = cudaEventRecord(mark[0][i & 1], c->stream[1]); }
            if (e == cudaSuccess && (mode & 2)) { e = cudaMemcpyAsync(h_down, d_down, bytes_per_copy, cudaMemcpyDeviceToHost, c->stream[2]); if (e == cudaSuccess) e = cudaEventRecord(mark[1][i & 1], c->stream[2]); }
            if (rep == 1) { done_up += (mode & 1) ? 1 : 0; done_down += (mode & 2) ? 1 : 0; }
        }
        if (e == cudaSuccess && (mode & 1)) e = cudaEventRecord(ev[1], c->stream[1]);
        if (e == cudaSuccess && (mode & 2)) e = cudaEventRecord(ev[3], c->stream[2]);
        if (e == cudaSuccess && (mode & 1)) e = cudaEventSynchronize(ev[1]);
        if (e == cudaSuccess && (mode & 2)) e = cudaEventSynchronize(ev[3]);
    }
    if (e == cudaSuccess) {
        float ms = 0.f;
        if (mode & 1) { e = cudaEventElapsedTime(&ms, ev[0], ev[1]); if (e == cudaSuccess && ms > 0) out->h2d_gbs = (double)bytes_per_copy * done_up / (ms * 1e-3) / 1e9; }
        if (e == cudaSuccess && (mode & 2)) { e = cudaEventElapsedTime(&ms, ev[2], ev[3]); if (e == cudaSuccess && ms > 0) out->d2h_gbs = (double)bytes_per_copy * done_down / (ms * 1e-3) / 1e9; }
    }
    for (int d = 0; d < 2; d++) for (int k = 0; k < 2; k++) if (mark[d][k]) cudaEventDestroy(mark[d][k]);
    for (int i = 0; i < 4; i++) if (ev[i]) cudaEventDestroy(ev[i]);
    if (h_up) cudaFreeHost(h_up);
    if (h_down) cudaFreeHost(h_down);
    if (d_up) cudaFree(d_up);
    if (d_down) cudaFree(d_down);
    if (e != cudaSuccess) return jmc_cuda_fail(e, "jmc_link_probe");
    return JMC_OK;
}

/* ---- events ---------------------------------------------------------------------------------- */
int jmc_event_create(jmc_ctx *c, jmc_event **out)
{
    JMC_BIND(c);
    if (!out) return JMC_ERR_INVALID;
    cudaEvent_t e;
    JMC_CUDA(cudaEventCreate(&e));
    *out = (jmc_event *)e;
    return JMC_OK;
}

int jmc_event_destroy(jmc_ctx *c, jmc_event *ev)
{
    JMC_BIND(c);
    JMC_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return JMC_OK;
}

int jmc_event_record(jmc_ctx *c, jmc_event *ev, int which)
{
    JMC_BIND(c);
    if (which < 0 || which > 2) return JMC_ERR_INVALID;
    JMC_CUDA(cudaEventRecord((cudaEvent_t)ev, c->stream[which]));
    return JMC_OK;
}

int jmc_event_elapsed_ms(jmc_ctx *c, jmc_event *start, jmc_event *stop, float *ms)
{
    JMC_BIND(c);
    if (!ms) return JMC_ERR_INVALID;
    /* both ends: `start` may sit on another stream than `stop`, and nothing orders it before the work `stop` follows
     * (a start recorded on an otherwise idle stream) */
    JMC_CUDA(cudaEventSynchronize((cudaEvent_t)start));
    JMC_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
    JMC_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return JMC_OK;
}

} /* extern "C" */

/* ---- host-delivery pipeline ------------------------------------------------------------------ */
struct jmc_slot {
    void *d_in, *d_out, *d_out2;
    cudaEvent_t uploaded, converted, delivered;
    bool in_flight;
};

struct jmc_pipeline {
    jmc_ctx *ctx;
    jmc_job shape;
    size_t in_bytes, out_bytes, out2_bytes;   /* per frame */
    int depth, next;
    jmc_slot *slot;
    uint64_t h2d, d2h;
};

static bool op_is_decode_side(int op)
{
    return op == JMC_OP_NV12_TO_NV12 || op == JMC_OP_NV12_TO_I420 || op == JMC_OP_NV12_TO_RGB24 || op == JMC_OP_NV12_TO_I420_RGB24 ||
           op == JMC_OP_NV12_TO_ARGB32;
}

extern "C" {

int jmc_pipeline_create(jmc_ctx *c, const jmc_job *shape, size_t surf_bytes, int depth, jmc_pipeline **out)
{
    JMC_BIND(c);
    if (!shape || !out || depth < 1 || shape->n_frames < 1) { jmc_set_error("jmc_pipeline_create: bad arguments"); return JMC_ERR_INVALID; }
    *out = nullptr;
    const size_t tight = (size_t)jmc_tight_bytes(shape->width, shape->height);
    const size_t rgb = (size_t)shape->rgb_pitch * shape->height;
    jmc_pipeline *p = new (std::nothrow) jmc_pipeline();
    if (!p) return JMC_ERR_NOMEM;
    p->ctx = c; p->shape = *shape; p->depth = depth; p->next = 0; p->h2d = p->d2h = 0;
    switch (shape->op) {
    case JMC_OP_NV12_TO_NV12: case JMC_OP_NV12_TO_I420: p->in_bytes = surf_bytes; p->out_bytes = tight; p->out2_bytes = 0; break;
    case JMC_OP_NV12_TO_RGB24: case JMC_OP_NV12_TO_ARGB32: p->in_bytes = surf_bytes; p->out_bytes = rgb; p->out2_bytes = 0; break;
    case JMC_OP_NV12_TO_I420_RGB24: p->in_bytes = surf_bytes; p->out_bytes = tight; p->out2_bytes = rgb; break;
    case JMC_OP_NV12_TO_SURF: case JMC_OP_I420_TO_SURF: p->in_bytes = tight; p->out_bytes = surf_bytes; p->out2_bytes = 0; break;
    case JMC_OP_RGB24_TO_SURF: p->in_bytes = rgb; p->out_bytes = surf_bytes; p->out2_bytes = 0; break;
    default: delete p; jmc_set_error("jmc_pipeline_create: unknown op"); return JMC_ERR_INVALID;
    }
    if (surf_bytes == 0) { delete p; jmc_set_error("jmc_pipeline_create: surf_bytes == 0"); return JMC_ERR_INVALID; }
    p->slot = (jmc_slot *)calloc((size_t)depth, sizeof(jmc_slot));
    if (!p->slot) { delete p; return JMC_ERR_NOMEM; }
    const size_t n = (size_t)shape->n_frames;
    for (int i = 0; i < depth; i++) {
        jmc_slot &s = p->slot[i];
        cudaError_t e = cudaMalloc(&s.d_in, p->in_bytes * n);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_out, p->out_bytes * n);
        if (e == cudaSuccess && p->out2_bytes) e = cudaMalloc(&s.d_out2, p->out2_bytes * n);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.converted, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.delivered, cudaEventDisableTiming);
        /* encode side: the surface padding is never written by the kernel; give it a defined value once */
        if (e == cudaSuccess && !op_is_decode_side(shape->op)) e = cudaMemset(s.d_out, 0, p->out_bytes * n);
        if (e != cudaSuccess) { int r = jmc_cuda_fail(e, "pipeline slot allocation"); jmc_pipeline_destroy(p); return r; }
    }
    *out = p;
    return JMC_OK;
}

int jmc_pipeline_destroy(jmc_pipeline *p)
{
    if (!p) return JMC_ERR_INVALID;
    JMC_BIND(p->ctx);
    jmc_ctx_sync(p->ctx);
    for (int i = 0; p->slot && i < p->depth; i++) {
        jmc_slot &s = p->slot[i];
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_out2) cudaFree(s.d_out2);
        if (s.uploaded) cudaEventDestroy(s.uploaded);
        if (s.converted) cudaEventDestroy(s.converted);
        if (s.delivered) cudaEventDestroy(s.delivered);
    }
    free(p->slot);
    delete p;
    return JMC_OK;
}

static int pipeline_submit_impl(jmc_pipeline *p, const void *host_in, const void *dev_in, void *host_out, void *host_out2, int n_frames);

int jmc_pipeline_submit(jmc_pipeline *p, const void *host_in, const void *dev_in, void *host_out, void *host_out2, int n_frames)
{
    const int r = pipeline_submit_impl(p, host_in, dev_in, host_out, host_out2, n_frames);
    if (r < 0 && p) {
        /* a submit that failed half way may have queued copies from / to the caller's buffers and will never be waited
         * for through its slot: nothing of it is left in flight when the error is returned (the message stays) */
        jmc_device_guard g(p->ctx);
        if (!g.err) for (int i = 0; i < 3; i++) if (cudaStreamSynchronize(p->ctx->stream[i]) != cudaSuccess) cudaGetLastError();
    }
    return r;
}

static int pipeline_submit_impl(jmc_pipeline *p, const void *host_in, const void *dev_in, void *host_out, void *host_out2, int n_frames)
{
    if (!p) return JMC_ERR_INVALID;
    JMC_BIND(p->ctx);
    if (n_frames < 1 || n_frames > p->shape.n_frames || (!host_in && !dev_in)) { jmc_set_error("jmc_pipeline_submit: bad arguments"); return JMC_ERR_INVALID; }
    jmc_ctx *c = p->ctx;
    const int si = p->next;
    jmc_slot &s = p->slot[si];
    if (s.in_flight) { JMC_CUDA(cudaEventSynchronize(s.delivered)); s.in_flight = false; }
    const size_t n = (size_t)n_frames;
    const void *src = dev_in;
    if (host_in) {
        /* upload stream: the slot's input buffer is free once its previous conversion has run */
        JMC_CUDA(cudaStreamWaitEvent(c->stream[1], s.converted, 0));
        const jmc_job &g = p->shape;
        const bool rows_only = jmc_env().pipeline_h2d_2d && op_is_decode_side(g.op) && g.width < g.pitch && g.surf_y_off == 0 &&
                               g.surf_uv_off == (int64_t)g.pitch * g.height && p->in_bytes % (size_t)g.pitch == 0;
        if (rows_only) {
            /* NV12 surfaces back to back = one 2-D array of `pitch`-byte rows: the DMA engine skips the
             * padding, so only width/pitch of the bytes cross PCIe (measured: same GB/s of useful bytes). */
            const size_t rows = p->in_bytes / (size_t)g.pitch * n;
            JMC_CUDA(cudaMemcpy2DAsync(s.d_in, (size_t)g.pitch, host_in, (size_t)g.pitch, (size_t)g.width, rows,
                                       cudaMemcpyHostToDevice, c->stream[1]));
            p->h2d += (size_t)g.width * rows;
        } else {
            JMC_CUDA(cudaMemcpyAsync(s.d_in, host_in, p->in_bytes * n, cudaMemcpyHostToDevice, c->stream[1]));
            p->h2d += p->in_bytes * n;
        }
        JMC_CUDA(cudaEventRecord(s.uploaded, c->stream[1]));
        JMC_CUDA(cudaStreamWaitEvent(c->stream[0], s.uploaded, 0));
        src = s.d_in;
    }
    /* convert stream: the slot's output buffers are free once their previous delivery has run */
    JMC_CUDA(cudaStreamWaitEvent(c->stream[0], s.delivered, 0));
    jmc_job j = p->shape;
    j.n_frames = n_frames;
    const bool dec = op_is_decode_side(j.op);
    j.surf.list = j.tight.list = j.rgb.list = nullptr;
    if (dec) {
        j.surf.base = (void *)src; j.surf.stride = p->in_bytes;
        if (j.op == JMC_OP_NV12_TO_RGB24 || j.op == JMC_OP_NV12_TO_ARGB32) { j.rgb.base = s.d_out; j.rgb.stride = p->out_bytes; }
        else { j.tight.base = s.d_out; j.tight.stride = p->out_bytes; }
        if (j.op == JMC_OP_NV12_TO_I420_RGB24) { j.rgb.base = s.d_out2; j.rgb.stride = p->out2_bytes; }
    } else {
        if (j.op == JMC_OP_RGB24_TO_SURF) { j.rgb.base = (void *)src; j.rgb.stride = p->in_bytes; }
        else { j.tight.base = (void *)src; j.tight.stride = p->in_bytes; }
        j.surf.base = s.d_out; j.surf.stride = p->out_bytes;
    }
    int r = jmc_launch_job(c, &j, c->stream[0]);
    if (r) return r;
    JMC_CUDA(cudaEventRecord(s.converted, c->stream[0]));
    /* delivery stream */
    JMC_CUDA(cudaStreamWaitEvent(c->stream[2], s.converted, 0));
    if (host_out) { JMC_CUDA(cudaMemcpyAsync(host_out, s.d_out, p->out_bytes * n, cudaMemcpyDeviceToHost, c->stream[2])); p->d2h += p->out_bytes * n; }
    if (host_out2 && p->out2_bytes) { JMC_CUDA(cudaMemcpyAsync(host_out2, s.d_out2, p->out2_bytes * n, cudaMemcpyDeviceToHost, c->stream[2])); p->d2h += p->out2_bytes * n; }
    JMC_CUDA(cudaEventRecord(s.delivered, c->stream[2]));
    s.in_flight = true;
    p->next = (si + 1) % p->depth;
    return si;
}

int jmc_pipeline_wait(jmc_pipeline *p, int slot)
{
    if (!p || slot < 0 || slot >= p->depth) return JMC_ERR_INVALID;
    JMC_BIND(p->ctx);
    if (p->slot[slot].in_flight) { JMC_CUDA(cudaEventSynchronize(p->slot[slot].delivered)); p->slot[slot].in_flight = false; }
    return JMC_OK;
}

int jmc_pipeline_drain(jmc_pipeline *p)
{
    if (!p) return JMC_ERR_INVALID;
    for (int i = 0; i < p->depth; i++) { int r = jmc_pipeline_wait(p, i); if (r) return r; }
    return JMC_OK;
}

uint64_t jmc_pipeline_h2d_bytes(const jmc_pipeline *p) { return p ? p->h2d : 0; }
uint64_t jmc_pipeline_d2h_bytes(const jmc_pipeline *p) { return p ? p->d2h : 0; }

} /* extern "C" */
