/*
 * jm_nv_dec.cu -- the jm_nvdec_* drop-in API (include/jm_nv_dec.h) over the jmc_* layer.
 *
 * Mirrors the control flow of the reference's nv_dec/nv_dec.cpp with the surface-format path moved
 * onto the device:
 *
 *   reference                                         here
 *   ------------------------------------------------  ---------------------------------------------
 *   decode_frame: parse -> display queue (:368-403)   decode_frame: front-end -> display queue
 *     pop 1 frame, cuvidMapVideoFrame (:439)            pop 1 frame (device surface + pitch)
 *     cuMemcpyDtoH pitch*h*3/2, SYNC (:452)             ONE kernel: NV12 -> tight NV12 / I420 in HBM
 *   output_frame: CPU strip/de-interleave (:782-820)  output_frame: D2H of the TIGHT frame only
 *
 * Front-ends: JM_NVDEC_CODEC_RAW_NV12 (decoded surfaces as packets).  Bitstream codecs need the
 * NVDEC parser library; see jm_nvdec_init.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <deque>
#include <new>

#include "jm_nv_dec.h"
#include "jmc_internal.h"

#define NVDEC_MAX_FRAMES 10         /* display queue depth / decode surfaces, nv_dec/nv_dec.h:32 */
#define MAX_LEN_DEC_INFO 1024       /* nv_dec/nv_dec.h:33 */

namespace {

struct decoded_surface {            /* what cuvidMapVideoFrame yields: device pointer + pitch */
    uint8_t *dptr;
    int pitch, width, height;
    int pool_slot;                  /* >= 0: one of our upload surfaces; -1: caller-owned device memory */
};

struct nvdec_b200 {
    int device;
    int codec_type;
    int out_fmt;                    /* 0: NV12, else "YV12" = I420 (nv_dec.h:94) */
    bool inited, is_eof, is_exit;
    jmc_ctx *ctx;

    /* display queue (nv_dec.h:88-91) */
    std::deque<decoded_surface> *queue;

    /* upload surfaces standing in for the decoder's surfaces (RAW front-end, host payloads) */
    uint8_t *pool[NVDEC_MAX_FRAMES];
    bool pool_busy[NVDEC_MAX_FRAMES];
    size_t pool_bytes;

    /* current output frame (nv_dec.h:119-123): tight frame in device memory */
    uint8_t *d_tight;
    size_t d_tight_bytes;
    bool have_cur;
    int cur_w, cur_h;
    int disp_w, disp_h;             /* dec_create_info.ulTargetWidth/Height */

    uint32_t num_frames;
    struct timespec t_start;
    bool started;
    char dec_info[MAX_LEN_DEC_INFO];
};

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

int ensure_tight(nvdec_b200 *c, int w, int h)
{
    size_t need = (size_t)jmc_tight_bytes(w, h);
    if (need == 0) need = 1;
    if (c->d_tight && c->d_tight_bytes >= need) return 0;
    if (c->d_tight) { jmc_ctx_sync(c->ctx); jmc_free_device(c->ctx, c->d_tight); c->d_tight = nullptr; }
    void *p = nullptr;
    int r = jmc_alloc_device(c->ctx, need, &p);     /* replaces cuMemAllocHost(pitch*h*3/2), nv_dec.cpp:569 */
    if (r) return r;
    c->d_tight = (uint8_t *)p;
    c->d_tight_bytes = need;
    return 0;
}

/* RAW front-end: one packet = one decoded surface -> display queue */
int raw_packet(nvdec_b200 *c, const unsigned char *buf, int len)
{
    if (len < (int)sizeof(jm_nvdec_raw_packet)) return -1;
    jm_nvdec_raw_packet h;
    memcpy(&h, buf, sizeof(h));
    if (h.magic != JM_NVDEC_RAW_MAGIC || h.width < 0 || h.height < 0 || h.pitch < h.width) return -1;
    if ((int)c->queue->size() >= NVDEC_MAX_FRAMES) return -1;            /* all decode surfaces in use */
    decoded_surface s;
    s.width = h.width; s.height = h.height; s.pitch = h.pitch;
    if (h.flags & JM_NVDEC_RAW_DEVICE_PTR) {
        s.dptr = (uint8_t *)(uintptr_t)h.device_ptr;
        s.pool_slot = -1;
    } else {
        const size_t bytes = (size_t)h.pitch * h.height * 3 / 2;         /* nv_dec.cpp:453 */
        if ((size_t)len < sizeof(h) + bytes) return -1;
        if (bytes > c->pool_bytes) {                                     /* geometry grew: new surfaces */
            jmc_ctx_sync(c->ctx);
            if (!c->queue->empty()) return -1;
            for (int i = 0; i < NVDEC_MAX_FRAMES; i++) if (c->pool[i]) { jmc_free_device(c->ctx, c->pool[i]); c->pool[i] = nullptr; }
            c->pool_bytes = bytes;
        }
        int slot = -1;
        for (int i = 0; i < NVDEC_MAX_FRAMES; i++) if (!c->pool_busy[i]) { slot = i; break; }
        if (slot < 0) return -1;
        if (!c->pool[slot]) {
            void *p = nullptr;
            if (jmc_alloc_device(c->ctx, c->pool_bytes ? c->pool_bytes : 1, &p)) return -1;
            c->pool[slot] = (uint8_t *)p;
        }
        /* the "decode": the surface lands in HBM.  in_buf is consumed before we return. */
        cudaStream_t st = (cudaStream_t)jmc_ctx_stream(c->ctx, 0);
        if (cudaMemcpyAsync(c->pool[slot], buf + sizeof(h), bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
        c->pool_busy[slot] = true;
        s.dptr = c->pool[slot];
        s.pool_slot = slot;
    }
    if (!c->started) { clock_gettime(CLOCK_MONOTONIC, &c->t_start); c->started = true; }   /* nv_dec.cpp:537 */
    c->disp_w = h.width; c->disp_h = h.height;
    c->num_frames += 1;                                                   /* nv_dec.cpp:48 */
    c->queue->push_back(s);
    return 0;
}

/* nvdec_decode_output_frame, nv_dec.cpp:406-478 */
void output_stage(nvdec_b200 *c, int *got_frame)
{
    if (c->queue->empty()) {
        if (c->is_eof) { c->is_exit = true; show_info(c); }               /* :460-466 */
        return;
    }
    decoded_surface s = c->queue->front();
    c->queue->pop_front();
    if (ensure_tight(c, s.width, s.height) == 0) {
        jmc_job j;
        memset(&j, 0, sizeof(j));
        jmc_job_nvdec(&j, s.width, s.height, s.pitch, c->out_fmt);
        j.n_frames = 1;
        j.surf.base = s.dptr;
        j.tight.base = c->d_tight;
        /* same stream as the upload, so the surface slot can be recycled right away */
        if (jmc_convert(c->ctx, &j, nullptr) == JMC_OK) {
            c->have_cur = true;
            c->cur_w = s.width; c->cur_h = s.height;
            *got_frame = 1;                                               /* :455 */
        }
    }
    if (s.pool_slot >= 0) c->pool_busy[s.pool_slot] = false;              /* nvdec_frame_item_release, :458 */
}

} /* namespace */

extern "C" {

handle_nvdec jm_nvdec_create_handle(void)
{
    nvdec_b200 *c = (nvdec_b200 *)calloc(1, sizeof(nvdec_b200));
    if (!c) return nullptr;
    const char *e = getenv("JMC_DEVICE");
    c->device = e ? atoi(e) : 0;
    return c;
}

int jm_nvdec_set_device(int device, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || c->inited) return -1;
    c->device = device;
    return 0;
}

int jm_nvdec_init(int codec_type, int out_fmt, char *extra_data, int len, handle_nvdec handle)
{
    (void)extra_data; (void)len;
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c) return -1;
    c->out_fmt = out_fmt;
    c->codec_type = codec_type;
    int r = jmc_ctx_create(c->device, &c->ctx);
    if (r == JMC_ERR_NO_DEVICE) return jmc_device_count() <= 0 ? -2 : -3;   /* nvdec_cuda_init, nv_dec.cpp:219-231 */
    if (r) return -1;
    c->queue = new (std::nothrow) std::deque<decoded_surface>();
    if (!c->queue) return -1;
    c->inited = true;
    if (codec_type != JM_NVDEC_CODEC_RAW_NV12) {
        /* Bitstream codecs go through the NVDEC parser/decoder (cuvidCreateVideoParser, nv_dec.cpp:278-366);
         * that front-end is not wired in this build. */
        jmc_set_error("jm_nvdec_init: codec %d needs the NVDEC parser front-end (not available); use JM_NVDEC_CODEC_RAW_NV12", codec_type);
        return -4;
    }
    return 0;
}

int jm_nvdec_deinit(handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c) return -1;
    if (c->ctx) {
        jmc_ctx_sync(c->ctx);
        for (int i = 0; i < NVDEC_MAX_FRAMES; i++) if (c->pool[i]) jmc_free_device(c->ctx, c->pool[i]);
        if (c->d_tight) jmc_free_device(c->ctx, c->d_tight);
        jmc_ctx_destroy(c->ctx);
    }
    delete c->queue;
    free(c);
    return 0;
}

int jm_nvdec_decode_frame(unsigned char *in_buf, int in_data_len, int *got_frame, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (got_frame) *got_frame = 0;
    if (!c || !got_frame) return 0;
    if (!c->inited || !c->ctx || !c->queue) return 0;         /* decoder never created: the reference swallows -1 (nv_dec.cpp:414-417,491-493) */
    if (!c->is_eof) {                                          /* nv_dec.cpp:486-488 */
        if (in_buf && in_data_len > 0) {
            if (c->codec_type == JM_NVDEC_CODEC_RAW_NV12) raw_packet(c, in_buf, in_data_len);
        } else {
            c->is_eof = true;                                  /* CUVID_PKT_ENDOFSTREAM, nv_dec.cpp:389-392 */
        }
    }
    output_stage(c, got_frame);
    return 0;
}

int jm_nvdec_output_frame(unsigned char *out_buf, int *out_len, handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c || !c->have_cur) return -1;                         /* nv_dec.cpp:757-758 */
    if (!c->d_tight || !out_buf || !out_len) return -1;        /* :768-771 */
    const int need = c->cur_w * c->cur_h * 3 / 2;
    if (*out_len < need) return -2;                            /* :773-774 */
    *out_len = 0;                                              /* :776 */
    cudaStream_t st = (cudaStream_t)jmc_ctx_stream(c->ctx, 0);
    /* Only the tight frame crosses PCIe.  Pinned out_buf: direct DMA; pageable: the driver stages it.
     * For odd sizes the reference writes fewer than w*h*3/2 bytes and leaves the rest of out_buf
     * untouched (h>>1 chroma rows, w>>1 samples: nv_dec.cpp:792-796,807-818): copy exactly those. */
    const size_t luma = (size_t)c->cur_w * c->cur_h;
    const size_t written = c->out_fmt == 0 ? luma + (size_t)(c->cur_h >> 1) * c->cur_w
                                           : luma + 2 * (size_t)(c->cur_w >> 1) * (c->cur_h >> 1);
    if (written > 0 && cudaMemcpyAsync(out_buf, c->d_tight, written, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    *out_len = need;                                           /* :824 */
    return need;                                               /* :827 */
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

} /* extern "C" */
