/*
 * jmnv_enc.cu -- the jm_nvenc_* drop-in API (include/jmnv_enc.h): encoder INPUT path only.
 *
 *   reference (nv_enc/nv_enc.cpp)                       here
 *   --------------------------------------------------  -------------------------------------------
 *   nvenc_register_frame: cuMemAllocPitch per surface   same pool (10 surfaces), no per-surface
 *     + cuMemAlloc U,V temps re-allocated 10x (:954-)     temp leak: one staging frame per handle
 *   NV12 : cuMemcpy2D H->D (:1029-1040)                 cudaMemcpy2DAsync H->D (DMA adds the pitch)
 *   YV12 : cuMemcpy2D Y + 2x cuMemcpyHtoD + byte-wise   ONE H->D of the tight frame + ONE kernel
 *          InterleaveUV<<<32x16>>> (:1041-1081)           (16-byte vector copy + prmt interleave)
 *   ARGB : flat cuMemcpyHtoD ignoring pitch (:1096)     cudaMemcpy2DAsync honouring the pitch
 *   nvEncMapInputResource / nvEncEncodePicture          absent on B200 (no NVENC engine)
 */
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "jmc_internal.h"
#include "jmnv_enc.h"

namespace {

struct enc_surface {
    uint8_t *dptr;          /* in_cuda_surf, nv_enc.h:112 */
    size_t pitch;           /* in_cuda_stride */
    int lock_count;         /* nv_enc.h:118 */
    uint64_t order;         /* upload sequence number, for release-oldest */
};

struct nvenc_b200 {
    int device;
    bool inited, surface_only;
    jmc_ctx *ctx;
    nv_enc_param param;
    int width, height, format;
    enc_surface surf[JM_NVENC_NUM_SURFACES];
    int rows;               /* allocated rows per surface */
    uint8_t *d_stage;       /* tight input frame in device memory */
    size_t stage_bytes;
    int last;               /* surface filled by the most recent enc_frame */
    uint64_t seq;
};

bool is_rgb(int f) { return f == JM_NVENC_FMT_ARGB || f == JM_NVENC_FMT_ABGR; }

/* Everything jm_nvenc_init allocated: surfaces, staging frame, context.  Used by deinit, by a failing init (the
 * reference's nvenc_register_frame leaks on both paths, nv_enc.cpp:954-1007) and by init on a live handle. */
void release_all(nvenc_b200 *c)
{
    if (c->ctx) {
        jmc_ctx_sync(c->ctx);
        for (int i = 0; i < JM_NVENC_NUM_SURFACES; i++) {
            if (c->surf[i].dptr) jmc_free_device(c->ctx, c->surf[i].dptr);
            c->surf[i] = enc_surface();
        }
        if (c->d_stage) jmc_free_device(c->ctx, c->d_stage);
        c->d_stage = nullptr;
        c->stage_bytes = 0;
        jmc_ctx_destroy(c->ctx);
        c->ctx = nullptr;
    }
    c->inited = false;
    c->last = -1;
}

} /* namespace */

extern "C" {

handle_nvenc jm_nvenc_create_handle(void)
{
    nvenc_b200 *c = (nvenc_b200 *)calloc(1, sizeof(nvenc_b200));      /* new + memset, nv_enc.cpp:28-43 */
    if (!c) return nullptr;
    const char *e = getenv("JMC_DEVICE");
    c->device = e ? atoi(e) : 0;
    c->last = -1;
    return c;
}

int jm_nvenc_set_device(int device, handle_nvenc handle)
{
    nvenc_b200 *c = (nvenc_b200 *)handle;
    if (!c || c->inited) return -1;
    c->device = device;
    return 0;
}

int jm_nvenc_init(nv_enc_param *in_param, handle_nvenc handle)
{
    nvenc_b200 *c = (nvenc_b200 *)handle;
    if (!c || !in_param) return JM_NVENC_ERR_INVALID_PARAM;
    if (c->ctx) release_all(c);                                       /* init on a live handle: start over, leak nothing */
    c->param = *in_param;
    c->width = in_param->src_width;
    c->height = in_param->src_height;
    c->format = in_param->in_fmt;
    if (c->width <= 0 || c->height <= 0) return JM_NVENC_ERR_INVALID_PARAM;
    if (!(c->format == JM_NVENC_FMT_NV12 || c->format == JM_NVENC_FMT_YV12 || is_rgb(c->format))) return JM_NVENC_ERR_INVALID_PARAM;

    int r = jmc_ctx_create(c->device, &c->ctx);                       /* nvenc_cuda_init, nv_enc.cpp:232-276 */
    if (r == JMC_ERR_NO_DEVICE) return JM_NVENC_ERR_NO_ENCODE_DEVICE;
    if (r) return JM_NVENC_ERR_GENERIC;

    const char *env = getenv("JMC_NVENC_SURFACE_ONLY");
    c->surface_only = in_param->codec_id == JM_NVENC_CODEC_SURFACE_ONLY || (env && atoi(env) != 0);
    if (!c->surface_only) {
        /* nvenc_loading_libraries (nv_enc.cpp:340-380): no NVENC engine on B200, the driver ships no usable encoder. */
        void *lib = dlopen("libnvidia-encode.so.1", RTLD_LAZY | RTLD_LOCAL);
        if (lib) dlclose(lib);
        jmc_set_error("jm_nvenc_init: no NVENC engine on this device (B200); use JM_NVENC_CODEC_SURFACE_ONLY for the input path");
        jmc_ctx_destroy(c->ctx);
        c->ctx = nullptr;
        return JM_NVENC_ERR_NO_ENCODE_DEVICE;
    }

    /* nvenc_register_frame (nv_enc.cpp:966-980): NV12/YV12 w x h*3/2, ARGB w*4 x h */
    const size_t wbytes = is_rgb(c->format) ? (size_t)c->width * 4 : (size_t)c->width;
    c->rows = is_rgb(c->format) ? c->height : c->height * 3 / 2;
    for (int i = 0; i < JM_NVENC_NUM_SURFACES; i++) {
        void *p = nullptr;
        size_t pitch = 0;
        if (jmc_alloc_pitched(c->ctx, wbytes, (size_t)c->rows, &p, &pitch) != JMC_OK) { release_all(c); return JM_NVENC_ERR_GENERIC; }
        c->surf[i].dptr = (uint8_t *)p;
        c->surf[i].pitch = pitch;
        if (jmc_memset_device(c->ctx, p, 0, pitch * (size_t)c->rows) != JMC_OK) { release_all(c); return JM_NVENC_ERR_GENERIC; }
    }
    if (c->format == JM_NVENC_FMT_YV12) {
        /* one staging frame; the reference stages U and V separately in uv_tmp_ptr[0..1] (:972-973) */
        c->stage_bytes = (size_t)c->width * c->height * 3 / 2 + 16;
        void *p = nullptr;
        if (jmc_alloc_device(c->ctx, c->stage_bytes, &p) != JMC_OK) { release_all(c); return JM_NVENC_ERR_GENERIC; }
        c->d_stage = (uint8_t *)p;
    }
    c->inited = true;
    return JM_NVENC_SUCCESS;
}

int jm_nvenc_deinit(handle_nvenc handle)
{
    nvenc_b200 *c = (nvenc_b200 *)handle;
    if (!c) return -1;
    release_all(c);
    free(c);
    return 0;
}

int jm_nvenc_enc_frame(const unsigned char *in_yuv_buf, const int yuv_len, int *got_packet, handle_nvenc handle)
{
    nvenc_b200 *c = (nvenc_b200 *)handle;
    if (got_packet) *got_packet = 0;
    if (!c || !c->inited) return -1;
    if (!in_yuv_buf || yuv_len <= 0) return 0;                        /* EOS, nv_enc.cpp:113-117 */
    jmc_device_guard guard(c->ctx);                                    /* CCudaAutoLock, nv_enc.cpp:1025; the caller's device is restored on return */
    if (guard.err) return JM_NVENC_ERR_GENERIC;
    /* a buffer shorter than the format needs would be over-read by the DMA (the reference does over-read,
     * nv_enc.cpp:1029-1040,1096): refuse it instead.  YV12 keeps the reference's clamp-to-yuv_len behaviour. */
    {
        const int64_t px = (int64_t)c->width * c->height;
        const int64_t want = is_rgb(c->format) ? px * 4 : (c->format == JM_NVENC_FMT_NV12 ? (int64_t)c->width * (c->height * 3 / 2) : 0);
        if (want > (int64_t)yuv_len) {
            jmc_set_error("jm_nvenc_enc_frame: yuv_len %d is shorter than the %lld bytes a %dx%d frame of this format holds", yuv_len, (long long)want, c->width, c->height);
            return JM_NVENC_ERR_INVALID_PARAM;
        }
    }

    int idx = -1;                                                     /* nvenc_get_free_frame, :916-927 */
    for (int i = 0; i < JM_NVENC_NUM_SURFACES; i++) if (!c->surf[i].lock_count) { idx = i; break; }
    if (idx < 0) return -1;                                           /* :90-93 */
    enc_surface &s = c->surf[idx];
    s.lock_count = 1;
    s.order = ++c->seq;

    cudaStream_t st = (cudaStream_t)jmc_ctx_stream(c->ctx, 0);
    cudaError_t e = cudaSuccess;
    const int w = c->width, h = c->height;
    if (c->format == JM_NVENC_FMT_NV12) {                             /* :1029-1040 */
        e = cudaMemcpy2DAsync(s.dptr, s.pitch, in_yuv_buf, (size_t)w, (size_t)w, (size_t)(h * 3 / 2), cudaMemcpyHostToDevice, st);
    } else if (c->format == JM_NVENC_FMT_YV12) {                      /* :1041-1081 */
        const size_t y_len = (size_t)w * h;
        /* bytes the reference touches: Y, then y_len/4 at y_len and y_len/4 at y_len*5/4 (:1055-1056) */
        size_t n = y_len * 5 / 4 + y_len / 4;
        if (n > (size_t)yuv_len) n = (size_t)yuv_len;
        if (n > c->stage_bytes) n = c->stage_bytes;
        e = cudaMemcpyAsync(c->d_stage, in_yuv_buf, n, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            jmc_job j;
            memset(&j, 0, sizeof(j));
            jmc_job_nvenc(&j, w, h, (int)s.pitch, JM_NVENC_FMT_YV12);
            j.n_frames = 1;
            j.tight.base = c->d_stage;
            j.surf.base = s.dptr;
            /* the pool's surfaces are ours: zeroed at init and only ever written here, so zeroing the padding behind a row
             * end again changes nothing -- and saves the kernel the read-merge of every row's last sector */
            j.flags = JMC_JOB_PAD_ZERO;
            if (jmc_convert(c->ctx, &j, nullptr) != JMC_OK) {
                cudaStreamSynchronize(st);                            /* the upload may still be reading in_yuv_buf */
                s.lock_count = 0;
                return JM_NVENC_ERR_GENERIC;
            }
        }
    } else {                                                          /* ARGB/ABGR, :1083-1097 (pitch honoured) */
        e = cudaMemcpy2DAsync(s.dptr, s.pitch, in_yuv_buf, (size_t)w * 4, (size_t)w * 4, (size_t)h, cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);              /* in_yuv_buf is consumed before return */
    if (e != cudaSuccess) {
        jmc_cuda_fail(e, "jm_nvenc_enc_frame upload");
        cudaStreamSynchronize(st);                                    /* nothing of ours reads in_yuv_buf after we return */
        cudaGetLastError();
        s.lock_count = 0;
        return JM_NVENC_ERR_GENERIC;
    }
    c->last = idx;
    /* no NVENC: nothing is encoded, so no packet ever becomes ready */
    return 0;
}

int jm_nvenc_get_bitstream(unsigned char *out_buf, int *out_data_len, int *is_keyframe, handle_nvenc handle)
{
    (void)out_buf; (void)is_keyframe; (void)handle;
    if (out_data_len) *out_data_len = 0;
    return -1;                                                        /* no packet ready, nv_enc.cpp:175-178 */
}

int jm_nvenc_get_spspps_len(int *sps_len, int *pps_len, handle_nvenc handle)
{
    (void)handle;
    if (sps_len) *sps_len = 0;
    if (pps_len) *pps_len = 0;
    return 0;
}

int jm_nvenc_get_spspps(unsigned char *out_buf, handle_nvenc handle)
{
    (void)out_buf; (void)handle;
    return 0;                                                         /* copies sps_len+pps_len = 0 bytes */
}

int jm_nvenc_memory_alloc_host(void **buf, int buf_len, handle_nvenc handle)
{
    nvenc_b200 *c = (nvenc_b200 *)handle;
    if (!c || !c->ctx || !buf || buf_len < 0) return JM_NVENC_ERR_GENERIC;
    return jmc_alloc_host(c->ctx, (size_t)buf_len, 1, buf) == JMC_OK ? 0 : JM_NVENC_ERR_GENERIC;   /* WRITECOMBINED, :1305 */
}

int jm_nvenc_memory_release_host(void *buf, handle_nvenc handle)
{
    nvenc_b200 *c = (nvenc_b200 *)handle;
    if (!c || !c->ctx) return JM_NVENC_ERR_GENERIC;
    return jmc_free_host(c->ctx, buf) == JMC_OK ? 0 : JM_NVENC_ERR_GENERIC;
}

int jm_nvenc_peek_surface(void **dptr, int *pitch, int *rows, handle_nvenc handle)
{
    nvenc_b200 *c = (nvenc_b200 *)handle;
    if (!c || c->last < 0) return -1;
    if (dptr) *dptr = c->surf[c->last].dptr;
    if (pitch) *pitch = (int)c->surf[c->last].pitch;
    if (rows) *rows = c->rows;
    return 0;
}

int jm_nvenc_release_surface(handle_nvenc handle)
{
    nvenc_b200 *c = (nvenc_b200 *)handle;
    if (!c) return -1;
    int idx = -1;
    for (int i = 0; i < JM_NVENC_NUM_SURFACES; i++)
        if (c->surf[i].lock_count && (idx < 0 || c->surf[i].order < c->surf[idx].order)) idx = i;
    if (idx < 0) return -1;
    c->surf[idx].lock_count = 0;                                      /* nv_enc.cpp:225 */
    return 0;
}

} /* extern "C" */
