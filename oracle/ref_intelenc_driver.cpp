/*
 * TEST INFRASTRUCTURE ONLY (checker / CPU baseline).  Never linked into the product library.
 *
 * extern "C" driver around the UNMODIFIED reference translation unit
 * /root/reference/intel_enc/intel_enc.cpp.  Calls the reference's own
 *     intel_enc_input_yuv_frame()    intel_enc/intel_enc.cpp:251-314  (tight NV12 -> pitched NV12)
 *     intel_enc_input_yuv_yuv420()   intel_enc/intel_enc.cpp:316-387  (tight I420 -> pitched NV12)
 * on a hand-built system-memory mfxFrameSurface1.  The second one is the reference's only CPU
 * implementation of the I420->NV12 pack, i.e. the oracle for the encoder-input kernel.
 */
#include "intel_enc.h"

#include <time.h>

static intel_enc_ctx *make_ctx(mfxFrameSurface1 *s, mfxFrameSurface1 **arr,
                               unsigned char *surf_y, unsigned char *surf_uv, int pitch,
                               int info_w, int info_h, int crop_x, int crop_y, int crop_w, int crop_h)
{
    intel_enc_ctx *ctx = (intel_enc_ctx *)calloc(1, sizeof(intel_enc_ctx));
    memset(s, 0, sizeof(*s));
    s->Info.FourCC = MFX_FOURCC_NV12;
    s->Info.Width = (mfxU16)info_w;
    s->Info.Height = (mfxU16)info_h;
    s->Info.CropX = (mfxU16)crop_x;
    s->Info.CropY = (mfxU16)crop_y;
    s->Info.CropW = (mfxU16)crop_w;
    s->Info.CropH = (mfxU16)crop_h;
    s->Data.Y = surf_y;
    s->Data.UV = surf_uv;
    s->Data.Pitch = (mfxU16)pitch;
    arr[0] = s;
    ctx->surfaces = arr;
    ctx->num_surfaces = 1;
    ctx->in_surf_queue = new std::queue<mfxFrameSurface1 *>;
    return ctx;
}

static void drop_ctx(intel_enc_ctx *ctx)
{
    delete ctx->in_surf_queue;
    free(ctx);
}

extern "C" {

/* is_i420 = 0: intel_enc_input_yuv_frame (NV12 in); 1: intel_enc_input_yuv_yuv420 (I420 in).
 * surface_free==0 marks the only surface locked -> reference returns -1 (no free surface). */
__attribute__((visibility("default")))
int jmref_intelenc_input(const unsigned char *yuv, int len, int is_i420,
                         unsigned char *surf_y, unsigned char *surf_uv, int pitch,
                         int info_w, int info_h, int crop_x, int crop_y, int crop_w, int crop_h,
                         int surface_free)
{
    mfxFrameSurface1 s, *arr[1];
    intel_enc_ctx *ctx = make_ctx(&s, arr, surf_y, surf_uv, pitch, info_w, info_h, crop_x, crop_y, crop_w, crop_h);
    if (!surface_free) s.Data.Locked = 1;
    int r = is_i420 ? intel_enc_input_yuv_yuv420((uint8_t *)yuv, len, ctx)
                    : intel_enc_input_yuv_frame((uint8_t *)yuv, len, ctx);
    drop_ctx(ctx);
    return r;
}

/* Timed single-thread loop for the pack CPU baseline: frame f reads tight input f % n_in and
 * writes surface f % n_surf (UV plane at surf + pitch*height).  Returns seconds or -1. */
__attribute__((visibility("default")))
double jmref_intelenc_run(const unsigned char *in_base, size_t in_stride, int n_in,
                          unsigned char *surf_base, size_t surf_stride, int n_surf,
                          int pitch, int width, int height, int is_i420, int frames)
{
    mfxFrameSurface1 s, *arr[1];
    intel_enc_ctx *ctx = make_ctx(&s, arr, NULL, NULL, pitch, width, height, 0, 0, width, height);
    struct timespec t0, t1;
    int bad = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int f = 0; f < frames; f++) {
        unsigned char *surf = surf_base + (size_t)(f % n_surf) * surf_stride;
        s.Data.Y = surf;
        s.Data.UV = surf + (size_t)pitch * height;
        s.reserved[INDEX_OF_RESERVED_IN_USE] = 0;
        const unsigned char *in = in_base + (size_t)(f % n_in) * in_stride;
        int r = is_i420 ? intel_enc_input_yuv_yuv420((uint8_t *)in, width * height * 3 / 2, ctx)
                        : intel_enc_input_yuv_frame((uint8_t *)in, width * height * 3 / 2, ctx);
        if (r != 0) bad++;
        if (!ctx->in_surf_queue->empty()) ctx->in_surf_queue->pop();
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    drop_ctx(ctx);
    if (bad) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

} /* extern "C" */
