/*
 * TEST INFRASTRUCTURE ONLY (checker).  Never linked into the product library.
 *
 * extern "C" driver around the UNMODIFIED reference translation unit
 * /root/reference/intel_dec/intel_dec.cpp.  Calls the reference's own
 *     intel_dec_output_yuv_frame()   intel_dec/intel_dec.cpp:244-332
 * on a hand-built system-memory mfxFrameSurface1 (crop rectangle included), i.e. the
 * mirror of the nv_dec loop with CropX/CropY source offsets.  No Intel engine is
 * touched: the MFX* entry points are link-time stubs (oracle/ref_mfx_stubs.cpp).
 */
#include "intel_dec.h"
#include "jm_intel_dec.h"

extern "C" {

/* surf_y / surf_uv: the two plane pointers of the surface (mfxFrameData.Y / .UV).
 * have_surface==0 -> empty queue: reference sets *out_len=0 and returns -1 (intel_dec.cpp:251-255). */
__attribute__((visibility("default")))
int jmref_inteldec_output_frame(unsigned char *surf_y, unsigned char *surf_uv, int pitch,
                                int crop_x, int crop_y, int crop_w, int crop_h,
                                int out_fmt, int have_surface, unsigned char *out_buf, int *out_len)
{
    intel_ctx *ctx = (intel_ctx *)calloc(1, sizeof(intel_ctx));
    ctx->out_fmt = out_fmt;
    ctx->out_surf_queue = new std::queue<mfxFrameSurface1 *>;
    mfxFrameSurface1 s;
    memset(&s, 0, sizeof(s));
    s.Info.FourCC = MFX_FOURCC_NV12;
    s.Info.CropX = (mfxU16)crop_x;
    s.Info.CropY = (mfxU16)crop_y;
    s.Info.CropW = (mfxU16)crop_w;
    s.Info.CropH = (mfxU16)crop_h;
    s.Data.Y = surf_y;
    s.Data.UV = surf_uv;
    s.Data.Pitch = (mfxU16)pitch;
    if (have_surface) ctx->out_surf_queue->push(&s);
    int r = intel_dec_output_yuv_frame(out_buf, out_len, ctx);
    delete ctx->out_surf_queue;
    free(ctx);
    return r;
}

} /* extern "C" */
