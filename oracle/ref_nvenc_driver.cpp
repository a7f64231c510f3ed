/*
 * TEST INFRASTRUCTURE ONLY (checker).  Never linked into the product library.
 *
 * extern "C" driver around the UNMODIFIED reference translation unit /root/reference/nv_enc/nv_enc.cpp.
 * It runs the reference's own
 *     nvenc_convert_yuv_data_to_nv12()      nv_enc/nv_enc.cpp:1023-1103
 * against a FAKE CUDA DRIVER: the reference calls CUDA through global function pointers
 * (nv_sdk/src/dynlink_cuda.cpp), which this file points at host-memory emulations, so "device"
 * pointers are plain host pointers.  Everything the reference computes on the host -- which bytes of
 * the input go where, plane offsets (y_len, y_len*5/4), copy extents, the 8 kernel arguments and the
 * launch grid -- is therefore executed, not restated.  The one thing that cannot be executed is the
 * InterleaveUV kernel itself (its PTX, preproc32_lowlat.ptx, is not in the reference tree): the
 * emulation below implements the NVENC-SDK sample semantics
 *     dst[y*nv12Pitch + 2x] = U[y*cbPitch + x];  dst[y*nv12Pitch + 2x + 1] = V[y*crPitch + x]
 * for every thread (x, y) of the launched grid with x < chromaWidth, y < chromaHeight.
 */
#include <tchar.h>      /* the shim, exactly as nv_enc.cpp:13 pulls it in */
#undef cuMemAllocPitch
#include "nv_enc.h"

/* defined in nv_enc/nv_enc.cpp:1023, not declared in nv_enc.h */
int nvenc_convert_yuv_data_to_nv12(const unsigned char *in_yuv_buf, const int yuv_len, nvenc_surface *nv_frame, nvenc_ctx *ctx);

static CUresult CUDAAPI fake_push(CUcontext) { return CUDA_SUCCESS; }
static CUresult CUDAAPI fake_pop(CUcontext *p) { if (p) *p = NULL; return CUDA_SUCCESS; }
static CUresult CUDAAPI fake_stream_query(CUstream) { return CUDA_SUCCESS; }

static CUresult CUDAAPI fake_memcpy_htod(CUdeviceptr dst, const void *src, size_t n)
{
    memcpy((void *)(uintptr_t)dst, src, n);
    return CUDA_SUCCESS;
}

static CUresult CUDAAPI fake_memcpy2d(const CUDA_MEMCPY2D *c)
{
    if (c->srcMemoryType != CU_MEMORYTYPE_HOST || c->dstMemoryType != CU_MEMORYTYPE_DEVICE) return CUDA_ERROR_INVALID_VALUE;
    const unsigned char *s = (const unsigned char *)c->srcHost + c->srcY * c->srcPitch + c->srcXInBytes;
    unsigned char *d = (unsigned char *)(uintptr_t)c->dstDevice + c->dstY * c->dstPitch + c->dstXInBytes;
    for (size_t r = 0; r < c->Height; r++) memcpy(d + r * c->dstPitch, s + r * c->srcPitch, c->WidthInBytes);
    return CUDA_SUCCESS;
}

#define FAKE_INTERLEAVE_FN ((CUfunction)(uintptr_t)0x1A7E)
static int g_launches = 0;

static CUresult CUDAAPI fake_launch(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                    unsigned smem, CUstream, void **args, void **)
{
    if (f != FAKE_INTERLEAVE_FN || gz != 1 || bz != 1 || smem != 0) return CUDA_ERROR_INVALID_VALUE;
    /* argument order at nv_enc.cpp:1070 */
    const unsigned char *u = (const unsigned char *)(uintptr_t)*(CUdeviceptr *)args[0];
    const unsigned char *v = (const unsigned char *)(uintptr_t)*(CUdeviceptr *)args[1];
    unsigned char *dst = (unsigned char *)(uintptr_t)*(CUdeviceptr *)args[2];
    const int cw = *(int *)args[3], ch = *(int *)args[4], cb_pitch = *(int *)args[5], cr_pitch = *(int *)args[6];
    const unsigned nv12_pitch = *(uint32_t *)args[7];
    for (unsigned y = 0; y < gy * by; y++)
        for (unsigned x = 0; x < gx * bx; x++)
            if ((int)x < cw && (int)y < ch) {
                dst[(size_t)y * nv12_pitch + 2 * x] = u[(size_t)y * cb_pitch + x];
                dst[(size_t)y * nv12_pitch + 2 * x + 1] = v[(size_t)y * cr_pitch + x];
            }
    g_launches++;
    return CUDA_SUCCESS;
}

int jmref_allocpitch_compat(void *dptr, void *pitch_u32, size_t width_bytes, size_t height, unsigned)
{
    /* only reachable from nvenc_register_frame (needs NVENC); keep the link complete */
    size_t pitch = (width_bytes + 511) & ~(size_t)511;
    *(CUdeviceptr *)dptr = (CUdeviceptr)(uintptr_t)calloc(pitch, height ? height : 1);
    *(uint32_t *)pitch_u32 = (uint32_t)pitch;
    return 0;
}

extern "C" {

/* fmt: raw NV_ENC_BUFFER_FORMAT value.  surf: the pitched "device" surface (host memory here).
 * Returns the reference's return value; *launches = InterleaveUV launches it issued. */
__attribute__((visibility("default")))
int jmref_nvenc_convert(const unsigned char *in_yuv, int yuv_len, int fmt, int width, int height,
                        unsigned char *surf, int stride, int *launches)
{
    cuCtxPushCurrent = fake_push;
    cuCtxPopCurrent = fake_pop;
    cuStreamQuery = fake_stream_query;
    cuMemcpyHtoD = fake_memcpy_htod;
    cuMemcpy2D = fake_memcpy2d;
    cuLaunchKernel = fake_launch;

    nvenc_ctx *ctx = (nvenc_ctx *)calloc(1, sizeof(nvenc_ctx));
    ctx->format = (NV_ENC_BUFFER_FORMAT)fmt;
    ctx->width = width;
    ctx->height = height;
    ctx->cuInterleaveUVFunction = FAKE_INTERLEAVE_FN;
    /* the reference allocates width*height/4 bytes for each chroma temp (nv_enc.cpp:972-973) */
    const size_t tmp = (size_t)width * height / 4 + 1;
    unsigned char *t0 = (unsigned char *)malloc(tmp), *t1 = (unsigned char *)malloc(tmp);
    ctx->uv_tmp_ptr[0] = (CUdeviceptr)(uintptr_t)t0;
    ctx->uv_tmp_ptr[1] = (CUdeviceptr)(uintptr_t)t1;
    nvenc_surface s;
    memset(&s, 0, sizeof(s));
    s.in_cuda_surf = (CUdeviceptr)(uintptr_t)surf;
    s.in_cuda_stride = (uint32_t)stride;
    g_launches = 0;
    int r = nvenc_convert_yuv_data_to_nv12(in_yuv, yuv_len, &s, ctx);
    if (launches) *launches = g_launches;
    free(t0); free(t1); free(ctx);
    return r;
}

} /* extern "C" */
