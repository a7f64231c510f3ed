/*
 * jm_oracle.c -- TEST INFRASTRUCTURE ONLY (see jm_oracle.h for the rules and parity status).
 *
 * Plain-C restatement of the reference's CPU surface-format loops.  Each function names the
 * reference lines it follows; the arithmetic (integer truncation, plane offsets, which bytes
 * are left untouched, return codes) is the reference's, the code shape is not.
 */
#define _POSIX_C_SOURCE 200809L
#include "jm_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* rows x nbytes from a pitched source into a pitched destination */
static void copy_rows(uint8_t *dst, ptrdiff_t dst_pitch, const uint8_t *src, ptrdiff_t src_pitch,
                      int nbytes, int rows)
{
    for (int r = 0; r < rows; r++)
        memcpy(dst + r * dst_pitch, src + r * src_pitch, (size_t)nbytes);
}

/* interleaved UV rows -> two planar rows of `cw` samples each */
static void split_uv_rows(uint8_t *u, uint8_t *v, int cw, const uint8_t *uv, ptrdiff_t uv_pitch, int rows)
{
    for (int r = 0; r < rows; r++) {
        const uint8_t *s = uv + r * uv_pitch;
        for (int x = 0; x < cw; x++) {
            u[r * cw + x] = s[2 * x];
            v[r * cw + x] = s[2 * x + 1];
        }
    }
}

/* two planar chroma streams -> interleaved UV rows of `cw` pairs each */
static void merge_uv_rows(uint8_t *uv, ptrdiff_t uv_pitch, const uint8_t *u, const uint8_t *v, int cw, int rows)
{
    size_t k = 0;
    for (int r = 0; r < rows; r++) {
        uint8_t *d = uv + r * uv_pitch;
        for (int x = 0; x < cw; x++, k++) {
            d[2 * x] = u[k];
            d[2 * x + 1] = v[k];
        }
    }
}

/* nv_dec/nv_dec.cpp:750-828 */
int jmo_nvdec_output_frame(const uint8_t *surf, int pitch, int width, int height,
                           int out_fmt, int have_frame, uint8_t *out_buf, int *out_len)
{
    if (!have_frame) return -1;                       /* :757-758 */
    if (!surf) return -1;                             /* :768-771 */
    const int need = width * height * 3 / 2;
    if (*out_len < need) return -2;                   /* :773-774, *out_len not yet cleared */
    *out_len = 0;                                     /* :776 */

    const uint8_t *uv = surf + pitch * height;        /* :765 */
    const int luma = width * height;
    copy_rows(out_buf, width, surf, pitch, width, height);            /* :787-790 / :801-804 */
    if (out_fmt == 0) {
        copy_rows(out_buf + luma, width, uv, pitch, width, height >> 1);   /* :792-796 */
    } else {
        const int cw = width >> 1, ch = height >> 1;                  /* :807-808 */
        split_uv_rows(out_buf + luma, out_buf + luma + cw * ch, cw, uv, pitch, ch);   /* :812-818 */
    }
    *out_len = need;                                  /* :824 */
    return need;                                      /* :827 -- the byte count, not 0 */
}

/* intel_dec/intel_dec.cpp:244-332 */
int jmo_inteldec_output_frame(const uint8_t *surf_y, const uint8_t *surf_uv, int pitch,
                              int crop_x, int crop_y, int crop_w, int crop_h,
                              int out_fmt, int have_surface, uint8_t *out_buf, int *out_len)
{
    if (!have_surface) { *out_len = 0; return -1; }   /* :251-255 */
    const int y_len = crop_w * crop_h;                /* :261 */
    const int uv_len = y_len / 2;                     /* :262 */
    if (*out_len < y_len + uv_len) { *out_len = 0; return -2; }   /* :264-268 */

    copy_rows(out_buf, crop_w, surf_y + crop_y * pitch + crop_x, pitch, crop_w, crop_h);   /* :284-287 */
    /* chroma origin: rows crop_y/2, and crop_x/2 BYTES (upstream quirk, :292-293 / :303-304) */
    const uint8_t *uv = surf_uv + (crop_y / 2) * pitch + (crop_x / 2);
    if (out_fmt == 0) {
        copy_rows(out_buf + y_len, crop_w, uv, pitch, crop_w, crop_h / 2);                 /* :294-299 */
    } else {
        uint8_t *pu = out_buf + y_len;
        uint8_t *pv = pu + uv_len / 2;                /* :306-307: (W*H/2)/2, not (W/2)*(H/2) */
        split_uv_rows(pu, pv, crop_w / 2, uv, pitch, crop_h / 2);                          /* :308-314 */
    }
    *out_len = y_len + uv_len;                        /* :317 */
    return 0;
}

/* intel_enc/intel_enc.cpp:251-314 and :316-387 */
int jmo_intelenc_input(const uint8_t *yuv, int len, int is_i420,
                       uint8_t *surf_y, uint8_t *surf_uv, int pitch_in,
                       int info_w, int info_h, int crop_x, int crop_y, int crop_w, int crop_h,
                       int surface_free)
{
    (void)len;                                        /* never read by the reference */
    if (!surface_free) return -1;                     /* :254-259 / :319-324 */
    uint16_t w, h;                                    /* the reference computes in uint16_t (:265 / :330) */
    const uint16_t pitch = (uint16_t)pitch_in;
    if (crop_w > 0 && crop_h > 0) { w = (uint16_t)crop_w; h = (uint16_t)crop_h; }   /* :271-278 */
    else                          { w = (uint16_t)info_w; h = (uint16_t)info_h; }
    const int y_len = w * h;                          /* :280 */

    copy_rows(surf_y + crop_y * pitch + crop_x, pitch, yuv, w, w, h);               /* :291-295 */
    crop_x /= 2; crop_y /= 2; h /= 2;                 /* :298-300 */
    uint8_t *uv = surf_uv + crop_y * pitch + crop_x;
    if (!is_i420) {
        copy_rows(uv, pitch, yuv + y_len, w, w, h);   /* :303-307 */
    } else {
        w /= 2;                                       /* :366 */
        const uint8_t *pu = yuv + y_len;              /* :368-369 */
        const uint8_t *pv = pu + w * h;               /* :370 */
        merge_uv_rows(uv, pitch, pu, pv, w, h);       /* :375-380 */
    }
    return 0;
}

/* nv_enc/nv_enc.cpp:1023-1103 */
int jmo_nvenc_upload(const uint8_t *in_buf, int fmt, int width, int height,
                     uint8_t *surf, int stride)
{
    if (fmt == 0x1) {                                 /* NV_ENC_BUFFER_FORMAT_NV12, :1029-1040 */
        copy_rows(surf, stride, in_buf, width, width, height * 3 / 2);
        return 0;
    }
    if (fmt == 0x10) {                                /* NV_ENC_BUFFER_FORMAT_YV12, :1041-1081 */
        copy_rows(surf, stride, in_buf, width, width, height);        /* luma 2-D copy, :1043-1051 */
        const int y_len = width * height;             /* :1054 */
        const uint8_t *pu = in_buf + y_len;           /* :1055, y_len/4 bytes staged */
        const uint8_t *pv = in_buf + y_len * 5 / 4;   /* :1056 */
        const int ch = height / 2, cw = width / 2;    /* :1062-1063 */
        /* InterleaveUV(U, V, dst, cw, ch, cbPitch=cw, crPitch=cw, nv12Pitch=stride), :1069-1075 */
        merge_uv_rows(surf + (size_t)stride * height, stride, pu, pv, cw, ch);
        return 0;
    }
    if (fmt == 0x01000000 || fmt == 0x10000000) {     /* ARGB / ABGR, :1083-1097: flat, pitch ignored */
        memcpy(surf, in_buf, (size_t)width * height * 4);
        return 0;
    }
    return -1;
}

static inline uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* builder-defined spec, see jm_oracle.h (PARITY UNPINNED) */
int jmo_nv12_to_rgb24(const uint8_t *surf, int pitch, int width, int height,
                      uint8_t *rgb, int rgb_pitch)
{
    const int cw = width >> 1, ch = height >> 1;
    if (cw < 1 || ch < 1) return -1;
    const uint8_t *uvp = surf + (size_t)pitch * height;
    for (int y = 0; y < height; y++) {
        int cy = y >> 1; if (cy > ch - 1) cy = ch - 1;
        const uint8_t *yr = surf + (size_t)y * pitch;
        const uint8_t *cr = uvp + (size_t)cy * pitch;
        uint8_t *o = rgb + (size_t)y * rgb_pitch;
        for (int x = 0; x < width; x++) {
            int cx = x >> 1; if (cx > cw - 1) cx = cw - 1;
            const int c = yr[x] - 16, d = cr[2 * cx] - 128, e = cr[2 * cx + 1] - 128;
            o[3 * x + 0] = clip8((298 * c + 409 * e + 128) >> 8);
            o[3 * x + 1] = clip8((298 * c - 100 * d - 208 * e + 128) >> 8);
            o[3 * x + 2] = clip8((298 * c + 516 * d + 128) >> 8);
        }
    }
    return 0;
}

/* builder-defined spec, see jm_oracle.h (PARITY UNPINNED) */
int jmo_nv12_to_argb32(const uint8_t *surf, int pitch, int width, int height,
                       uint8_t *argb, int argb_pitch)
{
    const int cw = width >> 1, ch = height >> 1;
    if (cw < 1 || ch < 1) return -1;
    const uint8_t *uvp = surf + (size_t)pitch * height;
    for (int y = 0; y < height; y++) {
        int cy = y >> 1; if (cy > ch - 1) cy = ch - 1;
        const uint8_t *yr = surf + (size_t)y * pitch;
        const uint8_t *cr = uvp + (size_t)cy * pitch;
        uint8_t *o = argb + (size_t)y * argb_pitch;
        for (int x = 0; x < width; x++) {
            int cx = x >> 1; if (cx > cw - 1) cx = cw - 1;
            const int c = yr[x] - 16, d = cr[2 * cx] - 128, e = cr[2 * cx + 1] - 128;
            o[4 * x + 0] = clip8((298 * c + 516 * d + 128) >> 8);               /* B */
            o[4 * x + 1] = clip8((298 * c - 100 * d - 208 * e + 128) >> 8);     /* G */
            o[4 * x + 2] = clip8((298 * c + 409 * e + 128) >> 8);               /* R */
            o[4 * x + 3] = 0xFF;                                                 /* A */
        }
    }
    return 0;
}

/* builder-defined spec, see jm_oracle.h (PARITY UNPINNED) */
int jmo_rgb24_to_nv12(const uint8_t *rgb, int rgb_pitch, int width, int height,
                      uint8_t *surf, int pitch)
{
    const int cw = width >> 1, ch = height >> 1;
    if (width < 1 || height < 1) return -1;
    for (int y = 0; y < height; y++) {
        const uint8_t *p = rgb + (size_t)y * rgb_pitch;
        uint8_t *yr = surf + (size_t)y * pitch;
        for (int x = 0; x < width; x++)
            yr[x] = (uint8_t)((66 * p[3 * x] + 129 * p[3 * x + 1] + 25 * p[3 * x + 2] + 128 + 16 * 256) >> 8);
    }
    uint8_t *uvp = surf + (size_t)pitch * height;
    for (int cy = 0; cy < ch; cy++) {
        const uint8_t *p0 = rgb + (size_t)(2 * cy) * rgb_pitch, *p1 = p0 + rgb_pitch;
        uint8_t *o = uvp + (size_t)cy * pitch;
        for (int cx = 0; cx < cw; cx++) {
            const int a = 6 * cx;                                   /* pixels 2cx, 2cx+1 of both rows */
            const int r = p0[a] + p0[a + 3] + p1[a] + p1[a + 3];
            const int g = p0[a + 1] + p0[a + 4] + p1[a + 1] + p1[a + 4];
            const int b = p0[a + 2] + p0[a + 5] + p1[a + 2] + p1[a + 5];
            /* + 128*1024 keeps the sum positive, so >> is a floor whatever the compiler does with negatives */
            o[2 * cx] = (uint8_t)((-38 * r - 74 * g + 112 * b + 512 + 128 * 1024) >> 10);
            o[2 * cx + 1] = (uint8_t)((112 * r - 94 * g - 18 * b + 512 + 128 * 1024) >> 10);
        }
    }
    return 0;
}

/* ---- "port" CPU baseline loop ---- */
typedef struct {
    const uint8_t *surf_base; size_t surf_stride; int n_surf;
    uint8_t *out_base; size_t out_stride; int n_out;
    int pitch, width, height, out_fmt, frames, tid, nthreads;
    long bad;
} run_job;

static void *run_worker(void *p)
{
    run_job *j = (run_job *)p;
    const int need = j->width * j->height * 3 / 2;
    for (int f = j->tid; f < j->frames; f += j->nthreads) {
        int len = (int)j->out_stride;
        int r = jmo_nvdec_output_frame(j->surf_base + (size_t)(f % j->n_surf) * j->surf_stride,
                                       j->pitch, j->width, j->height, j->out_fmt, 1,
                                       j->out_base + (size_t)(f % j->n_out) * j->out_stride, &len);
        if (r != need) j->bad++;
    }
    return NULL;
}

double jmo_nvdec_run(const uint8_t *surf_base, size_t surf_stride, int n_surf,
                     uint8_t *out_base, size_t out_stride, int n_out,
                     int pitch, int width, int height, int out_fmt, int frames, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    run_job *jobs = (run_job *)calloc((size_t)nthreads, sizeof(run_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        run_job j = { surf_base, surf_stride, n_surf, out_base, out_stride, n_out,
                      pitch, width, height, out_fmt, frames, t, nthreads, 0 };
        jobs[t] = j;
        if (nthreads == 1) run_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, run_worker, &jobs[t]);
    }
    long bad = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        bad += jobs[t].bad;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(jobs);
    free(th);
    if (bad) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
