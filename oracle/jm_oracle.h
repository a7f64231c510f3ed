/*
 * jm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the surface-format loops of mojing1999/jmcodec, used as the
 * bit-exact checker for the CUDA path and, when oracle/_ref is unavailable, as the "port"
 * CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product library (libjmcodec_b200.so)
 * never links or calls it and has no CPU fallback.
 *
 * Parity status: NV12->NV12/I420, NV12/I420->pitched NV12 are PINNED against the unmodified
 * reference translation units compiled into oracle/_ref/libjmref.so (tests/test_oracle.py)
 * and against SHA-256 known-answer vectors generated from them (tests/golden/).
 * jmo_nvenc_upload restates nv_enc/nv_enc.cpp:1023-1103: PINNED against the reference's own
 * nvenc_convert_yuv_data_to_nv12() executed over a fake CUDA driver (oracle/ref_nvenc_driver.cpp);
 * only the InterleaveUV kernel body (PTX absent from the reference tree) is emulated there, from the
 * 8 launch arguments at nv_enc.cpp:1070.  jmo_nv12_to_rgb24, jmo_nv12_to_argb32 and jmo_rgb24_to_nv12 are
 * builder-defined BT.601 specs: PARITY UNPINNED (the reference has no YUV<->RGB code; SDL2 does it).
 *
 * All citations are relative to /root/reference.
 */
#ifndef JM_ORACLE_H
#define JM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* nv_dec/nv_dec.cpp:750-828  jm_nvdec_output_frame().
 * surf: pitched NV12 (Y at surf, UV at surf + pitch*height, nv_dec.cpp:765).
 * have_frame==0 models cur_out_frame==NULL.  Returns the reference's return value:
 * -1 (no frame / NULL surface), -2 (*out_len too small, *out_len untouched), else w*h*3/2
 * with *out_len set to the same.  out_fmt 0 = tight NV12, anything else = I420 (U first). */
int jmo_nvdec_output_frame(const uint8_t *surf, int pitch, int width, int height,
                           int out_fmt, int have_frame, uint8_t *out_buf, int *out_len);

/* intel_dec/intel_dec.cpp:244-332  intel_dec_output_yuv_frame() for an NV12 surface.
 * Returns -1 (*out_len=0) with no surface, -2 (*out_len=0) if the buffer is short, else 0
 * with *out_len = W*H + W*H/2.  Mirrors the crop quirks: UV x-offset is crop_x/2 BYTES
 * (:292,:303) and the V plane starts at (W*H/2)/2 past U (:306), not (W/2)*(H/2). */
int jmo_inteldec_output_frame(const uint8_t *surf_y, const uint8_t *surf_uv, int pitch,
                              int crop_x, int crop_y, int crop_w, int crop_h,
                              int out_fmt, int have_surface, uint8_t *out_buf, int *out_len);

/* intel_enc/intel_enc.cpp:251-314 (is_i420=0, tight NV12 in) and :316-387 (is_i420=1, tight
 * I420 in) -> pitched NV12 surface.  Geometry follows the reference's uint16_t arithmetic.
 * Returns -1 when no surface is free, else 0.  `len` is ignored, as in the reference. */
int jmo_intelenc_input(const uint8_t *yuv, int len, int is_i420,
                       uint8_t *surf_y, uint8_t *surf_uv, int pitch,
                       int info_w, int info_h, int crop_x, int crop_y, int crop_w, int crop_h,
                       int surface_free);

/* nv_enc/nv_enc.cpp:1023-1103  nvenc_convert_yuv_data_to_nv12(), device side restated on the CPU.
 * fmt uses the raw NV_ENC_BUFFER_FORMAT values carried in nv_enc_param.in_fmt
 * (nv_sdk/inc/nvEncodeAPI.h:306-315): 0x1 NV12, 0x10 YV12 (read as I420: first chroma plane ->
 * even bytes, V plane at y_len*5/4, :1055-1056), 0x01000000 ARGB / 0x10000000 ABGR (flat w*h*4
 * copy that ignores the pitch, :1096).  surf: pitched surface, UV plane at surf + stride*height
 * (:1069).  InterleaveUV semantics per the 8 launch arguments at :1070.  Returns 0, or -1 for an
 * unknown format (the reference silently does nothing). */
int jmo_nvenc_upload(const uint8_t *in_buf, int fmt, int width, int height,
                     uint8_t *surf, int stride);

/* Builder-defined (PARITY UNPINNED): BT.601 limited range, nearest chroma, integer
 *   C=Y-16 D=U-128 E=V-128
 *   R=clip8((298C+409E+128)>>8) G=clip8((298C-100D-208E+128)>>8) B=clip8((298C+516D+128)>>8)
 * pixel (x,y) uses chroma sample (min(x>>1,w2-1), min(y>>1,h2-1)), w2=w>>1, h2=h>>1.
 * Output packed R,G,B with row pitch rgb_pitch (>= 3*w).  Returns -1 if w<2 or h<2, else 0. */
int jmo_nv12_to_rgb24(const uint8_t *surf, int pitch, int width, int height,
                      uint8_t *rgb, int rgb_pitch);

/* Builder-defined (PARITY UNPINNED): the same integer BT.601 as jmo_nv12_to_rgb24, written as packed
 * ARGB8888 words, i.e. bytes B,G,R,0xFF per pixel -- the layout the reference's disabled
 * NV12ToARGB_drvapi hook targets (nv_dec/nv_dec.cpp:244-265).  argb_pitch >= 4*w. */
int jmo_nv12_to_argb32(const uint8_t *surf, int pitch, int width, int height,
                       uint8_t *argb, int argb_pitch);

/* Builder-defined (PARITY UNPINNED; SURVEY 8f rank 4 "RGB -> NV12 would complete a transcode loop"): the
 * forward BT.601 limited-range integer transform, the counterpart of jmo_nv12_to_rgb24:
 *   Y = ((66R + 129G + 25B + 128) >> 8) + 16                      per pixel
 *   U = floor((-38Rs - 74Gs + 112Bs + 512) / 1024) + 128          Rs,Gs,Bs = sums over the 2x2 block
 *   V = floor((112Rs - 94Gs - 18Bs + 512) / 1024) + 128
 * into an nv_enc-style pitched NV12 surface (Y at 0, UV at pitch*height, nv_enc.cpp:1069); chroma has
 * (w>>1) x (h>>1) samples, an odd last column / row contributes luma only; padding is not written.
 * Returns -1 if w<1 or h<1, else 0. */
int jmo_rgb24_to_nv12(const uint8_t *rgb, int rgb_pitch, int width, int height,
                      uint8_t *surf, int pitch);

/* "port" CPU baseline: frames calls of jmo_nvdec_output_frame, frame f reading surface
 * f % n_surf and writing slot f % n_out, round-robin over nthreads.  Returns seconds or -1. */
double jmo_nvdec_run(const uint8_t *surf_base, size_t surf_stride, int n_surf,
                     uint8_t *out_base, size_t out_stride, int n_out,
                     int pitch, int width, int height, int out_fmt, int frames, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
