/*
 * jmnv_enc.h -- drop-in for the reference's nv_enc/jmnv_enc.h (the jm_nvenc_* encoder API).
 *
 * Same nine entry points and the same nv_enc_param (nv_enc/jmnv_enc.h:23-67, bodies
 * nv_enc/nv_enc.cpp:1329-1380).  What this library implements behind them is the encoder INPUT
 * path: the pitched NV12 surface pool (nvenc_register_frame, nv_enc.cpp:954-1007) and the upload /
 * I420->NV12 pack (nvenc_convert_yuv_data_to_nv12, nv_enc.cpp:1023-1103) as one H2D copy plus one
 * sm_100a kernel, replacing 1 cuMemcpy2D + 2 cuMemcpyHtoD + the byte-granular InterleaveUV PTX.
 *
 * B200 has no NVENC engine.  jm_nvenc_init therefore returns JM_NVENC_ERR_NO_ENCODE_DEVICE unless
 * surface-only mode is requested (codec_id = JM_NVENC_CODEC_SURFACE_ONLY, or env
 * JMC_NVENC_SURFACE_ONLY=1): then frames are uploaded and packed exactly as the reference would
 * hand them to nvEncMapInputResource, got_packet stays 0 and jm_nvenc_get_bitstream reports "no
 * packet" (-1).  jm_nvenc_peek_surface exposes the packed device surface for verification.
 */
#ifndef _JMNV_ENC_H_
#define _JMNV_ENC_H_

#include <stdint.h>

#ifndef JMDLL_FUNC
#if defined(__GNUC__)
#define JMDLL_FUNC __attribute__((visibility("default")))
#else
#define JMDLL_FUNC
#endif
#define JMDLL_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void *handle_nvenc;

/* codec_id (nv_enc/nv_enc.h:61-65) */
#define JM_NVENC_H264 0
#define JM_NVENC_HEVC 1
#define JM_NVENC_CODEC_SURFACE_ONLY (-1)     /* extension: input path only, no bitstream */

/* in_fmt carries raw NV_ENC_BUFFER_FORMAT values (nv_sdk/inc/nvEncodeAPI.h:306-315) */
#define JM_NVENC_FMT_NV12 0x00000001
#define JM_NVENC_FMT_YV12 0x00000010         /* the reference reads it U-plane-first, i.e. as I420 (nv_enc.cpp:1055-1056) */
#define JM_NVENC_FMT_ARGB 0x01000000
#define JM_NVENC_FMT_ABGR 0x10000000

/* return codes that are NVENCSTATUS values in the reference (nvEncodeAPI.h:370-508) */
#define JM_NVENC_SUCCESS                0
#define JM_NVENC_ERR_NO_ENCODE_DEVICE   1
#define JM_NVENC_ERR_INVALID_PARAM      8
#define JM_NVENC_ERR_OUT_OF_MEMORY     10
#define JM_NVENC_ERR_GENERIC           20    /* what __cu() maps any CUDA failure to (nv_enc.h:57) */

#define JM_NVENC_NUM_SURFACES 10             /* MAX_NV_ENC_FRAME_NUM, nv_enc/nv_enc.h:99 */

/* Field order and types are the ABI (nv_enc/jmnv_enc.h:23-53): thirteen ints. */
typedef struct _nv_enc_param {
    int codec_id;           /* JM_NVENC_H264 / JM_NVENC_HEVC / JM_NVENC_CODEC_SURFACE_ONLY */
    int in_fmt;             /* an NV_ENC_BUFFER_FORMAT value: JM_NVENC_FMT_* below */
    int preset;             /* index into the NVENC preset GUIDs; unused without an encoder */
    int src_width;          /* size of the frames handed to jm_nvenc_enc_frame */
    int src_height;
    int dst_width;          /* encoded size; unused without an encoder */
    int dst_height;
    int fps;
    int bitrate_kb;
    int gop_len;
    int num_bframe;
    int is_external_alloc;  /* 1: CUDA surface path (the one implemented here) */
    int qp;
} nv_enc_param;

JMDLL_FUNC handle_nvenc jm_nvenc_create_handle(void);
JMDLL_FUNC int jm_nvenc_init(nv_enc_param *in_param, handle_nvenc handle);
JMDLL_FUNC int jm_nvenc_deinit(handle_nvenc handle);
/* NULL / 0 length = end of stream (nv_enc.cpp:87,113-117).  -1: no free surface (nv_enc.cpp:90-93).
 * In surface-only mode there is no encoder to hand a surface back after its bitstream was fetched
 * (nv_enc.cpp:204-225), so every uploaded surface stays locked until jm_nvenc_release_surface() is called:
 * a caller following the reference's call sequence gets -1 on the 11th frame unless it releases.
 * JM_NVENC_ERR_INVALID_PARAM: yuv_len is shorter than an NV12 / ARGB frame of the configured size (the
 * reference over-reads the buffer instead, nv_enc.cpp:1029-1040,1096; the YV12 path clamps like the reference). */
JMDLL_FUNC int jm_nvenc_enc_frame(const unsigned char *in_yuv_buf, const int yuv_len, int *got_packet, handle_nvenc handle);
/* -1: no packet ready (nv_enc.cpp:175-178) -- always, in surface-only mode */
JMDLL_FUNC int jm_nvenc_get_bitstream(unsigned char *out_buf, int *out_data_len, int *is_keyframe, handle_nvenc handle);

JMDLL_FUNC int jm_nvenc_get_spspps_len(int *sps_len, int *pps_len, handle_nvenc handle);
JMDLL_FUNC int jm_nvenc_get_spspps(unsigned char *out_buf, handle_nvenc handle);

/* pinned, write-combined host memory for in_yuv_buf (cuMemHostAlloc(WRITECOMBINED), nv_enc.cpp:1301-1310) */
JMDLL_FUNC int jm_nvenc_memory_alloc_host(void **buf, int buf_len, handle_nvenc handle);
JMDLL_FUNC int jm_nvenc_memory_release_host(void *buf, handle_nvenc handle);

/* ---- extensions ---------------------------------------------------------------------------- */
JMDLL_FUNC int jm_nvenc_set_device(int device, handle_nvenc handle);
/* Device pointer / pitch / allocated rows of the surface filled by the most recent
 * jm_nvenc_enc_frame, after its upload has completed.  -1 if no frame was uploaded yet. */
JMDLL_FUNC int jm_nvenc_peek_surface(void **dptr, int *pitch, int *rows, handle_nvenc handle);
/* Hand the oldest in-flight surface back to the pool (what nvEncUnmapInputResource does after the
 * bitstream is fetched, nv_enc.cpp:204-222).  Returns -1 if none is held. */
JMDLL_FUNC int jm_nvenc_release_surface(handle_nvenc handle);

#ifdef __cplusplus
}
#endif
#endif  /* _JMNV_ENC_H_ */
