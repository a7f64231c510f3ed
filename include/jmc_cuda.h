/*
 * jmc_cuda.h -- thin C-ABI layer under the jm_nvdec_* / jm_nvenc_* drop-in API.
 *
 * This is the boundary a maintainer of mojing1999/jmcodec binds to replace the reference's
 * decoded-surface format path (CUDA driver calls + CPU loops) with sm_100a kernels:
 * plain pointers and sizes, no C++/torch types, one shared library (libjmcodec_b200.so).
 * Every entry point names the reference call site it replaces (paths relative to the reference
 * repository root).  There is NO CPU fallback anywhere behind this header: without a CUDA
 * device every call fails with JMC_ERR_NO_DEVICE / JMC_ERR_CUDA.
 *
 * Vocabulary: a "surface" is a pitched NV12 image as NVDEC/NVENC/MFX hold it (Y rows of `pitch`
 * bytes, then interleaved UV rows); a "tight" frame is the caller-side packed buffer
 * (w*h*3/2 bytes: NV12 or planar I420).  A "batch" is n_frames surfaces converted by ONE launch.
 */
#ifndef JMC_CUDA_H
#define JMC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define JMC_API __attribute__((visibility("default")))
#else
#define JMC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (negative = failure) ------------------------------------------------- */
#define JMC_OK              0
#define JMC_ERR_INVALID    (-1)   /* bad argument / geometry                                 */
#define JMC_ERR_CUDA       (-2)   /* a CUDA runtime call failed; see jmc_last_error()         */
#define JMC_ERR_NO_DEVICE  (-3)   /* no CUDA device / bad device id (nv_dec.cpp:219-231)      */
#define JMC_ERR_NOMEM      (-4)
#define JMC_ERR_BUSY       (-5)   /* pipeline slot still in flight                            */

typedef struct jmc_ctx jmc_ctx;           /* one per (host thread, device): streams + scratch  */
typedef struct jmc_pipeline jmc_pipeline; /* host-delivery ring: H2D -> convert -> pinned D2H  */

/* Human-readable description of the last failure on the calling thread. */
JMC_API const char *jmc_last_error(void);
JMC_API const char *jmc_version(void);
/* The JMC_* environment switches (kernel-variant A/B: JMC_NO_BULK, JMC_NO_ROWS, JMC_RGB_FLAT ...) are read once per
 * process, not per launch; call this after changing them in a running process. */
JMC_API void jmc_reload_env(void);

/* ---- device + context ------------------------------------------------------------------
 * Replaces nvdec_cuda_init()/nvenc_cuda_init(): cuInit + cuDeviceGetCount + cuCtxCreate on the
 * hard-coded device 0 (nv_dec/nv_dec.cpp:202-273, nv_enc/nv_enc.cpp:232-276).  The device is
 * now selectable so that streams can be sharded over 1..8 GPUs. */
JMC_API int jmc_device_count(void);                              /* <=0: no usable device      */
/* The calling thread's current CUDA device as this library's (statically linked) runtime sees it.  Every jmc_* /
 * jm_nvdec_* / jm_nvenc_* entry point switches to its own device and RESTORES this one before returning, as the
 * reference pushes / pops its context around each call (nv_dec.cpp:378,398,423,471). */
JMC_API int jmc_current_device(void);
JMC_API int jmc_set_current_device(int device);
JMC_API int jmc_ctx_create(int device, jmc_ctx **out);
JMC_API int jmc_ctx_destroy(jmc_ctx *ctx);                       /* cuCtxDestroy, nv_dec.cpp:104 */
JMC_API int jmc_ctx_device(const jmc_ctx *ctx);
JMC_API int jmc_ctx_sm_count(const jmc_ctx *ctx);
/* which: 0 = convert stream, 1 = upload (H2D) stream, 2 = delivery (D2H) stream.
 * Returns the cudaStream_t as void* (for callers that time with their own events). */
JMC_API void *jmc_ctx_stream(jmc_ctx *ctx, int which);
JMC_API int jmc_ctx_sync(jmc_ctx *ctx);                          /* all three streams          */

/* ---- memory ---------------------------------------------------------------------------- */
/* cuMemAlloc (nv_enc.cpp:972-973) */
JMC_API int jmc_alloc_device(jmc_ctx *ctx, size_t bytes, void **dptr);
/* cuMemAllocPitch(w, rows, 16) (nv_enc.cpp:967,970,979).  min_pitch: 0 = driver's choice. */
JMC_API int jmc_alloc_pitched(jmc_ctx *ctx, size_t width_bytes, size_t rows, void **dptr, size_t *pitch);
JMC_API int jmc_free_device(jmc_ctx *ctx, void *dptr);
/* cuMemAllocHost (nv_dec.cpp:569) / cuMemHostAlloc(WRITECOMBINED) (nv_enc.cpp:1305) */
JMC_API int jmc_alloc_host(jmc_ctx *ctx, size_t bytes, int write_combined, void **hptr);
JMC_API int jmc_free_host(jmc_ctx *ctx, void *hptr);             /* cuMemFreeHost, nv_dec.cpp:607 */
/* Synchronous helpers on the convert stream (tests, setup).  Hot paths use the pipeline. */
JMC_API int jmc_memcpy_h2d(jmc_ctx *ctx, void *dptr, const void *hptr, size_t bytes);
JMC_API int jmc_memcpy_d2h(jmc_ctx *ctx, void *hptr, const void *dptr, size_t bytes);
JMC_API int jmc_memset_device(jmc_ctx *ctx, void *dptr, int byte, size_t bytes);

/* ---- one batched conversion launch ------------------------------------------------------ */
typedef enum jmc_op {
    /* decode side: pitched NV12 surface -> tight frame */
    JMC_OP_NV12_TO_NV12 = 0,   /* strip the pitch          nv_dec.cpp:782-797, intel_dec.cpp:284-299 */
    JMC_OP_NV12_TO_I420 = 1,   /* + U/V de-interleave      nv_dec.cpp:798-820, intel_dec.cpp:301-314 */
    /* encode side: tight frame -> pitched NV12 surface */
    JMC_OP_NV12_TO_SURF = 2,   /* add the pitch            nv_enc.cpp:1029-1040, intel_enc.cpp:291-307 */
    JMC_OP_I420_TO_SURF = 3,   /* + U/V interleave         nv_enc.cpp:1041-1081 (InterleaveUV), intel_enc.cpp:366-380 */
    /* display side (test_player.cpp:283-288 hands I420 to SDL2; builder-defined integer BT.601) */
    JMC_OP_NV12_TO_RGB24 = 4,
    JMC_OP_NV12_TO_I420_RGB24 = 5, /* fused: one read of the surface, both outputs */
    /* what the reference's disabled NV12ToARGB_drvapi hook would have produced on the device
     * (nv_dec.cpp:244-265): packed 32-bit ARGB8888, i.e. bytes B,G,R,0xFF per pixel; same integer BT.601 */
    JMC_OP_NV12_TO_ARGB32 = 6,
    /* encode side from RGB ("RGB -> NV12 would complete a transcode loop"): packed R,G,B rows
     * in job->rgb -> pitched NV12 surface in job->surf; builder-defined forward integer BT.601, chroma from the
     * 2x2 block sums (see oracle/jm_oracle.h jmo_rgb24_to_nv12) */
    JMC_OP_RGB24_TO_SURF = 7
} jmc_op;

/* Where frame f of a batch lives: base + f*stride, or list[f] (a DEVICE array of n_frames device
 * pointers, e.g. surfaces mapped from a decoder) when list != NULL. */
typedef struct jmc_frames {
    void        *base;
    size_t       stride;
    void *const *list;
} jmc_frames;

/* Memory contract of a job.
 *  - Surfaces are allocated as whole rows: every row of `pitch` bytes exists up to its end, as cudaMallocPitch /
 *    cuMemAllocPitch (nv_enc.cpp:961-975) and decoder-mapped surfaces guarantee.  When a surface is 16-byte
 *    aligned (base, pitch, plane offsets) the kernels may READ a source row up to the next multiple of 16 past
 *    `width`, inside the pitch; padding is never written.
 *  - Tight / RGB frames: exactly the bytes the reference writes are written.  Reads of a tight or RGB source may
 *    start up to 15 bytes before a row (aligned 16-byte loads): that is the previous row or, for the first row,
 *    still inside the device allocation (allocations are at least 256-byte aligned); nothing is read past the
 *    last byte of a frame. */
typedef struct jmc_job {
    int32_t    op;            /* jmc_op                                                         */
    int32_t    n_frames;      /* frames converted by this launch                                */
    int32_t    width, height; /* active picture size in pixels                                  */
    /* pitched NV12 side: source of the decode/display ops, destination of the encode ops */
    jmc_frames surf;
    int32_t    pitch;         /* bytes per surface row                                          */
    int64_t    surf_y_off;    /* first active Y byte, from the frame pointer (crop folded in)   */
    int64_t    surf_uv_off;   /* first active UV byte (nv_dec: pitch*height, nv_dec.cpp:765)    */
    /* tight side */
    jmc_frames tight;
    int64_t    tight_u_off;   /* I420 ops: U plane offset in the tight frame (w*h)              */
    int64_t    tight_v_off;   /* I420 ops: V plane offset (nv_dec: w*h+(w>>1)*(h>>1), :815)     */
    /* packed RGB24 / ARGB32 destination of the display ops, RGB24 source of RGB24_TO_SURF */
    jmc_frames rgb;
    int32_t    rgb_pitch;     /* bytes per RGB row, >= 3*width (>= 4*width for ARGB32)          */
    uint32_t   flags;         /* JMC_JOB_*                                                      */
} jmc_job;

/* flags: the caller guarantees every pointer in surf.list / tight.list / rgb.list is 16-byte
 * aligned (true for cudaMalloc'ed and decoder-mapped surfaces), so the 16-byte-vector kernel can be
 * chosen without reading the lists.  Without it, pointer lists take the any-alignment kernel. */
#define JMC_JOB_ALIGNED16 1u
/* every non-NULL surf.list / tight.list / rgb.list is a HOST array of at most JMC_INLINE_LIST_MAX device pointers:
 * they are passed to the kernel as arguments, nothing is uploaded first (how jm_nvdec_* feeds the few decoder
 * surfaces it has mapped into one launch).  Alignment is then checked on the host; ALIGNED16 is not needed. */
#define JMC_JOB_LIST_ON_HOST 2u
/* encode ops (tight / RGB frame -> pitched surface): the caller does not care about the surface's pitch padding -- true for
 * encoder input surfaces, whose padding nobody reads -- so the kernel may ZERO the padding of every row up to the next
 * 16-byte boundary instead of preserving it.  Default (flag clear): every padding byte keeps its value, as after the
 * reference's cuMemcpy2D / row memcpy (nv_enc.cpp:1029-1040, intel_enc.cpp:291-307). */
#define JMC_JOB_PAD_ZERO 4u
#define JMC_INLINE_LIST_MAX 8

/* Geometry fillers: set width/height/pitch and every offset exactly as the named reference
 * function computes them (odd sizes included).  They leave op-independent fields (frames) alone.
 *   nvdec    : nv_dec.cpp:765,807-815      out_fmt 0 -> NV12_TO_NV12, else NV12_TO_I420
 *   inteldec : intel_dec.cpp:261-307       crop offsets incl. the crop_x/2-bytes UV quirk; V plane at (W*H/2)/2
 *   intelenc : intel_enc.cpp:271-300,366-370   surf_rows = allocated surface height (UV plane row)
 *   nvenc    : nv_enc.cpp:1029-1069        in_fmt = NV_ENC_BUFFER_FORMAT value (0x1 NV12, 0x10 YV12-read-as-I420) */
JMC_API int jmc_job_nvdec(jmc_job *job, int width, int height, int pitch, int out_fmt);
JMC_API int jmc_job_inteldec(jmc_job *job, int pitch, int surf_rows, int crop_x, int crop_y, int crop_w, int crop_h, int out_fmt);
JMC_API int jmc_job_intelenc(jmc_job *job, int pitch, int surf_rows, int crop_x, int crop_y, int crop_w, int crop_h, int is_i420);
JMC_API int jmc_job_nvenc(jmc_job *job, int width, int height, int stride, int in_fmt);
/* NV12_TO_RGB24 / NV12_TO_I420_RGB24 on an nv_dec-style surface (fused != 0 adds the I420 output). */
JMC_API int jmc_job_rgb(jmc_job *job, int width, int height, int pitch, int rgb_pitch, int fused);
/* NV12_TO_ARGB32 on an nv_dec-style surface; the ARGB frames go to job->rgb, row pitch argb_pitch >= 4*width. */
JMC_API int jmc_job_argb(jmc_job *job, int width, int height, int pitch, int argb_pitch);
/* RGB24_TO_SURF into an nv_enc-style surface (Y at 0, UV at stride*height, nv_enc.cpp:1069); the RGB frames are
 * read from job->rgb, row pitch rgb_pitch >= 3*width (a pipeline batch takes rgb_pitch*height bytes per frame). */
JMC_API int jmc_job_rgb_to_nv12(jmc_job *job, int width, int height, int rgb_pitch, int stride);
/* Bytes of one tight frame as the reference computes it: w*h*3/2 (nv_dec.cpp:773,824). */
JMC_API int64_t jmc_tight_bytes(int width, int height);
/* Algorithmic bytes (read + write, padding excluded) one frame of `job` moves: the roofline numerator. */
JMC_API int64_t jmc_job_algorithmic_bytes(const jmc_job *job);

/* Enqueue ONE kernel launch converting the whole batch.  stream: a cudaStream_t, or NULL for the
 * context's convert stream.  Asynchronous; all pointers are device pointers.  Replaces, per frame:
 * cuMemcpyDtoH + the CPU loop (decode side) or cuMemcpy2D/cuMemcpyHtoD/cuLaunchKernel (encode side). */
JMC_API int jmc_convert(jmc_ctx *ctx, const jmc_job *job, void *stream);
/* Number of conversion kernels this context has launched (bench.py's gpu_launches). */
JMC_API uint64_t jmc_ctx_launch_count(const jmc_ctx *ctx);

/* Device-timed repeat of one launch: CUDA events on the launching stream around `iters`
 * back-to-back launches of `job`; returns average milliseconds per launch in *ms_per_launch. */
JMC_API int jmc_convert_timed(jmc_ctx *ctx, const jmc_job *job, int iters, float *ms_per_launch);

/* ---- host link probe ------------------------------------------------------------------------
 * What the PCIe link between this context's device and pinned host memory delivers right now: `copies` copies of
 * bytes_per_copy bytes, device-timed (CUDA events on the upload / delivery stream).  mode 1: host -> device only,
 * 2: device -> host only, 3: both directions at once (what the host-delivery pipeline does); +4: the upload reads a
 * write-combined host buffer (cuMemHostAlloc(WRITECOMBINED), what nv_enc.cpp:1305 hands its callers).  Run it on every GPU
 * of a box between two barriers to measure the whole-box host ceiling the e2e leg is bounded by (bench.py). */
typedef struct jmc_link_rates { double h2d_gbs, d2h_gbs; } jmc_link_rates;
JMC_API int jmc_link_probe(jmc_ctx *ctx, size_t bytes_per_copy, int copies, int mode, jmc_link_rates *out);

/* ---- device-side timing -------------------------------------------------------------------
 * CUDA events recorded on one of the context's own streams (which: as jmc_ctx_stream), so that a
 * caller can bracket any sequence of conversions / pipeline submissions on the device timeline. */
typedef struct jmc_event jmc_event;
JMC_API int jmc_event_create(jmc_ctx *ctx, jmc_event **out);
JMC_API int jmc_event_destroy(jmc_ctx *ctx, jmc_event *ev);
JMC_API int jmc_event_record(jmc_ctx *ctx, jmc_event *ev, int which);
/* Waits for `stop`, then returns the time between the two events in milliseconds. */
JMC_API int jmc_event_elapsed_ms(jmc_ctx *ctx, jmc_event *start, jmc_event *stop, float *ms);

/* ---- host-delivery pipeline --------------------------------------------------------------
 * The reference moves every decoded surface to the host with one synchronous cuMemcpyDtoH of
 * pitch*h*3/2 bytes and converts on the CPU (nv_dec.cpp:452, :750-828); nothing overlaps.
 * The pipeline converts on the device and delivers the TIGHT result with cudaMemcpyAsync into
 * pinned memory on a side stream, overlapped with the next batch (depth >= 2 slots).
 *
 * `shape` fixes op/geometry/batch size; its surf/tight/rgb frame sets are ignored (the pipeline
 * owns device staging).  For decode/display ops the host input of a batch is n_frames contiguous
 * surfaces of surf_bytes each and the output n_frames tight (and/or RGB) frames; for encode ops
 * the direction is reversed.  Host buffers should be pinned (jmc_alloc_host) for full overlap. */
JMC_API int jmc_pipeline_create(jmc_ctx *ctx, const jmc_job *shape, size_t surf_bytes, int depth, jmc_pipeline **out);
JMC_API int jmc_pipeline_destroy(jmc_pipeline *p);
/* Enqueue one batch: upload host_in (NULL: input already in device memory at dev_in), convert,
 * deliver to host_out (and host_out2 for the fused op's RGB plane).  n_frames <= shape->n_frames.
 * Returns the slot index (>=0) or an error.  Blocks only if the slot is still in flight. */
JMC_API int jmc_pipeline_submit(jmc_pipeline *p, const void *host_in, const void *dev_in,
                                void *host_out, void *host_out2, int n_frames);
JMC_API int jmc_pipeline_wait(jmc_pipeline *p, int slot);        /* result of that slot is on the host */
JMC_API int jmc_pipeline_drain(jmc_pipeline *p);
JMC_API uint64_t jmc_pipeline_h2d_bytes(const jmc_pipeline *p);  /* totals since creation             */
JMC_API uint64_t jmc_pipeline_d2h_bytes(const jmc_pipeline *p);

#ifdef __cplusplus
}
#endif
#endif /* JMC_CUDA_H */
