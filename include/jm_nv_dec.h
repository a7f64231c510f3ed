/*
 * jm_nv_dec.h -- drop-in for the reference's nv_dec/jm_nv_dec.h (the jm_nvdec_* decoder API).
 *
 * Same eleven entry points, argument meaning and return conventions as the reference
 * (nv_dec/jm_nv_dec.h:27-88, bodies nv_dec/nv_dec.cpp:695-870), so the call sequences of
 * test_nv_dec/test_nv_dec.cpp:163-259 and test_player/test_player.cpp:204-258 compile and behave
 * the same.  Differences, all additive:
 *   - exports are extern "C" with default visibility (the reference's are MSVC C++-mangled
 *     _declspec(dllexport), jm_nv_dec.h:14-17, so its binary ABI was never portable);
 *   - the decoded-surface format path behind jm_nvdec_decode_frame / jm_nvdec_output_frame runs
 *     as sm_100a CUDA kernels (NV12 -> tight NV12 / I420 on the device, only the tight frame
 *     crosses PCIe) instead of cuMemcpyDtoH of the padded surface + a CPU loop
 *     (nv_dec.cpp:452, :782-820);
 *   - bitstream codecs (0..7) go through NVDEC exactly as in the reference (parser + decoder
 *     callbacks, nv_dec.cpp:23-52,278-403,496-540), bound at run time from libnvcuvid.so.1 (or
 *     $JMC_NVCUVID_LIB); the mapped device surface feeds the kernel directly;
 *   - codec_type JM_NVDEC_CODEC_RAW_NV12 accepts already-decoded pitched NV12 surfaces as
 *     "packets" (host bytes or a device pointer), which is how the path is exercised where no
 *     NVDEC engine is exposed;
 *   - jm_nvdec_set_device() picks the GPU (the reference hard-codes device 0, nv_dec.cpp:209).
 * There is no CPU fallback: without a CUDA device jm_nvdec_init fails.
 */
#ifndef _JM_NV_DECODER_H_
#define _JM_NV_DECODER_H_

#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

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

typedef void *handle_nvdec;

/* codec_type values of jm_nvdec_init (enum NV_CODEC_TYPE, nv_dec/nv_dec.h:37-46) */
#define JM_NVDEC_CODEC_AVC    0
#define JM_NVDEC_CODEC_HEVC   1
#define JM_NVDEC_CODEC_MJPEG  2
#define JM_NVDEC_CODEC_MPEG4  3
#define JM_NVDEC_CODEC_MPEG2  4
#define JM_NVDEC_CODEC_VP8    5
#define JM_NVDEC_CODEC_VP9    6
#define JM_NVDEC_CODEC_VC1    7
/* extension: every "packet" is one decoded surface (struct jm_nvdec_raw_packet) */
#define JM_NVDEC_CODEC_RAW_NV12 100

#define JM_NVDEC_RAW_MAGIC 0x53524D4Au      /* "JMRS" */
#define JM_NVDEC_RAW_DEVICE_PTR 1u          /* flags: surface already in device memory at device_ptr */
#define JM_NVDEC_RAW_SYNC       2u          /* flags: do not return before the conversion has READ the surface */
#define JM_NVDEC_RAW_WAIT_EVENT 4u          /* flags: packet is a jm_nvdec_raw_packet_ex; order the conversion after ready_event */

/* What cuvidMapVideoFrame hands the reference (a device pointer + pitch, nv_dec.cpp:439-442),
 * as a packet.  Without JM_NVDEC_RAW_DEVICE_PTR the header is followed by pitch*height*3/2
 * bytes of pitched NV12 (Y rows, then UV rows from row `height`, nv_dec.cpp:453,765). */
typedef struct jm_nvdec_raw_packet {
    uint32_t magic;
    int32_t  width, height, pitch;
    uint32_t flags;
    uint32_t reserved;
    uint64_t device_ptr;
} jm_nvdec_raw_packet;

/* Device-pointer packets and ownership.  The header is consumed before jm_nvdec_decode_frame returns; the SURFACE
 * it points to is read by a kernel that is only enqueued by then.  So, like a decoder's own surface ring, it must
 * (a) be completely written when the call is made -- or carry the CUDA event that marks its completion
 *     (JM_NVDEC_RAW_WAIT_EVENT, ready_event = a cudaEvent_t / CUevent recorded on the producing stream), and
 * (b) stay untouched until jm_nvdec_output_frame has returned the frame made from it -- or the packet sets
 *     JM_NVDEC_RAW_SYNC, then jm_nvdec_decode_frame itself waits until the surface has been read. */
typedef struct jm_nvdec_raw_packet_ex {
    jm_nvdec_raw_packet base;
    uint64_t ready_event;
} jm_nvdec_raw_packet_ex;

/** create decode handle (new + zero, no CUDA; nv_dec.cpp:54-60) */
JMDLL_FUNC handle_nvdec jm_nvdec_create_handle(void);

/**
 *   Init decode before use (nv_dec.cpp:62-80).
 *   codec_type: 0 - H.264, 1 - H.265 ... 7 - VC1, or JM_NVDEC_CODEC_RAW_NV12
 *   out_fmt:    0 - NV12, anything else - "YV12", which the reference writes U-plane-first,
 *               i.e. I420 (nv_dec.cpp:814-815)
 *   extra_data: sps/pps buffer, NULL is OK
 *   return: 0 - successful, else failed (-2 no CUDA device, -3 bad device id as
 *           nvdec_cuda_init nv_dec.cpp:219-231; -4 NVDEC library/engine unavailable for a bitstream codec)
 */
JMDLL_FUNC int jm_nvdec_init(int codec_type, int out_fmt, char *extra_data, int len, handle_nvdec handle);

/** destroy decode handle (nv_dec.cpp:82-110); the handle is invalid afterwards */
JMDLL_FUNC int jm_nvdec_deinit(handle_nvdec handle);

/**
 *   decode video frame (nv_dec.cpp:481-494): consumes in_buf before returning; in_data_len == 0
 *   (or in_buf == NULL) flushes / signals end of stream; *got_frame = 1 if a frame is ready for
 *   jm_nvdec_output_frame.  At most one frame per call.  Returns 0 like the reference, with ONE
 *   exception: -1 when decoded frames had to be dropped because the caller stopped fetching (the handle
 *   keeps at most 64 converted frames; the reference's queue is unbounded and overwrites decode surfaces
 *   instead).  Every surface that became available in this call is converted by one launch and its
 *   delivery to the host is started before the call returns; nothing waits for it here.
 *   With a display delay of n (jm_nvdec_set_display_delay) the frame announced is the one decoded n calls
 *   earlier, as with the parser's own ulMaxDisplayDelay (nv_dec.cpp:346); flush with in_data_len == 0 until
 *   jm_nvdec_is_exit(), as test_nv_dec.cpp:232-246 does.
 */
JMDLL_FUNC int jm_nvdec_decode_frame(unsigned char *in_buf, int in_data_len, int *got_frame, handle_nvdec handle);

/**
 *   fetch the frame announced by got_frame (nv_dec.cpp:750-828).
 *   out_len [in] capacity of out_buf, [out] frame size w*h*3/2.
 *   return: w*h*3/2 (>0, NOT 0 -- nv_dec.cpp:827) on success; -1 no frame / NULL buffer;
 *           -2 capacity too small (*out_len left untouched, :773-774).
 *   out_buf may be pageable (as in the reference): the frame is copied out of the handle's pinned delivery
 *   ring, where the DMA started by jm_nvdec_decode_frame has put (or is putting) it; or pinned
 *   (jm_nvdec_memory_alloc_host / jm_nvdec_memory_register_host): it receives the frame by direct DMA; or a
 *   device pointer: device-to-device copy.  May be called again for the same frame.
 */
JMDLL_FUNC int jm_nvdec_output_frame(unsigned char *out_buf, int *out_len, handle_nvdec handle);

/** display size of the stream, valid after the first decoded frame (nv_dec.cpp:839-846) */
JMDLL_FUNC int jm_nvdec_stream_info(int *disp_width, int *disp_height, handle_nvdec handle);

JMDLL_FUNC void jm_nvdec_set_eof(bool is_eof, handle_nvdec handle);
JMDLL_FUNC bool jm_nvdec_is_exit(handle_nvdec handle);
/** info block built when the stream is drained (nv_dec.cpp:663-683); owned by the handle */
JMDLL_FUNC char *jm_nvdec_show_dec_info(handle_nvdec handle);
JMDLL_FUNC bool jm_nvdec_is_hw_support(void);

/* ---- extensions ---------------------------------------------------------------------------- */
/** choose the CUDA device before jm_nvdec_init (default 0, or env JMC_DEVICE) */
JMDLL_FUNC int jm_nvdec_set_device(int device, handle_nvdec handle);
/** pinned host memory for out_buf (mirror of jm_nvenc_memory_alloc_host, jmnv_enc.h:65-66) */
JMDLL_FUNC int jm_nvdec_memory_alloc_host(void **buf, int buf_len, handle_nvdec handle);
JMDLL_FUNC int jm_nvdec_memory_release_host(void *buf, handle_nvdec handle);
/** page-lock a buffer the caller already owns (malloc'ed out_buf / packet buffer) so that it is reached by direct
 *  DMA; unregister it BEFORE freeing it.  Only the whole pages INSIDE the buffer are locked (its partial first / last
 *  page may hold other allocations of the caller); those < 4 KB edges are copied through a pinned bounce buffer.
 *  Buffers with less than 64 KB of whole pages are left pageable (returns -1).  (JMC_NVDEC_LAZY_PIN=1 / option "lazy_pin" does this automatically for an
 *  OUT_BUF it sees for the second time -- opt-in, because the library cannot see the caller free it: a buffer that is
 *  freed while registered and whose address is handed out again would receive its frames in the old pages.) */
JMDLL_FUNC int jm_nvdec_memory_register_host(void *buf, int buf_len, handle_nvdec handle);
JMDLL_FUNC int jm_nvdec_memory_unregister_host(void *buf, handle_nvdec handle);
/** frames held back before they are announced (0..20, default 0 or env JMC_NVDEC_DISPLAY_DELAY): with n >= 1 the
 *  delivery of frame k overlaps the upload / decode / conversion of frame k+1 */
JMDLL_FUNC int jm_nvdec_set_display_delay(int frames, handle_nvdec handle);
/** "display_delay", "lazy_pin" (0/1), "copy_threads" (helper threads for copies from/to PAGEABLE caller memory, 0..16; default
 *  min(4, cores/4) or env JMC_NVDEC_COPY_THREADS; started only when such a copy happens), "map_limit" (decoder surfaces mapped and converted per launch, 1..8) */
JMDLL_FUNC int jm_nvdec_set_option(const char *name, int value, handle_nvdec handle);
/** zero-copy fetch: *frame points at the current frame inside the handle's pinned delivery ring (same bytes
 *  jm_nvdec_output_frame would write), valid until the next jm_nvdec_decode_frame call.  Returns w*h*3/2 or -1. */
JMDLL_FUNC int jm_nvdec_output_frame_ref(const unsigned char **frame, int *frame_len, handle_nvdec handle);
/** decoded frames dropped so far because the caller did not fetch (see jm_nvdec_decode_frame) */
JMDLL_FUNC int jm_nvdec_dropped_frames(handle_nvdec handle);
/** conversion kernels this handle has launched (one per batch of surfaces, not one per frame) */
JMDLL_FUNC long long jm_nvdec_launch_count(handle_nvdec handle);
/** diagnostic: delivery copies (device-to-host copies of converted frames) that the handles of this process have queued on
 *  `device` (0..63) and not yet seen complete -- the count the per-device cap works on; 0 once every handle is gone */
JMDLL_FUNC int jm_nvdec_deliveries_in_flight(int device);

#ifdef __cplusplus
}
#endif
#endif /* _JM_NV_DECODER_H_ */
