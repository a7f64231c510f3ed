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
 * Front-ends:
 *   JM_NVDEC_CODEC_RAW_NV12  decoded surfaces as packets (host bytes or device pointers);
 *   bitstream codecs         NVDEC through libnvcuvid.so.1, bound at run time (dlopen): the same
 *                            parser/decoder callback structure as nv_dec.cpp:23-52,278-403,496-540, with
 *                            the mapped device surface fed straight into the conversion kernel
 *                            (no intermediate copy).  If the library or an NVDEC engine is not
 *                            available jm_nvdec_init fails (-4); nothing falls back to a CPU decoder.
 */
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <deque>
#include <new>

#include "cuvid_min.h"
#include "jm_nv_dec.h"
#include "jmc_internal.h"

#define NVDEC_MAX_FRAMES 10         /* display queue depth / decode surfaces, nv_dec/nv_dec.h:32 */
#define MAX_LEN_DEC_INFO 1024       /* nv_dec/nv_dec.h:33 */

namespace {

struct decoded_surface {            /* what cuvidMapVideoFrame yields: device pointer + pitch */
    uint8_t *dptr;
    int pitch, width, height;
    int pool_slot;                  /* >= 0: one of our upload surfaces; -1: caller-owned device memory;
                                       -2: a CUVID picture still to be mapped (disp valid) */
    CUVIDPARSERDISPINFO disp;       /* copied, not pointed to (the reference queues the parser's pointer, nv_dec.cpp:156) */
};

struct cuvid_api {
    void *lib;
    tcuvidCreateVideoParser create_parser;
    tcuvidParseVideoData parse;
    tcuvidDestroyVideoParser destroy_parser;
    tcuvidCreateDecoder create_decoder;
    tcuvidDestroyDecoder destroy_decoder;
    tcuvidDecodePicture decode_picture;
    tcuvidMapVideoFrame64 map_frame;
    tcuvidUnmapVideoFrame64 unmap_frame;
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

    /* NVDEC front-end (nv_dec.h:80-86) */
    cuvid_api nv;
    CUvideoparser parser;
    CUvideodecoder decoder;
    CUVIDEOFORMATEX parse_ext;
    int cuvid_codec;
    bool decoder_failed;
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

/* ---- NVDEC front-end ------------------------------------------------------------------------ */
int cuvid_codec_of(int t)           /* nv_dec.cpp:295-333 */
{
    switch (t) {
    case JM_NVDEC_CODEC_AVC: return CUVID_CODEC_H264;
    case JM_NVDEC_CODEC_HEVC: return CUVID_CODEC_HEVC;
    case JM_NVDEC_CODEC_MJPEG: return CUVID_CODEC_JPEG;
    case JM_NVDEC_CODEC_MPEG4: return CUVID_CODEC_MPEG4;
    case JM_NVDEC_CODEC_MPEG2: return CUVID_CODEC_MPEG2;
    case JM_NVDEC_CODEC_VP8: return CUVID_CODEC_VP8;
    case JM_NVDEC_CODEC_VP9: return CUVID_CODEC_VP9;
    case JM_NVDEC_CODEC_VC1: return CUVID_CODEC_VC1;
    default: return CUVID_CODEC_H264;                                    /* :330-332 */
    }
}

bool cuvid_load(nvdec_b200 *c)
{
    const char *name = getenv("JMC_NVCUVID_LIB");
    c->nv.lib = dlopen(name && *name ? name : "libnvcuvid.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!c->nv.lib) { jmc_set_error("cannot load the NVDEC library: %s", dlerror()); return false; }
    void *l = c->nv.lib;
    c->nv.create_parser = (tcuvidCreateVideoParser)dlsym(l, "cuvidCreateVideoParser");
    c->nv.parse = (tcuvidParseVideoData)dlsym(l, "cuvidParseVideoData");
    c->nv.destroy_parser = (tcuvidDestroyVideoParser)dlsym(l, "cuvidDestroyVideoParser");
    c->nv.create_decoder = (tcuvidCreateDecoder)dlsym(l, "cuvidCreateDecoder");
    c->nv.destroy_decoder = (tcuvidDestroyDecoder)dlsym(l, "cuvidDestroyDecoder");
    c->nv.decode_picture = (tcuvidDecodePicture)dlsym(l, "cuvidDecodePicture");
    c->nv.map_frame = (tcuvidMapVideoFrame64)dlsym(l, "cuvidMapVideoFrame64");
    c->nv.unmap_frame = (tcuvidUnmapVideoFrame64)dlsym(l, "cuvidUnmapVideoFrame64");
    if (!c->nv.create_parser || !c->nv.parse || !c->nv.destroy_parser || !c->nv.create_decoder || !c->nv.destroy_decoder ||
        !c->nv.decode_picture || !c->nv.map_frame || !c->nv.unmap_frame) {
        jmc_set_error("the NVDEC library lacks a required cuvid* entry point");
        return false;
    }
    /* Is there an engine behind it?  (cuvidGetDecoderCaps exists since SDK 8; CUVIDDECODECAPS starts with
     * codec, chroma, bit depth, 3 reserved words, then bIsSupported.) */
    typedef int (*tcaps)(void *);
    tcaps caps = (tcaps)dlsym(l, "cuvidGetDecoderCaps");
    if (caps) {
        struct { int codec, chroma; unsigned int depth_minus8, r1[3]; unsigned char supported, n_engines; unsigned char rest[80]; } q;
        memset(&q, 0, sizeof(q));
        q.codec = c->cuvid_codec; q.chroma = CUVID_CHROMA_420;
        int r = caps(&q);
        if (r != 0 || !q.supported) {
            jmc_set_error("NVDEC is not usable here: cuvidGetDecoderCaps returned %d, supported=%d", r, (int)q.supported);
            return false;
        }
    }
    return true;
}

/* nvdec_create_decoder, nv_dec.cpp:496-540, called from the parser on the caller's thread */
int cuvid_on_sequence(void *user, CUVIDEOFORMAT *f)
{
    nvdec_b200 *c = (nvdec_b200 *)user;
    if (c->decoder) { c->nv.destroy_decoder(c->decoder); c->decoder = nullptr; }
    unsigned surfaces = NVDEC_MAX_FRAMES;                                 /* :526 */
    if (f->min_num_decode_surfaces > surfaces) surfaces = f->min_num_decode_surfaces;
    CUVIDDECODECREATEINFO ci;
    memset(&ci, 0, sizeof(ci));
    ci.CodecType = f->codec;
    ci.ChromaFormat = f->chroma_format;
    ci.OutputFormat = CUVID_SURFACE_NV12;                                 /* :507 */
    ci.DeinterlaceMode = f->progressive_sequence ? CUVID_DEINTERLACE_WEAVE : CUVID_DEINTERLACE_ADAPTIVE;   /* :508 */
    ci.bitDepthMinus8 = f->bit_depth_luma_minus8;
    ci.ulWidth = f->coded_width;                                          /* :510-511 */
    ci.ulHeight = f->coded_height;
    ci.ulTargetWidth = (unsigned long)(f->display_area.right - f->display_area.left);    /* :513-514 */
    ci.ulTargetHeight = (unsigned long)(f->display_area.bottom - f->display_area.top);
    ci.display_area.left = (short)f->display_area.left;
    ci.display_area.top = (short)f->display_area.top;
    ci.display_area.right = (short)f->display_area.right;
    ci.display_area.bottom = (short)f->display_area.bottom;
    ci.ulNumDecodeSurfaces = surfaces;
    ci.ulNumOutputSurfaces = 2;                                           /* the reference maps one at a time (:527) */
    ci.ulCreationFlags = CUVID_CREATE_PREFER_CUVID;                       /* :528 */
    ci.vidLock = nullptr;                                                 /* as the reference: never created (nv_dec.h:96) */
    int r = c->nv.create_decoder(&c->decoder, &ci);
    if (r != 0 || f->bit_depth_luma_minus8 != 0 || f->chroma_format != CUVID_CHROMA_420) {
        jmc_set_error("cuvidCreateDecoder failed (%d) or unsupported format (bit depth %d, chroma %d)", r,
                      8 + f->bit_depth_luma_minus8, f->chroma_format);
        if (c->decoder) { c->nv.destroy_decoder(c->decoder); c->decoder = nullptr; }
        c->decoder_failed = true;
        return 0;                                                         /* stop the parser */
    }
    c->decoder_failed = false;
    c->disp_w = (int)ci.ulTargetWidth;
    c->disp_h = (int)ci.ulTargetHeight;
    if (!c->started) { clock_gettime(CLOCK_MONOTONIC, &c->t_start); c->started = true; }   /* :537 */
    return (int)surfaces;                                                 /* > 1: tells newer parsers the surface count */
}

int cuvid_on_decode(void *user, void *pic)                                /* nv_dec.cpp:33-41 */
{
    nvdec_b200 *c = (nvdec_b200 *)user;
    if (!c->decoder) return 0;
    return c->nv.decode_picture(c->decoder, pic) == 0 ? 1 : 0;
}

int cuvid_on_display(void *user, CUVIDPARSERDISPINFO *d)                  /* nv_dec.cpp:44-52,151-161 */
{
    nvdec_b200 *c = (nvdec_b200 *)user;
    if (!d) return 1;                                                     /* newer parsers signal EOS with NULL */
    c->num_frames += 1;
    decoded_surface s;
    memset(&s, 0, sizeof(s));
    s.pool_slot = -2;
    s.disp = *d;
    c->queue->push_back(s);
    return 1;
}

int cuvid_open(nvdec_b200 *c, const char *extra, int len)                 /* nvdec_create_parser, nv_dec.cpp:278-366 */
{
    c->cuvid_codec = cuvid_codec_of(c->codec_type);
    if (!cuvid_load(c)) return -4;
    CUVIDPARSERPARAMS pp;
    memset(&pp, 0, sizeof(pp));
    memset(&c->parse_ext, 0, sizeof(c->parse_ext));
    pp.CodecType = c->cuvid_codec;
    pp.pExtVideoInfo = &c->parse_ext;
    c->parse_ext.format.chroma_format = CUVID_CHROMA_420;                 /* :335-336 */
    c->parse_ext.format.progressive_sequence = 1;
    if (extra && len > 0) {                                               /* :339-342 */
        const int n = len < (int)sizeof(c->parse_ext.raw_seqhdr_data) ? len : (int)sizeof(c->parse_ext.raw_seqhdr_data);
        c->parse_ext.format.seqhdr_data_length = (unsigned)n;
        memcpy(c->parse_ext.raw_seqhdr_data, extra, (size_t)n);
    }
    pp.ulMaxNumDecodeSurfaces = NVDEC_MAX_FRAMES;                         /* :345 */
    pp.ulMaxDisplayDelay = 2;                                             /* :346 */
    pp.pUserData = c;
    pp.pfnSequenceCallback = cuvid_on_sequence;
    pp.pfnDecodePicture = cuvid_on_decode;
    pp.pfnDisplayPicture = cuvid_on_display;
    int r = c->nv.create_parser(&c->parser, &pp);
    if (r != 0 || !c->parser) { jmc_set_error("cuvidCreateVideoParser failed (%d)", r); return -4; }
    if (c->parse_ext.format.seqhdr_data_length > 0) {                     /* :357-362: prime the parser with the sequence header */
        CUVIDSOURCEDATAPACKET pkt;
        memset(&pkt, 0, sizeof(pkt));
        pkt.payload = c->parse_ext.raw_seqhdr_data;
        pkt.payload_size = c->parse_ext.format.seqhdr_data_length;
        c->nv.parse(c->parser, &pkt);
    }
    return 0;
}

void cuvid_packet(nvdec_b200 *c, const unsigned char *buf, int len)       /* nvdec_decode_packet, nv_dec.cpp:368-403 */
{
    CUVIDSOURCEDATAPACKET pkt;
    memset(&pkt, 0, sizeof(pkt));
    if (buf && len > 0) {
        pkt.payload = buf;
        pkt.payload_size = (unsigned long)len;
        pkt.flags = CUVID_PKT_TIMESTAMP;                                  /* :386-387 */
    } else {
        pkt.flags = CUVID_PKT_ENDOFSTREAM;                                /* :390 */
    }
    c->nv.parse(c->parser, &pkt);
}

void cuvid_close(nvdec_b200 *c)
{
    if (c->parser) { c->nv.destroy_parser(c->parser); c->parser = nullptr; }     /* nv_dec.cpp:95-101 */
    if (c->decoder) { c->nv.destroy_decoder(c->decoder); c->decoder = nullptr; }
    if (c->nv.lib) { dlclose(c->nv.lib); c->nv.lib = nullptr; }
}

/* RAW front-end: one packet = one decoded surface -> display queue */
int raw_packet(nvdec_b200 *c, const unsigned char *buf, int len)
{
    if (len < (int)sizeof(jm_nvdec_raw_packet)) return -1;
    jm_nvdec_raw_packet h;
    memcpy(&h, buf, sizeof(h));
    if (h.magic != JM_NVDEC_RAW_MAGIC || h.width < 0 || h.height < 0 || h.pitch < h.width) return -1;
    if ((int64_t)h.pitch * h.height * 3 / 2 > 0x7fffffffll) return -1;     /* the API counts frame bytes in int (nv_dec.cpp:773) */
    if ((int)c->queue->size() >= NVDEC_MAX_FRAMES) return -1;            /* all decode surfaces in use */
    decoded_surface s;
    memset(&s, 0, sizeof(s));
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
    unsigned long long mapped = 0;
    if (s.pool_slot == -2) {
        /* cuvidMapVideoFrame (nv_dec.cpp:427-442): post-processed NV12 surface, produced on OUR convert
         * stream so the kernel below is ordered after it without a host sync */
        if (!c->decoder) return;                                          /* :414-417 */
        CUVIDPROCPARAMS pp;
        memset(&pp, 0, sizeof(pp));
        pp.progressive_frame = s.disp.progressive_frame;
        pp.top_field_first = s.disp.top_field_first;
        pp.unpaired_field = s.disp.repeat_first_field < 0;
        pp.output_stream = jmc_ctx_stream(c->ctx, 0);
        unsigned int pitch = 0;
        if (c->nv.map_frame(c->decoder, s.disp.picture_index, &mapped, &pitch, &pp) != 0 || !mapped) return;
        s.dptr = (uint8_t *)(uintptr_t)mapped;
        s.pitch = (int)pitch;
        s.width = c->disp_w;                                              /* ulTargetWidth/Height, :440-442 */
        s.height = c->disp_h;
    }
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
    if (mapped) {
        /* the surface goes back to the decoder only after the kernel has consumed it (:469) */
        cudaStreamSynchronize((cudaStream_t)jmc_ctx_stream(c->ctx, 0));
        c->nv.unmap_frame(c->decoder, mapped);
    }
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
        /* bitstream codecs: NVDEC parser + decoder (nvdec_create_parser, nv_dec.cpp:278-366) */
        int r2 = cuvid_open(c, extra_data, len);
        if (r2) { cuvid_close(c); return r2; }
    }
    return 0;
}

int jm_nvdec_deinit(handle_nvdec handle)
{
    nvdec_b200 *c = (nvdec_b200 *)handle;
    if (!c) return -1;
    if (c->ctx) {
        jmc_ctx_sync(c->ctx);
        cuvid_close(c);
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
    if (jmc_bind_thread(c->ctx)) return 0;                     /* the reference pushes its context around every call (nv_dec.cpp:378,423) */
    if (!c->is_eof) {                                          /* nv_dec.cpp:486-488 */
        const bool cuvid = c->codec_type != JM_NVDEC_CODEC_RAW_NV12;
        if (in_buf && in_data_len > 0) {
            if (!cuvid) raw_packet(c, in_buf, in_data_len);
            else if (c->parser) cuvid_packet(c, in_buf, in_data_len);      /* no parser: init failed, packet dropped */
        } else {
            if (cuvid && c->parser) cuvid_packet(c, nullptr, 0);           /* flush: the parser hands out its delayed pictures */
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
    if (jmc_bind_thread(c->ctx)) return -1;
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
