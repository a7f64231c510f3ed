/*
 * fake_nvcuvid.cu -- TEST INFRASTRUCTURE ONLY: a stand-in for libnvcuvid.so.1.
 *
 * The B200 boxes of this pool ship libnvcuvid but expose no NVDEC engine to the container
 * (cuvidGetDecoderCaps -> CUDA_ERROR_NO_DEVICE, profiles/r1_nvdec_probe.txt), so the NVDEC front-end
 * of jm_nv_dec.cu cannot meet real hardware here.  This library implements the same C ABI
 * (parser + decoder + map/unmap, synchronous callbacks on the caller's thread, display delay,
 * end-of-stream flush, post-processing enqueued on CUVIDPROCPARAMS.output_stream) for a trivial
 * "codec", so that the front-end's glue is exercised end to end on the GPU:
 *
 *   Annex-B byte stream; NAL = start code (00 00 01 | 00 00 00 01) + type byte + RBSP with H.264-style
 *   emulation prevention (00 00 03).  type 0x67: sequence header {u32 width, height} (little endian);
 *   type 0x65: one picture = tight NV12 frame (width*height*3/2 bytes).
 *
 * Select it with JMC_NVCUVID_LIB=<this .so>.  FAKE_NVCUVID_NO_ENGINE=1 makes cuvidGetDecoderCaps fail
 * like the real boxes do.
 *
 * Hazards a client can get wrong are made to HURT, so that the parity tests catch them:
 *   - at most ulNumOutputSurfaces frames can be mapped at a time (a further cuvidMapVideoFrame fails);
 *   - cuvidUnmapVideoFrame POISONS the output surface at once, on a private stream that is not ordered with the
 *     client's: a client that unmaps before its kernel has read the surface converts garbage;
 *   - cuvidDecodePicture overwrites the decode surface immediately (it does not wait for a post-processing copy
 *     the client has only enqueued): a client that lets the parser reuse a picture index whose frame it has not
 *     consumed yet gets the wrong picture;
 *   - fake_nvcuvid_stats() reports how many frames were mapped at once, for tests of the batch drain.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <deque>
#include <vector>

#include "cuvid_min.h"

#define FAKE_API extern "C" __attribute__((visibility("default")))

namespace {

struct pic_params {                 /* leading fields of CUVIDPICPARAMS (cuviddec.h) */
    int PicWidthInMbs, FrameHeightInMbs, CurrPicIdx, field_pic_flag, bottom_field_flag, second_field;
    unsigned int nBitstreamDataLen;
    const unsigned char *pBitstreamData;
    unsigned int nNumSlices;
    const unsigned int *pSliceDataOffsets;
    int ref_pic_flag, intra_pic_flag;
    unsigned int Reserved[30];
    unsigned int CodecReserved[1024];
};

struct fake_parser {
    CUVIDPARSERPARAMS p;
    int width, height;
    unsigned n_surfaces, next_idx;
    std::deque<int> delayed;        /* decoded, not yet displayed (display delay) */
    std::vector<unsigned char> rbsp;
};

struct stats_t { int maps, unmaps, cur_mapped, max_mapped, map_refused, decoders; } g_stats;
/* fault injection (fake_nvcuvid_fail): the next `count` calls of one kind, after `after` good ones, fail */
struct fault_t { int what, after, count; } g_fault = { 0, 0, 0 };
bool inject(int what)
{
    if (g_fault.what != what || g_fault.count <= 0) return false;
    if (g_fault.after > 0) { g_fault.after--; return false; }
    g_fault.count--;
    return true;
}
cudaStream_t g_poison = nullptr;

struct fake_decoder {
    CUVIDDECODECREATEINFO ci;
    size_t pitch, rows;
    std::vector<unsigned char *> decode_surf;
    std::vector<unsigned char *> out_surf;
    std::vector<bool> out_busy;
};

void unescape(const unsigned char *s, size_t n, std::vector<unsigned char> &out)
{
    out.clear();
    out.reserve(n);
    int zeros = 0;
    for (size_t i = 0; i < n; i++) {
        if (zeros >= 2 && s[i] == 3) { zeros = 0; continue; }          /* emulation prevention byte */
        out.push_back(s[i]);
        zeros = s[i] == 0 ? zeros + 1 : 0;
    }
}

void display(fake_parser *ps, int idx)
{
    CUVIDPARSERDISPINFO d;
    memset(&d, 0, sizeof(d));
    d.picture_index = idx;
    d.progressive_frame = 1;
    d.top_field_first = 1;
    ps->p.pfnDisplayPicture(ps->p.pUserData, &d);
}

void handle_nal(fake_parser *ps, const unsigned char *nal, size_t n)
{
    if (n < 1) return;
    const unsigned char type = nal[0];
    unescape(nal + 1, n - 1, ps->rbsp);
    if (type == 0x67 && ps->rbsp.size() >= 8) {
        /* a new sequence: the pictures of the old one are displayed first, as the real parser does */
        while (!ps->delayed.empty()) { display(ps, ps->delayed.front()); ps->delayed.pop_front(); }
        uint32_t w, h;
        memcpy(&w, &ps->rbsp[0], 4);
        memcpy(&h, &ps->rbsp[4], 4);
        ps->width = (int)w; ps->height = (int)h;
        CUVIDEOFORMAT f;
        memset(&f, 0, sizeof(f));
        f.codec = ps->p.CodecType;
        f.progressive_sequence = 1;
        f.min_num_decode_surfaces = 6;
        f.coded_width = (w + 15) & ~15u;
        f.coded_height = (h + 15) & ~15u;
        f.display_area.right = (int)w;
        f.display_area.bottom = (int)h;
        f.chroma_format = CUVID_CHROMA_420;
        int r = ps->p.pfnSequenceCallback(ps->p.pUserData, &f);
        ps->n_surfaces = r > 1 ? (unsigned)r : ps->p.ulMaxNumDecodeSurfaces;
        if (r == 0) ps->width = ps->height = 0;                         /* client refused the format */
    } else if (type == 0x65 && ps->width > 0) {
        const size_t need = (size_t)ps->width * ps->height * 3 / 2;
        if (ps->rbsp.size() < need) return;                             /* damaged picture: dropped */
        pic_params pp;
        memset(&pp, 0, sizeof(pp));
        pp.PicWidthInMbs = (ps->width + 15) / 16;
        pp.FrameHeightInMbs = (ps->height + 15) / 16;
        pp.CurrPicIdx = (int)(ps->next_idx++ % ps->n_surfaces);
        pp.nBitstreamDataLen = (unsigned)need;
        pp.pBitstreamData = ps->rbsp.data();
        pp.nNumSlices = 1;
        pp.intra_pic_flag = 1;
        if (!ps->p.pfnDecodePicture(ps->p.pUserData, &pp)) return;
        ps->delayed.push_back(pp.CurrPicIdx);
        while (ps->delayed.size() > ps->p.ulMaxDisplayDelay) { display(ps, ps->delayed.front()); ps->delayed.pop_front(); }
    }
}

} /* namespace */

FAKE_API int cuvidGetDecoderCaps(void *caps)
{
    const char *e = getenv("FAKE_NVCUVID_NO_ENGINE");
    if (e && atoi(e)) return 100;                                       /* CUDA_ERROR_NO_DEVICE, as on the real boxes */
    CUVIDDECODECAPS *c = (CUVIDDECODECAPS *)caps;
    c->bIsSupported = 1;
    c->nNumNVDECs = 1;
    c->nMaxWidth = c->nMaxHeight = 8192;
    return 0;
}

FAKE_API int cuvidCreateVideoParser(CUvideoparser *out, CUVIDPARSERPARAMS *p)
{
    if (!out || !p || !p->pfnSequenceCallback || !p->pfnDecodePicture || !p->pfnDisplayPicture) return 1;
    fake_parser *ps = new fake_parser();
    ps->p = *p;
    ps->width = ps->height = 0;
    ps->n_surfaces = p->ulMaxNumDecodeSurfaces ? p->ulMaxNumDecodeSurfaces : 1;
    ps->next_idx = 0;
    *out = ps;
    return 0;
}

FAKE_API int cuvidDestroyVideoParser(CUvideoparser h)
{
    delete (fake_parser *)h;
    return 0;
}

FAKE_API int cuvidParseVideoData(CUvideoparser h, CUVIDSOURCEDATAPACKET *pkt)
{
    fake_parser *ps = (fake_parser *)h;
    if (!ps || !pkt) return 1;
    const unsigned char *b = pkt->payload;
    const size_t n = b ? pkt->payload_size : 0;
    /* split at start codes; bytes before the first start code are ignored */
    size_t i = 0, nal_start = (size_t)-1;
    while (i + 3 <= n) {
        if (b[i] == 0 && b[i + 1] == 0 && b[i + 2] == 1) {
            if (nal_start != (size_t)-1) {
                size_t end = i;
                while (end > nal_start && b[end - 1] == 0) end--;       /* trailing zero of a 4-byte start code */
                handle_nal(ps, b + nal_start, end - nal_start);
            }
            nal_start = i + 3;
            i += 3;
        } else {
            i++;
        }
    }
    if (nal_start != (size_t)-1 && nal_start <= n) handle_nal(ps, b + nal_start, n - nal_start);
    if (pkt->flags & CUVID_PKT_ENDOFSTREAM) {
        while (!ps->delayed.empty()) { display(ps, ps->delayed.front()); ps->delayed.pop_front(); }
    }
    return 0;
}

FAKE_API int cuvidCreateDecoder(CUvideodecoder *out, CUVIDDECODECREATEINFO *ci)
{
    if (!out || !ci || ci->OutputFormat != CUVID_SURFACE_NV12 || ci->ulNumDecodeSurfaces < 1 || ci->ulNumOutputSurfaces < 1) return 1;
    if (inject(3)) return 2;
    fake_decoder *d = new fake_decoder();
    d->ci = *ci;
    d->pitch = (ci->ulTargetWidth + 511) & ~(size_t)511;                /* decoder-chosen pitch, like the real one */
    d->rows = ci->ulTargetHeight * 3 / 2 + 2;
    bool ok = true;
    for (unsigned long i = 0; ok && i < ci->ulNumDecodeSurfaces; i++) {
        unsigned char *p = nullptr;
        ok = cudaMalloc(&p, d->pitch * d->rows) == cudaSuccess;
        if (!ok) break;
        cudaMemset(p, 0xCD, d->pitch * d->rows);
        d->decode_surf.push_back(p);
    }
    for (unsigned long i = 0; ok && i < ci->ulNumOutputSurfaces; i++) {
        unsigned char *p = nullptr;
        ok = cudaMalloc(&p, d->pitch * d->rows) == cudaSuccess;
        if (!ok) break;
        d->out_surf.push_back(p);
        d->out_busy.push_back(false);
    }
    if (!ok) {                                                          /* out of device memory: nothing is kept */
        for (unsigned char *p : d->decode_surf) cudaFree(p);
        for (unsigned char *p : d->out_surf) cudaFree(p);
        delete d;
        return 2;
    }
    g_stats.decoders++;
    *out = d;
    return 0;
}

FAKE_API int cuvidDestroyDecoder(CUvideodecoder h)
{
    fake_decoder *d = (fake_decoder *)h;
    if (!d) return 1;
    cudaDeviceSynchronize();
    for (unsigned char *p : d->decode_surf) cudaFree(p);
    for (unsigned char *p : d->out_surf) cudaFree(p);
    delete d;
    return 0;
}

FAKE_API int cuvidDecodePicture(CUvideodecoder h, void *pic)
{
    fake_decoder *d = (fake_decoder *)h;
    pic_params *pp = (pic_params *)pic;
    if (!d || !pp || pp->CurrPicIdx < 0 || (size_t)pp->CurrPicIdx >= d->decode_surf.size()) return 1;
    const size_t w = d->ci.ulTargetWidth, hgt = d->ci.ulTargetHeight;
    if (pp->nBitstreamDataLen < w * hgt * 3 / 2) return 1;
    if (inject(2)) return 2;
    /* "decode": the tight NV12 picture lands in the pitched decode surface (rows h + h/2) */
    if (cudaMemcpy2D(d->decode_surf[pp->CurrPicIdx], d->pitch, pp->pBitstreamData, w, w, hgt + hgt / 2, cudaMemcpyHostToDevice) != cudaSuccess) return 2;
    return 0;
}

FAKE_API int cuvidMapVideoFrame64(CUvideodecoder h, int idx, unsigned long long *dptr, unsigned int *pitch, CUVIDPROCPARAMS *pp)
{
    fake_decoder *d = (fake_decoder *)h;
    if (!d || !dptr || !pitch || idx < 0 || (size_t)idx >= d->decode_surf.size()) return 1;
    if (inject(1)) return 2;
    size_t slot = 0;
    while (slot < d->out_surf.size() && d->out_busy[slot]) slot++;
    if (slot == d->out_surf.size()) { g_stats.map_refused++; return 3; }    /* more frames mapped than ulNumOutputSurfaces */
    /* post-processing is ENQUEUED on the caller's stream and not waited for, like the real decoder:
     * a client that reads the surface on another stream without ordering sees stale data */
    cudaStream_t st = pp ? (cudaStream_t)pp->output_stream : 0;
    if (cudaMemsetAsync(d->out_surf[slot], 0xEE, d->pitch * d->rows, st) != cudaSuccess) return 2;
    if (cudaMemcpyAsync(d->out_surf[slot], d->decode_surf[idx], d->pitch * d->rows, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return 2;
    d->out_busy[slot] = true;
    g_stats.maps++;
    if (++g_stats.cur_mapped > g_stats.max_mapped) g_stats.max_mapped = g_stats.cur_mapped;
    *dptr = (unsigned long long)(uintptr_t)d->out_surf[slot];
    *pitch = (unsigned int)d->pitch;
    return 0;
}

FAKE_API int cuvidUnmapVideoFrame64(CUvideodecoder h, unsigned long long dptr)
{
    fake_decoder *d = (fake_decoder *)h;
    if (!d) return 1;
    for (size_t i = 0; i < d->out_surf.size(); i++)
        if ((unsigned long long)(uintptr_t)d->out_surf[i] == dptr && d->out_busy[i]) {
            /* the surface goes back to the decoder NOW: poison it without waiting for anything the client enqueued */
            if (!g_poison) cudaStreamCreateWithFlags(&g_poison, cudaStreamNonBlocking);
            cudaMemsetAsync(d->out_surf[i], 0xDD, d->pitch * d->rows, g_poison);
            cudaStreamSynchronize(g_poison);
            d->out_busy[i] = false;
            g_stats.unmaps++;
            g_stats.cur_mapped--;
            return 0;
        }
    return 1;
}

/* test hook (not part of the NVDEC ABI): make the next `count` calls of one kind fail after `after` good ones.
 * what: 1 cuvidMapVideoFrame, 2 cuvidDecodePicture, 3 cuvidCreateDecoder; 0 clears */
FAKE_API void fake_nvcuvid_fail(int what, int after, int count)
{
    g_fault.what = what; g_fault.after = after; g_fault.count = count;
}

/* test hook (not part of the NVDEC ABI): v[0..5] = maps, unmaps, currently mapped, max mapped at once, refused maps,
 * decoders created; reset != 0 clears the counters afterwards */
FAKE_API void fake_nvcuvid_stats(int *v, int reset)
{
    if (v) { v[0] = g_stats.maps; v[1] = g_stats.unmaps; v[2] = g_stats.cur_mapped; v[3] = g_stats.max_mapped; v[4] = g_stats.map_refused; v[5] = g_stats.decoders; }
    if (reset) { const int cur = g_stats.cur_mapped; memset(&g_stats, 0, sizeof(g_stats)); g_stats.cur_mapped = cur; }
}
