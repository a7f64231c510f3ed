/*
 * cuvid_min.h -- the slice of NVIDIA's NVDEC (CUVID) C ABI that the decoder front-end binds at run
 * time from libnvcuvid.so.1 (dlopen; nothing is linked).  Only what jm_nv_dec.cu touches is declared:
 * the parser (cuvidCreateVideoParser / cuvidParseVideoData), the decoder (cuvidCreateDecoder /
 * cuvidDecodePicture) and frame mapping (cuvidMapVideoFrame64).  Layouts follow the public
 * nvcuvid.h / cuviddec.h ABI (binary compatible from Video Codec SDK 7 to 12: newer fields live in
 * what older headers call "reserved"); the reference drives the same calls through its vendored
 * nv_sdk/inc/dynlink_nvcuvid.h / dynlink_cuviddec.h (nv_dec/nv_dec.cpp:23-52,278-403,496-540).
 */
#pragma once
#include <stdint.h>

extern "C" {

typedef int CUVID_RESULT;                       /* CUresult */
typedef void *CUvideoparser;
typedef void *CUvideodecoder;
typedef void *CUvideoctxlock;
typedef long long CUvideotimestamp;

enum {                                          /* cudaVideoCodec */
    CUVID_CODEC_MPEG1 = 0, CUVID_CODEC_MPEG2 = 1, CUVID_CODEC_MPEG4 = 2, CUVID_CODEC_VC1 = 3, CUVID_CODEC_H264 = 4,
    CUVID_CODEC_JPEG = 5, CUVID_CODEC_HEVC = 8, CUVID_CODEC_VP8 = 9, CUVID_CODEC_VP9 = 10
};
enum { CUVID_CHROMA_420 = 1 };                  /* cudaVideoChromaFormat */
enum { CUVID_SURFACE_NV12 = 0 };                /* cudaVideoSurfaceFormat */
enum { CUVID_DEINTERLACE_WEAVE = 0, CUVID_DEINTERLACE_ADAPTIVE = 2 };
enum { CUVID_CREATE_PREFER_CUVID = 4 };
enum { CUVID_PKT_ENDOFSTREAM = 1, CUVID_PKT_TIMESTAMP = 2 };

typedef struct {
    int codec;
    struct { unsigned int numerator, denominator; } frame_rate;
    unsigned char progressive_sequence, bit_depth_luma_minus8, bit_depth_chroma_minus8;
    unsigned char min_num_decode_surfaces;      /* "reserved1" in SDK 7 headers */
    unsigned int coded_width, coded_height;
    struct { int left, top, right, bottom; } display_area;
    int chroma_format;
    unsigned int bitrate;
    struct { int x, y; } display_aspect_ratio;
    struct { unsigned char flags, color_primaries, transfer_characteristics, matrix_coefficients; } video_signal_description;
    unsigned int seqhdr_data_length;
} CUVIDEOFORMAT;

typedef struct {
    CUVIDEOFORMAT format;
    unsigned char raw_seqhdr_data[1024];
} CUVIDEOFORMATEX;

typedef struct {
    unsigned long flags;
    unsigned long payload_size;
    const unsigned char *payload;
    CUvideotimestamp timestamp;
} CUVIDSOURCEDATAPACKET;

typedef struct {
    int picture_index, progressive_frame, top_field_first, repeat_first_field;
    CUvideotimestamp timestamp;
} CUVIDPARSERDISPINFO;

/* Opaque to us: the parser fills it, we hand it to cuvidDecodePicture unchanged.  Only the leading
 * fields are named (CurrPicIdx is useful for diagnostics); the size is never needed on our side. */
typedef struct {
    int PicWidthInMbs, FrameHeightInMbs, CurrPicIdx;
} CUVIDPICPARAMS_HEAD;

typedef int (*PFNVIDSEQUENCECALLBACK)(void *, CUVIDEOFORMAT *);
typedef int (*PFNVIDDECODECALLBACK)(void *, void * /* CUVIDPICPARAMS* */);
typedef int (*PFNVIDDISPLAYCALLBACK)(void *, CUVIDPARSERDISPINFO *);

typedef struct {
    int CodecType;
    unsigned int ulMaxNumDecodeSurfaces, ulClockRate, ulErrorThreshold, ulMaxDisplayDelay;
    unsigned int uReserved1[5];
    void *pUserData;
    PFNVIDSEQUENCECALLBACK pfnSequenceCallback;
    PFNVIDDECODECALLBACK pfnDecodePicture;
    PFNVIDDISPLAYCALLBACK pfnDisplayPicture;
    void *pvReserved2[7];
    CUVIDEOFORMATEX *pExtVideoInfo;
} CUVIDPARSERPARAMS;

typedef struct {
    unsigned long ulWidth, ulHeight, ulNumDecodeSurfaces;
    int CodecType, ChromaFormat;
    unsigned long ulCreationFlags, bitDepthMinus8;
    unsigned long Reserved1[4];                 /* ulIntraDecodeOnly, ulMaxWidth, ulMaxHeight, reserved in newer SDKs */
    struct { short left, top, right, bottom; } display_area;
    int OutputFormat, DeinterlaceMode;
    unsigned long ulTargetWidth, ulTargetHeight, ulNumOutputSurfaces;
    CUvideoctxlock vidLock;
    struct { short left, top, right, bottom; } target_rect;
    unsigned long Reserved2[5];
} CUVIDDECODECREATEINFO;

typedef struct {
    int progressive_frame, second_field, top_field_first, unpaired_field;
    unsigned int reserved_flags, reserved_zero;
    unsigned long long raw_input_dptr;
    unsigned int raw_input_pitch, raw_input_format;
    unsigned long long raw_output_dptr;
    unsigned int raw_output_pitch, Reserved1;
    void *output_stream;                        /* CUstream the post-processed surface is produced on (SDK >= 8) */
    unsigned int Reserved[46];
    void *Reserved3[3];
} CUVIDPROCPARAMS;

/* cuvidGetDecoderCaps (Video Codec SDK >= 8.0; the headers the reference vendors predate it, so this layout is
 * restated from the published cuviddec.h and pinned by the static_asserts below: three IN words, three reserved
 * words, then bIsSupported at byte 24; 88 bytes in every SDK from 8.0 to 12.x -- later SDKs only renamed
 * reserved words at the end). */
typedef struct {
    int eCodecType, eChromaFormat;              /* IN: cudaVideoCodec, cudaVideoChromaFormat */
    unsigned int nBitDepthMinus8;               /* IN */
    unsigned int reserved1[3];
    unsigned char bIsSupported;                 /* OUT */
    unsigned char nNumNVDECs;                   /* OUT ("reserved2" in SDK 8) */
    unsigned short nOutputFormatMask;
    unsigned int nMaxWidth, nMaxHeight, nMaxMBCount;
    unsigned short nMinWidth, nMinHeight;
    unsigned char bIsHistogramSupported, nCounterBitDepth;
    unsigned short nMaxHistogramBins;
    unsigned int reserved3[10];
} CUVIDDECODECAPS;
#ifdef __cplusplus
static_assert(sizeof(CUVIDDECODECAPS) == 88, "CUVIDDECODECAPS is 88 bytes");
static_assert(__builtin_offsetof(CUVIDDECODECAPS, bIsSupported) == 24, "bIsSupported follows six 32-bit words");
static_assert(__builtin_offsetof(CUVIDDECODECAPS, nMaxWidth) == 28 && __builtin_offsetof(CUVIDDECODECAPS, nMinWidth) == 40, "caps limits");
#endif
typedef CUVID_RESULT (*tcuvidGetDecoderCaps)(CUVIDDECODECAPS *);

typedef CUVID_RESULT (*tcuvidCreateVideoParser)(CUvideoparser *, CUVIDPARSERPARAMS *);
typedef CUVID_RESULT (*tcuvidParseVideoData)(CUvideoparser, CUVIDSOURCEDATAPACKET *);
typedef CUVID_RESULT (*tcuvidDestroyVideoParser)(CUvideoparser);
typedef CUVID_RESULT (*tcuvidCreateDecoder)(CUvideodecoder *, CUVIDDECODECREATEINFO *);
typedef CUVID_RESULT (*tcuvidDestroyDecoder)(CUvideodecoder);
typedef CUVID_RESULT (*tcuvidDecodePicture)(CUvideodecoder, void * /* CUVIDPICPARAMS* */);
typedef CUVID_RESULT (*tcuvidMapVideoFrame64)(CUvideodecoder, int, unsigned long long *, unsigned int *, CUVIDPROCPARAMS *);
typedef CUVID_RESULT (*tcuvidUnmapVideoFrame64)(CUvideodecoder, unsigned long long);

} /* extern "C" */
