/* Compile-time check (tests/test_cuvid_abi.py): jmcodec_b200/csrc/cuvid_min.h declares the same C ABI as the
 * NVDEC headers the reference vendors (nv_sdk/inc/dynlink_nvcuvid.h, dynlink_cuviddec.h) -- sizes, field offsets
 * and enum values of everything jm_nv_dec.cu touches.  Fields that newer SDKs carved out of "reserved" space
 * (min_num_decode_surfaces, output_stream) are checked against the reserved field they replace.
 * Nothing is executed; the reference headers are only included, in place. */
#include <stddef.h>
#include <stdint.h>

#include "dynlink_nvcuvid.h"          /* the reference's (pulls in dynlink_cuviddec.h and dynlink_cuda.h) */

namespace mine {
#include "cuvid_min.h"
}

#define SAME_SIZE(T) static_assert(sizeof(::T) == sizeof(mine::T), "sizeof " #T)
#define SAME_OFF(T, F) static_assert(offsetof(::T, F) == offsetof(mine::T, F), "offsetof " #T "." #F)
#define SAME_OFF2(T, F, G) static_assert(offsetof(::T, F) == offsetof(mine::T, G), "offsetof " #T "." #F " vs " #G)

SAME_SIZE(CUVIDEOFORMAT);
SAME_OFF(CUVIDEOFORMAT, codec);
SAME_OFF(CUVIDEOFORMAT, frame_rate);
SAME_OFF(CUVIDEOFORMAT, progressive_sequence);
SAME_OFF(CUVIDEOFORMAT, bit_depth_luma_minus8);
SAME_OFF(CUVIDEOFORMAT, coded_width);
SAME_OFF(CUVIDEOFORMAT, coded_height);
SAME_OFF(CUVIDEOFORMAT, display_area);
SAME_OFF(CUVIDEOFORMAT, chroma_format);
SAME_OFF(CUVIDEOFORMAT, seqhdr_data_length);
SAME_SIZE(CUVIDEOFORMATEX);
SAME_OFF(CUVIDEOFORMATEX, raw_seqhdr_data);

SAME_SIZE(CUVIDSOURCEDATAPACKET);
SAME_OFF(CUVIDSOURCEDATAPACKET, flags);
SAME_OFF(CUVIDSOURCEDATAPACKET, payload_size);
SAME_OFF(CUVIDSOURCEDATAPACKET, payload);
SAME_OFF(CUVIDSOURCEDATAPACKET, timestamp);

SAME_SIZE(CUVIDPARSERDISPINFO);
SAME_OFF(CUVIDPARSERDISPINFO, picture_index);
SAME_OFF(CUVIDPARSERDISPINFO, progressive_frame);
SAME_OFF(CUVIDPARSERDISPINFO, top_field_first);
SAME_OFF(CUVIDPARSERDISPINFO, repeat_first_field);
SAME_OFF(CUVIDPARSERDISPINFO, timestamp);

SAME_SIZE(CUVIDPARSERPARAMS);
SAME_OFF(CUVIDPARSERPARAMS, CodecType);
SAME_OFF(CUVIDPARSERPARAMS, ulMaxNumDecodeSurfaces);
SAME_OFF(CUVIDPARSERPARAMS, ulMaxDisplayDelay);
SAME_OFF(CUVIDPARSERPARAMS, pUserData);
SAME_OFF(CUVIDPARSERPARAMS, pfnSequenceCallback);
SAME_OFF(CUVIDPARSERPARAMS, pfnDecodePicture);
SAME_OFF(CUVIDPARSERPARAMS, pfnDisplayPicture);
SAME_OFF(CUVIDPARSERPARAMS, pExtVideoInfo);

SAME_SIZE(CUVIDDECODECREATEINFO);
SAME_OFF(CUVIDDECODECREATEINFO, ulWidth);
SAME_OFF(CUVIDDECODECREATEINFO, ulHeight);
SAME_OFF(CUVIDDECODECREATEINFO, ulNumDecodeSurfaces);
SAME_OFF(CUVIDDECODECREATEINFO, CodecType);
SAME_OFF(CUVIDDECODECREATEINFO, ChromaFormat);
SAME_OFF(CUVIDDECODECREATEINFO, ulCreationFlags);
SAME_OFF(CUVIDDECODECREATEINFO, display_area);
SAME_OFF(CUVIDDECODECREATEINFO, OutputFormat);
SAME_OFF(CUVIDDECODECREATEINFO, DeinterlaceMode);
SAME_OFF(CUVIDDECODECREATEINFO, ulTargetWidth);
SAME_OFF(CUVIDDECODECREATEINFO, ulTargetHeight);
SAME_OFF(CUVIDDECODECREATEINFO, ulNumOutputSurfaces);
SAME_OFF(CUVIDDECODECREATEINFO, vidLock);
SAME_OFF(CUVIDDECODECREATEINFO, target_rect);

SAME_SIZE(CUVIDPROCPARAMS);
SAME_OFF(CUVIDPROCPARAMS, progressive_frame);
SAME_OFF(CUVIDPROCPARAMS, second_field);
SAME_OFF(CUVIDPROCPARAMS, top_field_first);
SAME_OFF(CUVIDPROCPARAMS, unpaired_field);
SAME_OFF(CUVIDPROCPARAMS, raw_input_dptr);
SAME_OFF(CUVIDPROCPARAMS, raw_output_dptr);

static_assert((int)cudaVideoCodec_MPEG2 == (int)mine::CUVID_CODEC_MPEG2 && (int)cudaVideoCodec_MPEG4 == (int)mine::CUVID_CODEC_MPEG4 &&
              (int)cudaVideoCodec_VC1 == (int)mine::CUVID_CODEC_VC1 && (int)cudaVideoCodec_H264 == (int)mine::CUVID_CODEC_H264 &&
              (int)cudaVideoCodec_JPEG == (int)mine::CUVID_CODEC_JPEG && (int)cudaVideoCodec_HEVC == (int)mine::CUVID_CODEC_HEVC &&
              (int)cudaVideoCodec_VP8 == (int)mine::CUVID_CODEC_VP8 && (int)cudaVideoCodec_VP9 == (int)mine::CUVID_CODEC_VP9, "codec ids");
static_assert((int)cudaVideoChromaFormat_420 == (int)mine::CUVID_CHROMA_420, "chroma 4:2:0");
static_assert((int)cudaVideoSurfaceFormat_NV12 == (int)mine::CUVID_SURFACE_NV12, "NV12 surface format");
static_assert((int)cudaVideoDeinterlaceMode_Weave == (int)mine::CUVID_DEINTERLACE_WEAVE &&
              (int)cudaVideoDeinterlaceMode_Adaptive == (int)mine::CUVID_DEINTERLACE_ADAPTIVE, "deinterlace modes");
static_assert((int)cudaVideoCreate_PreferCUVID == (int)mine::CUVID_CREATE_PREFER_CUVID, "creation flag");
static_assert((int)CUVID_PKT_ENDOFSTREAM == (int)mine::CUVID_PKT_ENDOFSTREAM && (int)CUVID_PKT_TIMESTAMP == (int)mine::CUVID_PKT_TIMESTAMP, "packet flags");

int main() { return 0; }
