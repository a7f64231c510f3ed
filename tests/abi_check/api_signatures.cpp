/* Compile-time check (tests/test_cuvid_abi.py): include/jm_nv_dec.h and include/jmnv_enc.h declare every reference
 * entry point with the reference's exact function type, and nv_enc_param with the reference's layout.  The
 * reference headers (nv_dec/jm_nv_dec.h, nv_enc/jmnv_enc.h) are included in place inside a namespace; nothing runs. */
#include <stddef.h>
#include <stdint.h>

#define _declspec(x)
namespace ref {
#include "nv_dec/jm_nv_dec.h"
#undef JMDLL_FUNC
#undef JMDLL_API
#include "nv_enc/jmnv_enc.h"
}
#undef JMDLL_FUNC
#undef JMDLL_API
#undef _JM_NV_DECODER_H_                /* both sides use the reference's include guards */
#undef _JMNV_ENC_H_
namespace mine {
#include "jm_nv_dec.h"
#include "jmnv_enc.h"
}

template <class A, class B> struct same { static const bool value = false; };
template <class A> struct same<A, A> { static const bool value = true; };
#define SAME_FN(F) static_assert(same<decltype(ref::F), decltype(mine::F)>::value, "signature of " #F)
#define SAME_OFF(T, F) static_assert(offsetof(ref::T, F) == offsetof(mine::T, F), "offsetof " #T "." #F)

SAME_FN(jm_nvdec_create_handle);
SAME_FN(jm_nvdec_init);
SAME_FN(jm_nvdec_deinit);
SAME_FN(jm_nvdec_decode_frame);
SAME_FN(jm_nvdec_output_frame);
SAME_FN(jm_nvdec_stream_info);
SAME_FN(jm_nvdec_set_eof);
SAME_FN(jm_nvdec_is_exit);
SAME_FN(jm_nvdec_show_dec_info);
SAME_FN(jm_nvdec_is_hw_support);

SAME_FN(jm_nvenc_create_handle);
/* nv_enc_param is a distinct type per namespace: compare shape, then layout below */
static_assert(same<decltype(ref::jm_nvenc_init), int(ref::nv_enc_param *, void *)>::value &&
              same<decltype(mine::jm_nvenc_init), int(mine::nv_enc_param *, void *)>::value, "signature of jm_nvenc_init");
SAME_FN(jm_nvenc_deinit);
SAME_FN(jm_nvenc_enc_frame);
SAME_FN(jm_nvenc_get_bitstream);
SAME_FN(jm_nvenc_get_spspps_len);
SAME_FN(jm_nvenc_get_spspps);
SAME_FN(jm_nvenc_memory_alloc_host);
SAME_FN(jm_nvenc_memory_release_host);

static_assert(sizeof(ref::nv_enc_param) == sizeof(mine::nv_enc_param), "sizeof nv_enc_param");
SAME_OFF(nv_enc_param, codec_id);
SAME_OFF(nv_enc_param, in_fmt);
SAME_OFF(nv_enc_param, preset);
SAME_OFF(nv_enc_param, src_width);
SAME_OFF(nv_enc_param, src_height);
SAME_OFF(nv_enc_param, dst_width);
SAME_OFF(nv_enc_param, dst_height);
SAME_OFF(nv_enc_param, fps);
SAME_OFF(nv_enc_param, bitrate_kb);
SAME_OFF(nv_enc_param, gop_len);
SAME_OFF(nv_enc_param, num_bframe);
SAME_OFF(nv_enc_param, is_external_alloc);
SAME_OFF(nv_enc_param, qp);

int main() { return 0; }
