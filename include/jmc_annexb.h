/*
 * jmc_annexb.h -- Annex-B (start-code) NAL unit splitter: the step in front of jm_nvdec_decode_frame.
 *
 * The reference keeps this in its test program (test_nv_dec/test_nv_dec.cpp:30-86, find_nalu_prefix /
 * find_nalu) and every caller of jm_nvdec_decode_frame needs it, so it is offered here with the same
 * results, minus the reference's one-byte over-read at the end of the buffer (:44-48 reads buf[3]
 * when only three bytes remain).
 */
#ifndef JMC_ANNEXB_H
#define JMC_ANNEXB_H

#include <stddef.h>

#if defined(__GNUC__)
#define JMC_API __attribute__((visibility("default")))
#else
#define JMC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Offset of the first start code (00 00 01 or 00 00 00 01) in buf[0, size), or -1.
 * *prefix_len = 3 or 4 (0 when not found).  test_nv_dec.cpp:30-61. */
JMC_API int jmc_annexb_find_prefix(const unsigned char *buf, int size, int *prefix_len);

/* The NAL unit that starts at the first start code of buf: returns its address (start code included)
 * and *nalu_len = bytes up to the NEXT start code.  NULL and *nalu_len = 0 when no second start code
 * is in the buffer yet (caller refills; at end of stream the remainder is the last NAL).
 * test_nv_dec.cpp:63-86. */
JMC_API const unsigned char *jmc_annexb_find_nalu(const unsigned char *buf, int size, int *nalu_len);

#ifdef __cplusplus
}
#endif
#endif
