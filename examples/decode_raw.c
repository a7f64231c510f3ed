/*
 * examples/decode_raw.c -- the jm_nvdec_* call sequence of test_nv_dec.cpp:163-259 in plain C99, fed
 * with decoded surfaces (JM_NVDEC_CODEC_RAW_NV12) instead of a bitstream.
 *
 *   gcc -std=c99 -Iinclude examples/decode_raw.c -Ljmcodec_b200 -ljmcodec_b200 -Wl,-rpath,$PWD/jmcodec_b200 -o decode_raw
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jm_nv_dec.h"
#include "jmc_annexb.h"
#include "jmc_cuda.h"
#include "jmnv_enc.h"

int main(void)
{
    const int w = 1920, h = 1080, pitch = 2048, frames = 8;
    const size_t surf = (size_t)pitch * h * 3 / 2;
    const int need = w * h * 3 / 2;
    unsigned char *pkt = malloc(sizeof(jm_nvdec_raw_packet) + surf);
    unsigned char *out = malloc((size_t)need);
    jm_nvdec_raw_packet hdr = { JM_NVDEC_RAW_MAGIC, w, h, pitch, 0, 0, 0 };
    handle_nvdec dec;
    int f, got = 0, len, decoded = 0;

    if (!jm_nvdec_is_hw_support()) { fprintf(stderr, "no CUDA device: %s\n", jmc_last_error()); return 1; }
    dec = jm_nvdec_create_handle();
    if (jm_nvdec_init(JM_NVDEC_CODEC_RAW_NV12, 1 /* "YV12" = I420 */, NULL, 0, dec) != 0) return 1;
    memcpy(pkt, &hdr, sizeof hdr);
    for (f = 0; f < frames; f++) {
        memset(pkt + sizeof hdr, 16 + f, surf);                        /* a flat grey surface */
        jm_nvdec_decode_frame(pkt, (int)(sizeof hdr + surf), &got, dec);
        if (got == 1) {
            len = need;
            if (jm_nvdec_output_frame(out, &len, dec) == need && out[0] == 16 + f && out[need - 1] == 16 + f) decoded++;
        }
    }
    jm_nvdec_decode_frame(NULL, 0, &got, dec);                          /* end of stream */
    printf("%s", jm_nvdec_show_dec_info(dec));
    printf("decoded %d of %d frames, exit=%d\n", decoded, frames, (int)jm_nvdec_is_exit(dec));
    jm_nvdec_deinit(dec);
    free(pkt);
    free(out);
    return decoded == frames ? 0 : 1;
}
