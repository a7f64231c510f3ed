/* jmc_annexb.cpp -- see include/jmc_annexb.h (host-side helper; no CUDA) */
#include "jmc_annexb.h"

extern "C" {

int jmc_annexb_find_prefix(const unsigned char *buf, int size, int *prefix_len)
{
    if (prefix_len) *prefix_len = 0;
    if (!buf) return -1;
    for (int off = 0; off + 3 <= size; off++) {                          /* test_nv_dec.cpp:37 */
        const unsigned char *b = buf + off;
        if (b[0] != 0 || b[1] != 0) continue;
        if (b[2] == 1) { if (prefix_len) *prefix_len = 3; return off; }  /* :40-46 */
        if (b[2] == 0 && off + 4 <= size && b[3] == 1) { if (prefix_len) *prefix_len = 4; return off; }   /* :47-54 */
    }
    return -1;
}

const unsigned char *jmc_annexb_find_nalu(const unsigned char *buf, int size, int *nalu_len)
{
    int prefix1 = 0, prefix2 = 0;
    if (nalu_len) *nalu_len = 0;
    const int off1 = jmc_annexb_find_prefix(buf, size, &prefix1);        /* :70 */
    if (off1 < 0) return nullptr;                                        /* the reference assumes a start code is there (:72) */
    const int off2 = jmc_annexb_find_prefix(buf + off1 + prefix1, size - prefix1 - off1, &prefix2);   /* :73 */
    if (off2 < 0) return nullptr;                                        /* :75-78: need more data */
    if (nalu_len) *nalu_len = off2 + prefix1;                            /* :80 */
    return buf + off1;                                                   /* :81 */
}

} /* extern "C" */
