/* TEST INFRASTRUCTURE ONLY: force-included in front of the reference's test_nv_dec.cpp, whose input
 * and output paths are hard-coded Windows paths (test_nv_dec.cpp:115,119).  Reads go to
 * $JM_TEST_INPUT, writes to $JM_TEST_OUTPUT (default /dev/null). */
#ifndef JMC_TEST_REDIRECT_H
#define JMC_TEST_REDIRECT_H
#include <stdio.h>
#include <stdlib.h>
static inline FILE *jmshim_fopen(const char *path, const char *mode)
{
    const char *in = getenv("JM_TEST_INPUT"), *out = getenv("JM_TEST_OUTPUT");
    if (mode[0] == 'r') return (fopen)(in ? in : path, mode);
    return (fopen)(out ? out : "/dev/null", mode);
}
#define fopen(p, m) jmshim_fopen(p, m)
#endif
