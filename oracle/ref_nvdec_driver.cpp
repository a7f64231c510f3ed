/*
 * TEST INFRASTRUCTURE ONLY (checker / CPU baseline).  Never linked into the product library.
 *
 * Thin extern "C" driver around the UNMODIFIED reference translation unit
 * /root/reference/nv_dec/nv_dec.cpp (compiled in place by oracle/Makefile into
 * oracle/_ref/libjmref.so).  It calls the reference's own exported
 *     jm_nvdec_output_frame()        nv_dec/nv_dec.cpp:750-828
 * on a hand-built nv_frame_buf (nv_dec/nv_dec.h:58-66) so that the real CPU
 * compaction / de-interleave loop runs with no GPU, driver or CUVID.
 *
 * No reference source is copied here: this file only includes the reference
 * headers from where they lie and pokes the ctx fields the loop reads.
 */
#include "nv_dec.h"
#include "jm_nv_dec.h"

#include <pthread.h>
#include <time.h>

extern "C" {

/* One call of the reference function on one pitched surface.
 * have_frame==0 -> ctx->cur_out_frame stays NULL (reference returns -1, nv_dec.cpp:757-758).
 * surf==NULL    -> big_buf NULL (reference returns -1, nv_dec.cpp:768-771).           */
__attribute__((visibility("default")))
int jmref_nvdec_output_frame(const unsigned char *surf, int pitch, int width, int height,
                             int out_fmt, int have_frame, unsigned char *out_buf, int *out_len)
{
    handle_nvdec h = jm_nvdec_create_handle();           /* new + memset, no CUDA (nv_dec.cpp:54-60) */
    nvdec_ctx *ctx = (nvdec_ctx *)h;
    nv_frame_buf fb;
    memset(&fb, 0, sizeof(fb));
    fb.big_buf = (unsigned char *)surf;
    fb.pitch = pitch;
    fb.width = width;
    fb.height = height;
    fb.big_buf_len = fb.data_len = pitch * height * 3 / 2;
    ctx->out_fmt = out_fmt;
    ctx->cur_out_frame = have_frame ? &fb : NULL;
    int r = jm_nvdec_output_frame(out_buf, out_len, h);
    delete ctx;                                           /* not jm_nvdec_deinit: that would call CUDA */
    return r;
}

struct jmref_job {
    const unsigned char *surf_base; size_t surf_stride; int n_surf;
    unsigned char *out_base; size_t out_stride; int n_out;
    int pitch, width, height, out_fmt, frames, tid, nthreads;
    long long bad;
};

static void *jmref_worker(void *p)
{
    jmref_job *j = (jmref_job *)p;
    handle_nvdec h = jm_nvdec_create_handle();           /* one handle per thread, as the reference is single-threaded per handle */
    nvdec_ctx *ctx = (nvdec_ctx *)h;
    nv_frame_buf fb;
    memset(&fb, 0, sizeof(fb));
    fb.pitch = j->pitch; fb.width = j->width; fb.height = j->height;
    ctx->out_fmt = j->out_fmt;
    ctx->cur_out_frame = &fb;
    const int need = j->width * j->height * 3 / 2;
    for (int f = j->tid; f < j->frames; f += j->nthreads) {
        fb.big_buf = (unsigned char *)(j->surf_base + (size_t)(f % j->n_surf) * j->surf_stride);
        int len = (int)j->out_stride;
        int r = jm_nvdec_output_frame(j->out_base + (size_t)(f % j->n_out) * j->out_stride, &len, h);
        if (r != need) j->bad++;
    }
    delete ctx;
    return NULL;
}

/* Timed loop for the CPU baseline: `frames` calls of the reference function, frame f reading
 * surface f % n_surf and writing output slot f % n_out, split round-robin over `nthreads`
 * handles.  Returns elapsed seconds (CLOCK_MONOTONIC), or -1 if any call misbehaved. */
__attribute__((visibility("default")))
double jmref_nvdec_run(const unsigned char *surf_base, size_t surf_stride, int n_surf,
                       unsigned char *out_base, size_t out_stride, int n_out,
                       int pitch, int width, int height, int out_fmt, int frames, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    jmref_job *jobs = new jmref_job[nthreads];
    pthread_t *th = new pthread_t[nthreads];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        jmref_job j = { surf_base, surf_stride, n_surf, out_base, out_stride, n_out,
                        pitch, width, height, out_fmt, frames, t, nthreads, 0 };
        jobs[t] = j;
        if (nthreads == 1) jmref_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, jmref_worker, &jobs[t]);
    }
    long long bad = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        bad += jobs[t].bad;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    delete[] jobs;
    delete[] th;
    if (bad) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

} /* extern "C" */

/* NAL splitter of the reference's test program (test_nv_dec/test_nv_dec.cpp:30-86), compiled from
 * the unmodified source into this library with main() renamed. */
int find_nalu_prefix(unsigned char *buf_start, int buf_size, int *prefix_len);
unsigned char *find_nalu(unsigned char *buf, int size, int *nalu_len);

extern "C" {
__attribute__((visibility("default")))
int jmref_find_nalu_prefix(unsigned char *buf, int size, int *prefix_len) { return find_nalu_prefix(buf, size, prefix_len); }

/* returns the offset of the NAL in buf, or -1 when the reference returns NULL */
__attribute__((visibility("default")))
int jmref_find_nalu(unsigned char *buf, int size, int *nalu_len)
{
    unsigned char *p = find_nalu(buf, size, nalu_len);
    return p ? (int)(p - buf) : -1;
}
}
