/* jmc_internal.h -- shared between the translation units of libjmcodec_b200.so (not installed). */
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "jmc_cuda.h"

struct jmc_ctx {
    int device;
    int sm_count;
    cudaStream_t stream[3];   /* 0 convert, 1 upload, 2 delivery */
    uint64_t launches;
};

void jmc_set_error(const char *fmt, ...);
int jmc_cuda_fail(cudaError_t e, const char *what);

#define JMC_CUDA(call)                                         \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) return jmc_cuda_fail(e_, #call); \
    } while (0)

/* Makes a device current on the calling thread for the lifetime of the guard and RESTORES the caller's
 * device afterwards: the reference pushes and pops its own CUcontext around every call
 * (nv_dec.cpp:378,398,423,471), so a host application that drives several GPUs (or PyTorch next to us)
 * never finds its current device changed by a jmc_* / jm_nvdec_* / jm_nvenc_* call.  err != 0: the
 * device could not be made current (already reported through jmc_set_error). */
struct jmc_device_guard {
    int prev, err;
    bool switched;
    explicit jmc_device_guard(int device);
    explicit jmc_device_guard(const jmc_ctx *c);
    ~jmc_device_guard();
    jmc_device_guard(const jmc_device_guard &) = delete;
    jmc_device_guard &operator=(const jmc_device_guard &) = delete;
private:
    void enter(int device);
};
#define JMC_BIND(c) jmc_device_guard guard_(c); do { if (guard_.err) return guard_.err; } while (0)

/* The JMC_* environment switches (kernel-variant A/B, pipeline upload shape), read ONCE per process --
 * not per launch -- and again only when jmc_reload_env() is called (tests flip them in-process). */
struct jmc_env_flags {
    bool no_bulk, no_rows, rows_single, rows_always, rgb_bulk_always, pipeline_h2d_2d;
    int rgb_flat;             /* -1 unset (heuristic), 0 off, 1 on */
    int rgb2_flat;            /* same, for the RGB24 -> NV12 kernel */
    bool pad_zero;            /* JMC_PAD_ZERO=1: as if every job carried JMC_JOB_PAD_ZERO (measurements) */
    int rgb_bulk_pairs;       /* JMC_RGB_BULK_PAIRS: row pairs per CTA of the re-aligning bulk colour kernel (0: two for the fused op, RGB alone stays on the warp-per-task kernel) */
    int brows_rows;           /* > 0: rows per tile of the bulk-loaded row kernels (tuning) */
};
const jmc_env_flags &jmc_env();

/* kernels (jmc_kernels.cu) */
int jmc_launch_job(jmc_ctx *ctx, const jmc_job *job, cudaStream_t stream);
