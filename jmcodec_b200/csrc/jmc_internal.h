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

/* Make ctx's device (its primary CUDA context) current on the calling thread.  Every entry point that
 * touches CUDA outside jmc_* calls this first: a handle may be used from any one thread at a time. */
int jmc_bind_thread(const jmc_ctx *ctx);

/* kernels (jmc_kernels.cu) */
int jmc_launch_job(jmc_ctx *ctx, const jmc_job *job, cudaStream_t stream);
