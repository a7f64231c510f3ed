/*
 * jmc_launch_job for the CPU simulation of the host layer (tests/host_logic): stands in for jmc_kernels.cu.  The
 * "kernel" is the oracle's restatement of the reference function, enqueued on the simulated stream so that it runs
 * when the stream gets there -- not when it is launched.  Test infrastructure: the product never links this file.
 */
#include <string.h>

#include <atomic>
#include <vector>

#include "jmc_internal.h"
#include "../../oracle/jm_oracle.h"

std::atomic<int> g_fake_launches{0};        /* launches seen (the tests check batching) */
std::atomic<int> g_fake_frames{0};          /* frames converted */
std::atomic<int> g_fake_max_batch{0};

static void *frame_of(const jmc_frames &f, const std::vector<void *> &list, int i)
{
    if (!list.empty()) return list[(size_t)i];
    return f.base ? (uint8_t *)f.base + (size_t)i * f.stride : nullptr;
}

int jmc_launch_job(jmc_ctx *ctx, const jmc_job *job, cudaStream_t stream)
{
    if (!ctx || !job || job->n_frames < 1) { jmc_set_error("fake launch: bad job"); return JMC_ERR_INVALID; }
    const jmc_job j = *job;
    const bool on_host = (j.flags & JMC_JOB_LIST_ON_HOST) != 0;
    if (on_host && j.n_frames > JMC_INLINE_LIST_MAX) { jmc_set_error("fake launch: too many inline frames"); return JMC_ERR_INVALID; }
    /* pointer lists passed as kernel arguments are read NOW; device-resident lists when the kernel runs */
    std::vector<void *> surf_l, tight_l;
    if (on_host) {
        if (j.surf.list) surf_l.assign(j.surf.list, j.surf.list + j.n_frames);
        if (j.tight.list) tight_l.assign(j.tight.list, j.tight.list + j.n_frames);
    } else if (j.surf.list || j.tight.list) {
        jmc_set_error("fake launch: device-resident pointer lists are not simulated");
        return JMC_ERR_INVALID;
    }
    if (j.op != JMC_OP_NV12_TO_NV12 && j.op != JMC_OP_NV12_TO_I420 && j.op != JMC_OP_I420_TO_SURF && j.op != JMC_OP_NV12_TO_SURF) {
        jmc_set_error("fake launch: op %d is not simulated", j.op);
        return JMC_ERR_INVALID;
    }
    if (fake_cuda_launch_should_fail()) { jmc_set_error("fake launch: injected launch failure"); return JMC_ERR_CUDA; }
    g_fake_launches++;
    g_fake_frames += j.n_frames;
    for (int seen = g_fake_max_batch.load(); j.n_frames > seen && !g_fake_max_batch.compare_exchange_weak(seen, j.n_frames);) {}
    fake_cuda_enqueue(stream, [j, surf_l, tight_l] {
        for (int f = 0; f < j.n_frames; f++) {
            uint8_t *surf = (uint8_t *)frame_of(j.surf, surf_l, f), *tight = (uint8_t *)frame_of(j.tight, tight_l, f);
            const size_t tight_bytes = (size_t)j.width * j.height * 3 / 2;
            if (!surf || !tight) continue;
            if (j.width > 0 && j.height > 0 && (!fake_cuda_is_device_range(surf, (size_t)j.pitch * j.height * 3 / 2) || !fake_cuda_is_device_range(tight, tight_bytes ? tight_bytes : 1))) {
                fake_cuda_complain("kernel: a frame pointer is not inside a live device allocation");
                continue;
            }
            if (j.op == JMC_OP_NV12_TO_NV12 || j.op == JMC_OP_NV12_TO_I420) {
                int len = (int)tight_bytes;
                jmo_nvdec_output_frame(surf, j.pitch, j.width, j.height, j.op == JMC_OP_NV12_TO_I420 ? 1 : 0, 1, tight, &len);
            } else {
                jmo_nvenc_upload(tight, j.op == JMC_OP_I420_TO_SURF ? 0x10 : 0x1, j.width, j.height, surf, j.pitch);
            }
        }
    }, "conversion kernel");
    ctx->launches++;
    return JMC_OK;
}
