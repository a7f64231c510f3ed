/*
 * tools/pcie.cu -- host-link microbenchmark (not part of the product library).
 * What can the PCIe link of this box deliver to the delivery pipeline?  H2D / D2H alone and
 * concurrently, 1-D vs 2-D (1920 of 2048 bytes per row), default vs write-combined pinned memory.
 *   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/pcie.cu -o tools/pcie
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

int main()
{
    const size_t pitch = 2048, w = 1920, rows = 1620, frames = 30;      /* one 1080p NV12 surface = 1620 rows */
    const size_t surf = pitch * rows, tight = w * rows, n_in = surf * frames, n_out = tight * frames;
    void *h_in, *h_in_wc, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc(&h_in, n_in, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_in_wc, n_in, cudaHostAllocWriteCombined));
    CK(cudaHostAlloc(&h_out, n_out, cudaHostAllocDefault));
    CK(cudaMalloc(&d_in, n_in));
    CK(cudaMalloc(&d_out, n_out));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t a, b, c;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreate(&c));
    const int it = 20;
    auto report = [&](const char *name, double bytes, float ms) { printf("%-34s %8.2f GB/s\n", name, bytes * it / ms / 1e6); };
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(a, s1));
        for (int i = 0; i < it; i++) CK(cudaMemcpyAsync(d_in, h_in, n_in, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(b, s1)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
        if (rep) report("H2D 1-D pinned", n_in, ms);
        CK(cudaEventRecord(a, s1));
        for (int i = 0; i < it; i++) CK(cudaMemcpyAsync(d_in, h_in_wc, n_in, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(b, s1)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
        if (rep) report("H2D 1-D write-combined", n_in, ms);
        CK(cudaEventRecord(a, s1));
        for (int i = 0; i < it; i++) CK(cudaMemcpy2DAsync(d_in, pitch, h_in, pitch, w, rows * frames, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(b, s1)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
        if (rep) report("H2D 2-D 1920/2048 (useful bytes)", (double)w * rows * frames, ms);
        CK(cudaEventRecord(a, s2));
        for (int i = 0; i < it; i++) CK(cudaMemcpyAsync(h_out, d_out, n_out, cudaMemcpyDeviceToHost, s2));
        CK(cudaEventRecord(b, s2)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
        if (rep) report("D2H 1-D pinned", n_out, ms);
        /* concurrent */
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, s1)); CK(cudaStreamWaitEvent(s2, a, 0));
        for (int i = 0; i < it; i++) {
            CK(cudaMemcpyAsync(d_in, h_in, n_in, cudaMemcpyHostToDevice, s1));
            CK(cudaMemcpyAsync(h_out, d_out, n_out, cudaMemcpyDeviceToHost, s2));
        }
        CK(cudaEventRecord(b, s1)); CK(cudaEventRecord(c, s2));
        CK(cudaEventSynchronize(b)); CK(cudaEventSynchronize(c));
        float m1, m2; CK(cudaEventElapsedTime(&m1, a, b)); CK(cudaEventElapsedTime(&m2, a, c));
        if (rep) { report("bidir: H2D 1-D", n_in, m1); report("bidir: D2H 1-D", n_out, m2); }
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, s1)); CK(cudaStreamWaitEvent(s2, a, 0));
        for (int i = 0; i < it; i++) {
            CK(cudaMemcpy2DAsync(d_in, pitch, h_in_wc, pitch, w, rows * frames, cudaMemcpyHostToDevice, s1));
            CK(cudaMemcpyAsync(h_out, d_out, n_out, cudaMemcpyDeviceToHost, s2));
        }
        CK(cudaEventRecord(b, s1)); CK(cudaEventRecord(c, s2));
        CK(cudaEventSynchronize(b)); CK(cudaEventSynchronize(c));
        CK(cudaEventElapsedTime(&m1, a, b)); CK(cudaEventElapsedTime(&m2, a, c));
        if (rep) { report("bidir: H2D 2-D WC (useful bytes)", (double)w * rows * frames, m1); report("bidir: D2H 1-D (with 2-D H2D)", n_out, m2); }
    }
    return 0;
}
