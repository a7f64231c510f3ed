/* tools/pcie_thp.cu -- does backing pinned memory with transparent huge pages change the host-link rate?
 * (cudaHostAlloc vs 2 MB-aligned mmap + MADV_HUGEPAGE + cudaHostRegister), H2D and D2H concurrently. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <sys/mman.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static void *thp_alloc(size_t n)
{
    n = (n + (2u << 20) - 1) & ~((size_t)(2u << 20) - 1);
    void *p = mmap(nullptr, n + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    void *a = (void *)(((uintptr_t)p + (2u << 20) - 1) & ~((uintptr_t)(2u << 20) - 1));
    madvise(a, n, MADV_HUGEPAGE);
    memset(a, 1, n);
    return a;
}

int main(int argc, char **argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1;      /* 0: cudaHostAlloc, 1: THP, -1: both */
    const int iters = argc > 2 ? atoi(argv[2]) : 10;
    const size_t n = (size_t)960 << 20;
    void *d_in, *d_out;
    CK(cudaMalloc(&d_in, n)); CK(cudaMalloc(&d_out, n));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t a, b, c;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreate(&c));
    for (int mode = 0; mode < 2; mode++) {
        if (only >= 0 && mode != only) continue;
        void *h_in, *h_out;
        if (mode == 0) { CK(cudaHostAlloc(&h_in, n, cudaHostAllocDefault)); CK(cudaHostAlloc(&h_out, n, cudaHostAllocDefault)); }
        else {
            h_in = thp_alloc(n); h_out = thp_alloc(n);
            if (!h_in || !h_out) { printf("mmap failed\n"); return 1; }
            CK(cudaHostRegister(h_in, n, cudaHostRegisterDefault)); CK(cudaHostRegister(h_out, n, cudaHostRegisterDefault));
        }
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a, s1)); CK(cudaStreamWaitEvent(s2, a, 0));
            for (int i = 0; i < iters; i++) {
                CK(cudaMemcpyAsync(d_in, h_in, n, cudaMemcpyHostToDevice, s1));
                CK(cudaMemcpyAsync(h_out, d_out, n, cudaMemcpyDeviceToHost, s2));
            }
            CK(cudaEventRecord(b, s1)); CK(cudaEventRecord(c, s2));
            CK(cudaEventSynchronize(b)); CK(cudaEventSynchronize(c));
            float m1, m2; CK(cudaEventElapsedTime(&m1, a, b)); CK(cudaEventElapsedTime(&m2, a, c));
            if (rep) printf("%s: bidir H2D %.2f GB/s, D2H %.2f GB/s\n", mode ? "THP+cudaHostRegister" : "cudaHostAlloc      ", n * (double)iters / m1 / 1e6, n * (double)iters / m2 / 1e6);
        }
    }
    FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r"); char buf[128] = "";
    if (f) { fgets(buf, sizeof buf, f); fclose(f); } printf("THP enabled: %s", buf);
    f = fopen("/proc/meminfo", "r"); while (f && fgets(buf, sizeof buf, f)) if (strstr(buf, "AnonHugePages") || strstr(buf, "HugePages_Total")) printf("%s", buf); if (f) fclose(f);
    return 0;
}
