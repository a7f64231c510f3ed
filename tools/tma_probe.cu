/* tools/tma_probe.cu -- which tensor-map accesses does the TMA unit accept?  (byte-granular box starts, clipped stores) */
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include "tma_probe_kernels.cuh"
using namespace jmc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("  %s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
static PFN_cuTensorMapEncodeTiled_v12000 enc;
static bool mk(CUtensorMap *m, void *base, int rank, std::vector<cuuint64_t> dims, std::vector<cuuint64_t> strides, std::vector<cuuint32_t> box)
{
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, base, dims.data(), strides.data(), box.data(), es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("  encode failed: %d\n", (int)r);
    return r == CUDA_SUCCESS;
}
__global__ void k_load2d(const __grid_constant__ CUtensorMap tm, int x, int y, uint8_t *out, int nbytes)
{
    __shared__ __align__(128) uint8_t s[4096];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, nbytes); tma_load_2d(s, &tm, x, y, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) out[i] = s[i];
}
__global__ void k_store3d(const __grid_constant__ CUtensorMap tm, int x, int y, int z, int nbytes)
{
    __shared__ __align__(128) uint8_t s[8192];
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) s[i] = (uint8_t)(i * 7 + 1);
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) { tma_store_3d(&tm, x, y, z, s); bulk_commit_wait_read(); }
}
int main()
{
    void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    const size_t N = 1 << 20;
    uint8_t *d, *o; CK(cudaMalloc(&d, N)); CK(cudaMalloc(&o, 8192));
    std::vector<uint8_t> h(N); for (size_t i = 0; i < N; i++) h[i] = (uint8_t)(i * 13 + (i >> 8));
    CK(cudaMemcpy(d, h.data(), N, cudaMemcpyHostToDevice));
    std::vector<uint8_t> got(8192);
    struct { const char *name; cuuint64_t d0, d1, s0; cuuint32_t b0, b1; int x, y; } L[] = {
        {"2d u8 box 256x1 x=0", 614880, 4, 614880, 256, 1, 0, 0}, {"2d u8 box 256x1 x=256 y=1", 614880, 4, 614880, 256, 1, 256, 1},
        {"2d u8 box 256x1 x=854 (unaligned)", 614880, 4, 614880, 256, 1, 854, 0}, {"2d u8 box 256x1 x=3 (unaligned)", 614880, 4, 614880, 256, 1, 3, 2},
        {"2d u8 box 128x1 x=5", 614880, 4, 614880, 128, 1, 5, 0}, {"2d u8 box 64x1 x=5", 614880, 4, 614880, 64, 1, 5, 0},
        {"2d u8 box 16x1 x=5", 614880, 4, 614880, 16, 1, 5, 0}, {"2d u8 box 256x4 x=0", 1024, 64, 1024, 256, 4, 0, 0}, {"2d u8 box 256x4 x=7", 1024, 64, 1024, 256, 4, 7, 1},
    };
    for (auto &t : L) {
        printf("%s\n", t.name);
        CUtensorMap m;
        if (!mk(&m, d, 2, {t.d0, t.d1}, {t.s0}, {t.b0, t.b1})) continue;
        const int nb = (int)(t.b0 * t.b1);
        k_load2d<<<1, 128>>>(m, t.x, t.y, o, nb);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  kernel -> %s\n", cudaGetErrorString(e)); return 2; }
        CK(cudaMemcpy(got.data(), o, nb, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (cuuint32_t r = 0; r < t.b1; r++) for (cuuint32_t c = 0; c < t.b0; c++) {
            const size_t src = (size_t)(t.y + r) * t.s0 + t.x + c;
            const uint8_t want = (t.x + c < t.d0 && t.y + r < t.d1) ? h[src] : 0;
            bad += got[r * t.b0 + c] != want;
        }
        printf("  ok, mismatches %d\n", bad);
    }
    printf("3d u8 store box 256x8x1 into {854, 20, 2} pitch 1024, x=768 (clipped at 854), y=16 (clipped at 20)\n");
    {
        CUtensorMap m;
        CK(cudaMemset(d, 0xCD, N));
        if (mk(&m, d, 3, {854, 20, 2}, {1024, 1024 * 32}, {256, 8, 1})) {
            k_store3d<<<1, 128>>>(m, 768, 16, 1, 2048);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  kernel -> %s\n", cudaGetErrorString(e)); return 2; }
            CK(cudaMemcpy(h.data(), d, N, cudaMemcpyDeviceToHost));
            int bad = 0, written = 0;
            for (size_t i = 0; i < N; i++) {
                const size_t z = i / (1024 * 32), rem = i % (1024 * 32), yy = rem / 1024, xx = rem % 1024;
                const bool in = z == 1 && yy >= 16 && yy < 20 && xx >= 768 && xx < 854;
                const uint8_t want = in ? (uint8_t)((((yy - 16) * 256 + (xx - 768)) * 7 + 1)) : 0xCD;
                bad += h[i] != want; written += h[i] != 0xCD;
            }
            printf("  ok, mismatches %d, bytes written %d (expected %d)\n", bad, written, 4 * 86);
        }
    }
    return 0;
}
