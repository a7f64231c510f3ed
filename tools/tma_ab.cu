/*
 * tools/tma_ab.cu -- the A/B SURVEY.md section 7 asks for: tensor-map TMA (cuTensorMapEncodeTiled + cp.async.bulk.tensor.3d)
 * against the product's 1-D bulk-copy kernel (bulk_planes_kernel, cp.async.bulk per row), NV12 -> I420 and
 * I420 -> NV12, at 1 / 8 / 64 / 300 frames per launch, plus the launch-latency knobs for the small-batch regime
 * (programmatic dependent launch, CUDA graph of launches).  Developer tool, not part of the product library.
 *
 * Tensor-map variant: the whole stride-mode batch is ONE 3-D tensor per plane, (x in 8-byte elements, row, frame):
 *   surface luma   {w/8,   h,   n} strides {pitch, surf_stride}        tight luma {w/8,  h,   n} strides {w,   tight_stride}
 *   surface chroma {w/8,   h/2, n} (interleaved UV rows)               tight U, V {w/16, h/2, n} strides {w/2, tight_stride}
 * A tile is a box of <= 256 elements x 8 rows x 1 frame.  One CTA per tile, one elected thread:
 *   decode  luma: 1 tensor load  -> 1 tensor store            (product: 8 row loads + 1 bulk store)
 *           chroma: 1 tensor load -> prmt in smem -> 2 tensor stores   (product: 8 row loads + 2 bulk stores)
 *   encode  luma: 1 tensor load  -> 1 tensor store            (product: 1 bulk load + 8 row stores)
 *           chroma: 2 tensor loads -> prmt -> 1 tensor store  (product: 2 bulk loads + 8 row stores)
 * Rows / columns past the plane are clipped by the TMA unit itself.
 *
 *   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I jmcodec_b200/csrc -I include tools/tma_ab.cu -o tools/tma_ab
 *   tools/tma_ab            one CSV line per (geometry, frames per launch, op, variant)
 */
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "jmc_kernels.cuh"

using namespace jmc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

/* ---- tensor-map kernels ---------------------------------------------------------------------------- */
__device__ __forceinline__ void tma_load_3d(void *smem, const CUtensorMap *tm, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem)),
                 "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, int x, int y, int z, const void *smem)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(x), "r"(y), "r"(z),
                 "r"(smem_u32(smem)) : "memory");
}
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct TmParams {
    uint32_t n_frames;
    uint32_t box_x;          /* luma box width in 8-byte elements (<= 256) */
    uint32_t boxes_x;        /* boxes per luma row */
    uint32_t tiles_y[2];     /* 8-row tiles of luma / chroma */
    uint32_t rows;           /* 8 */
    uint32_t pdl;            /* issue griddepcontrol.launch_dependents first */
};

constexpr int TM_THREADS = 128;

/* TO_TIGHT: surface -> tight I420; else tight I420 -> surface.  Tensor maps: surface Y, surface UV, tight Y, tight U, tight V. */
template <bool TO_TIGHT>
__global__ void __launch_bounds__(TM_THREADS) tmap_planes_kernel(const __grid_constant__ CUtensorMap sy, const __grid_constant__ CUtensorMap suv,
                                                                const __grid_constant__ CUtensorMap ty, const __grid_constant__ CUtensorMap tu,
                                                                const __grid_constant__ CUtensorMap tv, const TmParams p)
{
    extern __shared__ __align__(128) uint8_t tm_smem[];
    __shared__ __align__(8) uint64_t bar;
    if (p.pdl) griddep_launch_dependents();
    const uint32_t per_frame = (p.tiles_y[0] + p.tiles_y[1]) * p.boxes_x;
    const uint32_t f = blockIdx.x / per_frame;
    uint32_t r = blockIdx.x - f * per_frame;
    const bool chroma = r >= p.tiles_y[0] * p.boxes_x;
    if (chroma) r -= p.tiles_y[0] * p.boxes_x;
    const uint32_t ty_i = r / p.boxes_x, bx = r - ty_i * p.boxes_x;
    const int x = (int)(bx * p.box_x), y = (int)(ty_i * p.rows);
    const uint32_t box_bytes = p.box_x * 8 * p.rows;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (!chroma) {
        if (threadIdx.x != 0) return;
        mbar_expect_tx(&bar, box_bytes);
        tma_load_3d(tm_smem, TO_TIGHT ? &sy : &ty, x, y, (int)f, &bar);
        mbar_wait(&bar, 0);
        tma_store_3d(TO_TIGHT ? &ty : &sy, x, y, (int)f, tm_smem);
        bulk_commit_wait_read();
        return;
    }
    /* chroma: the interleaved box is box_x elements wide, each planar box box_x/2 */
    uint8_t *s_uv = tm_smem, *s_u = tm_smem + box_bytes, *s_v = s_u + box_bytes / 2;
    const uint32_t nvec = box_bytes / 32;
    if (TO_TIGHT) {
        if (threadIdx.x == 0) { mbar_expect_tx(&bar, box_bytes); tma_load_3d(s_uv, &suv, x, y, (int)f, &bar); }
        mbar_wait(&bar, 0);
        for (uint32_t v = threadIdx.x; v < nvec; v += TM_THREADS) {
            const uint4 a = *(const uint4 *)(s_uv + (size_t)v * 32), b = *(const uint4 *)(s_uv + (size_t)v * 32 + 16);
            uint4 u, w;
            u.x = __byte_perm(a.x, a.y, 0x6420); w.x = __byte_perm(a.x, a.y, 0x7531);
            u.y = __byte_perm(a.z, a.w, 0x6420); w.y = __byte_perm(a.z, a.w, 0x7531);
            u.z = __byte_perm(b.x, b.y, 0x6420); w.z = __byte_perm(b.x, b.y, 0x7531);
            u.w = __byte_perm(b.z, b.w, 0x6420); w.w = __byte_perm(b.z, b.w, 0x7531);
            *(uint4 *)(s_u + (size_t)v * 16) = u;
            *(uint4 *)(s_v + (size_t)v * 16) = w;
        }
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            tma_store_3d(&tu, x / 2, y, (int)f, s_u);
            tma_store_3d(&tv, x / 2, y, (int)f, s_v);
            bulk_commit_wait_read();
        }
    } else {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, box_bytes);
            tma_load_3d(s_u, &tu, x / 2, y, (int)f, &bar);
            tma_load_3d(s_v, &tv, x / 2, y, (int)f, &bar);
        }
        mbar_wait(&bar, 0);
        for (uint32_t v = threadIdx.x; v < nvec; v += TM_THREADS) {
            const uint4 u = *(const uint4 *)(s_u + (size_t)v * 16), w = *(const uint4 *)(s_v + (size_t)v * 16);
            uint4 a, b;
            a.x = __byte_perm(u.x, w.x, 0x5140); a.y = __byte_perm(u.x, w.x, 0x7362);
            a.z = __byte_perm(u.y, w.y, 0x5140); a.w = __byte_perm(u.y, w.y, 0x7362);
            b.x = __byte_perm(u.z, w.z, 0x5140); b.y = __byte_perm(u.z, w.z, 0x7362);
            b.z = __byte_perm(u.w, w.w, 0x5140); b.w = __byte_perm(u.w, w.w, 0x7362);
            *(uint4 *)(s_uv + (size_t)v * 32) = a;
            *(uint4 *)(s_uv + (size_t)v * 32 + 16) = b;
        }
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) { tma_store_3d(&suv, x, y, (int)f, s_uv); bulk_commit_wait_read(); }
    }
}

/* product kernel with the PDL trigger in front (same body otherwise) -- only for the launch-latency experiment */
template <bool TO_TIGHT, int KIND1>
__global__ void __launch_bounds__(BULK_THREADS) bulk_planes_pdl_kernel(const __grid_constant__ BulkParams p)
{
    griddep_launch_dependents();
    extern __shared__ __align__(128) uint8_t bulk_smem2[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t tpf = p.tiles[0] + p.tiles[1];
    const uint32_t f = blockIdx.x / tpf;
    uint32_t r = blockIdx.x - f * tpf;
    const bool second = r >= p.tiles[0];
    if (second) r -= p.tiles[0];
    const Part &pt = second ? p.part[1] : p.part[0];
    uint8_t *pp = frame_ptr(p.pitched, f) + pt.p_off;
    uint8_t *tp = frame_ptr(p.tight, f);
    const uint32_t r0 = r * p.rows_per_tile;
    const uint32_t nr = min(p.rows_per_tile, pt.rows - r0);
    const uint32_t re = pt.row_elems;
    const size_t pitch = (size_t)pt.p_pitch;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (!second) {
        if (threadIdx.x != 0) return;
        uint8_t *t = tp + pt.a_off + (size_t)r0 * re;
        mbar_expect_tx(&bar, nr * re);
        for (uint32_t i = 0; i < nr; i++) bulk_g2s(bulk_smem2 + (size_t)i * re, pp + (size_t)(r0 + i) * pitch, re, &bar);
        mbar_wait(&bar, 0);
        bulk_s2g(t, bulk_smem2, nr * re);
        bulk_commit_wait_read();
    } else {
        uint8_t *s_uv = bulk_smem2;
        uint8_t *s_u = bulk_smem2 + (size_t)p.rows_per_tile * 2 * re;
        uint8_t *s_v = s_u + (size_t)p.rows_per_tile * re;
        uint8_t *tu = tp + pt.a_off + (size_t)r0 * re, *tv = tp + pt.b_off + (size_t)r0 * re;
        const uint32_t nvec = nr * re / 16;
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, nr * 2 * re);
            for (uint32_t i = 0; i < nr; i++) bulk_g2s(s_uv + (size_t)i * 2 * re, pp + (size_t)(r0 + i) * pitch, 2 * re, &bar);
        }
        mbar_wait(&bar, 0);
        for (uint32_t v = threadIdx.x; v < nvec; v += BULK_THREADS) {
            const uint4 a = *(const uint4 *)(s_uv + (size_t)v * 32), b = *(const uint4 *)(s_uv + (size_t)v * 32 + 16);
            uint4 u, w;
            u.x = __byte_perm(a.x, a.y, 0x6420); w.x = __byte_perm(a.x, a.y, 0x7531);
            u.y = __byte_perm(a.z, a.w, 0x6420); w.y = __byte_perm(a.z, a.w, 0x7531);
            u.z = __byte_perm(b.x, b.y, 0x6420); w.z = __byte_perm(b.x, b.y, 0x7531);
            u.w = __byte_perm(b.z, b.w, 0x6420); w.w = __byte_perm(b.z, b.w, 0x7531);
            *(uint4 *)(s_u + (size_t)v * 16) = u;
            *(uint4 *)(s_v + (size_t)v * 16) = w;
        }
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) { bulk_s2g(tu, s_u, nr * re); bulk_s2g(tv, s_v, nr * re); bulk_commit_wait_read(); }
    }
}

/* ---- host ---------------------------------------------------------------------------------------------- */
static PFN_cuTensorMapEncodeTiled_v12000 g_encode;

static CUtensorMap make_map(void *base, uint64_t w_elems, uint64_t rows, uint64_t frames, uint64_t row_stride, uint64_t frame_stride,
                            uint32_t box_x, uint32_t box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[3] = { w_elems, rows, frames };
    cuuint64_t strides[2] = { row_stride, frame_stride };
    cuuint32_t box[3] = { box_x, box_rows, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d (w %llu rows %llu frames %llu)\n", (int)r, (unsigned long long)w_elems,
                                     (unsigned long long)rows, (unsigned long long)frames); exit(1); }
    return m;
}

struct Geom { const char *name; int w, h, pitch; int counts[4]; };
static const Geom GEOMS[] = { { "1080p_p2048", 1920, 1080, 2048, { 1, 8, 64, 300 } }, { "4k_p4096", 3840, 2160, 4096, { 1, 8, 64, 0 } } };

static Part mk_part(int kind, uint32_t rows, uint32_t row_elems, int64_t p_off, int32_t pitch, int64_t a_off, int64_t b_off)
{
    Part p;
    memset(&p, 0, sizeof(p));
    p.kind = kind; p.rows = rows; p.row_elems = row_elems;
    p.p_off = p_off; p.p_pitch = pitch; p.a_off = a_off; p.b_off = b_off;
    return p;
}

template <class F> static float time_it(F launch, int iters, cudaStream_t st)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 5; i++) launch();
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < iters; i++) launch();
    CK(cudaEventRecord(b, st));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms / iters;
}

int main(int argc, char **argv)
{
    const bool verify = argc > 1 && !strcmp(argv[1], "--verify");
    CK(cudaSetDevice(0));
    cudaDriverEntryPointQueryResult qr;
    void *fn = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) { fprintf(stderr, "cuTensorMapEncodeTiled not available\n"); return 1; }
    g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));

    printf("geometry,frames_per_launch,op,variant,us_per_launch,GB/s,frac_of_6555.5\n");
    for (const Geom &g : GEOMS) {
        const size_t surf = (size_t)g.pitch * g.h * 3 / 2, tight = (size_t)g.w * g.h * 3 / 2;
        int nmax = 0;
        for (int c : g.counts) if (c > nmax) nmax = c;
        uint8_t *d_surf, *d_tight, *d_back, *d_tight2;
        CK(cudaMalloc(&d_surf, surf * nmax)); CK(cudaMalloc(&d_tight, tight * nmax));
        CK(cudaMalloc(&d_back, surf * nmax)); CK(cudaMalloc(&d_tight2, tight * nmax));
        {
            std::vector<uint8_t> h(surf * (size_t)(nmax < 8 ? nmax : 8));
            unsigned x = 12345u;
            for (auto &b : h) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
            for (int f = 0; f < nmax; f += 8) CK(cudaMemcpy(d_surf + (size_t)f * surf, h.data(), surf * (size_t)((nmax - f) < 8 ? (nmax - f) : 8), cudaMemcpyHostToDevice));
        }
        const int64_t u_off = (int64_t)g.w * g.h, v_off = u_off + (int64_t)(g.w / 2) * (g.h / 2);
        const uint32_t w8 = (uint32_t)g.w / 8;
        uint32_t boxes_x = (w8 + 255) / 256, box_x = w8 / boxes_x;
        while (w8 % boxes_x || (box_x & 1)) { boxes_x++; box_x = w8 / boxes_x; }      /* equal, even-width boxes */
        for (int n : g.counts) {
            if (!n) continue;
            const double bytes = 3.0 * g.w * g.h * n;
            const int iters = n >= 64 ? 20 : 200;
            for (int dir = 0; dir < 2; dir++) {            /* 0: NV12 -> I420, 1: I420 -> NV12 */
                uint8_t *surf_p = dir == 0 ? d_surf : d_back, *tight_p = d_tight;
                /* --- A: product kernel --- */
                BulkParams b;
                memset(&b, 0, sizeof(b));
                b.pitched.base = surf_p; b.pitched.stride = surf;
                b.tight.base = tight_p; b.tight.stride = tight;
                b.n_frames = (uint32_t)n;
                b.rows_per_tile = 8;
                b.part[0] = mk_part(PART_COPY, g.h, g.w, 0, g.pitch, 0, 0);
                b.part[1] = mk_part(dir == 0 ? PART_SPLIT : PART_MERGE, g.h / 2, g.w / 2, (int64_t)g.pitch * g.h, g.pitch, u_off, v_off);
                b.tiles[0] = (g.h + 7) / 8; b.tiles[1] = (g.h / 2 + 7) / 8;
                const uint32_t gridA = (b.tiles[0] + b.tiles[1]) * n;
                const size_t smemA = (size_t)8 * 4 * (g.w / 2);
                CK(cudaFuncSetAttribute(bulk_planes_kernel<true, PART_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                CK(cudaFuncSetAttribute(bulk_planes_kernel<false, PART_MERGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                CK(cudaFuncSetAttribute(bulk_planes_pdl_kernel<true, PART_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                auto launchA = [&]() {
                    if (dir == 0) bulk_planes_kernel<true, PART_SPLIT><<<gridA, BULK_THREADS, smemA, st>>>(b);
                    else bulk_planes_kernel<false, PART_MERGE><<<gridA, BULK_THREADS, smemA, st>>>(b);
                };
                float ms = time_it(launchA, iters, st);
                printf("%s,%d,%s,bulk_1d(product),%.2f,%.1f,%.3f\n", g.name, n, dir == 0 ? "nv12_to_i420" : "i420_to_nv12", ms * 1e3, bytes / ms / 1e6, bytes / ms / 1e6 / 6555.5);
                /* --- B: tensor maps --- */
                CUtensorMap sy = make_map(surf_p, w8, g.h, n, g.pitch, surf, box_x, 8);
                CUtensorMap suv = make_map(surf_p + (size_t)g.pitch * g.h, w8, g.h / 2, n, g.pitch, surf, box_x, 8);
                uint8_t *tp = dir == 0 ? d_tight2 : d_tight;
                CUtensorMap tyy = make_map(tp, w8, g.h, n, g.w, tight, box_x, 8);
                CUtensorMap tu = make_map(tp + u_off, w8 / 2, g.h / 2, n, g.w / 2, tight, box_x / 2, 8);
                CUtensorMap tv = make_map(tp + v_off, w8 / 2, g.h / 2, n, g.w / 2, tight, box_x / 2, 8);
                TmParams p;
                p.n_frames = (uint32_t)n; p.box_x = box_x; p.boxes_x = boxes_x; p.rows = 8; p.pdl = 0;
                p.tiles_y[0] = (g.h + 7) / 8; p.tiles_y[1] = (g.h / 2 + 7) / 8;
                const uint32_t gridB = (p.tiles_y[0] + p.tiles_y[1]) * boxes_x * n;
                const size_t smemB = (size_t)box_x * 8 * 8 * 2;
                CK(cudaFuncSetAttribute(tmap_planes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                CK(cudaFuncSetAttribute(tmap_planes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                auto launchB = [&]() {
                    if (dir == 0) tmap_planes_kernel<true><<<gridB, TM_THREADS, smemB, st>>>(sy, suv, tyy, tu, tv, p);
                    else tmap_planes_kernel<false><<<gridB, TM_THREADS, smemB, st>>>(sy, suv, tyy, tu, tv, p);
                };
                if (dir == 0) {
                    CK(cudaMemsetAsync(d_tight, 0xA5, tight * n, st)); CK(cudaMemsetAsync(d_tight2, 0x5A, tight * n, st));
                    launchA(); launchB();
                    CK(cudaStreamSynchronize(st));
                    if (verify || n <= 8) {                 /* B must reproduce A byte for byte */
                        std::vector<uint8_t> ha(tight * n), hb(tight * n);
                        CK(cudaMemcpy(ha.data(), d_tight, tight * n, cudaMemcpyDeviceToHost));
                        CK(cudaMemcpy(hb.data(), d_tight2, tight * n, cudaMemcpyDeviceToHost));
                        if (memcmp(ha.data(), hb.data(), tight * n) != 0) { fprintf(stderr, "MISMATCH %s n=%d decode\n", g.name, n); return 2; }
                    }
                } else if (verify || n <= 8) {              /* pack of d_tight (A's I420) must give back the surfaces' active bytes */
                    CK(cudaMemsetAsync(d_back, 0xCD, surf * n, st));
                    launchB();
                    CK(cudaStreamSynchronize(st));
                    std::vector<uint8_t> hs(surf), hb(surf);
                    CK(cudaMemcpy(hs.data(), d_surf, surf, cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(hb.data(), d_back, surf, cudaMemcpyDeviceToHost));
                    for (int y = 0; y < g.h * 3 / 2; y++) {
                        if (memcmp(hs.data() + (size_t)y * g.pitch, hb.data() + (size_t)y * g.pitch, (size_t)g.w) != 0) { fprintf(stderr, "MISMATCH %s n=%d encode row %d\n", g.name, n, y); return 2; }
                        for (int x2 = g.w; x2 < g.pitch; x2++) if (hb[(size_t)y * g.pitch + x2] != 0xCD) { fprintf(stderr, "PADDING WRITTEN %s row %d\n", g.name, y); return 2; }
                    }
                }
                ms = time_it(launchB, iters, st);
                printf("%s,%d,%s,tensor_map_3d,%.2f,%.1f,%.3f\n", g.name, n, dir == 0 ? "nv12_to_i420" : "i420_to_nv12", ms * 1e3, bytes / ms / 1e6, bytes / ms / 1e6 / 6555.5);
                if (dir == 0 && n <= 8) {
                    /* --- launch-latency knobs for small batches: PDL, CUDA graph --- */
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof(cfg));
                    cfg.gridDim = dim3(gridA); cfg.blockDim = dim3(BULK_THREADS); cfg.dynamicSmemBytes = smemA; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
                    auto launchPdl = [&]() { CK(cudaLaunchKernelEx(&cfg, bulk_planes_pdl_kernel<true, PART_SPLIT>, b)); };
                    ms = time_it(launchPdl, iters, st);
                    printf("%s,%d,nv12_to_i420,bulk_1d+PDL,%.2f,%.1f,%.3f\n", g.name, n, ms * 1e3, bytes / ms / 1e6, bytes / ms / 1e6 / 6555.5);
                    TmParams pp = p; pp.pdl = 1;
                    cfg.gridDim = dim3(gridB); cfg.blockDim = dim3(TM_THREADS); cfg.dynamicSmemBytes = smemB;
                    auto launchPdlB = [&]() { CK(cudaLaunchKernelEx(&cfg, tmap_planes_kernel<true>, sy, suv, tyy, tu, tv, pp)); };
                    ms = time_it(launchPdlB, iters, st);
                    printf("%s,%d,nv12_to_i420,tensor_map_3d+PDL,%.2f,%.1f,%.3f\n", g.name, n, ms * 1e3, bytes / ms / 1e6, bytes / ms / 1e6 / 6555.5);
                    /* graph of 50 product launches */
                    cudaGraph_t graph; cudaGraphExec_t exec;
                    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                    for (int i = 0; i < 50; i++) launchA();
                    CK(cudaStreamEndCapture(st, &graph));
                    CK(cudaGraphInstantiate(&exec, graph, 0));
                    auto launchG = [&]() { CK(cudaGraphLaunch(exec, st)); };
                    ms = time_it(launchG, 10, st) / 50;
                    printf("%s,%d,nv12_to_i420,bulk_1d_in_graph_of_50,%.2f,%.1f,%.3f\n", g.name, n, ms * 1e3, bytes / ms / 1e6, bytes / ms / 1e6 / 6555.5);
                    cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
                }
                fflush(stdout);
            }
        }
        cudaFree(d_surf); cudaFree(d_tight); cudaFree(d_back); cudaFree(d_tight2);
    }
    return 0;
}
