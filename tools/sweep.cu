/*
 * tools/sweep.cu -- tuning harness (not part of the product library).
 *
 * Instantiates the kernels of jmcodec_b200/csrc/jmc_kernels.cuh under many configurations
 * (threads per CTA, 16-byte vectors in flight per thread, CTAs per SM, load/store cache policy)
 * and times each on the BASELINE.json geometries with CUDA events, so that one gpurun call
 * answers "which configuration is closest to the HBM roofline".  Output: one CSV line per variant.
 *
 *   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I jmcodec_b200/csrc -I include \
 *        tools/sweep.cu -o tools/sweep && tools/sweep
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "jmc_kernels.cuh"

using namespace jmc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int T, int U, int L, int S, int B> struct Cfg {
    static constexpr int THREADS = T, UNROLL = U, LDP = L, STP = S, BLOCKS_PER_SM = B;
};

struct Geom { const char *name; int w, h, pitch, n; };
static const Geom GEOMS[] = { {"1080p_x300_p2048", 1920, 1080, 2048, 300}, {"4k_x64_p4096", 3840, 2160, 4096, 64}, {"4k_x64_p3840", 3840, 2160, 3840, 64} };

static int g_sms = 148;
static uint8_t *g_a, *g_b, *g_c;      /* big device buffers */

template <class C> Part mk_part(int kind, uint32_t rows, uint32_t row_elems, int64_t p_off, int32_t pitch, int64_t a_off, int64_t b_off)
{
    Part p;
    memset(&p, 0, sizeof(p));
    p.kind = kind; p.rows = rows; p.row_elems = row_elems;
    const uint64_t total = (uint64_t)rows * row_elems;
    p.tiles = (uint32_t)((total + TileGeom<C>::TILE_ELEMS - 1) / TileGeom<C>::TILE_ELEMS);
    p.p_off = p_off; p.p_pitch = pitch; p.a_off = a_off; p.b_off = b_off;
    uint32_t d = row_elems, s = 0;
    while ((1ull << s) < d) s++;
    p.rdiv.d = d; p.rdiv.sh = 31 + s; p.rdiv.m = (uint32_t)(((1ull << (31 + s)) + d - 1) / d);
    return p;
}

static float time_launches(void (*launch)(void *), void *arg, int iters)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; i++) launch(arg);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; i++) launch(arg);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms / iters;
}

struct PlaneRun { PlaneParams p; uint32_t grid; };

template <class C, bool TT, int K1> static void launch_planes(void *arg)
{
    PlaneRun *r = (PlaneRun *)arg;
    planes_kernel<C, TT, K1, true><<<r->grid, C::THREADS>>>(r->p);
}

/* op: 0 = NV12->I420 (split), 1 = I420->NV12 (merge), 2 = NV12->NV12 (copy) */
template <class C> static void run_planes(const char *tag, int op, int grid_mult_override = 0)
{
    for (const Geom &g : GEOMS) {
        PlaneRun r;
        memset(&r, 0, sizeof(r));
        const size_t surf = (size_t)g.pitch * g.h * 3 / 2, tight = (size_t)g.w * g.h * 3 / 2;
        r.p.pitched = { op == 1 ? g_b : g_a, surf, nullptr };
        r.p.tight = { op == 1 ? g_a : g_b, tight, nullptr };
        r.p.n_frames = g.n;
        r.p.to_tight = op != 1;
        r.p.part[0] = mk_part<C>(PART_COPY, g.h, g.w, 0, g.pitch, 0, 0);
        if (op == 2) r.p.part[1] = mk_part<C>(PART_COPY, g.h / 2, g.w, (int64_t)g.pitch * g.h, g.pitch, (int64_t)g.w * g.h, 0);
        else r.p.part[1] = mk_part<C>(op == 0 ? PART_SPLIT : PART_MERGE, g.h / 2, g.w / 2, (int64_t)g.pitch * g.h, g.pitch,
                                      (int64_t)g.w * g.h, (int64_t)g.w * g.h * 5 / 4);
        r.p.tiles_per_frame = r.p.part[0].tiles + r.p.part[1].tiles;
        r.p.total_tiles = r.p.tiles_per_frame * g.n;
        const int mult = grid_mult_override ? grid_mult_override : C::BLOCKS_PER_SM;
        r.grid = (uint32_t)g_sms * mult;
        if (r.grid > r.p.total_tiles) r.grid = r.p.total_tiles;
        float ms = op == 0 ? time_launches(launch_planes<C, true, PART_SPLIT>, &r, 20)
                 : op == 1 ? time_launches(launch_planes<C, false, PART_MERGE>, &r, 20)
                           : time_launches(launch_planes<C, true, PART_COPY>, &r, 20);
        const double bytes = 3.0 * g.w * g.h * g.n;
        printf("planes,%s,%s,op%d,T%d,U%d,L%d,S%d,B%d,grid%u,%.4f ms,%.1f GB/s\n", tag, g.name, op, C::THREADS, C::UNROLL, C::LDP,
               C::STP, C::BLOCKS_PER_SM, r.grid, ms, bytes / ms / 1e6);
        fflush(stdout);
    }
}

struct RgbRun { RgbParams p; uint32_t grid; };
template <class C> static void launch_rgb(void *arg)
{
    RgbRun *r = (RgbRun *)arg;
    rgb_kernel<C, false><<<r->grid, C::THREADS>>>(r->p);
}

template <int T, int B, int L, int S> struct RCfg { static constexpr int THREADS = T, BLOCKS_PER_SM = B, LDP = L, STP = S; };

template <class C> static void run_rgb(const char *tag, int fused, bool grid_all = false)
{
    for (const Geom &g : GEOMS) {
        RgbRun r;
        memset(&r, 0, sizeof(r));
        const size_t surf = (size_t)g.pitch * g.h * 3 / 2, tight = (size_t)g.w * g.h * 3 / 2;
        r.p.surf = { g_a, surf, nullptr };
        r.p.tight = { g_b, tight, nullptr };
        r.p.rgb = { g_c, (size_t)3 * g.w * g.h, nullptr };
        r.p.n_frames = g.n; r.p.width = g.w; r.p.height = g.h; r.p.pitch = g.pitch;
        r.p.y_off = 0; r.p.uv_off = (int64_t)g.pitch * g.h;
        r.p.u_off = (int64_t)g.w * g.h; r.p.v_off = (int64_t)g.w * g.h * 5 / 4;
        r.p.rgb_pitch = 3 * g.w; r.p.fused = fused;
        r.p.segs_per_row = (g.w + 511) / 512; r.p.row_pairs = (g.h + 1) / 2;
        r.p.tasks_per_frame = r.p.segs_per_row * r.p.row_pairs;
        r.p.total_tasks = r.p.tasks_per_frame * g.n;
        r.grid = (uint32_t)g_sms * C::BLOCKS_PER_SM;
        if (grid_all) r.grid = (r.p.total_tasks + C::THREADS / 32 - 1) / (C::THREADS / 32);
        float ms = time_launches(launch_rgb<C>, &r, 20);
        const double bytes = (fused ? 6.0 : 4.5) * g.w * g.h * g.n;
        printf("rgb,%s,%s,fused%d,T%d,B%d,L%d,S%d,grid%u,%.4f ms,%.1f GB/s\n", tag, g.name, fused, C::THREADS, C::BLOCKS_PER_SM, C::LDP, C::STP,
               r.grid, ms, bytes / ms / 1e6);
        fflush(stdout);
    }
}

/* reference points: cudaMemcpyAsync D2D and a plain grid-stride uint4 copy */
__global__ void plain_copy(const uint4 *__restrict__ s, uint4 *__restrict__ d, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}


/* ------------------------------------------------------------------------------------------------
 * Experiment: the same NV12 -> I420 conversion with the bulk-copy engine (cp.async.bulk, "1-D TMA").
 * Luma tile: ROWS row loads (pitch -> contiguous smem) + ONE contiguous bulk store.  Chroma tile:
 * row loads, threads de-interleave smem -> smem with prmt, two bulk stores.  One CTA per tile.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void *dst, const void *src_smem, uint32_t bytes, uint64_t pol)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol) : "memory");
}

struct XBulkParams {
    const uint8_t *src; size_t src_stride; uint8_t *dst; size_t dst_stride;
    int w, h, pitch, n_frames, rows_y, rows_c;
    uint32_t tiles_y, tiles_c;
};

/* HINT bit 0: evict_first on loads, bit 1: evict_first on stores */
template <int THREADS, int HINT = 0>
__global__ void __launch_bounds__(THREADS) bulk_i420_kernel(const __grid_constant__ XBulkParams p)
{
    const uint64_t pol = HINT ? policy_evict_first() : 0;
#define G2S(d, s_, n, b) do { if (HINT & 1) bulk_g2s_hint(d, s_, n, b, pol); else bulk_g2s(d, s_, n, b); } while (0)
#define S2G(d, s_, n) do { if (HINT & 2) bulk_s2g_hint(d, s_, n, pol); else bulk_s2g(d, s_, n); } while (0)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t tpf = p.tiles_y + p.tiles_c;
    const uint32_t f = blockIdx.x / tpf;
    uint32_t r = blockIdx.x - f * tpf;
    const uint8_t *sp = p.src + (size_t)f * p.src_stride;
    uint8_t *dp = p.dst + (size_t)f * p.dst_stride;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (r < p.tiles_y) {
        if (threadIdx.x != 0) return;
        const int r0 = r * p.rows_y, nr = min(p.rows_y, p.h - r0);
        mbar_expect_tx(&bar, (uint32_t)nr * p.w);
        for (int i = 0; i < nr; i++) G2S(smem + (size_t)i * p.w, sp + (size_t)(r0 + i) * p.pitch, p.w, &bar);
        mbar_wait(&bar, 0);
        S2G(dp + (size_t)r0 * p.w, smem, (uint32_t)nr * p.w);
        bulk_commit_wait_read();
    } else {
        r -= p.tiles_y;
        const int ch = p.h >> 1, cw = p.w >> 1;
        const int r0 = r * p.rows_c, nr = min(p.rows_c, ch - r0);
        uint8_t *s_uv = smem, *s_u = smem + (size_t)p.rows_c * p.w, *s_v = s_u + (size_t)p.rows_c * cw;
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, (uint32_t)nr * p.w);
            for (int i = 0; i < nr; i++) G2S(s_uv + (size_t)i * p.w, sp + (size_t)p.pitch * p.h + (size_t)(r0 + i) * p.pitch, p.w, &bar);
        }
        mbar_wait(&bar, 0);
        const int nvec = nr * cw / 16;                       /* 16 output bytes of U (and V) per step */
        for (int v = threadIdx.x; v < nvec; v += THREADS) {
            const uint4 a = *(const uint4 *)(s_uv + (size_t)v * 32), b = *(const uint4 *)(s_uv + (size_t)v * 32 + 16);
            uint4 u, w4;
            u.x = __byte_perm(a.x, a.y, 0x6420); w4.x = __byte_perm(a.x, a.y, 0x7531);
            u.y = __byte_perm(a.z, a.w, 0x6420); w4.y = __byte_perm(a.z, a.w, 0x7531);
            u.z = __byte_perm(b.x, b.y, 0x6420); w4.z = __byte_perm(b.x, b.y, 0x7531);
            u.w = __byte_perm(b.z, b.w, 0x6420); w4.w = __byte_perm(b.z, b.w, 0x7531);
            *(uint4 *)(s_u + (size_t)v * 16) = u;
            *(uint4 *)(s_v + (size_t)v * 16) = w4;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const size_t luma = (size_t)p.w * p.h;
            S2G(dp + luma + (size_t)r0 * cw, s_u, (uint32_t)nr * cw);
            S2G(dp + luma + (size_t)cw * ch + (size_t)r0 * cw, s_v, (uint32_t)nr * cw);
            bulk_commit_wait_read();
        }
    }
}
#undef G2S
#undef S2G

struct XBulkRun { XBulkParams p; uint32_t grid; size_t smem; };
template <int THREADS, int HINT> static void launch_bulk(void *arg)
{
    XBulkRun *r = (XBulkRun *)arg;
    bulk_i420_kernel<THREADS, HINT><<<r->grid, THREADS, r->smem>>>(r->p);
}

template <int THREADS, int HINT = 0> static void run_bulk(int rows_y, int rows_c, bool verify)
{
    for (const Geom &g : GEOMS) {
        XBulkRun r;
        memset(&r, 0, sizeof(r));
        const size_t surf = (size_t)g.pitch * g.h * 3 / 2, tight = (size_t)g.w * g.h * 3 / 2;
        r.p = { g_a, surf, g_b, tight, g.w, g.h, g.pitch, g.n, rows_y, rows_c,
                (uint32_t)((g.h + rows_y - 1) / rows_y), (uint32_t)((g.h / 2 + rows_c - 1) / rows_c) };
        r.grid = (r.p.tiles_y + r.p.tiles_c) * g.n;
        const size_t sy = (size_t)rows_y * g.w, sc = (size_t)rows_c * g.w * 2;
        r.smem = sy > sc ? sy : sc;
        CK(cudaFuncSetAttribute(bulk_i420_kernel<THREADS, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r.smem));
        float ms = time_launches(launch_bulk<THREADS, HINT>, &r, 20);
        const double bytes = 3.0 * g.w * g.h * g.n;
        printf("bulk,%s,T%d,hint%d,rowsY%d,rowsC%d,smem%zu,grid%u,%.4f ms,%.1f GB/s", g.name, THREADS, HINT, rows_y, rows_c, r.smem, r.grid, ms, bytes / ms / 1e6);
        if (verify) {
            /* compare with the plane kernel's output on frames 0 and n-1 */
            std::vector<uint8_t> a(tight), b(tight);
            int bad = 0;
            for (int f : {0, g.n - 1}) {
                CK(cudaMemcpy(a.data(), g_b + (size_t)f * tight, tight, cudaMemcpyDeviceToHost));
                PlaneRun pr;
                memset(&pr, 0, sizeof(pr));
                typedef Cfg<256, 4, 1, 1, 4> C;
                pr.p.pitched = { g_a, surf, nullptr };
                pr.p.tight = { g_c, tight, nullptr };
                pr.p.n_frames = g.n; pr.p.to_tight = 1;
                pr.p.part[0] = mk_part<C>(PART_COPY, g.h, g.w, 0, g.pitch, 0, 0);
                pr.p.part[1] = mk_part<C>(PART_SPLIT, g.h / 2, g.w / 2, (int64_t)g.pitch * g.h, g.pitch, (int64_t)g.w * g.h, (int64_t)g.w * g.h * 5 / 4);
                pr.p.tiles_per_frame = pr.p.part[0].tiles + pr.p.part[1].tiles;
                pr.p.total_tiles = pr.p.tiles_per_frame * g.n;
                pr.grid = pr.p.total_tiles;
                launch_planes<C, true, PART_SPLIT>(&pr);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(b.data(), g_c + (size_t)f * tight, tight, cudaMemcpyDeviceToHost));
                bad += memcmp(a.data(), b.data(), tight) != 0;
            }
            printf(",%s", bad ? "MISMATCH" : "bit-exact vs planes_kernel");
        }
        printf("\n");
        fflush(stdout);
    }
}

int main(int argc, char **argv)
{
    const char *only = argc > 1 ? argv[1] : "all";
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("# %s, %d SMs\n", prop.name, g_sms);
    const size_t big = (size_t)4096 * 2160 * 3 / 2 * 64 + (1 << 20);           /* >= 300 x 1080p surfaces too */
    const size_t big1080 = (size_t)2048 * 1080 * 3 / 2 * 300 + (1 << 20);
    const size_t sz = big > big1080 ? big : big1080;
    CK(cudaMalloc(&g_a, sz)); CK(cudaMalloc(&g_b, sz)); CK(cudaMalloc(&g_c, (size_t)3 * 1920 * 1080 * 300 + (1 << 20)));
    CK(cudaMemset(g_a, 0x5a, sz)); CK(cudaMemset(g_b, 0x3c, sz));

    if (!strcmp(only, "all") || !strcmp(only, "ref")) {
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        const size_t n = (size_t)768 << 20;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(a));
            for (int i = 0; i < 10; i++) CK(cudaMemcpyAsync(g_b, g_a, n, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            printf("ref,cudaMemcpyD2D,768MiB,%.4f ms,%.1f GB/s\n", ms / 10, 2.0 * n / (ms / 10) / 1e6);
        }
        for (int mult : {2, 4, 8, 16}) {
            CK(cudaEventRecord(a));
            for (int i = 0; i < 10; i++) plain_copy<<<g_sms * mult, 512>>>((const uint4 *)g_a, (uint4 *)g_b, n / 16);
            CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            printf("ref,plain_copy_512x%d,768MiB,%.4f ms,%.1f GB/s\n", mult, ms / 10, 2.0 * n / (ms / 10) / 1e6);
        }
    }
    if (!strcmp(only, "all") || !strcmp(only, "planes")) {
        /* one CTA per tile (grid = all tiles): threads x unroll x policies */
        const int ALL = 1 << 20;
        run_planes<Cfg<128, 1, 1, 0, 8>>("ga", 0, ALL);
        run_planes<Cfg<128, 2, 1, 0, 8>>("ga", 0, ALL);
        run_planes<Cfg<128, 4, 1, 0, 8>>("ga", 0, ALL);
        run_planes<Cfg<128, 8, 1, 0, 4>>("ga", 0, ALL);
        run_planes<Cfg<256, 1, 1, 0, 8>>("ga", 0, ALL);
        run_planes<Cfg<256, 2, 1, 0, 8>>("ga", 0, ALL);
        run_planes<Cfg<256, 2, 1, 0, 4>>("ga", 0, ALL);
        run_planes<Cfg<256, 4, 1, 0, 4>>("ga", 0, ALL);
        run_planes<Cfg<256, 4, 1, 0, 5>>("ga", 0, ALL);
        run_planes<Cfg<256, 4, 1, 0, 2>>("ga", 0, ALL);
        run_planes<Cfg<512, 1, 1, 0, 4>>("ga", 0, ALL);
        run_planes<Cfg<512, 2, 1, 0, 4>>("ga", 0, ALL);
        run_planes<Cfg<512, 4, 1, 0, 2>>("ga", 0, ALL);
        run_planes<Cfg<1024, 1, 1, 0, 2>>("ga", 0, ALL);
        run_planes<Cfg<1024, 2, 1, 0, 1>>("ga", 0, ALL);
        run_planes<Cfg<256, 4, 0, 0, 4>>("gp", 0, ALL);
        run_planes<Cfg<256, 4, 2, 0, 4>>("gp", 0, ALL);
        run_planes<Cfg<256, 4, 1, 1, 4>>("gp", 0, ALL);
        run_planes<Cfg<256, 4, 1, 2, 4>>("gp", 0, ALL);
        run_planes<Cfg<256, 4, 2, 1, 4>>("gp", 0, ALL);
        run_planes<Cfg<256, 2, 2, 1, 8>>("gp", 0, ALL);
        run_planes<Cfg<256, 2, 0, 0, 8>>("gp", 0, ALL);
        run_planes<Cfg<256, 4, 1, 0, 4>>("gops", 1, ALL);
        run_planes<Cfg<256, 4, 1, 0, 4>>("gops", 2, ALL);
        run_planes<Cfg<256, 2, 1, 0, 8>>("gops", 1, ALL);
        run_planes<Cfg<256, 2, 1, 0, 8>>("gops", 2, ALL);
    }
    if (!strcmp(only, "all") || !strcmp(only, "bulk")) {
        /* distinct content so that the comparison means something */
        {
            std::vector<uint8_t> h((size_t)64 << 20);
            uint32_t x = 12345;
            for (auto &v : h) { x = x * 1664525u + 1013904223u; v = (uint8_t)(x >> 24); }
            for (size_t off = 0; off + h.size() <= sz; off += h.size()) CK(cudaMemcpy(g_a + off, h.data(), h.size(), cudaMemcpyHostToDevice));
        }
        run_bulk<128>(8, 8, true);
        run_bulk<128>(4, 4, false);
        run_bulk<128>(16, 8, false);
        run_bulk<128>(8, 4, false);
        run_bulk<256>(8, 8, false);
        run_bulk<64>(8, 8, false);
        run_bulk<128>(2, 2, false);
        run_bulk<128, 1>(8, 8, true);
        run_bulk<128, 2>(8, 8, true);
        run_bulk<128, 3>(8, 8, true);
        run_bulk<128, 0>(8, 8, false);
        run_planes<Cfg<256, 4, 1, 1, 4>>("cmp", 0, 1 << 20);
    }
    if (!strcmp(only, "all") || !strcmp(only, "rgb")) {
        for (int fused = 0; fused < 2; fused++) {
            run_rgb<RCfg<256, 4, 1, 0>>("r", fused);
            run_rgb<RCfg<256, 4, 1, 0>>("rga", fused, true);
            run_rgb<RCfg<256, 3, 1, 0>>("rga", fused, true);
            run_rgb<RCfg<256, 2, 1, 0>>("rga", fused, true);
            run_rgb<RCfg<128, 8, 1, 0>>("rga", fused, true);
            run_rgb<RCfg<128, 6, 1, 0>>("rga", fused, true);
            run_rgb<RCfg<128, 4, 1, 0>>("rga", fused, true);
            run_rgb<RCfg<64, 8, 1, 0>>("rga", fused, true);
            run_rgb<RCfg<512, 2, 1, 0>>("rga", fused, true);
            run_rgb<RCfg<256, 4, 0, 0>>("rga", fused, true);
            run_rgb<RCfg<256, 4, 1, 1>>("rga", fused, true);
            run_rgb<RCfg<256, 4, 2, 1>>("rga", fused, true);
        }
    }
    return 0;
}
