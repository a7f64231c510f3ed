/*
 * jmc_k_planes.cuh -- NV12 <-> tight NV12 / I420 plane kernels for geometries whose every row is
 * directly addressable: planes_kernel (LDG/STG, any alignment, vector width chosen in-kernel) and
 * bulk_planes_kernel (copy engine, everything 16-byte aligned; the one the bench runs).
 */
#pragma once
#include "jmc_k_common.cuh"

namespace jmc {

/* Chosen by tools/sweep.cu on B200 (profiles/sweep_r1.md): 256 threads x 4 vectors = one 16 KB tile
 * per CTA, ONE CTA PER TILE (a persistent grid-stride loop measured 14% slower: the hardware CTA
 * scheduler overlaps the next tile's loads with this tile's draining stores better than a loop
 * does), L1::no_allocate loads, evict-first (.cs) stores. */
struct Cfg256x4 {
    static constexpr int THREADS = 256;
    static constexpr int UNROLL = 4;      /* 16-byte vectors per thread in flight */
    static constexpr int LDP = 1;
    static constexpr int STP = 1;
    static constexpr int BLOCKS_PER_SM = 4;
};

template <class C> struct TileGeom {
    static constexpr uint32_t TILE_ELEMS = (uint32_t)C::THREADS * C::UNROLL * 16u;
};

/* ---- COPY: rows x row_elems bytes between a pitched and a contiguous plane ----------------
 * Element e (a byte of the contiguous side) sits at row e / row_elems, column e % row_elems of
 * the pitched side.  Loads of a thread's UNROLL chunks are issued back to back (indices clamped
 * into the tile so no load is conditional), then the stores, predicated on the real bound. */
template <class C, int V, bool TO_TIGHT>
__device__ __forceinline__ void copy_tile(uint8_t *pitched, uint32_t pitch, uint8_t *tight,
                                          const FastDiv &rd, uint32_t e0, uint32_t e1)
{
    constexpr uint32_t STEP = (uint32_t)C::THREADS * V;
#pragma unroll 1
    for (uint32_t base = e0 + threadIdx.x * V; base < e1; base += STEP * C::UNROLL) {
        Chunk<V> r[C::UNROLL];
        uint32_t e[C::UNROLL];
        size_t poff[C::UNROLL];
#pragma unroll
        for (int k = 0; k < C::UNROLL; k++) {
            e[k] = base + k * STEP;
            const uint32_t ec = min(e[k], e1 - V);
            const uint32_t row = fast_div(ec, rd);
            poff[k] = (size_t)row * pitch + (ec - row * rd.d);
            r[k] = TO_TIGHT ? load_chunk<V, C::LDP>(pitched + poff[k]) : load_chunk<V, C::LDP>(tight + ec);
        }
#pragma unroll
        for (int k = 0; k < C::UNROLL; k++) {
            if (e[k] < e1) {
                if (TO_TIGHT) store_chunk<V, C::STP>(tight + e[k], r[k]);
                else store_chunk<V, C::STP>(pitched + poff[k], r[k]);
            }
        }
    }
}

/* ---- SPLIT: interleaved UV rows -> two contiguous chroma planes (V bytes per plane per chunk) */
template <int V> __device__ __forceinline__ void deinterleave(const Chunk<V> &lo, const Chunk<V> &hi, Chunk<V> &u, Chunk<V> &v)
{
    /* lo|hi hold 2V interleaved bytes U0 V0 U1 V1 ...; V >= 4 here */
#pragma unroll
    for (int i = 0; i < V / 4; i++) {
        const uint32_t a = (2 * i < V / 4) ? lo.w[2 * i] : hi.w[2 * i - V / 4];
        const uint32_t b = (2 * i + 1 < V / 4) ? lo.w[2 * i + 1] : hi.w[2 * i + 1 - V / 4];
        u.w[i] = __byte_perm(a, b, 0x6420);
        v.w[i] = __byte_perm(a, b, 0x7531);
    }
}
template <int V> __device__ __forceinline__ void interleave(const Chunk<V> &u, const Chunk<V> &v, Chunk<V> &lo, Chunk<V> &hi)
{
#pragma unroll
    for (int i = 0; i < V / 4; i++) {
        const uint32_t a = __byte_perm(u.w[i], v.w[i], 0x5140);
        const uint32_t b = __byte_perm(u.w[i], v.w[i], 0x7362);
        if (2 * i < V / 4) lo.w[2 * i] = a; else hi.w[2 * i - V / 4] = a;
        if (2 * i + 1 < V / 4) lo.w[2 * i + 1] = b; else hi.w[2 * i + 1 - V / 4] = b;
    }
}

/* elements are chroma sample pairs: pair e is bytes 2*(e % row_elems), +1 of UV row e / row_elems */
template <class C, int V>
__device__ __forceinline__ void split_tile(const uint8_t *uv, uint32_t pitch, uint8_t *pu, uint8_t *pv,
                                           const FastDiv &rd, uint32_t e0, uint32_t e1)
{
    constexpr uint32_t STEP = (uint32_t)C::THREADS * V;
    constexpr int U2 = (V == 16) ? (C::UNROLL + 1) / 2 : C::UNROLL;   /* 2V bytes are loaded per chunk */
#pragma unroll 1
    for (uint32_t base = e0 + threadIdx.x * V; base < e1; base += STEP * U2) {
        Chunk<V> lo[U2], hi[U2];
        uint32_t e[U2];
#pragma unroll
        for (int k = 0; k < U2; k++) {
            e[k] = base + k * STEP;
            const uint32_t ec = min(e[k], e1 - V);
            const uint32_t row = fast_div(ec, rd);
            const uint8_t *s = uv + (size_t)row * pitch + 2 * (size_t)(ec - row * rd.d);
            if (V >= 4) {
                lo[k] = load_chunk<V, C::LDP>(s);
                hi[k] = load_chunk<V, C::LDP>(s + V);
            } else if (V == 2) {
                lo[k].w[0] = __ldg((const uint32_t *)s);
            } else {
                lo[k].w[0] = __ldg(s);
                hi[k].w[0] = __ldg(s + 1);
            }
        }
#pragma unroll
        for (int k = 0; k < U2; k++) {
            if (e[k] < e1) {
                Chunk<V> u, v;
                if (V >= 4) deinterleave<V>(lo[k], hi[k], u, v);
                else if (V == 2) { u.w[0] = __byte_perm(lo[k].w[0], 0, 0x4420); v.w[0] = __byte_perm(lo[k].w[0], 0, 0x4431); }
                else { u.w[0] = lo[k].w[0]; v.w[0] = hi[k].w[0]; }
                store_chunk<V, C::STP>(pu + e[k], u);
                store_chunk<V, C::STP>(pv + e[k], v);
            }
        }
    }
}

/* ---- MERGE: two contiguous chroma planes -> interleaved UV rows ----------------------------- */
template <class C, int V>
__device__ __forceinline__ void merge_tile(uint8_t *uv, uint32_t pitch, const uint8_t *pu, const uint8_t *pv,
                                           const FastDiv &rd, uint32_t e0, uint32_t e1)
{
    constexpr uint32_t STEP = (uint32_t)C::THREADS * V;
    constexpr int U2 = (V == 16) ? (C::UNROLL + 1) / 2 : C::UNROLL;
#pragma unroll 1
    for (uint32_t base = e0 + threadIdx.x * V; base < e1; base += STEP * U2) {
        Chunk<V> u[U2], v[U2];
        uint32_t e[U2];
#pragma unroll
        for (int k = 0; k < U2; k++) {
            e[k] = base + k * STEP;
            const uint32_t ec = min(e[k], e1 - V);
            u[k] = load_chunk<V, C::LDP>(pu + ec);
            v[k] = load_chunk<V, C::LDP>(pv + ec);
        }
#pragma unroll
        for (int k = 0; k < U2; k++) {
            if (e[k] < e1) {
                const uint32_t row = fast_div(e[k], rd);
                uint8_t *d = uv + (size_t)row * pitch + 2 * (size_t)(e[k] - row * rd.d);
                if (V >= 4) {
                    Chunk<V> lo, hi;
                    interleave<V>(u[k], v[k], lo, hi);
                    store_chunk<V, C::STP>(d, lo);
                    store_chunk<V, C::STP>(d + V, hi);
                } else if (V == 2) {
                    *(uint32_t *)d = __byte_perm(u[k].w[0], v[k].w[0], 0x5140);
                } else {
                    d[0] = (uint8_t)u[k].w[0];
                    d[1] = (uint8_t)v[k].w[0];
                }
            }
        }
    }
}

#define JMC_DISPATCH_V(vw, CALL)            \
    switch (vw) {                           \
    case 16: { constexpr int V = 16; CALL; } break; \
    case 8:  { constexpr int V = 8;  CALL; } break; \
    case 4:  { constexpr int V = 4;  CALL; } break; \
    case 2:  { constexpr int V = 2;  CALL; } break; \
    default: { constexpr int V = 1;  CALL; } break; \
    }

/* vector width usable for a SPLIT/MERGE part: chunks of V bytes on every side, except V == 2
 * which moves one 4-byte word on the interleaved side */
__device__ __forceinline__ int chroma_vec_width(uint64_t pbits, uint64_t tbits)
{
    int vw = vec_width(pbits | tbits);
    if (vw == 2 && (pbits & 3)) vw = 1;
    return vw;
}

/* TO_TIGHT: 1 = pitched -> tight (decode side), 0 = tight -> pitched (encode side).
 * KIND1: what part[1] is (PART_COPY, PART_SPLIT or PART_MERGE); part[0] is always a COPY.
 * WIDE_ONLY: the host has proved every address/pitch/size 16-byte aligned (the 1080p / 4K case):
 * only the 16-byte path is compiled in, which keeps the register count low. */
template <class C, bool TO_TIGHT, int KIND1, bool WIDE_ONLY>
__global__ void __launch_bounds__(C::THREADS, WIDE_ONLY ? C::BLOCKS_PER_SM : 2) planes_kernel(const __grid_constant__ PlaneParams p)
{
    constexpr uint32_t TILE = TileGeom<C>::TILE_ELEMS;
    for (uint32_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const uint32_t f = t / p.tiles_per_frame;
        uint32_t r = t - f * p.tiles_per_frame;
        const bool second = r >= p.part[0].tiles;
        if (second) r -= p.part[0].tiles;
        uint8_t *pf = frame_ptr(p.pitched, f);
        uint8_t *tp = frame_ptr(p.tight, f);
        const uint32_t e0 = r * TILE;
        if (!second || KIND1 == PART_COPY) {
            const Part &pt = second ? p.part[1] : p.part[0];
            const uint32_t e1 = min(e0 + TILE, pt.rows * pt.row_elems);
            const uint32_t pitch = (uint32_t)pt.p_pitch;
            uint8_t *pp = pf + pt.p_off, *a = tp + pt.a_off;
            if (WIDE_ONLY) {
                copy_tile<C, 16, TO_TIGHT>(pp, pitch, a, pt.rdiv, e0, e1);
            } else {
                const int vw = vec_width((uint64_t)(uintptr_t)pp | (uint64_t)(uintptr_t)a | pitch | pt.row_elems);
                JMC_DISPATCH_V(vw, (copy_tile<C, V, TO_TIGHT>(pp, pitch, a, pt.rdiv, e0, e1)))
            }
        } else {
            const Part &pt = p.part[1];
            const uint32_t e1 = min(e0 + TILE, pt.rows * pt.row_elems);
            const uint32_t pitch = (uint32_t)pt.p_pitch;
            uint8_t *pp = pf + pt.p_off, *a = tp + pt.a_off, *b = tp + pt.b_off;
            if (WIDE_ONLY) {
                if (KIND1 == PART_SPLIT) split_tile<C, 16>(pp, pitch, a, b, pt.rdiv, e0, e1);
                else merge_tile<C, 16>(pp, pitch, a, b, pt.rdiv, e0, e1);
            } else {
                const int vw = chroma_vec_width((uint64_t)(uintptr_t)pp | pitch,
                                                (uint64_t)(uintptr_t)a | (uint64_t)(uintptr_t)b | pt.row_elems);
                if (KIND1 == PART_SPLIT) { JMC_DISPATCH_V(vw, (split_tile<C, V>(pp, pitch, a, b, pt.rdiv, e0, e1))) }
                else                     { JMC_DISPATCH_V(vw, (merge_tile<C, V>(pp, pitch, a, b, pt.rdiv, e0, e1))) }
            }
        }
    }
}

/* ========================================================================================== */
/* Bulk-copy-engine variant of the plane kernel (cp.async.bulk, SASS UBLKCP: the 1-D form of TMA). */
/* ========================================================================================== */
/* Used whenever the host has proved 16-byte alignment of everything (1080p, 4K, ...).  A tile is
 * `rows_per_tile` rows of one part; one CTA per tile:
 *   pitched -> tight : one bulk load per row (global, pitched) into CONTIGUOUS shared memory, then
 *                      ONE bulk store of the whole tile (the tight side is contiguous);
 *   tight -> pitched : one bulk load of the whole tile, one bulk store per row;
 *   SPLIT / MERGE    : the same, with the threads de-/interleaving shared -> shared (prmt) in between.
 * No register staging, no per-thread address arithmetic for the copies; measured +0.8 % (1080p) to
 * +1.7 % (4K) over the LDG/STG kernel (profiles/r1_sweep3_bulk_copy.csv). */
struct BulkParams {
    FrameSet pitched, tight;
    uint32_t n_frames;
    uint32_t rows_per_tile;
    uint32_t tiles[2];        /* tiles per frame of part 0 / part 1 */
    Part part[2];             /* Part::tiles unused here */
};

constexpr int BULK_THREADS = 128;

template <bool TO_TIGHT, int KIND1>
__global__ void __launch_bounds__(BULK_THREADS) bulk_planes_kernel(const __grid_constant__ BulkParams p)
{
    extern __shared__ __align__(128) uint8_t bulk_smem[];
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

    if (!second || KIND1 == PART_COPY) {
        if (threadIdx.x != 0) return;                     /* the copy engine does all the work */
        uint8_t *t = tp + pt.a_off + (size_t)r0 * re;
        mbar_expect_tx(&bar, nr * re);
        if (TO_TIGHT) {
            for (uint32_t i = 0; i < nr; i++) bulk_g2s(bulk_smem + (size_t)i * re, pp + (size_t)(r0 + i) * pitch, re, &bar);
            mbar_wait(&bar, 0);
            bulk_s2g(t, bulk_smem, nr * re);
        } else {
            bulk_g2s(bulk_smem, t, nr * re, &bar);
            mbar_wait(&bar, 0);
            for (uint32_t i = 0; i < nr; i++) bulk_s2g(pp + (size_t)(r0 + i) * pitch, bulk_smem + (size_t)i * re, re);
        }
        bulk_commit_wait_read();
    } else {
        /* chroma: re = pairs per row, 2*re interleaved bytes per pitched row */
        uint8_t *s_uv = bulk_smem;
        uint8_t *s_u = bulk_smem + (size_t)p.rows_per_tile * 2 * re;
        uint8_t *s_v = s_u + (size_t)p.rows_per_tile * re;
        uint8_t *tu = tp + pt.a_off + (size_t)r0 * re, *tv = tp + pt.b_off + (size_t)r0 * re;
        const uint32_t nvec = nr * re / 16;               /* 16 bytes of U and of V per step */
        if (KIND1 == PART_SPLIT) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bar, nr * 2 * re);
                for (uint32_t i = 0; i < nr; i++) bulk_g2s(s_uv + (size_t)i * 2 * re, pp + (size_t)(r0 + i) * pitch, 2 * re, &bar);
            }
            mbar_wait_cta(&bar, 0);
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
            fence_async_smem();                           /* generic-proxy writes -> visible to the copy engine */
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_s2g(tu, s_u, nr * re);
                bulk_s2g(tv, s_v, nr * re);
                bulk_commit_wait_read();
            }
        } else {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bar, nr * 2 * re);
                bulk_g2s(s_u, tu, nr * re, &bar);
                bulk_g2s(s_v, tv, nr * re, &bar);
            }
            mbar_wait_cta(&bar, 0);
            for (uint32_t v = threadIdx.x; v < nvec; v += BULK_THREADS) {
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
            if (threadIdx.x == 0) {
                for (uint32_t i = 0; i < nr; i++) bulk_s2g(pp + (size_t)(r0 + i) * pitch, s_uv + (size_t)i * 2 * re, 2 * re);
                bulk_commit_wait_read();
            }
        }
    }
}

} /* namespace jmc */
