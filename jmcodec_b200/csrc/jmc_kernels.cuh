/*
 * jmc_kernels.cuh -- sm_100a kernels for the decoded-surface format path.
 *
 * Everything here is HBM-bound byte movement (arithmetic intensity ~0; ~6 int-ops/B for RGB),
 * so the design rules are: 16-byte coalesced vector accesses, several independent loads in
 * flight per thread before the first store, fixed-size tiles with one CTA per tile (measured
 * faster than a persistent loop, see Cfg256x4), ONE launch per batch of frames, no tensor cores.
 *
 * Two kernels:
 *   planes_kernel : 2-D copy (strip/add pitch), U/V de-interleave (prmt 0x6420/0x7531) and
 *                   interleave (prmt 0x5140/0x7362).  Replaces the CPU loops of
 *                   nv_dec/nv_dec.cpp:782-820, intel_dec/intel_dec.cpp:284-314,
 *                   intel_enc/intel_enc.cpp:291-307,366-380 and the InterleaveUV launch of
 *                   nv_enc/nv_enc.cpp:1041-1081.
 *   rgb_kernel    : NV12 -> RGB24 (integer BT.601, dp2a + cvt.pack.sat) with optional fused I420
 *                   output; RGB rows are staged through shared memory so that every global
 *                   store instruction writes 512 contiguous bytes.
 *
 * The vector width of every (frame, plane) is chosen INSIDE the kernel from the actual
 * addresses/pitches (block-uniform branch), so odd sizes, odd crops and arbitrary pointer lists
 * are always correct and aligned geometries (1080p, 4K) always take the 16-byte path.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jmc {

struct FrameSet {
    uint8_t *base;
    size_t stride;
    uint8_t *const *list;
};

__device__ __forceinline__ uint8_t *frame_ptr(const FrameSet &s, uint32_t f)
{
    return s.list ? s.list[f] : s.base + (size_t)f * s.stride;
}

/* ---- division by a launch-invariant divisor ------------------------------------------------
 * floor(n / d) for n < 2^31 as (n * m) >> sh with m = ceil(2^sh / d), sh = 31 + ceil(log2 d):
 * the error term n*e/(d*2^sh), e < d <= 2^(sh-31), stays below 1/d.  Two instructions instead of
 * the ~20 of a generic 32-bit divide, four times per thread per tile. */
struct FastDiv {
    uint32_t m, sh, d, pad_;
};
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv &f)
{
    return (uint32_t)(((uint64_t)n * f.m) >> f.sh);
}

enum PartKind : int32_t { PART_NONE = 0, PART_COPY = 1, PART_SPLIT = 2, PART_MERGE = 3 };

/* One plane-level piece of work per frame.  "Elements" are bytes of a row (COPY) or chroma
 * sample pairs of a row (SPLIT / MERGE).  The tight side is always contiguous: element e of the
 * part lives at tight_frame + a_off + e (COPY; SPLIT/MERGE first chroma plane) and b_off + e
 * (second chroma plane). */
struct Part {
    int32_t kind;
    uint32_t rows;
    uint32_t row_elems;
    uint32_t tiles;      /* ceil(rows*row_elems / TILE_ELEMS) */
    int64_t p_off;       /* pitched side: offset of the part's first byte from the frame pointer */
    int32_t p_pitch;
    int32_t pad_;
    int64_t a_off;
    int64_t b_off;
    FastDiv rdiv;        /* division by row_elems */
};

struct PlaneParams {
    FrameSet pitched;
    FrameSet tight;
    uint32_t n_frames;
    int32_t to_tight;    /* 1: pitched -> tight (decode side), 0: tight -> pitched (encode side) */
    uint32_t tiles_per_frame;
    uint32_t total_tiles;
    Part part[2];
};

/* ------------------------------------------------------------------------------------------ */
/* memory access helpers.  LD policy 0: ld.global.nc  1: + L1::no_allocate  2: ld.global.cs      */
/*                         ST policy 0: st.global     1: st.global.cs       2: L1::no_allocate   */
template <int POL> __device__ __forceinline__ uint4 ld16(const void *p)
{
    uint4 r;
    if (POL == 1)
        asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (POL == 2)
        asm("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else
        asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int POL> __device__ __forceinline__ uint2 ld8(const void *p)
{
    uint2 r;
    if (POL == 1)
        asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    else if (POL == 2)
        asm("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    else
        asm("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
template <int POL> __device__ __forceinline__ void st16(void *p, uint4 v)
{
    if (POL == 1)
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (POL == 2)
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else
        asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int POL> __device__ __forceinline__ void st8(void *p, uint2 v)
{
    if (POL == 1)
        asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
    else if (POL == 2)
        asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
    else
        asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

/* A V-byte chunk held in registers (V = 16, 8, 4, 2, 1). */
template <int V> struct Chunk {
    uint32_t w[(V + 3) / 4];
};

template <int V, int LDP> __device__ __forceinline__ Chunk<V> load_chunk(const uint8_t *p)
{
    Chunk<V> c;
    if (V == 16) { uint4 t = ld16<LDP>(p); c.w[0] = t.x; c.w[1] = t.y; c.w[2] = t.z; c.w[3] = t.w; }
    else if (V == 8) { uint2 t = ld8<LDP>(p); c.w[0] = t.x; c.w[1] = t.y; }
    else if (V == 4) c.w[0] = __ldg((const uint32_t *)p);
    else if (V == 2) c.w[0] = __ldg((const uint16_t *)p);
    else c.w[0] = __ldg(p);
    return c;
}
template <int V, int STP> __device__ __forceinline__ void store_chunk(uint8_t *p, const Chunk<V> &c)
{
    if (V == 16) st16<STP>(p, make_uint4(c.w[0], c.w[1], c.w[2], c.w[3]));
    else if (V == 8) st8<STP>(p, make_uint2(c.w[0], c.w[1]));
    else if (V == 4) *(uint32_t *)p = c.w[0];
    else if (V == 2) *(uint16_t *)p = (uint16_t)c.w[0];
    else *p = (uint8_t)c.w[0];
}

/* largest power-of-two vector width (<=16) dividing every bit set in `bits` */
__device__ __forceinline__ int vec_width(uint64_t bits)
{
    uint32_t low = (uint32_t)bits & 15u;
    if (low == 0) return 16;
    return (int)(low & (0u - low));
}

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
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      /* smem may be released once it has been read */
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
            mbar_wait(&bar, 0);
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

/* store the first nbytes (<= 4*NW) of a register chunk at dst, as wide as dst's alignment allows */
template <int NW> __device__ __forceinline__ void store_prefix(uint8_t *dst, const uint32_t (&wd)[NW], uint32_t nbytes)
{
    const uint32_t a = (uint32_t)(uintptr_t)dst;
    if (NW == 4 && nbytes == 16 && (a & 7) == 0) {
        if ((a & 15) == 0) *(uint4 *)dst = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        else { *(uint2 *)dst = make_uint2(wd[0], wd[1]); *(uint2 *)(dst + 8) = make_uint2(wd[2], wd[3]); }
        return;
    }
    if (NW == 2 && nbytes == 8 && (a & 3) == 0) {
        if ((a & 7) == 0) *(uint2 *)dst = make_uint2(wd[0], wd[1]);
        else { *(uint32_t *)dst = wd[0]; *(uint32_t *)(dst + 4) = wd[1]; }
        return;
    }
    if ((a & 3) == 0) {
        /* row ends on aligned surfaces: whole words, then the last 1-3 bytes of the word that follows them
         * (a dozen instructions; the byte loop below costs ~50 issue slots even when it stores nothing) */
        const uint32_t nfull = nbytes >> 2, rem = nbytes & 3;
        uint32_t last = wd[0];
#pragma unroll
        for (int i = 0; i < NW; i++) {
            if ((uint32_t)i < nfull) *(uint32_t *)(dst + 4 * i) = wd[i];
            if ((uint32_t)i == nfull) last = wd[i];
        }
        uint8_t *q = dst + 4 * nfull;
        if (rem & 2) *(uint16_t *)q = (uint16_t)last;
        if (rem == 1) q[0] = (uint8_t)last;
        if (rem == 3) q[2] = (uint8_t)(last >> 16);
        return;
    }
#pragma unroll
    for (int i = 0; i < 4 * NW; i++)
        if ((uint32_t)i < nbytes) dst[i] = (uint8_t)(wd[i >> 2] >> (8 * (i & 3)));
}

/* ========================================================================================== */
/* Row kernels: full-width accesses for sizes that are NOT multiples of 16                        */
/* ========================================================================================== */
/* Decoder/encoder surfaces are always 16-byte aligned with a 16-byte-multiple pitch, whatever the
 * picture width; only the tight side (rows of w or w/2 bytes back to back) lands on odd addresses
 * when w is not a multiple of 16/32 (1366, 854, 1080-wide portrait chroma, odd sizes).  Common idea of
 * the kernels below: 16-byte accesses on the surface side, a pass through shared memory, and on the
 * tight side 16-byte accesses to the ALIGNED body of each row, re-aligned by a funnel shift (classic
 * unaligned memcpy), with byte accesses only for the <16-byte head and tail.
 * rows_kernel (this one, JMC_NO_BULK=1): LDG/STG, one warp per (row, 2 KB segment) or per group of short
 * rows.  bulk_rows_kernel / bulk_rows_pack_kernel (further down, the default): the copy engine loads. */
/* CTAs per SM the register allocation is sized for, decode / encode direction (A/B: tools/variants.sh,
 * profiles/r1_odd_sizes_minb.txt: 10 beats 8 and 12 on the decode side) */
#ifndef JMC_ROWS_MINB_DEC
#define JMC_ROWS_MINB_DEC 10
#endif
#ifndef JMC_ROWS_MINB_ENC
#define JMC_ROWS_MINB_ENC 8
#endif
constexpr int ROWS_THREADS = 128;
constexpr int ROWS_SEG = 2048;                       /* surface bytes per warp task */
constexpr int ROWS_MAX_RPT = 8;                      /* rows per warp task, upper bound (bounds the serial per-row store loop) */
constexpr int ROWS_SMEM_A = ROWS_SEG + 32, ROWS_SMEM_B = ROWS_SEG / 2 + 32;

struct RowsParams {
    FrameSet pitched, tight;
    uint32_t n_frames;
    uint32_t tasks[2];        /* warp tasks per frame of part 0 / part 1 */
    uint32_t segs[2];         /* segments per row (rows longer than ROWS_SEG) */
    uint32_t rpt[2];          /* rows per task (short rows: several rows share one warp task; 1 when segs > 1) */
    uint32_t rstride[2];      /* shared-memory stride of a staged row, surface bytes (multiple of 16; 32 for chroma pairs) */
    FastDiv cdiv[2];          /* division by rstride / 16 */
    uint32_t total_tasks;
    Part part[2];
};

/* bytes 4*WS + sh/8 .. +16 of the 32 bytes of two consecutive 16-byte chunks: words WS..WS+4, funnel-shifted.
 * WS is a template parameter and the callers branch on it ONCE per row (warp-uniform), outside their chunk
 * loops: as a run-time switch per chunk the compiler if-converts it into a dozen selects per 16 bytes, which
 * made the odd-width RGB kernels issue-bound (ncu: +48 % instructions, profiles/README.md). */
template <int WS> __device__ __forceinline__ uint4 shift_pair_ws(const uint4 &P, const uint4 &Q, uint32_t sh)
{
    const uint32_t x0 = WS == 0 ? P.x : WS == 1 ? P.y : WS == 2 ? P.z : P.w;
    const uint32_t x1 = WS == 0 ? P.y : WS == 1 ? P.z : WS == 2 ? P.w : Q.x;
    const uint32_t x2 = WS == 0 ? P.z : WS == 1 ? P.w : WS == 2 ? Q.x : Q.y;
    const uint32_t x3 = WS == 0 ? P.w : WS == 1 ? Q.x : WS == 2 ? Q.y : Q.z;
    const uint32_t x4 = WS == 0 ? Q.x : WS == 1 ? Q.y : WS == 2 ? Q.z : Q.w;
    uint4 o;
    o.x = __funnelshift_r(x0, x1, sh); o.y = __funnelshift_r(x1, x2, sh);
    o.z = __funnelshift_r(x2, x3, sh); o.w = __funnelshift_r(x3, x4, sh);
    return o;
}
__device__ __forceinline__ uint4 shift_pair(const uint4 &P, const uint4 &Q, uint32_t ws, uint32_t sh)
{
    switch (ws) {
    case 0: return shift_pair_ws<0>(P, Q, sh);
    case 1: return shift_pair_ws<1>(P, Q, sh);
    case 2: return shift_pair_ws<2>(P, Q, sh);
    default: return shift_pair_ws<3>(P, Q, sh);
    }
}

/* A staged buffer of nbytes -> dst (any alignment).  chunk(c) returns the shared-memory address of the
 * buffer's 16-byte chunk c (16-byte aligned; chunks up to nbytes/16 + 1 must be readable - the staging
 * buffers carry spare bytes).  Shared memory is read as whole chunks (conflict-free LDS.128), never as
 * strided words; global memory gets 16-byte stores on the aligned body, bytes on the < 16-byte head/tail. */
template <class ChunkMap>
__device__ __forceinline__ void warp_store_shifted_map(uint8_t *dst, ChunkMap chunk, uint32_t nbytes, uint32_t lane)
{
    const uint32_t head = min(nbytes, (16u - ((uint32_t)(uintptr_t)dst & 15u)) & 15u);
    const uint32_t body = (nbytes - head) & ~15u;
    if (lane < head) dst[lane] = ((const uint8_t *)chunk(0))[lane];
    const uint32_t sh = 8 * (head & 3);
#define JMC_SHIFTED_BODY(WS)                                                                                   \
    for (uint32_t j = lane; j < body / 16; j += 32) {                                                          \
        const uint4 P = *chunk(j), Q = *chunk(j + 1);                                                          \
        *(uint4 *)(dst + head + 16 * (size_t)j) = shift_pair_ws<WS>(P, Q, sh);                                 \
    }
    if (head == 0) {
        for (uint32_t j = lane; j < body / 16; j += 32) *(uint4 *)(dst + 16 * (size_t)j) = *chunk(j);
    } else {
        switch (head >> 2) {                                                     /* warp-uniform, once per row */
        case 0: JMC_SHIFTED_BODY(0) break;
        case 1: JMC_SHIFTED_BODY(1) break;
        case 2: JMC_SHIFTED_BODY(2) break;
        default: JMC_SHIFTED_BODY(3) break;
        }
    }
#undef JMC_SHIFTED_BODY
    const uint32_t t = head + body + lane;
    if (t < nbytes) dst[t] = ((const uint8_t *)chunk(t >> 4))[t & 15];
}

/* contiguous staging buffer sm[0..nbytes), 16-byte aligned, readable 32 bytes past nbytes */
__device__ __forceinline__ void warp_store_shifted(uint8_t *dst, const uint8_t *sm, uint32_t nbytes, uint32_t lane)
{
    warp_store_shifted_map(dst, [sm](uint32_t c) { return (const uint4 *)sm + c; }, nbytes, lane);
}

/* src (any alignment) -> smem[0..nbytes), nbytes <= 512*K.  Global memory is read as ALIGNED 16-byte
 * chunks, one load per lane and chunk, ALL issued before the first use; the neighbour chunk each output
 * needs comes from the next lane by shuffle (lane 31 takes lane 0's next chunk).  The first aligned chunk
 * starts up to 15 bytes before src: that is the end of the previous row / plane / frame, or - for the
 * first byte of a buffer - still inside the allocation (device allocations are at least 256-byte
 * aligned); nothing is ever read past src + nbytes. */
template <int K> struct ShiftedLoad {
    uint4 P[K + 1];
    uint32_t s, nfull, nout, t0, t1;

    /* phase 1: every global load of the row */
    __device__ __forceinline__ void issue(const uint8_t *src, uint32_t nbytes, uint32_t lane)
    {
        s = (uint32_t)(uintptr_t)src & 15u;
        const uint8_t *al = src - s;
        nfull = (nbytes + s) / 16;                           /* aligned chunks 0..nfull-1 end at or before src + nbytes */
        nout = s ? (nfull ? nfull - 1 : 0) : nfull;          /* output chunk j = bytes s.. of aligned chunks (j, j+1) */
#pragma unroll
        for (int k = 0; k < K; k++) {
            const uint32_t j = k * 32 + lane;
            P[k] = make_uint4(0, 0, 0, 0);
            if (j < nfull) P[k] = __ldg((const uint4 *)(al + 16 * (size_t)j));
        }
        P[K] = make_uint4(0, 0, 0, 0);
        const uint32_t i0 = nout * 16 + lane, i1 = i0 + 32;  /* the < 48 bytes after the last full output chunk */
        t0 = t1 = 0;
        if (i0 < nbytes) t0 = __ldg(src + i0);
        if (i1 < nbytes) t1 = __ldg(src + i1);
    }
    template <int WS> __device__ __forceinline__ void commit_ws(uint8_t *sm, uint32_t lane) const
    {
        const uint32_t sh = 8 * (s & 3), nxt = (lane + 1) & 31;
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (k * 32 >= (int)nout) break;                  /* warp-uniform */
            const uint32_t j = k * 32 + lane;
            /* lane l needs lane l+1's chunk; lane 31 needs lane 0's NEXT chunk.  Only the words that WS selects
             * travel: words WS.. of the neighbour chunk are never used when they fall beyond word WS+4. */
            const uint4 R = lane == 0 ? P[k + 1] : P[k];
            uint4 Q = make_uint4(0, 0, 0, 0);
            Q.x = __shfl_sync(0xffffffffu, R.x, nxt);
            if (WS >= 1) Q.y = __shfl_sync(0xffffffffu, R.y, nxt);
            if (WS >= 2) Q.z = __shfl_sync(0xffffffffu, R.z, nxt);
            if (WS >= 3) Q.w = __shfl_sync(0xffffffffu, R.w, nxt);
            if (j < nout) *(uint4 *)(sm + 16 * j) = shift_pair_ws<WS>(P[k], Q, sh);
        }
    }
    /* phase 2: re-align and store to shared memory */
    __device__ __forceinline__ void commit(uint8_t *sm, uint32_t nbytes, uint32_t lane) const
    {
        if (s == 0) {
#pragma unroll
            for (int k = 0; k < K; k++) { const uint32_t j = k * 32 + lane; if (j < nout) *(uint4 *)(sm + 16 * j) = P[k]; }
        } else {
            switch (s >> 2) {                                /* warp-uniform, once per row */
            case 0: commit_ws<0>(sm, lane); break;
            case 1: commit_ws<1>(sm, lane); break;
            case 2: commit_ws<2>(sm, lane); break;
            default: commit_ws<3>(sm, lane); break;
            }
        }
        const uint32_t i0 = nout * 16 + lane, i1 = i0 + 32;
        if (i0 < nbytes) sm[i0] = (uint8_t)t0;
        if (i1 < nbytes) sm[i1] = (uint8_t)t1;
    }
};

template <int K>
__device__ __forceinline__ void warp_load_shifted(uint8_t *sm, const uint8_t *src, uint32_t nbytes, uint32_t lane)
{
    ShiftedLoad<K> l;
    l.issue(src, nbytes, lane);
    l.commit(sm, nbytes, lane);
}

template <bool TO_TIGHT, int KIND1, bool MULTI>
__global__ void __launch_bounds__(ROWS_THREADS, TO_TIGHT ? JMC_ROWS_MINB_DEC : JMC_ROWS_MINB_ENC) rows_kernel(const __grid_constant__ RowsParams p)
{
    constexpr int WARPS = ROWS_THREADS / 32;
    __shared__ __align__(16) uint8_t sA[WARPS][ROWS_SMEM_A];
    __shared__ __align__(16) uint8_t sB[WARPS][ROWS_SMEM_B];
    __shared__ __align__(16) uint8_t sC[WARPS][ROWS_SMEM_B];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t task = blockIdx.x * WARPS + wib;
    if (task >= p.total_tasks) return;
    const uint32_t tpf = p.tasks[0] + p.tasks[1];
    const uint32_t f = task / tpf;
    uint32_t r = task - f * tpf;
    const bool second = r >= p.tasks[0];
    if (second) r -= p.tasks[0];
    const Part &pt = second ? p.part[1] : p.part[0];
    const uint32_t segs = second ? p.segs[1] : p.segs[0];
    /* MULTI: at least one part packs several rows into a task; otherwise the row arithmetic folds away */
    const uint32_t rpt = MULTI ? (second ? p.rpt[1] : p.rpt[0]) : 1u;
    const uint32_t rs = MULTI ? (second ? p.rstride[1] : p.rstride[0]) : (uint32_t)ROWS_SEG;
    const FastDiv &cdiv = second ? p.cdiv[1] : p.cdiv[0];
    /* a task is either one 2 KB segment of one row (segs >= 1, rpt == 1) or rpt whole rows (segs == 1) */
    uint32_t row, seg;
    if (rpt > 1) { row = r * rpt; seg = 0; } else { row = r / segs; seg = r - row * segs; }
    const uint32_t nr = min(rpt, pt.rows - row);
    const size_t pitch = (size_t)(uint32_t)pt.p_pitch;
    uint8_t *A = sA[wib], *B = sB[wib], *Cc = sC[wib];
    uint8_t *prow = frame_ptr(p.pitched, f) + pt.p_off + (size_t)row * pitch + (size_t)seg * ROWS_SEG;
    uint8_t *tp = frame_ptr(p.tight, f);
    /* surface side: slot s = 16 bytes at offset cc of staged row ri; shared-memory address A + 16 s */
    uint32_t ri[4], cc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t s = k * 32 + lane;
        ri[k] = MULTI ? fast_div(s, cdiv) : 0u;
        cc[k] = 16 * s - ri[k] * rs;
    }

    if (!second || KIND1 == PART_COPY) {
        const uint32_t nbytes = min((uint32_t)ROWS_SEG, pt.row_elems - seg * ROWS_SEG);
        uint8_t *trow = tp + pt.a_off + (size_t)row * pt.row_elems + (size_t)seg * ROWS_SEG;
        if (TO_TIGHT) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) if (ri[k] < nr && cc[k] < nbytes) v[k] = ld16<1>(prow + ri[k] * pitch + cc[k]);
#pragma unroll
            for (int k = 0; k < 4; k++) if (ri[k] < nr && cc[k] < nbytes) *(uint4 *)(A + 16 * (k * 32 + lane)) = v[k];
            __syncwarp();
            for (uint32_t i = 0; i < nr; i++) warp_store_shifted(trow + (size_t)i * pt.row_elems, A + i * rs, nbytes, lane);
        } else {
            for (uint32_t i = 0; i < nr; i++) warp_load_shifted<4>(A + i * rs, trow + (size_t)i * pt.row_elems, nbytes, lane);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (ri[k] >= nr) continue;
                uint8_t *d = prow + ri[k] * pitch + cc[k];
                const uint8_t *sm = A + 16 * (k * 32 + lane);
                if (cc[k] + 16 <= nbytes) *(uint4 *)d = *(const uint4 *)sm;
                else if (cc[k] < nbytes) { const uint4 t = *(const uint4 *)sm; const uint32_t wd[4] = {t.x, t.y, t.z, t.w}; store_prefix<4>(d, wd, nbytes - cc[k]); }
            }
        }
    } else {
        /* chroma: elements are pairs; a segment is ROWS_SEG interleaved bytes = ROWS_SEG/2 pairs; staged rows
         * are rs interleaved bytes apart in A (rs a multiple of 32) and rs/2 apart in B (U) and Cc (V) */
        const uint32_t npairs = min((uint32_t)ROWS_SEG / 2, pt.row_elems - seg * (ROWS_SEG / 2));
        const uint32_t nbytes = 2 * npairs;
        const uint32_t span = rpt > 1 ? nr * rs : nbytes;                  /* staged interleaved bytes of the task */
        uint8_t *tu = tp + pt.a_off + (size_t)row * pt.row_elems + (size_t)seg * (ROWS_SEG / 2);
        uint8_t *tv = tp + pt.b_off + (size_t)row * pt.row_elems + (size_t)seg * (ROWS_SEG / 2);
        if (KIND1 == PART_SPLIT) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) if (ri[k] < nr && cc[k] < nbytes) v[k] = ld16<1>(prow + ri[k] * pitch + cc[k]);
#pragma unroll
            for (int k = 0; k < 4; k++) if (ri[k] < nr && cc[k] < nbytes) *(uint4 *)(A + 16 * (k * 32 + lane)) = v[k];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t c = k * 32 + lane;                          /* 32-byte chunk -> 16 U + 16 V */
                if (32 * c < span) {
                    const uint4 a = *(const uint4 *)(A + 32 * c), b = *(const uint4 *)(A + 32 * c + 16);
                    uint4 u, w;
                    u.x = __byte_perm(a.x, a.y, 0x6420); w.x = __byte_perm(a.x, a.y, 0x7531);
                    u.y = __byte_perm(a.z, a.w, 0x6420); w.y = __byte_perm(a.z, a.w, 0x7531);
                    u.z = __byte_perm(b.x, b.y, 0x6420); w.z = __byte_perm(b.x, b.y, 0x7531);
                    u.w = __byte_perm(b.z, b.w, 0x6420); w.w = __byte_perm(b.z, b.w, 0x7531);
                    *(uint4 *)(B + 16 * c) = u;
                    *(uint4 *)(Cc + 16 * c) = w;
                }
            }
            __syncwarp();
            for (uint32_t i = 0; i < nr; i++) {
                warp_store_shifted(tu + (size_t)i * pt.row_elems, B + i * (rs / 2), npairs, lane);
                warp_store_shifted(tv + (size_t)i * pt.row_elems, Cc + i * (rs / 2), npairs, lane);
            }
        } else {
            for (uint32_t i = 0; i < nr; i++) {
                ShiftedLoad<2> lu, lv;                                      /* U and V loads in flight together */
                lu.issue(tu + (size_t)i * pt.row_elems, npairs, lane);
                lv.issue(tv + (size_t)i * pt.row_elems, npairs, lane);
                lu.commit(B + i * (rs / 2), npairs, lane);
                lv.commit(Cc + i * (rs / 2), npairs, lane);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t c = k * 32 + lane;
                if (32 * c < span) {
                    const uint4 u = *(const uint4 *)(B + 16 * c), w = *(const uint4 *)(Cc + 16 * c);
                    uint4 a, b;
                    a.x = __byte_perm(u.x, w.x, 0x5140); a.y = __byte_perm(u.x, w.x, 0x7362);
                    a.z = __byte_perm(u.y, w.y, 0x5140); a.w = __byte_perm(u.y, w.y, 0x7362);
                    b.x = __byte_perm(u.z, w.z, 0x5140); b.y = __byte_perm(u.z, w.z, 0x7362);
                    b.z = __byte_perm(u.w, w.w, 0x5140); b.w = __byte_perm(u.w, w.w, 0x7362);
                    *(uint4 *)(A + 32 * c) = a;
                    *(uint4 *)(A + 32 * c + 16) = b;
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (ri[k] >= nr) continue;
                uint8_t *d = prow + ri[k] * pitch + cc[k];
                const uint8_t *sm = A + 16 * (k * 32 + lane);
                if (cc[k] + 16 <= nbytes) *(uint4 *)d = *(const uint4 *)sm;
                else if (cc[k] < nbytes) { const uint4 t = *(const uint4 *)sm; const uint32_t wd[4] = {t.x, t.y, t.z, t.w}; store_prefix<4>(d, wd, nbytes - cc[k]); }
            }
        }
    }
}

/* ========================================================================================== */
/* Bulk-loaded rows: decode direction, aligned surface, any width                              */
/* ========================================================================================== */
/* The surface side of a width that is not a multiple of 16 is still bulk-copy friendly (aligned rows,
 * over-readable to the next multiple of 16 inside the pitch), so the copy engine loads a tile of rows
 * into shared memory - every byte of the tile in flight at once, no registers, no LDG issue slots - and
 * the four warps only do the re-aligned 16-byte stores of warp_store_shifted(), one tight row at a time
 * (chroma: after a shared -> shared prmt de-interleave).  rows_kernel's load half was what held 1366-
 * and 854-wide frames at 0.84-0.89 of peak: one row per warp leaves too few bytes in flight. */
constexpr int BROWS_THREADS = 128;

struct BulkRowsParams {
    FrameSet pitched, tight;
    uint32_t n_frames;
    uint32_t rows_per_tile;
    uint32_t tiles[2];        /* tiles per frame of part 0 / part 1 */
    uint32_t rstride[2];      /* shared-memory stride of a staged row (surface bytes; multiple of 16, of 32 for chroma pairs) */
    uint32_t ldbytes[2];      /* bytes per bulk row load: row bytes rounded up to 16 */
    Part part[2];
};

template <int KIND1>
__global__ void __launch_bounds__(BROWS_THREADS) bulk_rows_kernel(const __grid_constant__ BulkRowsParams p)
{
    extern __shared__ __align__(128) uint8_t bulk_smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t tpf = p.tiles[0] + p.tiles[1];
    const uint32_t f = blockIdx.x / tpf;
    uint32_t r = blockIdx.x - f * tpf;
    const bool second = r >= p.tiles[0];
    if (second) r -= p.tiles[0];
    const Part &pt = second ? p.part[1] : p.part[0];
    const uint32_t rs = second ? p.rstride[1] : p.rstride[0];
    const uint32_t ld = second ? p.ldbytes[1] : p.ldbytes[0];
    const uint8_t *pp = frame_ptr(p.pitched, f) + pt.p_off;
    uint8_t *tp = frame_ptr(p.tight, f);
    const uint32_t r0 = r * p.rows_per_tile;
    const uint32_t nr = min(p.rows_per_tile, pt.rows - r0);
    const uint32_t re = pt.row_elems;
    const size_t pitch = (size_t)(uint32_t)pt.p_pitch;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *A = bulk_smem;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, nr * ld);
        for (uint32_t i = 0; i < nr; i++) bulk_g2s(A + (size_t)i * rs, pp + (size_t)(r0 + i) * pitch, ld, &bar);
    }
    mbar_wait(&bar, 0);

    if (!second || KIND1 == PART_COPY) {
        uint8_t *t = tp + pt.a_off + (size_t)r0 * re;
        for (uint32_t i = warp; i < nr; i += BROWS_THREADS / 32) warp_store_shifted(t + (size_t)i * re, A + (size_t)i * rs, re, lane);
    } else {
        /* chroma: re = pairs per row; staged rows are rs interleaved bytes apart, rs/2 apart in the planar halves */
        uint8_t *B = A + (size_t)p.rows_per_tile * rs + 32;
        uint8_t *Cc = B + (size_t)p.rows_per_tile * (rs / 2) + 32;
        const uint32_t nvec = nr * rs / 32;                   /* 16 bytes of U and of V per step */
        for (uint32_t v = threadIdx.x; v < nvec; v += BROWS_THREADS) {
            const uint4 a = *(const uint4 *)(A + (size_t)v * 32), b = *(const uint4 *)(A + (size_t)v * 32 + 16);
            uint4 u, w;
            u.x = __byte_perm(a.x, a.y, 0x6420); w.x = __byte_perm(a.x, a.y, 0x7531);
            u.y = __byte_perm(a.z, a.w, 0x6420); w.y = __byte_perm(a.z, a.w, 0x7531);
            u.z = __byte_perm(b.x, b.y, 0x6420); w.z = __byte_perm(b.x, b.y, 0x7531);
            u.w = __byte_perm(b.z, b.w, 0x6420); w.w = __byte_perm(b.z, b.w, 0x7531);
            *(uint4 *)(B + (size_t)v * 16) = u;
            *(uint4 *)(Cc + (size_t)v * 16) = w;
        }
        __syncthreads();
        uint8_t *tu = tp + pt.a_off + (size_t)r0 * re, *tv = tp + pt.b_off + (size_t)r0 * re;
        for (uint32_t i = warp; i < nr; i += BROWS_THREADS / 32) {
            warp_store_shifted(tu + (size_t)i * re, B + (size_t)i * (rs / 2), re, lane);
            warp_store_shifted(tv + (size_t)i * re, Cc + (size_t)i * (rs / 2), re, lane);
        }
    }
}

/* Encode direction of the same idea.  The tight rows of a tile are ONE contiguous run at an arbitrary
 * address: its 16-byte-aligned interior is bulk-loaded into shared memory at the same alignment modulo
 * 16 (nothing outside the run is read), the < 16-byte head and tail come in through two warps, and
 * each surface row (16-byte aligned) is then assembled from two aligned shared-memory chunks with a
 * per-row funnel shift - U and V re-aligned separately and interleaved in registers for the packed
 * chroma plane.  Padding bytes are never written (the last chunk of a row is a prefix store). */
struct StagedRun {
    uint32_t a, head, body, len;      /* run byte i lives at S[a + i]; S + a + head is 16-byte aligned */
};
__device__ __forceinline__ StagedRun make_run(const uint8_t *src, uint32_t len)
{
    StagedRun r;
    r.a = (uint32_t)(uintptr_t)src & 15u;
    r.len = len;
    r.head = min(len, (16u - r.a) & 15u);
    r.body = (len - r.head) & ~15u;
    return r;
}
/* warps 0 and 1 bring in the head and the tail (thread 0 has already issued the bulk load of the body) */
__device__ __forceinline__ void run_edges(uint8_t *S, const uint8_t *src, const StagedRun &r, uint32_t lane, uint32_t warp)
{
    if (warp == 0 && lane < r.head) S[r.a + lane] = src[lane];
    const uint32_t t = r.head + r.body + lane;
    if (warp == 1 && t < r.len) S[r.a + t] = src[t];
}
/* 16 bytes of a staged run starting at byte offset off of S (any alignment) */
__device__ __forceinline__ uint4 staged16(const uint8_t *S, uint32_t off)
{
    const uint4 *q = (const uint4 *)S + (off >> 4);
    if ((off & 15) == 0) return q[0];
    return shift_pair(q[0], q[1], (off & 15) >> 2, 8 * (off & 3));
}

template <int KIND1>
__global__ void __launch_bounds__(BROWS_THREADS) bulk_rows_pack_kernel(const __grid_constant__ BulkRowsParams p)
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
    const uint8_t *tp = frame_ptr(p.tight, f);
    const uint32_t r0 = r * p.rows_per_tile;
    const uint32_t nr = min(p.rows_per_tile, pt.rows - r0);
    const uint32_t re = pt.row_elems;
    const size_t pitch = (size_t)(uint32_t)pt.p_pitch;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();

    if (!second || KIND1 == PART_COPY) {
        const uint8_t *src = tp + pt.a_off + (size_t)r0 * re;
        const StagedRun run = make_run(src, nr * re);
        uint8_t *S = bulk_smem;
        if (threadIdx.x == 0 && run.body) {
            mbar_expect_tx(&bar, run.body);
            bulk_g2s(S + run.a + run.head, src + run.head, run.body, &bar);
        }
        run_edges(S, src, run, lane, warp);
        __syncthreads();
        if (run.body) mbar_wait(&bar, 0);
        for (uint32_t i = warp; i < nr; i += BROWS_THREADS / 32) {
            uint8_t *d = pp + (size_t)(r0 + i) * pitch;
            const uint32_t off = run.a + i * re;
            const uint4 *q0 = (const uint4 *)S + (off >> 4);                  /* the row starts off & 15 bytes into this chunk */
            const uint32_t sh = 8 * (off & 3);
#define JMC_PACK_ROW(EXPR)                                                                                     \
            for (uint32_t c = 16 * lane; c < re; c += 512) {                                                   \
                const uint4 *q = q0 + (c >> 4);                                                                \
                const uint4 o = EXPR;                                                                          \
                if (c + 16 <= re) *(uint4 *)(d + c) = o;                                                       \
                else { const uint32_t wd[4] = {o.x, o.y, o.z, o.w}; store_prefix<4>(d + c, wd, re - c); }     \
            }
            if ((off & 15) == 0) { JMC_PACK_ROW(q[0]) }
            else switch ((off & 15) >> 2) {                                    /* warp-uniform, once per row */
            case 0: JMC_PACK_ROW(shift_pair_ws<0>(q[0], q[1], sh)) break;
            case 1: JMC_PACK_ROW(shift_pair_ws<1>(q[0], q[1], sh)) break;
            case 2: JMC_PACK_ROW(shift_pair_ws<2>(q[0], q[1], sh)) break;
            default: JMC_PACK_ROW(shift_pair_ws<3>(q[0], q[1], sh)) break;
            }
#undef JMC_PACK_ROW
        }
    } else {
        /* MERGE: re = pairs per row; U run and V run staged separately */
        const uint8_t *su = tp + pt.a_off + (size_t)r0 * re, *sv = tp + pt.b_off + (size_t)r0 * re;
        const StagedRun ru = make_run(su, nr * re), rv = make_run(sv, nr * re);
        uint8_t *Su = bulk_smem;
        uint8_t *Sv = bulk_smem + (((size_t)p.rows_per_tile * re + 63) & ~(size_t)15);
        if (threadIdx.x == 0 && (ru.body | rv.body)) {
            mbar_expect_tx(&bar, ru.body + rv.body);
            if (ru.body) bulk_g2s(Su + ru.a + ru.head, su + ru.head, ru.body, &bar);
            if (rv.body) bulk_g2s(Sv + rv.a + rv.head, sv + rv.head, rv.body, &bar);
        }
        run_edges(Su, su, ru, lane, warp);
        run_edges(Sv, sv, rv, lane, warp ^ 2);               /* warps 2 and 3 */
        __syncthreads();
        if (ru.body | rv.body) mbar_wait(&bar, 0);
        const uint32_t nbytes = 2 * re;
        for (uint32_t i = warp; i < nr; i += BROWS_THREADS / 32) {
            uint8_t *d = pp + (size_t)(r0 + i) * pitch;
            const uint32_t offu = ru.a + i * re, offv = rv.a + i * re;
            for (uint32_t c = 16 * lane; c < re; c += 512) {          /* 16 pairs -> 32 interleaved bytes at 2c */
                const uint4 u = staged16(Su, offu + c), w = staged16(Sv, offv + c);
                uint32_t lo[4], hi[4];
                lo[0] = __byte_perm(u.x, w.x, 0x5140); lo[1] = __byte_perm(u.x, w.x, 0x7362);
                lo[2] = __byte_perm(u.y, w.y, 0x5140); lo[3] = __byte_perm(u.y, w.y, 0x7362);
                hi[0] = __byte_perm(u.z, w.z, 0x5140); hi[1] = __byte_perm(u.z, w.z, 0x7362);
                hi[2] = __byte_perm(u.w, w.w, 0x5140); hi[3] = __byte_perm(u.w, w.w, 0x7362);
                const uint32_t rem = nbytes - 2 * c;                  /* > 0 */
                if (rem >= 32) {
                    *(uint4 *)(d + 2 * c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *(uint4 *)(d + 2 * c + 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                } else if (rem >= 16) {
                    *(uint4 *)(d + 2 * c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    if (rem > 16) store_prefix<4>(d + 2 * c + 16, hi, rem - 16);
                } else {
                    store_prefix<4>(d + 2 * c, lo, rem);
                }
            }
        }
    }
}

/* ========================================================================================== */
/* NV12 -> RGB24 (+ optional I420)                                                            */
/* ========================================================================================== */
struct RgbParams {
    FrameSet surf, tight, rgb;
    uint32_t n_frames;
    int32_t width, height, pitch;
    int64_t y_off, uv_off;
    int64_t u_off, v_off;      /* tight I420 plane offsets (fused only) */
    int32_t rgb_pitch;
    int32_t fused;
    int32_t argb;              /* 1: 4 bytes per pixel (B,G,R,0xFF) instead of packed R,G,B */
    uint32_t segs_per_row;     /* ceil(width / 512): one warp covers 512 pixels of a row pair */
    uint32_t row_pairs;        /* ceil(height / 2) */
    uint32_t tasks_per_frame;  /* row_pairs * segs_per_row */
    FastDiv tpf_div, seg_div;  /* division by tasks_per_frame / segs_per_row (a generic divide costs ~20 issue slots) */
    uint32_t total_tasks;
};

/* d = c + a.lo16 * b.byte[0|2] + a.hi16 * b.byte[1|3]   (signed 16-bit coefficients x unsigned bytes) */
__device__ __forceinline__ int dp2a_lo(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
/* (sat_u16(a) << 16) | sat_u16(b).  clip8(x >> 8) == sat_u16(x) >> 8, so the result bytes we want
 * are byte 3 (from a) and byte 1 (from b). */
__device__ __forceinline__ uint32_t pack_sat_u16(int a, int b)
{
    uint32_t d;
    asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

/* BT.601 limited range, integer:  x_r = 298(Y-16)+409(V-128)+128, etc. */
constexpr int RGB_CR = -298 * 16 - 409 * 128 + 128;
constexpr int RGB_CG = -298 * 16 + 100 * 128 + 208 * 128 + 128;
constexpr int RGB_CB = -298 * 16 - 516 * 128 + 128;
constexpr uint32_t COEF_RV = (409u << 16);                              /* 0*U + 409*V   */
constexpr uint32_t COEF_GUV = ((uint32_t)(uint16_t)(-100)) | ((uint32_t)(uint16_t)(-208) << 16);
constexpr uint32_t COEF_BU = 516u;                                      /* 516*U + 0*V   */
constexpr uint32_t COEF_Y_EVEN = 298u;                                  /* picks byte 0 / 2 */
constexpr uint32_t COEF_Y_ODD = (298u << 16);                           /* picks byte 1 / 3 */

/* 4 pixels (one Y word) + their 2 chroma pairs (one UV word) -> 12 RGB bytes in 3 words */
__device__ __forceinline__ void rgb4(uint32_t yw, int r0, int g0, int b0, int r1, int g1, int b1, uint32_t *out)
{
    const int R0 = dp2a_lo(COEF_Y_EVEN, yw, r0), G0 = dp2a_lo(COEF_Y_EVEN, yw, g0), B0 = dp2a_lo(COEF_Y_EVEN, yw, b0);
    const int R1 = dp2a_lo(COEF_Y_ODD, yw, r0), G1 = dp2a_lo(COEF_Y_ODD, yw, g0), B1 = dp2a_lo(COEF_Y_ODD, yw, b0);
    const int R2 = dp2a_hi(COEF_Y_EVEN, yw, r1), G2 = dp2a_hi(COEF_Y_EVEN, yw, g1), B2 = dp2a_hi(COEF_Y_EVEN, yw, b1);
    const int R3 = dp2a_hi(COEF_Y_ODD, yw, r1), G3 = dp2a_hi(COEF_Y_ODD, yw, g1), B3 = dp2a_hi(COEF_Y_ODD, yw, b1);
    out[0] = __byte_perm(pack_sat_u16(G0, R0), pack_sat_u16(R1, B0), 0x7531);   /* R0 G0 B0 R1 */
    out[1] = __byte_perm(pack_sat_u16(B1, G1), pack_sat_u16(G2, R2), 0x7531);   /* G1 B1 R2 G2 */
    out[2] = __byte_perm(pack_sat_u16(R3, B2), pack_sat_u16(B3, G3), 0x7531);   /* B2 R3 G3 B3 */
}

/* same 4 pixels -> 4 ARGB8888 words (bytes B,G,R,0xFF): sat_u16(65535) supplies the alpha byte */
__device__ __forceinline__ void argb4(uint32_t yw, int r0, int g0, int b0, int r1, int g1, int b1, uint32_t *out)
{
    const int R0 = dp2a_lo(COEF_Y_EVEN, yw, r0), G0 = dp2a_lo(COEF_Y_EVEN, yw, g0), B0 = dp2a_lo(COEF_Y_EVEN, yw, b0);
    const int R1 = dp2a_lo(COEF_Y_ODD, yw, r0), G1 = dp2a_lo(COEF_Y_ODD, yw, g0), B1 = dp2a_lo(COEF_Y_ODD, yw, b0);
    const int R2 = dp2a_hi(COEF_Y_EVEN, yw, r1), G2 = dp2a_hi(COEF_Y_EVEN, yw, g1), B2 = dp2a_hi(COEF_Y_EVEN, yw, b1);
    const int R3 = dp2a_hi(COEF_Y_ODD, yw, r1), G3 = dp2a_hi(COEF_Y_ODD, yw, g1), B3 = dp2a_hi(COEF_Y_ODD, yw, b1);
    out[0] = __byte_perm(pack_sat_u16(G0, B0), pack_sat_u16(65535, R0), 0x7531);
    out[1] = __byte_perm(pack_sat_u16(G1, B1), pack_sat_u16(65535, R1), 0x7531);
    out[2] = __byte_perm(pack_sat_u16(G2, B2), pack_sat_u16(65535, R2), 0x7531);
    out[3] = __byte_perm(pack_sat_u16(G3, B3), pack_sat_u16(65535, R3), 0x7531);
}

__device__ __forceinline__ uint8_t clip8_dev(int v) { return (uint8_t)min(max(v, 0), 255); }

struct RgbCfg {                           /* tools/sweep.cu: 128 x 8 CTAs/SM, one warp task per warp */
    static constexpr int THREADS = 128;
    static constexpr int BLOCKS_PER_SM = 8;
    static constexpr int LDP = 1;
    static constexpr int STP = 0;
};

/* copy nbytes from warp-private shared memory to global, V bytes per lane per step */
template <int V, int STP> __device__ __forceinline__ void warp_flush(uint8_t *g, const uint8_t *st, uint32_t nbytes, uint32_t lane)
{
    uint8_t *gl = g + lane * V;                       /* per-lane bases once, constant steps of 32*V */
    const uint8_t *sl = st + lane * V;
#pragma unroll
    for (int k = 0; k < 1536 / (32 * V); k++) {
        constexpr int STEP = 32 * V;
        if (k * STEP + lane * V < nbytes) {
            if (V == 16) st16<STP>(gl + k * STEP, *(const uint4 *)(sl + k * STEP));
            else if (V == 8) st8<STP>(gl + k * STEP, *(const uint2 *)(sl + k * STEP));
            else if (V == 4) *(uint32_t *)(gl + k * STEP) = *(const uint32_t *)(sl + k * STEP);
            else if (V == 2) *(uint16_t *)(gl + k * STEP) = *(const uint16_t *)(sl + k * STEP);
            else gl[k * STEP] = sl[k * STEP];
        }
    }
}

template <class C, bool ARGB>
__global__ void __launch_bounds__(C::THREADS, C::BLOCKS_PER_SM) rgb_kernel(const __grid_constant__ RgbParams p)
{
    constexpr int WARPS = C::THREADS / 32;
    __shared__ __align__(16) uint8_t stage[WARPS][32 * 80];      /* RGB24: 48 B per lane; ARGB32: 64 B at an 80-byte stride */
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warps_total = gridDim.x * WARPS;
    const int w = p.width, h = p.height, cw = w >> 1, ch = h >> 1;

    for (uint32_t task = blockIdx.x * WARPS + wib; task < p.total_tasks; task += warps_total) {
        const uint32_t f = fast_div(task, p.tpf_div);
        const uint32_t r = task - f * p.tasks_per_frame;
        const uint32_t rp = fast_div(r, p.seg_div), seg = r - rp * p.segs_per_row;
        const uint8_t *sp = frame_ptr(p.surf, f);
        uint8_t *rgbp = frame_ptr(p.rgb, f);
        uint8_t *tp = p.fused ? frame_ptr(p.tight, f) : nullptr;
        const uint32_t y0 = rp * 2;
        const bool two = (y0 + 1 < (uint32_t)h);
        const uint32_t cy = min(rp, (uint32_t)(ch - 1));
        const uint8_t *yrow = sp + p.y_off + (size_t)y0 * p.pitch;
        const uint8_t *crow = sp + p.uv_off + (size_t)cy * p.pitch;
        uint8_t *orow = rgbp + (size_t)y0 * p.rgb_pitch;

        /* vector path: 16-byte loads need an aligned surface whose rows can be over-read up to the next
         * multiple of 16 (always true for decoder surfaces); the stores adapt to whatever alignment the
         * tight / RGB rows have (1080-wide portrait video: 8-byte RGB rows, 4-byte chroma rows). */
        const uint64_t sbits = (uint64_t)(uintptr_t)(sp + p.y_off) | (uint64_t)(uintptr_t)(sp + p.uv_off) | (uint32_t)p.pitch;
        const bool vec = (sbits & 15) == 0 && (w & 1) == 0 && p.pitch >= ((w + 15) & ~15);

        if (vec) {
            const uint32_t px0 = (seg * 32 + lane) * 16;
            const uint32_t npx = px0 < (uint32_t)w ? min(16u, (uint32_t)w - px0) : 0u;     /* valid pixels of this lane (even) */
            const uint32_t seg_px = min(512u, (uint32_t)w - seg * 512);                    /* valid pixels of this warp */
            uint4 ya = make_uint4(0, 0, 0, 0), yb = ya, uv = ya;
            if (npx) {
                ya = ld16<C::LDP>(yrow + px0);
                uv = ld16<C::LDP>(crow + px0);
                if (two) yb = ld16<C::LDP>(yrow + p.pitch + px0);
            }
            uint8_t *st = stage[wib];
            if (p.fused) {
                uint8_t *ty = tp + (size_t)y0 * w + seg * 512;
                uint8_t *tu = tp + p.u_off + (size_t)rp * cw + seg * 256, *tv = tp + p.v_off + (size_t)rp * cw + seg * 256;
                const uint32_t u[2] = {__byte_perm(uv.x, uv.y, 0x6420), __byte_perm(uv.z, uv.w, 0x6420)};
                const uint32_t v[2] = {__byte_perm(uv.x, uv.y, 0x7531), __byte_perm(uv.z, uv.w, 0x7531)};
                const uint64_t tbits = (uint64_t)(uintptr_t)ty | (uint64_t)(uintptr_t)tu | (uint64_t)(uintptr_t)tv | (uint32_t)w | (uint32_t)cw;
                if ((tbits & 7) == 0) {                                   /* every lane's 16 luma / 8 chroma bytes land aligned */
                    if (npx) {
                        const uint32_t y_a[4] = {ya.x, ya.y, ya.z, ya.w}, y_b[4] = {yb.x, yb.y, yb.z, yb.w};
                        store_prefix<4>(ty + 16 * lane, y_a, npx);
                        if (two) store_prefix<4>(ty + w + 16 * lane, y_b, npx);
                        if (rp < (uint32_t)ch) { store_prefix<2>(tu + 8 * lane, u, npx >> 1); store_prefix<2>(tv + 8 * lane, v, npx >> 1); }
                    }
                } else {
                    /* tight rows on odd addresses (1366-wide, ...): stage each row in the spare 960 bytes behind the
                     * RGB staging area and write it with 16-byte stores re-aligned by a funnel shift */
                    uint8_t *sy = st + 1600, *sv = st + 1600 + 288;
                    *(uint4 *)(sy + 16 * lane) = ya;
                    __syncwarp();
                    warp_store_shifted(ty, sy, seg_px, lane);
                    __syncwarp();
                    if (two) {
                        *(uint4 *)(sy + 16 * lane) = yb;
                        __syncwarp();
                        warp_store_shifted(ty + w, sy, seg_px, lane);
                        __syncwarp();
                    }
                    if (rp < (uint32_t)ch) {
                        *(uint2 *)(sy + 8 * lane) = make_uint2(u[0], u[1]);
                        *(uint2 *)(sv + 8 * lane) = make_uint2(v[0], v[1]);
                        __syncwarp();
                        warp_store_shifted(tu, sy, seg_px >> 1, lane);
                        warp_store_shifted(tv, sv, seg_px >> 1, lane);
                    }
                }
            }
            /* chroma terms of the 8 pairs this thread owns */
            int cr[8], cg[8], cb[8];
            const uint32_t uvw[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                cr[2 * j] = dp2a_lo(COEF_RV, uvw[j], RGB_CR);  cr[2 * j + 1] = dp2a_hi(COEF_RV, uvw[j], RGB_CR);
                cg[2 * j] = dp2a_lo(COEF_GUV, uvw[j], RGB_CG); cg[2 * j + 1] = dp2a_hi(COEF_GUV, uvw[j], RGB_CG);
                cb[2 * j] = dp2a_lo(COEF_BU, uvw[j], RGB_CB);  cb[2 * j + 1] = dp2a_hi(COEF_BU, uvw[j], RGB_CB);
            }
#pragma unroll
            for (int row = 0; row < 2; row++) {
                if (row == 1 && !two) break;
                const uint4 yy = row ? yb : ya;
                const uint32_t yw[4] = {yy.x, yy.y, yy.z, yy.w};
                if constexpr (ARGB) {
                    __syncwarp();
                    uint4 *s4 = (uint4 *)(st + lane * 80);                        /* 80-byte stride: conflict-free 16-byte stores */
#pragma unroll
                    for (int j = 0; j < 4; j++) {                                 /* one luma word = 4 pixels = one 16-byte store */
                        uint32_t o[4];
                        argb4(yw[j], cr[2 * j], cg[2 * j], cb[2 * j], cr[2 * j + 1], cg[2 * j + 1], cb[2 * j + 1], o);
                        s4[j] = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    __syncwarp();
                    uint8_t *g = orow + (size_t)row * p.rgb_pitch + (size_t)seg * (32 * 64);
                    const uint32_t nb = 4 * seg_px;                               /* a multiple of 8 */
                    if ((((uint32_t)(uintptr_t)g | nb) & 15) == 0) {
                        /* 16-byte chunk c = 32k + lane lives at stage lane c/4, part c%4: per-lane bases once, constant steps */
                        const uint8_t *sl = st + (lane >> 2) * 80 + (lane & 3) * 16;
                        uint8_t *gl = g + 16 * lane;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (512 * k + 16 * lane < nb) st16<C::STP>(gl + 512 * k, *(const uint4 *)(sl + 640 * k));
                    } else {                                                      /* 8-byte aligned rows (w % 4 == 2) or any other pitch */
                        warp_store_shifted_map(g, [st](uint32_t c) { return (const uint4 *)(st + (c >> 2) * 80 + (c & 3) * 16); }, nb, lane);
                    }
                } else {
                    uint32_t o[12];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        rgb4(yw[j], cr[2 * j], cg[2 * j], cb[2 * j], cr[2 * j + 1], cg[2 * j + 1], cb[2 * j + 1], o + 3 * j);
                    __syncwarp();
                    uint4 *s4 = (uint4 *)(st + lane * 48);
                    s4[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    s4[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    s4[2] = make_uint4(o[8], o[9], o[10], o[11]);
                    __syncwarp();
                    uint8_t *g = orow + (size_t)row * p.rgb_pitch + (size_t)seg * (32 * 48);
                    const uint32_t nbytes = 3 * seg_px;                               /* a multiple of 6 */
                    /* aligned rows: straight 16-byte stores; any other alignment (1080- or 1366-wide video): 16-byte
                     * stores to the aligned body, re-aligned from shared memory by a funnel shift */
                    const uint32_t gb = (uint32_t)(uintptr_t)g | nbytes;
                    if ((gb & 15) == 0) warp_flush<16, C::STP>(g, st, nbytes, lane);
                    else if ((gb & 7) == 0) warp_flush<8, C::STP>(g, st, nbytes, lane);      /* 1080-wide: measured 0.82 vs 0.74 shifted */
                    else warp_store_shifted(g, st, nbytes, lane);
                }
            }
        } else {
            /* odd widths, unaligned or too-tight surfaces: one pixel per lane per step, byte accesses */
            const uint32_t x_begin = seg * 512, x_end = min((uint32_t)w, x_begin + 512);
            for (uint32_t x = x_begin + lane; x < x_end; x += 32) {
                const uint32_t cx = min(x >> 1, (uint32_t)(cw - 1));
                const int U = crow[2 * cx], V = crow[2 * cx + 1];
                const int d = U - 128, e = V - 128;
                for (uint32_t row = 0; row < (two ? 2u : 1u); row++) {
                    const int Y = yrow[(size_t)row * p.pitch + x];
                    const int c = Y - 16;
                    const uint8_t R = clip8_dev((298 * c + 409 * e + 128) >> 8);
                    const uint8_t G = clip8_dev((298 * c - 100 * d - 208 * e + 128) >> 8);
                    const uint8_t Bl = clip8_dev((298 * c + 516 * d + 128) >> 8);
                    if (ARGB) {
                        uint8_t *o = orow + (size_t)row * p.rgb_pitch + 4 * (size_t)x;
                        o[0] = Bl; o[1] = G; o[2] = R; o[3] = 0xFF;
                    } else {
                        uint8_t *o = orow + (size_t)row * p.rgb_pitch + 3 * (size_t)x;
                        o[0] = R; o[1] = G; o[2] = Bl;
                    }
                    if (p.fused) tp[(size_t)(y0 + row) * w + x] = (uint8_t)Y;
                }
                if (p.fused && rp < (uint32_t)ch && (x & 1) == 0 && (x >> 1) < (uint32_t)cw) {
                    tp[p.u_off + (size_t)rp * cw + (x >> 1)] = (uint8_t)U;
                    tp[p.v_off + (size_t)rp * cw + (x >> 1)] = (uint8_t)V;
                }
            }
        }
    }
}


/* ---- bulk-copy-engine variant of the RGB kernel ------------------------------------------------
 * One CTA per (frame, row pair, column segment of <= 2048 pixels): three bulk loads (two luma rows,
 * one chroma row) into shared memory, threads convert shared -> shared (same dp2a / cvt.pack.sat
 * arithmetic as above, 16 pixels x 2 rows per step), then two bulk stores of 3*seg bytes (plus, fused:
 * the two luma rows straight from the input buffer and the de-interleaved U / V rows).
 * 10*seg_w bytes of shared memory (<= 20 KB, ~11 CTAs per SM); everything 16-byte aligned, host-checked. */
struct RgbBulkParams {
    FrameSet surf, tight, rgb;
    uint32_t n_frames;
    int32_t width, height, pitch;
    int64_t y_off, uv_off;
    int64_t u_off, v_off;
    int32_t rgb_pitch;
    int32_t fused;
    uint32_t row_pairs;
    uint32_t segs;            /* column segments per row pair */
    uint32_t seg_w;           /* pixels per segment (multiple of 32); the last one takes the remainder */
};

constexpr int RGB_BULK_THREADS = 128;

/* ALIGNED: every RGB / tight row is a 16-byte-aligned multiple of 16 bytes and leaves through the copy engine.
 * !ALIGNED: only the surface is aligned (any even width): rows are loaded rounded up to 16 bytes (inside the
 * pitch) and the four warps write the RGB / tight rows with re-aligned 16-byte stores (warp_store_shifted). */
template <bool ALIGNED>
__global__ void __launch_bounds__(RGB_BULK_THREADS) rgb_bulk_kernel(const __grid_constant__ RgbBulkParams p)
{
    extern __shared__ __align__(128) uint8_t rs[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t per_frame = p.row_pairs * p.segs;
    const uint32_t f = blockIdx.x / per_frame;
    const uint32_t t = blockIdx.x - f * per_frame;
    const uint32_t rp = t / p.segs, seg = t - rp * p.segs;
    const uint32_t W = (uint32_t)p.width, h = (uint32_t)p.height, cw = W >> 1, ch = h >> 1;
    const uint32_t x0 = seg * p.seg_w;                     /* first pixel of this segment */
    const uint32_t w = min(p.seg_w, W - x0);               /* pixels in this segment (ALIGNED: a multiple of 16; else even) */
    const uint32_t lw = ALIGNED ? w : ((w + 15) & ~15u);   /* bytes loaded per row */
    const uint32_t y0 = rp * 2;
    const bool two = y0 + 1 < h;
    const uint32_t cy = min(rp, ch - 1);
    const uint8_t *sp = frame_ptr(p.surf, f);
    uint8_t *rgbp = frame_ptr(p.rgb, f) + (size_t)y0 * p.rgb_pitch + 3 * (size_t)x0;
    const uint32_t sw = p.seg_w;                           /* shared-memory row stride */
    uint8_t *s_y = rs;                    /* 2*sw : luma rows y0, y0+1 */
    uint8_t *s_uv = rs + 2 * (size_t)sw;  /* sw   */
    uint8_t *s_rgb = rs + 3 * (size_t)sw; /* 6*sw : two RGB rows */
    uint8_t *s_u = rs + 9 * (size_t)sw;   /* sw/2 + sw/2 (fused) */
    uint8_t *s_v = s_u + (sw >> 1);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, (two ? 3u : 2u) * lw);
        const uint8_t *yrow = sp + p.y_off + (size_t)y0 * p.pitch + x0;
        bulk_g2s(s_y, yrow, lw, &bar);
        if (two) bulk_g2s(s_y + sw, yrow + p.pitch, lw, &bar);
        bulk_g2s(s_uv, sp + p.uv_off + (size_t)cy * p.pitch + x0, lw, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    const bool do_uv = p.fused && rp < ch;
    for (uint32_t unit = threadIdx.x; unit < (lw >> 4); unit += RGB_BULK_THREADS) {
        const uint4 uv = *(const uint4 *)(s_uv + unit * 16);
        const uint32_t uvw[4] = {uv.x, uv.y, uv.z, uv.w};
        int cr[8], cg[8], cb[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            cr[2 * j] = dp2a_lo(COEF_RV, uvw[j], RGB_CR);  cr[2 * j + 1] = dp2a_hi(COEF_RV, uvw[j], RGB_CR);
            cg[2 * j] = dp2a_lo(COEF_GUV, uvw[j], RGB_CG); cg[2 * j + 1] = dp2a_hi(COEF_GUV, uvw[j], RGB_CG);
            cb[2 * j] = dp2a_lo(COEF_BU, uvw[j], RGB_CB);  cb[2 * j + 1] = dp2a_hi(COEF_BU, uvw[j], RGB_CB);
        }
        if (do_uv) {
            uint2 u, v;
            u.x = __byte_perm(uv.x, uv.y, 0x6420); v.x = __byte_perm(uv.x, uv.y, 0x7531);
            u.y = __byte_perm(uv.z, uv.w, 0x6420); v.y = __byte_perm(uv.z, uv.w, 0x7531);
            *(uint2 *)(s_u + unit * 8) = u;
            *(uint2 *)(s_v + unit * 8) = v;
        }
#pragma unroll
        for (int row = 0; row < 2; row++) {
            if (row == 1 && !two) break;
            const uint4 yy = *(const uint4 *)(s_y + (size_t)row * sw + unit * 16);
            const uint32_t yw[4] = {yy.x, yy.y, yy.z, yy.w};
            uint32_t o[12];
#pragma unroll
            for (int j = 0; j < 4; j++)
                rgb4(yw[j], cr[2 * j], cg[2 * j], cb[2 * j], cr[2 * j + 1], cg[2 * j + 1], cb[2 * j + 1], o + 3 * j);
            uint4 *d = (uint4 *)(s_rgb + (size_t)row * 3 * sw + unit * 48);
            d[0] = make_uint4(o[0], o[1], o[2], o[3]);
            d[1] = make_uint4(o[4], o[5], o[6], o[7]);
            d[2] = make_uint4(o[8], o[9], o[10], o[11]);
        }
    }
    if (ALIGNED) {
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_s2g(rgbp, s_rgb, 3 * w);
            if (two) bulk_s2g(rgbp + p.rgb_pitch, s_rgb + 3 * (size_t)sw, 3 * w);
            if (p.fused) {
                uint8_t *tp = frame_ptr(p.tight, f);
                bulk_s2g(tp + (size_t)y0 * W + x0, s_y, w);
                if (two) bulk_s2g(tp + (size_t)(y0 + 1) * W + x0, s_y + sw, w);
                if (do_uv) {
                    bulk_s2g(tp + p.u_off + (size_t)rp * cw + (x0 >> 1), s_u, w >> 1);
                    bulk_s2g(tp + p.v_off + (size_t)rp * cw + (x0 >> 1), s_v, w >> 1);
                }
            }
            bulk_commit_wait_read();
        }
    } else {
        __syncthreads();
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t row = warp & 1;                                 /* warps 0,2: first row; 1,3: second row */
        if (p.fused) {
            if (warp < 2) {
                if (row == 0 || two) warp_store_shifted(rgbp + (size_t)row * p.rgb_pitch, s_rgb + (size_t)row * 3 * sw, 3 * w, lane);
            } else {
                uint8_t *tp = frame_ptr(p.tight, f);
                if (row == 0 || two) warp_store_shifted(tp + (size_t)(y0 + row) * W + x0, s_y + (size_t)row * sw, w, lane);
                if (do_uv) warp_store_shifted(tp + (row ? p.v_off : p.u_off) + (size_t)rp * cw + (x0 >> 1), row ? s_v : s_u, w >> 1, lane);
            }
        } else if (row == 0 || two) {
            const uint32_t half = ((3 * w) >> 1) & ~15u;               /* each RGB row is shared by two warps */
            const uint32_t b0 = warp < 2 ? 0u : half, b1 = warp < 2 ? half : 3 * w;
            warp_store_shifted(rgbp + (size_t)row * p.rgb_pitch + b0, s_rgb + (size_t)row * 3 * sw + b0, b1 - b0, lane);
        }
    }
}

/* ========================================================================================== */
/* RGB24 -> pitched NV12 (forward integer BT.601, chroma from 2x2 block sums)                  */
/* ========================================================================================== */
/* One warp per (row pair, 512-pixel segment).  The RGB rows sit at arbitrary addresses (3*w bytes per
 * row), so they come in through ShiftedLoad (aligned 16-byte loads, all in flight together, re-aligned by
 * shuffle + funnel shift into warp-private shared memory); each lane then owns 16 pixels x 2 rows = 2 x 48
 * bytes.  A pixel is cut out of its three-word group with one prmt (the fourth byte meets a zero
 * coefficient), Y is one dp4a per pixel, U and V four dp4a each per 2x2 block (dp4a is linear, so the
 * block sum never has to be formed).  Surface rows get 16 bytes per lane; prefix stores at the row end
 * keep the padding untouched. */
struct Rgb2Params {
    FrameSet rgb, surf;
    uint32_t n_frames;
    int32_t width, height, pitch, rgb_pitch;
    int64_t y_off, uv_off;
    uint32_t row_pairs, segs_per_row, tasks_per_frame, total_tasks;
    FastDiv tpf_div, seg_div;  /* division by tasks_per_frame / segs_per_row */
};

constexpr uint32_t FWD_Y = 66u | (129u << 8) | (25u << 16);                 /* R,G,B -> Y, unsigned bytes */
constexpr uint32_t FWD_U = 0xDAu | (0xB6u << 8) | (0x70u << 16);            /* -38, -74, 112 as signed bytes */
constexpr uint32_t FWD_V = 0x70u | (0xA2u << 8) | (0xEEu << 16);            /* 112, -94, -18 */
constexpr int FWD_Y_BIAS = 128 + 16 * 256, FWD_C_BIAS = 512 + 128 * 1024;

__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}

/* the four pixels (R,G,B,x) of a 12-byte group a,b,c */
__device__ __forceinline__ void cut4(uint32_t a, uint32_t b, uint32_t c, uint32_t (&px)[4])
{
    px[0] = a;
    px[1] = __byte_perm(a, b, 0x6543);
    px[2] = __byte_perm(b, c, 0x5432);
    px[3] = c >> 8;
}

constexpr int RGB2_THREADS = 128;
constexpr int RGB2_ROW = 1536 + 32;                   /* staged bytes per RGB row segment + spare chunks */

__global__ void __launch_bounds__(RGB2_THREADS, 8) rgb_to_nv12_kernel(const __grid_constant__ Rgb2Params p)
{
    constexpr int WARPS = RGB2_THREADS / 32;
    __shared__ __align__(16) uint8_t stage[WARPS][2 * RGB2_ROW];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t task = blockIdx.x * WARPS + wib;
    if (task >= p.total_tasks) return;
    const uint32_t f = fast_div(task, p.tpf_div);
    const uint32_t r = task - f * p.tasks_per_frame;
    const uint32_t rp = fast_div(r, p.seg_div), seg = r - rp * p.segs_per_row;
    const uint32_t w = (uint32_t)p.width, h = (uint32_t)p.height, cw = w >> 1, ch = h >> 1;
    const uint32_t y0 = 2 * rp, x0 = seg * 512;
    const bool two = y0 + 1 < h, do_uv = rp < ch;
    const uint32_t seg_px = min(512u, w - x0);
    const uint8_t *src = frame_ptr(p.rgb, f) + (size_t)y0 * p.rgb_pitch + 3 * (size_t)x0;
    uint8_t *sp = frame_ptr(p.surf, f);
    uint8_t *A0 = stage[wib], *A1 = A0 + RGB2_ROW;
    {
        ShiftedLoad<3> l0, l1;
        l0.issue(src, 3 * seg_px, lane);
        if (two) l1.issue(src + p.rgb_pitch, 3 * seg_px, lane);
        l0.commit(A0, 3 * seg_px, lane);
        if (two) l1.commit(A1, 3 * seg_px, lane);
    }
    __syncwarp();
    const uint32_t px0 = 16 * lane;
    if (px0 >= seg_px) return;
    const uint32_t npx = min(16u, seg_px - px0);
    uint32_t ya[4], yb[4], uvw[4];
    uint32_t r0[12], r1[12];                                                /* this lane's 16 pixels of both rows */
#pragma unroll
    for (int m = 0; m < 3; m++) {                                           /* 48-byte lane stride: conflict-free LDS.128 */
        const uint4 a = *(const uint4 *)(A0 + 48 * lane + 16 * m);
        const uint4 b = two ? *(const uint4 *)(A1 + 48 * lane + 16 * m) : make_uint4(0, 0, 0, 0);
        r0[4 * m] = a.x; r0[4 * m + 1] = a.y; r0[4 * m + 2] = a.z; r0[4 * m + 3] = a.w;
        r1[4 * m] = b.x; r1[4 * m + 1] = b.y; r1[4 * m + 2] = b.z; r1[4 * m + 3] = b.w;
    }
#pragma unroll
    for (int g = 0; g < 4; g++) {                                           /* 4 pixels = 12 bytes per row */
        uint32_t pa[4], pb[4];
        cut4(r0[3 * g], r0[3 * g + 1], r0[3 * g + 2], pa);
        cut4(r1[3 * g], r1[3 * g + 1], r1[3 * g + 2], pb);
        uint32_t t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = __dp4a(pa[k], FWD_Y, (uint32_t)FWD_Y_BIAS);       /* Y in byte 1 */
        ya[g] = __byte_perm(__byte_perm(t[0], t[1], 0x0051), __byte_perm(t[2], t[3], 0x0051), 0x5410);
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = __dp4a(pb[k], FWD_Y, (uint32_t)FWD_Y_BIAS);
        yb[g] = __byte_perm(__byte_perm(t[0], t[1], 0x0051), __byte_perm(t[2], t[3], 0x0051), 0x5410);
        uint32_t c[4];                                                      /* U0 V0 U1 V1 of the two 2x2 blocks */
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int u = dp4a_us(pa[2 * b], FWD_U, dp4a_us(pa[2 * b + 1], FWD_U, dp4a_us(pb[2 * b], FWD_U, dp4a_us(pb[2 * b + 1], FWD_U, FWD_C_BIAS))));
            const int v = dp4a_us(pa[2 * b], FWD_V, dp4a_us(pa[2 * b + 1], FWD_V, dp4a_us(pb[2 * b], FWD_V, dp4a_us(pb[2 * b + 1], FWD_V, FWD_C_BIAS))));
            c[2 * b] = (uint32_t)u >> 10;
            c[2 * b + 1] = (uint32_t)v >> 10;
        }
        uvw[g] = __byte_perm(__byte_perm(c[0], c[1], 0x0040), __byte_perm(c[2], c[3], 0x0040), 0x5410);
    }
    uint8_t *yrow = sp + p.y_off + (size_t)y0 * p.pitch + x0 + px0;
    store_prefix<4>(yrow, ya, npx);
    if (two) store_prefix<4>(yrow + p.pitch, yb, npx);
    if (do_uv) {
        const uint32_t pair0 = (x0 + px0) >> 1;                             /* first chroma pair of this lane */
        if (pair0 < cw) store_prefix<4>(sp + p.uv_off + (size_t)rp * p.pitch + x0 + px0, uvw, 2 * min(8u, cw - pair0));
    }
}

} /* namespace jmc */
