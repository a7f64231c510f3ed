/*
 * jmc_k_common.cuh -- types and device helpers shared by every kernel of the surface-format path:
 * frame addressing, launch-invariant division, cache-policy loads/stores, the bulk-copy-engine
 * (cp.async.bulk + mbarrier) wrappers and the row-end prefix store.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jmc {


/* Where frame f of a batch lives: base + f*stride; or list[f], a DEVICE array of frame pointers; or, for small
 * batches (decoder-mapped surfaces drained a few at a time), inl[f]: the pointers themselves travel as kernel
 * arguments, so nothing has to be uploaded before the launch (JMC_JOB_LIST_ON_HOST). */
constexpr int INLINE_LIST_MAX = 8;
struct FrameSet {
    uint8_t *base;
    size_t stride;
    uint8_t *const *list;
    uint32_t n_inline, pad_;
    uint8_t *inl[INLINE_LIST_MAX];
};

__device__ __forceinline__ uint8_t *frame_ptr(const FrameSet &s, uint32_t f)
{
    if (s.n_inline) return s.inl[f];
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

/* ---- bulk-copy engine (cp.async.bulk, the 1-D form of TMA) + mbarrier ---------------------- */
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
/* JMC_MBAR_POLL: 0 = every thread of the CTA polls the mbarrier (round 1); 1 = the first warp polls, with the
 * hardware suspend-time hint, and the other warps sleep in the CTA barrier.  Polling warps execute instructions:
 * ncu counted ~1500 warp-instructions per CTA of pure try_wait / branch in the row kernels, competing for issue
 * slots with the CTAs that have data (profiles/README.md, round 2). */
#ifndef JMC_MBAR_POLL
#define JMC_MBAR_POLL 1
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
#if JMC_MBAR_POLL
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680) : "memory");
#else
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity) : "memory");
#endif
}
/* the whole CTA waits for the barrier's phase: one warp observes it, the CTA barrier hands the observation on */
__device__ __forceinline__ void mbar_wait_cta(uint64_t *bar, uint32_t parity)
{
#if JMC_MBAR_POLL
    if (threadIdx.x < 32) mbar_wait(bar, parity);
    __syncthreads();
#else
    mbar_wait(bar, parity);
#endif
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

} /* namespace jmc */
