/*
 * jmc_k_rows.cuh -- widths that are NOT multiples of 16 on aligned surfaces: the re-aligning
 * building blocks (shift_pair_ws, warp_store_shifted, ShiftedLoad), the warp-per-row LDG kernel
 * (rows_kernel) and the bulk-loaded tile kernels (bulk_rows_kernel, bulk_rows_pack_kernel).
 */
#pragma once
#include "jmc_k_common.cuh"

namespace jmc {

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
    uint32_t pad_zero;        /* encode direction: the padding behind a row may be zeroed up to the next 16-byte boundary (JMC_JOB_PAD_ZERO) */
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
    mbar_wait_cta(&bar, 0);

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

/* 16 bytes at chunk q (aligned) or at a row-uniform byte shift behind it; WS < 0: the row starts on a chunk boundary */
template <int WS> __device__ __forceinline__ uint4 staged16_ws(const uint4 *q, uint32_t sh)
{
    if (WS < 0) return q[0];
    return shift_pair_ws<(WS < 0 ? 0 : WS)>(q[0], q[1], sh);
}
/* word shift of a row that starts `off` bytes into the staging buffer: -1 when it starts on a 16-byte chunk boundary */
__device__ __forceinline__ int row_ws(uint32_t off) { return (off & 15u) == 0 ? -1 : (int)((off & 15u) >> 2); }

/* The first n (0..16) bytes of `data`, the rest from `old` */
__device__ __forceinline__ uint4 blend16(const uint4 &data, const uint4 &old, uint32_t n)
{
    const uint32_t dw[4] = {data.x, data.y, data.z, data.w}, ow[4] = {old.x, old.y, old.z, old.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t have = n > 4u * k ? min(n - 4u * k, 4u) : 0u;          /* data bytes in word k */
        const uint32_t mask = have >= 4 ? 0xffffffffu : ((1u << (8 * have)) - 1u);
        r[k] = (dw[k] & mask) | (ow[k] & ~mask);
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
}
/* plain (coherent) 16-byte global load / store: for bytes this kernel reads and then writes back */
__device__ __forceinline__ uint4 ld16_plain(const void *p)
{
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

/* One surface chroma row from its U and V rows (re pairs each, staged at byte offsets offu / offv): whole 32-byte
 * outputs in the loop -- the two row shifts are template parameters, so the loop body is 4 LDS.128, 8 SHF, 8 PRMT,
 * 2 STG.128 and nothing else; the < 32 bytes at the row end are the kernel's row-end pass.
 * Round 1 took the shifts as run-time values per chunk and ended every row with two store_prefix() calls under
 * divergence: ncu counted ~80 warp instructions per 512 output bytes (profiles/README.md, round 2). */
template <int WSU, int WSV>
__device__ __forceinline__ void merge_row_ws(uint8_t *d, const uint8_t *Su, uint32_t offu, const uint8_t *Sv, uint32_t offv, uint32_t re, uint32_t lane)
{
    const uint4 *qu = (const uint4 *)Su + (offu >> 4), *qv = (const uint4 *)Sv + (offv >> 4);
    const uint32_t shu = 8 * (offu & 3), shv = 8 * (offv & 3);
    const uint32_t nfull = re >> 4;
    for (uint32_t j = lane; j < nfull; j += 32) {
        const uint4 u = staged16_ws<WSU>(qu + j, shu), w = staged16_ws<WSV>(qv + j, shv);
        uint4 lo, hi;
        lo.x = __byte_perm(u.x, w.x, 0x5140); lo.y = __byte_perm(u.x, w.x, 0x7362);
        lo.z = __byte_perm(u.y, w.y, 0x5140); lo.w = __byte_perm(u.y, w.y, 0x7362);
        hi.x = __byte_perm(u.z, w.z, 0x5140); hi.y = __byte_perm(u.z, w.z, 0x7362);
        hi.z = __byte_perm(u.w, w.w, 0x5140); hi.w = __byte_perm(u.w, w.w, 0x7362);
        *(uint4 *)(d + 32 * (size_t)j) = lo;
        *(uint4 *)(d + 32 * (size_t)j + 16) = hi;
    }
}
template <int WSU>
__device__ __forceinline__ void merge_row_v(int wsv, uint8_t *d, const uint8_t *Su, uint32_t offu, const uint8_t *Sv, uint32_t offv, uint32_t re, uint32_t lane)
{
    switch (wsv) {                                                /* warp-uniform, once per row */
    case -1: merge_row_ws<WSU, -1>(d, Su, offu, Sv, offv, re, lane); break;
    case 0: merge_row_ws<WSU, 0>(d, Su, offu, Sv, offv, re, lane); break;
    case 1: merge_row_ws<WSU, 1>(d, Su, offu, Sv, offv, re, lane); break;
    case 2: merge_row_ws<WSU, 2>(d, Su, offu, Sv, offv, re, lane); break;
    default: merge_row_ws<WSU, 3>(d, Su, offu, Sv, offv, re, lane); break;
    }
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
        /* Row ends.  A row of re bytes ends re & 15 bytes into a 16-byte chunk of the surface; the rest of that chunk is
         * pitch padding, which must keep its value.  Writing just the re & 15 bytes costs far more than its share: rows
         * that end on a sub-16-byte store run at 0.87 of the roofline where rows ending on a chunk boundary reach 0.98
         * (profiles/r2_partial_sector_probe.txt) -- the L2 has to merge every such fragment.  So lane k of each warp
         * fetches the old chunk of the warp's k-th row NOW (the latency hides behind the bulk load), and after the
         * copy writes the chunk back whole: the row's last bytes blended over the padding's own bytes. */
        const uint32_t nfull = re >> 4, tail = re & 15u;
        const uint32_t my_row = warp + (BROWS_THREADS / 32) * lane;            /* the row whose end this lane owns */
        uint4 old = make_uint4(0, 0, 0, 0);
        if (tail && my_row < nr && !p.pad_zero) old = ld16_plain(pp + (size_t)(r0 + my_row) * pitch + 16 * (size_t)nfull);
#if JMC_MBAR_POLL
        if (run.body && warp == 0) mbar_wait(&bar, 0);
        __syncthreads();
#else
        __syncthreads();
        if (run.body) mbar_wait(&bar, 0);
#endif
        for (uint32_t i = warp; i < nr; i += BROWS_THREADS / 32) {
            uint8_t *d = pp + (size_t)(r0 + i) * pitch;
            const uint32_t off = run.a + i * re;
            const uint4 *q0 = (const uint4 *)S + (off >> 4);                  /* the row starts off & 15 bytes into this chunk */
            const uint32_t sh = 8 * (off & 3);
#define JMC_PACK_ROW(WS)                                                                                       \
            for (uint32_t j = lane; j < nfull; j += 32) *(uint4 *)(d + 16 * (size_t)j) = staged16_ws<WS>(q0 + j, sh);
            switch (row_ws(off)) {                                             /* warp-uniform, once per row */
            case -1: JMC_PACK_ROW(-1) break;
            case 0: JMC_PACK_ROW(0) break;
            case 1: JMC_PACK_ROW(1) break;
            case 2: JMC_PACK_ROW(2) break;
            default: JMC_PACK_ROW(3) break;
            }
#undef JMC_PACK_ROW
        }
        if (tail && my_row < nr) {
            const uint32_t off = run.a + my_row * re + 16 * nfull;           /* the row's last bytes in the staging buffer */
            const uint4 *q = (const uint4 *)S + (off >> 4);
            const uint4 data = (off & 15) ? shift_pair(q[0], q[1], (off & 15) >> 2, 8 * (off & 3)) : q[0];
            *(uint4 *)(pp + (size_t)(r0 + my_row) * pitch + 16 * (size_t)nfull) = blend16(data, old, tail);
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
        /* row ends as above: 2*(re & 15) interleaved bytes behind the last whole 32; the chunk they end in is read now
         * and written back whole after the merge */
        const uint32_t nfull = re >> 4, tail = 2 * (re & 15u);
        const uint32_t my_row = warp + (BROWS_THREADS / 32) * lane;
        const uint32_t mixed = tail > 16 ? 16u : 0u;          /* offset of the chunk that holds the row end */
        uint4 old = make_uint4(0, 0, 0, 0);
        if (tail && tail != 16 && my_row < nr && !p.pad_zero) old = ld16_plain(pp + (size_t)(r0 + my_row) * pitch + 32 * (size_t)nfull + mixed);
#if JMC_MBAR_POLL
        if ((ru.body | rv.body) && warp == 0) mbar_wait(&bar, 0);
        __syncthreads();
#else
        __syncthreads();
        if (ru.body | rv.body) mbar_wait(&bar, 0);
#endif
        for (uint32_t i = warp; i < nr; i += BROWS_THREADS / 32) {
            uint8_t *d = pp + (size_t)(r0 + i) * pitch;
            const uint32_t offu = ru.a + i * re, offv = rv.a + i * re;
            const int wsv = row_ws(offv);
            switch (row_ws(offu)) {                                            /* warp-uniform, once per row */
            case -1: merge_row_v<-1>(wsv, d, Su, offu, Sv, offv, re, lane); break;
            case 0: merge_row_v<0>(wsv, d, Su, offu, Sv, offv, re, lane); break;
            case 1: merge_row_v<1>(wsv, d, Su, offu, Sv, offv, re, lane); break;
            case 2: merge_row_v<2>(wsv, d, Su, offu, Sv, offv, re, lane); break;
            default: merge_row_v<3>(wsv, d, Su, offu, Sv, offv, re, lane); break;
            }
        }
        if (tail && my_row < nr) {
            const uint32_t offu = ru.a + my_row * re + 16 * nfull, offv = rv.a + my_row * re + 16 * nfull;
            const uint4 u = staged16(Su, offu), w = staged16(Sv, offv);
            uint4 lo, hi;
            lo.x = __byte_perm(u.x, w.x, 0x5140); lo.y = __byte_perm(u.x, w.x, 0x7362);
            lo.z = __byte_perm(u.y, w.y, 0x5140); lo.w = __byte_perm(u.y, w.y, 0x7362);
            hi.x = __byte_perm(u.z, w.z, 0x5140); hi.y = __byte_perm(u.z, w.z, 0x7362);
            hi.z = __byte_perm(u.w, w.w, 0x5140); hi.w = __byte_perm(u.w, w.w, 0x7362);
            uint8_t *d = pp + (size_t)(r0 + my_row) * pitch + 32 * (size_t)nfull;
            if (tail < 16) *(uint4 *)d = blend16(lo, old, tail);
            else {
                *(uint4 *)d = lo;
                if (tail > 16) *(uint4 *)(d + 16) = blend16(hi, old, tail - 16);
            }
        }
    }
}

} /* namespace jmc */
