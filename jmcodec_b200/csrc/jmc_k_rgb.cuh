/*
 * jmc_k_rgb.cuh -- colour kernels (builder-defined integer BT.601): NV12 -> RGB24 / ARGB32 with optional
 * fused I420 output (rgb_kernel, rgb_bulk_kernel) and RGB24 -> pitched NV12 (rgb_to_nv12_kernel).
 */
#pragma once
#include "jmc_k_common.cuh"
#include "jmc_k_rows.cuh"

namespace jmc {

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


/* ---- flattened variant: every lane slot busy whatever the width --------------------------------
 * rgb_kernel gives a warp 512 pixels of ONE row pair, so a 1366-wide row pair costs 96 lane slots for 85.4
 * sixteen-pixel units and a 1080-wide one 96 for 67.5 - and at that utilisation the kernel is issue-bound.
 * Here the (row pair, unit) space of a frame is numbered through and a warp takes 32 CONSECUTIVE units, which
 * may end one row pair and begin the next: loads are per lane anyway; the staged RGB bytes of a warp then form
 * one run per row pair it touches (usually one or two), each flushed like rgb_kernel flushes its single run.
 * Host guarantees: aligned surface, even width, rows over-readable to 16 bytes, not the fused op. */
struct RgbFlatParams {
    FrameSet surf, rgb;
    uint32_t n_frames;
    int32_t width, height, pitch;
    int64_t y_off, uv_off;
    int32_t rgb_pitch;
    uint32_t units_per_row;    /* ceil(width / 16) */
    uint32_t units_per_frame;  /* units_per_row * ceil(height / 2) */
    uint32_t tasks_per_frame;  /* ceil(units_per_frame / 32) */
    FastDiv tpf_div, unit_div; /* division by tasks_per_frame / units_per_row */
    uint32_t total_tasks;
};

template <class C, bool ARGB>
__global__ void __launch_bounds__(C::THREADS, C::BLOCKS_PER_SM) rgb_flat_kernel(const __grid_constant__ RgbFlatParams p)
{
    constexpr int WARPS = C::THREADS / 32;
    constexpr uint32_t BPP = ARGB ? 4 : 3;
    __shared__ __align__(16) uint8_t stage[WARPS][32 * 80 + 64];     /* RGB24: 48 B per lane; ARGB32: 64 B at an 80-byte stride */
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t task = blockIdx.x * WARPS + wib;
    if (task >= p.total_tasks) return;
    const uint32_t f = fast_div(task, p.tpf_div);
    const uint32_t u0 = (task - f * p.tasks_per_frame) * 32;         /* first unit of this warp */
    const uint32_t w = (uint32_t)p.width, h = (uint32_t)p.height, ch = h >> 1;
    const uint32_t U = p.units_per_row, ulast = min(u0 + 31, p.units_per_frame - 1);
    const bool valid = u0 + lane <= ulast;
    const uint32_t u = valid ? u0 + lane : ulast;
    const uint32_t rp = fast_div(u, p.unit_div), px0 = (u - rp * U) * 16;
    const uint32_t y0 = 2 * rp;
    const bool two = y0 + 1 < h;
    const uint8_t *sp = frame_ptr(p.surf, f);
    uint8_t *rgbp = frame_ptr(p.rgb, f);
    const uint8_t *yrow = sp + p.y_off + (size_t)y0 * p.pitch + px0;
    uint4 ya = make_uint4(0, 0, 0, 0), yb = ya, uv = ya;
    if (valid) {
        ya = ld16<C::LDP>(yrow);
        uv = ld16<C::LDP>(sp + p.uv_off + (size_t)min(rp, ch - 1) * p.pitch + px0);
        if (two) yb = ld16<C::LDP>(yrow + p.pitch);
    }
    int cr[8], cg[8], cb[8];
    const uint32_t uvw[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        cr[2 * j] = dp2a_lo(COEF_RV, uvw[j], RGB_CR);  cr[2 * j + 1] = dp2a_hi(COEF_RV, uvw[j], RGB_CR);
        cg[2 * j] = dp2a_lo(COEF_GUV, uvw[j], RGB_CG); cg[2 * j + 1] = dp2a_hi(COEF_GUV, uvw[j], RGB_CG);
        cb[2 * j] = dp2a_lo(COEF_BU, uvw[j], RGB_CB);  cb[2 * j + 1] = dp2a_hi(COEF_BU, uvw[j], RGB_CB);
    }
    uint8_t *st = stage[wib];
    const uint32_t rp_a = fast_div(u0, p.unit_div), rp_b = fast_div(ulast, p.unit_div);   /* row pairs this warp touches */
#pragma unroll
    for (int row = 0; row < 2; row++) {
        const uint4 yy = row ? yb : ya;
        const uint32_t yw[4] = {yy.x, yy.y, yy.z, yy.w};
        __syncwarp();
        if constexpr (ARGB) {
            uint4 *s4 = (uint4 *)(st + lane * 80);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t o[4];
                argb4(yw[j], cr[2 * j], cg[2 * j], cb[2 * j], cr[2 * j + 1], cg[2 * j + 1], cb[2 * j + 1], o);
                s4[j] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        } else {
            uint32_t o[12];
#pragma unroll
            for (int j = 0; j < 4; j++)
                rgb4(yw[j], cr[2 * j], cg[2 * j], cb[2 * j], cr[2 * j + 1], cg[2 * j + 1], cb[2 * j + 1], o + 3 * j);
            uint4 *s4 = (uint4 *)(st + lane * 48);
            s4[0] = make_uint4(o[0], o[1], o[2], o[3]);
            s4[1] = make_uint4(o[4], o[5], o[6], o[7]);
            s4[2] = make_uint4(o[8], o[9], o[10], o[11]);
        }
        __syncwarp();
        for (uint32_t r = rp_a; r <= rp_b; r++) {                        /* one run per row pair; all of it warp-uniform */
            if (2 * r + row >= h) continue;                              /* odd height: the last row pair has one row */
            const uint32_t ua = max(u0, r * U), ub = min(ulast + 1, (r + 1) * U);          /* units [ua, ub) of row pair r */
            const uint32_t lo = ua - u0;                                 /* first lane of the run */
            const uint32_t px_lo = (ua - r * U) * 16, px_hi = min(w, (ub - r * U) * 16);
            const uint32_t nb = BPP * (px_hi - px_lo);
            uint8_t *g = rgbp + (size_t)(2 * r + row) * p.rgb_pitch + (size_t)BPP * px_lo;
            const uint32_t gb = (uint32_t)(uintptr_t)g | nb;
            if constexpr (ARGB) {
                if ((gb & 15) == 0) {
                    const uint8_t *sl = st + (lo + (lane >> 2)) * 80 + (lane & 3) * 16;
                    uint8_t *gl = g + 16 * lane;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (512 * k + 16 * lane < nb) st16<C::STP>(gl + 512 * k, *(const uint4 *)(sl + 640 * k));
                } else {
                    warp_store_shifted_map(g, [st, lo](uint32_t c) { return (const uint4 *)(st + (lo + (c >> 2)) * 80 + (c & 3) * 16); }, nb, lane);
                }
            } else {
                const uint8_t *s0 = st + lo * 48;
                if ((gb & 15) == 0) warp_flush<16, C::STP>(g, s0, nb, lane);
                else if ((gb & 7) == 0) warp_flush<8, C::STP>(g, s0, nb, lane);
                else warp_store_shifted(g, s0, nb, lane);
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
    /* rgb_bulk_pairs_kernel only */
    uint32_t pairs;           /* row pairs per CTA (segs == 1) */
    uint32_t groups;          /* ceil(row_pairs / pairs) */
    FastDiv upr_div;          /* division by the 16-pixel units per row */
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
    mbar_wait_cta(&bar, 0);
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

/* The re-aligning variant (!ALIGNED above) with `pairs` consecutive row pairs of whole rows per CTA: every row of the
 * tile arrives on one barrier, the units of all pairs are dealt out to the threads as one flat range, and each warp
 * writes whole rows. */
__global__ void __launch_bounds__(RGB_BULK_THREADS) rgb_bulk_pairs_kernel(const __grid_constant__ RgbBulkParams p)
{
    extern __shared__ __align__(128) uint8_t rs[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t per_frame = p.groups * p.segs;
    const uint32_t f = blockIdx.x / per_frame;
    const uint32_t t = blockIdx.x - f * per_frame;
    const uint32_t grp = t / p.segs, seg = t - grp * p.segs;
    const uint32_t W = (uint32_t)p.width, h = (uint32_t)p.height, cw = W >> 1, ch = h >> 1;
    const uint32_t x0 = seg * p.seg_w;                     /* first pixel of this segment */
    const uint32_t w = min(p.seg_w, W - x0);               /* pixels in this segment (even) */
    const uint32_t lw = (w + 15) & ~15u;                   /* bytes loaded per row */
    const uint32_t pairs = p.pairs;
    const uint32_t rp0 = grp * pairs;                      /* first row pair of this CTA */
    const uint32_t np = min(pairs, p.row_pairs - rp0);     /* row pairs of this CTA */
    const uint8_t *sp = frame_ptr(p.surf, f);
    uint8_t *rgb0 = frame_ptr(p.rgb, f) + 3 * (size_t)x0;
    const uint32_t sw = p.seg_w;                           /* shared-memory row stride */
    const size_t pair_bytes = 10 * (size_t)sw;             /* per row pair: 2*sw luma, sw chroma, 6*sw RGB, sw U + V (fused) */
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        uint32_t rows = 0;
        for (uint32_t i = 0; i < np; i++) rows += (2 * (rp0 + i) + 1 < h) ? 3u : 2u;
        mbar_expect_tx(&bar, rows * lw);
        for (uint32_t i = 0; i < np; i++) {
            const uint32_t rp = rp0 + i, y0 = 2 * rp, cy = min(rp, ch - 1);
            uint8_t *s_y = rs + i * pair_bytes;
            const uint8_t *yrow = sp + p.y_off + (size_t)y0 * p.pitch + x0;
            bulk_g2s(s_y, yrow, lw, &bar);
            if (y0 + 1 < h) bulk_g2s(s_y + sw, yrow + p.pitch, lw, &bar);
            bulk_g2s(s_y + 2 * (size_t)sw, sp + p.uv_off + (size_t)cy * p.pitch + x0, lw, &bar);
        }
    }
    __syncthreads();
    mbar_wait_cta(&bar, 0);
    const uint32_t upr = lw >> 4;
    for (uint32_t u = threadIdx.x; u < np * upr; u += RGB_BULK_THREADS) {
        const uint32_t i = pairs > 1 ? fast_div(u, p.upr_div) : 0u;
        const uint32_t unit = u - i * upr;
        const uint32_t rp = rp0 + i;
        const bool two = 2 * rp + 1 < h, do_uv = p.fused && rp < ch;
        uint8_t *s_y = rs + i * pair_bytes, *s_uv = s_y + 2 * (size_t)sw, *s_rgb = s_y + 3 * (size_t)sw;
        uint8_t *s_u = s_y + 9 * (size_t)sw, *s_v = s_u + (sw >> 1);
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
            uint2 uu, vv;
            uu.x = __byte_perm(uv.x, uv.y, 0x6420); vv.x = __byte_perm(uv.x, uv.y, 0x7531);
            uu.y = __byte_perm(uv.z, uv.w, 0x6420); vv.y = __byte_perm(uv.z, uv.w, 0x7531);
            *(uint2 *)(s_u + unit * 8) = uu;
            *(uint2 *)(s_v + unit * 8) = vv;
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
    {
        /* whole rows are dealt out to the warps.  Jobs per row pair: RGB row 0, RGB row 1 and, fused, luma row 0, luma
         * row 1, U row, V row.  Two row pairs give the four warps one RGB row each (RGB alone) or 4.5 w bytes each
         * (fused: 3w + w/2 + w twice, w + 3w + w/2 twice); splitting rows between warps only adds partial head / tail
         * stores (profiles/r2_rgb_tile_shape_sweep.txt). */
        __syncthreads();
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t jobs_per_pair = p.fused ? 6u : 2u;
        uint8_t *tp = p.fused ? frame_ptr(p.tight, f) : nullptr;
        for (uint32_t j = warp; j < np * jobs_per_pair; j += RGB_BULK_THREADS / 32) {
            const uint32_t i = j / jobs_per_pair, kind = j - i * jobs_per_pair;
            const uint32_t rp = rp0 + i, y0 = 2 * rp;
            const bool two = y0 + 1 < h, do_uv = rp < ch;
            uint8_t *s_y = rs + i * pair_bytes, *s_rgb = s_y + 3 * (size_t)sw, *s_u = s_y + 9 * (size_t)sw, *s_v = s_u + (sw >> 1);
            uint8_t *rgbp = rgb0 + (size_t)y0 * p.rgb_pitch;
            if (kind == 0) warp_store_shifted(rgbp, s_rgb, 3 * w, lane);
            else if (kind == 1) { if (two) warp_store_shifted(rgbp + p.rgb_pitch, s_rgb + 3 * (size_t)sw, 3 * w, lane); }
            else if (kind == 2) warp_store_shifted(tp + (size_t)y0 * W + x0, s_y, w, lane);
            else if (kind == 3) { if (two) warp_store_shifted(tp + (size_t)(y0 + 1) * W + x0, s_y + sw, w, lane); }
            else if (kind == 4) { if (do_uv) warp_store_shifted(tp + p.u_off + (size_t)rp * cw + (x0 >> 1), s_u, w >> 1, lane); }
            else { if (do_uv) warp_store_shifted(tp + p.v_off + (size_t)rp * cw + (x0 >> 1), s_v, w >> 1, lane); }
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
    uint32_t pad_zero;         /* the padding behind a row may be zeroed up to the next 16-byte boundary (JMC_JOB_PAD_ZERO) */
};

/* the first n (0..16) bytes of a 16-byte lane result, zeros behind them */
__device__ __forceinline__ void keep_prefix(uint32_t (&w)[4], uint32_t n)
{
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t have = n > 4u * k ? min(n - 4u * k, 4u) : 0u;
        w[k] &= have >= 4 ? 0xffffffffu : ((1u << (8 * have)) - 1u);
    }
}

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
    if (p.pad_zero) {
        /* whole 16-byte chunks everywhere (the host has checked 16-byte alignment and pitch >= width rounded up to 16): a
         * row end written as a sub-16-byte fragment makes the L2 fetch the padding's sector from DRAM to merge it, which
         * costs 10 % of the roofline on 1366-wide frames (profiles/r2_partial_sector_probe.txt) */
        if (npx < 16) { keep_prefix(ya, npx); keep_prefix(yb, npx); }
        *(uint4 *)yrow = make_uint4(ya[0], ya[1], ya[2], ya[3]);
        if (two) *(uint4 *)(yrow + p.pitch) = make_uint4(yb[0], yb[1], yb[2], yb[3]);
        if (do_uv) {
            const uint32_t pair0 = (x0 + px0) >> 1;
            if (pair0 < cw) {
                const uint32_t nuv = 2 * min(8u, cw - pair0);
                if (nuv < 16) keep_prefix(uvw, nuv);
                *(uint4 *)(sp + p.uv_off + (size_t)rp * p.pitch + x0 + px0) = make_uint4(uvw[0], uvw[1], uvw[2], uvw[3]);
            }
        }
        return;
    }
    store_prefix<4>(yrow, ya, npx);
    if (two) store_prefix<4>(yrow + p.pitch, yb, npx);
    if (do_uv) {
        const uint32_t pair0 = (x0 + px0) >> 1;                             /* first chroma pair of this lane */
        if (pair0 < cw) store_prefix<4>(sp + p.uv_off + (size_t)rp * p.pitch + x0 + px0, uvw, 2 * min(8u, cw - pair0));
    }
}

} /* namespace jmc */
