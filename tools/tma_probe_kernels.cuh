/*
 * tools/tma_probe_kernels.cuh -- EXPERIMENT, not part of the product library (tools/tma_probe.cu).
 *
 * Idea tested in round 2: let the TMA unit do the re-alignment of the ENCODE direction for widths that are not
 * multiples of 16 (tight rows at arbitrary byte addresses) -- tile-mode loads with 1-byte elements whose box starts at
 * any byte of the tight frame, and tile-mode stores into the surface clipped at `width` by the hardware -- instead of
 * the funnel shifts of bulk_rows_pack_kernel (0.87-0.94 of the roofline on 854 / 1366-wide frames).
 *
 * Result (profiles/r2_tma_probe.txt): the hardware REJECTS box starts that are not 16-byte aligned in global memory
 * (illegal instruction at x = 854 with a 256 x 1 byte box; x = 0 and x = 256 work, clipped stores work).  Tensor-map
 * TMA therefore cannot replace the shared-memory re-alignment; the kernels below are kept only as the evidence.
 */
#pragma once
#include <cuda.h>

#include "jmc_k_common.cuh"

namespace jmc {

__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *tm, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem)),
                 "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, int x, int y, int z, const void *smem)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(x), "r"(y), "r"(z),
                 "r"(smem_u32(smem)) : "memory");
}

constexpr int TMAP_THREADS = 128;
constexpr uint32_t TMAP_BOX = 256;            /* bytes per box row: the TMA limit of 256 elements per dimension */

struct TmapPackParams {
    uint32_t n_frames;
    uint32_t rows_per_tile;                   /* R: rows of the store box */
    uint32_t tiles[2];                        /* tiles per frame of luma / chroma */
    uint32_t rows[2];                         /* rows of luma / chroma */
    uint32_t re[2];                           /* bytes per TIGHT row: luma w; chroma: w (NV12) or pairs per row (I420: U and V rows) */
    uint32_t nb[2];                           /* 256-byte boxes per tight row */
    uint32_t nbs;                             /* boxes per surface chroma row */
    uint32_t y_off, a_off, u_off, v_off;      /* tensor offsets of luma, the chroma plane (NV12) / the U and V planes (I420);
                                                 all include the skew of the tight base below its 16-byte boundary */
};

/* Tensor maps: tight = {tight_stride bytes, n_frames} (1-byte elements, box 256 x 1);
 * sy / suv = surface luma / chroma plane {row bytes, rows, n_frames}, strides {pitch, surf_stride}, box 256 x R x 1.
 * Shared memory is box-major: box b of row i at (b*R + i)*256, which is exactly the layout a 256 x R store box reads. */
template <int KIND1>
__global__ void __launch_bounds__(TMAP_THREADS) tmap_pack_kernel(const __grid_constant__ CUtensorMap tight, const __grid_constant__ CUtensorMap sy,
                                                                const __grid_constant__ CUtensorMap suv, const TmapPackParams p)
{
    extern __shared__ __align__(128) uint8_t tm_smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t tpf = p.tiles[0] + p.tiles[1];
    const uint32_t f = blockIdx.x / tpf;
    uint32_t r = blockIdx.x - f * tpf;
    const bool second = r >= p.tiles[0];
    if (second) r -= p.tiles[0];
    const uint32_t R = p.rows_per_tile;
    const uint32_t r0 = r * R;
    const uint32_t part = second ? 1u : 0u;
    const uint32_t nr = min(R, p.rows[part] - r0);
    const uint32_t re = p.re[part], nb = p.nb[part];
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();

    if (!second || KIND1 == PART_COPY) {
        if (threadIdx.x != 0) return;                                     /* the TMA unit does all the work */
        const uint32_t base = (second ? p.a_off : p.y_off) + r0 * re;
        mbar_expect_tx(&bar, nr * nb * TMAP_BOX);
        for (uint32_t i = 0; i < nr; i++)
            for (uint32_t b = 0; b < nb; b++)
                tma_load_2d(tm_smem + (size_t)(b * R + i) * TMAP_BOX, &tight, (int)(base + i * re + b * TMAP_BOX), (int)f, &bar);
        mbar_wait(&bar, 0);
        for (uint32_t b = 0; b < nb; b++)
            tma_store_3d(second ? &suv : &sy, (int)(b * TMAP_BOX), (int)r0, (int)f, tm_smem + (size_t)b * R * TMAP_BOX);
        bulk_commit_wait_read();
        return;
    }
    /* MERGE: U rows and V rows (re bytes each) -> interleaved rows of 2*re bytes */
    uint8_t *Su = tm_smem, *Sv = Su + (size_t)R * nb * TMAP_BOX, *Suv = Sv + (size_t)R * nb * TMAP_BOX;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 2 * nr * nb * TMAP_BOX);
        for (uint32_t i = 0; i < nr; i++)
            for (uint32_t b = 0; b < nb; b++) {
                const uint32_t x = (r0 + i) * re + b * TMAP_BOX;
                tma_load_2d(Su + (size_t)(b * R + i) * TMAP_BOX, &tight, (int)(p.u_off + x), (int)f, &bar);
                tma_load_2d(Sv + (size_t)(b * R + i) * TMAP_BOX, &tight, (int)(p.v_off + x), (int)f, &bar);
            }
    }
    mbar_wait(&bar, 0);
    const uint32_t cpr = (re + 15) / 16;                                  /* 16-sample chunks per row */
    for (uint32_t t = threadIdx.x; t < nr * cpr; t += TMAP_THREADS) {
        const uint32_t i = t / cpr, c = t - i * cpr;
        const uint32_t x0 = 16 * c;
        const size_t src = (size_t)((x0 >> 8) * R + i) * TMAP_BOX + (x0 & 255);
        const uint4 u = *(const uint4 *)(Su + src), w = *(const uint4 *)(Sv + src);
        uint4 a, b;
        a.x = __byte_perm(u.x, w.x, 0x5140); a.y = __byte_perm(u.x, w.x, 0x7362);
        a.z = __byte_perm(u.y, w.y, 0x5140); a.w = __byte_perm(u.y, w.y, 0x7362);
        b.x = __byte_perm(u.z, w.z, 0x5140); b.y = __byte_perm(u.z, w.z, 0x7362);
        b.z = __byte_perm(u.w, w.w, 0x5140); b.w = __byte_perm(u.w, w.w, 0x7362);
        const uint32_t y0 = 2 * x0;                                       /* multiple of 32: both halves in one box */
        uint8_t *dst = Suv + (size_t)((y0 >> 8) * R + i) * TMAP_BOX + (y0 & 255);
        *(uint4 *)dst = a;
        *(uint4 *)(dst + 16) = b;
    }
    fence_async_smem();                                                   /* generic-proxy writes -> visible to the TMA unit */
    __syncthreads();
    if (threadIdx.x == 0) {
        for (uint32_t b = 0; b < p.nbs; b++) tma_store_3d(&suv, (int)(b * TMAP_BOX), (int)r0, (int)f, Suv + (size_t)b * R * TMAP_BOX);
        bulk_commit_wait_read();
    }
}

} /* namespace jmc */
