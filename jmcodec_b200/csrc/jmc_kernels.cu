/*
 * jmc_kernels.cu -- turns a jmc_job into ONE kernel launch (see jmc_kernels.cuh for the kernels).
 */
#include <stdlib.h>

#include <algorithm>
#include <atomic>

#include "jmc_internal.h"
#include "jmc_kernels.cuh"

using namespace jmc;

typedef Cfg256x4 PlaneCfg;
static_assert(INLINE_LIST_MAX == JMC_INLINE_LIST_MAX, "include/jmc_cuda.h and jmc_k_common.cuh disagree on the inline list size");

/* JMC_JOB_LIST_ON_HOST: every non-NULL list of the job is a HOST array of <= INLINE_LIST_MAX device pointers,
 * which travel to the kernel as arguments (checked in jmc_launch_job). */
static FrameSet to_set(const jmc_job *j, const jmc_frames &f)
{
    FrameSet s;
    s.base = (uint8_t *)f.base;
    s.stride = f.stride;
    s.list = (uint8_t *const *)f.list;
    s.n_inline = 0;
    s.pad_ = 0;
    for (int i = 0; i < INLINE_LIST_MAX; i++) s.inl[i] = nullptr;
    if (f.list && (j->flags & JMC_JOB_LIST_ON_HOST)) {
        s.n_inline = (uint32_t)j->n_frames;
        for (int i = 0; i < j->n_frames && i < INLINE_LIST_MAX; i++) s.inl[i] = (uint8_t *)f.list[i];
        s.list = nullptr;
    }
    return s;
}

static bool frames_ok(const jmc_frames &f) { return f.base != nullptr || f.list != nullptr; }

/* OR of every address bit of a frame set the host can see: base | stride, or the pointers of a host-side
 * list.  A device-side list cannot be read here: *known is cleared unless the caller vouches for 16-byte
 * alignment with JMC_JOB_ALIGNED16 (then it contributes no bits). */
static uint64_t frames_bits(const jmc_job *j, const jmc_frames &f, bool *known)
{
    if (!f.list) return (uint64_t)(uintptr_t)f.base | (uint64_t)f.stride;
    if (j->flags & JMC_JOB_LIST_ON_HOST) {
        uint64_t bits = 0;
        for (int i = 0; i < j->n_frames && i < INLINE_LIST_MAX; i++) bits |= (uint64_t)(uintptr_t)f.list[i];
        return bits;
    }
    if (!(j->flags & JMC_JOB_ALIGNED16)) *known = false;
    return 0;
}

/* fast_div(): m = ceil(2^sh / d), sh = 31 + ceil(log2 d) */
static FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    if (d == 0) d = 1;
    uint32_t s = 0;
    while ((1ull << s) < d) s++;
    f.d = d;
    f.sh = 31 + s;
    f.m = (uint32_t)(((1ull << (31 + s)) + d - 1) / d);
    f.pad_ = 0;
    return f;
}

static Part make_part(int kind, uint32_t rows, uint32_t row_elems, int64_t p_off, int32_t pitch, int64_t a_off, int64_t b_off)
{
    Part p;
    p.kind = (rows && row_elems) ? kind : PART_NONE;
    p.rows = rows;
    p.row_elems = row_elems;
    const uint64_t total = (uint64_t)rows * row_elems;
    p.tiles = (uint32_t)((total + TileGeom<PlaneCfg>::TILE_ELEMS - 1) / TileGeom<PlaneCfg>::TILE_ELEMS);
    p.p_off = p_off;
    p.p_pitch = pitch;
    p.pad_ = 0;
    p.a_off = a_off;
    p.b_off = b_off;
    p.rdiv = make_fastdiv(row_elems);
    return p;
}

/* Opt a kernel in to more than 48 KB of dynamic shared memory, once per (kernel, device); safe to race */
#define JMC_SMEM_ONCE(BYTES, ...)                                                                                 \
    do {                                                                                                          \
        static std::atomic<bool> smem_done[64];                                                                   \
        const bool tracked = ctx->device >= 0 && ctx->device < 64;                                                \
        if (!tracked || !smem_done[ctx->device].load(std::memory_order_relaxed)) {                                \
            JMC_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES));      \
            if (tracked) smem_done[ctx->device].store(true, std::memory_order_relaxed);                           \
        }                                                                                                         \
    } while (0)

/* Bulk-copy-engine kernel (cp.async.bulk): returns 1 when the geometry does not fit it. */
static int launch_bulk(jmc_ctx *ctx, const PlaneParams &pp, int k1, cudaStream_t stream)
{
    BulkParams b;
    b.pitched = pp.pitched;
    b.tight = pp.tight;
    b.n_frames = pp.n_frames;
    b.part[0] = pp.part[0];
    b.part[1] = pp.part[1];
    /* shared memory per row of a tile: luma/NV12-chroma rows need row_elems bytes, SPLIT/MERGE rows
     * 2*row_elems interleaved + row_elems per planar half */
    const uint32_t row0 = pp.part[0].row_elems;
    const uint32_t row1 = pp.part[1].kind == PART_NONE ? 0 : (pp.part[1].kind == PART_COPY ? pp.part[1].row_elems : 4 * pp.part[1].row_elems);
    const uint32_t per_row = row0 > row1 ? row0 : row1;
    if (per_row == 0 || per_row > 96 * 1024) return 1;
    uint32_t rows = 8;                               /* tools/sweep.cu: 8-row tiles are the optimum at 1080p and 4K */
    while (rows > 1 && (size_t)rows * per_row > 96 * 1024) rows >>= 1;
    b.rows_per_tile = rows;
    b.tiles[0] = pp.part[0].kind == PART_NONE ? 0 : (pp.part[0].rows + rows - 1) / rows;
    b.tiles[1] = pp.part[1].kind == PART_NONE ? 0 : (pp.part[1].rows + rows - 1) / rows;
    const uint64_t total = (uint64_t)(b.tiles[0] + b.tiles[1]) * b.n_frames;
    if (total == 0) return JMC_OK;
    if (total > 0x7fffffffull) return 1;
    const size_t smem = (size_t)rows * per_row;
    const uint32_t grid = (uint32_t)total;
#define JMC_BULK(TT, K1)                                                                                          \
    do {                                                                                                          \
        JMC_SMEM_ONCE(96 * 1024, bulk_planes_kernel<TT, K1>);                                                     \
        bulk_planes_kernel<TT, K1><<<grid, BULK_THREADS, smem, stream>>>(b);                                      \
    } while (0)
    if (pp.to_tight) {
        if (k1 == PART_SPLIT) JMC_BULK(true, PART_SPLIT); else JMC_BULK(true, PART_COPY);
    } else {
        if (k1 == PART_MERGE) JMC_BULK(false, PART_MERGE); else JMC_BULK(false, PART_COPY);
    }
#undef JMC_BULK
    JMC_CUDA(cudaGetLastError());
    ctx->launches++;
    return JMC_OK;
}

/* Only consulted with JMC_NO_BULK=1 (the LDG kernels): would the any-alignment kernel be reduced to <= 4-byte
 * luma or <= 2-byte chroma accesses on this job?  Then the warp-per-row kernel wins (0.41-0.59 -> 0.84-0.98 of
 * peak on 1366/854/1918/1919-wide frames); with 8-byte luma / 4-byte chroma vectors - 1080-wide portrait,
 * 720x480 - the plain kernel is faster (0.96 vs 0.78-0.87).  Pointer lists count as aligned only with
 * JMC_JOB_ALIGNED16. */
static bool narrow_vectors(const jmc_job *j, const PlaneParams &p)
{
    bool known = true;
    const uint64_t common = frames_bits(j, j->surf, &known) | frames_bits(j, j->tight, &known);
    if (!known) return false;                        /* unknown: keep the per-frame in-kernel choice */
    auto width = [](uint64_t bits) { const uint32_t low = (uint32_t)bits & 15u; return low == 0 ? 16u : (low & (0u - low)); };
    const Part &y = p.part[0], &c = p.part[1];
    const uint32_t vy = y.kind == PART_NONE ? 16u : width(common | (uint64_t)y.p_off | (uint32_t)y.p_pitch | (uint64_t)y.a_off | y.row_elems);
    uint32_t vc = 16u;
    if (c.kind == PART_COPY) vc = width(common | (uint64_t)c.p_off | (uint32_t)c.p_pitch | (uint64_t)c.a_off | c.row_elems) / 2;
    else if (c.kind != PART_NONE) vc = width(common | (uint64_t)c.p_off | (uint32_t)c.p_pitch | (uint64_t)c.a_off | (uint64_t)c.b_off | c.row_elems);
    return vy <= 4 || vc <= 2;
}

/* Warp-per-row LDG kernel for widths that are not multiples of 16 (the A/B partner of launch_bulk_rows):
 * needs only the SURFACE side aligned (base, pitch, plane offsets multiples of 16; rows over-readable up
 * to the next multiple of 16).  Returns 1 when that does not hold. */
static int launch_rows(jmc_ctx *ctx, const jmc_job *j, const PlaneParams &pp, int k1, cudaStream_t stream)
{
    {
        bool known = true;
        if ((frames_bits(j, j->surf, &known) & 15) || !known) return 1;
    }
    RowsParams r;
    r.pitched = pp.pitched;
    r.tight = pp.tight;
    r.n_frames = pp.n_frames;
    uint64_t per_frame = 0;
    for (int i = 0; i < 2; i++) {
        const Part &pt = pp.part[i];
        r.part[i] = pt;
        r.segs[i] = r.tasks[i] = 0;
        r.rpt[i] = 1;
        r.rstride[i] = ROWS_SEG;
        r.cdiv[i] = make_fastdiv(ROWS_SEG / 16);
        if (pt.kind == PART_NONE) continue;
        const uint32_t row_bytes = pt.kind == PART_COPY ? pt.row_elems : 2 * pt.row_elems;      /* surface bytes per row */
        if (((uint64_t)pt.p_off | (uint32_t)pt.p_pitch) & 15) return 1;
        if ((uint32_t)pt.p_pitch < ((row_bytes + 15) & ~15u)) return 1;
        r.segs[i] = (row_bytes + ROWS_SEG - 1) / ROWS_SEG;
        /* short rows: several whole rows per warp task, so that each lane still has up to four loads in flight */
        const uint32_t align = pt.kind == PART_COPY ? 16u : 32u;
        const uint32_t rs = (row_bytes + align - 1) & ~(align - 1);
        if (r.segs[i] == 1 && 2 * rs <= (uint32_t)ROWS_SEG && !jmc_env().rows_single) {
            r.rpt[i] = std::min<uint32_t>(ROWS_SEG / rs, ROWS_MAX_RPT);
            r.rstride[i] = rs;
            r.cdiv[i] = make_fastdiv(rs / 16);
        }
        const uint64_t t = r.rpt[i] > 1 ? ((uint64_t)pt.rows + r.rpt[i] - 1) / r.rpt[i] : (uint64_t)pt.rows * r.segs[i];
        if (t > 0x3fffffffull) return 1;
        r.tasks[i] = (uint32_t)t;
        per_frame += t;
    }
    const uint64_t total = per_frame * r.n_frames;
    if (total == 0) return JMC_OK;
    if (total > 0x7fffffffull) return 1;
    r.total_tasks = (uint32_t)total;
    const uint32_t grid = (r.total_tasks + ROWS_THREADS / 32 - 1) / (ROWS_THREADS / 32);
    const bool multi = r.rpt[0] > 1 || r.rpt[1] > 1;
#define JMC_ROWS(TT, K1)                                                                      \
    do {                                                                                      \
        if (multi) rows_kernel<TT, K1, true><<<grid, ROWS_THREADS, 0, stream>>>(r);           \
        else rows_kernel<TT, K1, false><<<grid, ROWS_THREADS, 0, stream>>>(r);                \
    } while (0)
    if (pp.to_tight) {
        if (k1 == PART_SPLIT) JMC_ROWS(true, PART_SPLIT); else JMC_ROWS(true, PART_COPY);
    } else {
        if (k1 == PART_MERGE) JMC_ROWS(false, PART_MERGE); else JMC_ROWS(false, PART_COPY);
    }
#undef JMC_ROWS
    JMC_CUDA(cudaGetLastError());
    ctx->launches++;
    return JMC_OK;
}

/* The same geometries through the bulk-copy engine (bulk_rows_kernel / bulk_rows_pack_kernel).  Same
 * preconditions as launch_rows(); returns 1 when they do not hold or the tile does not fit. */
static int launch_bulk_rows(jmc_ctx *ctx, const jmc_job *j, const PlaneParams &pp, int k1, cudaStream_t stream)
{
    {
        bool known = true;
        if ((frames_bits(j, j->surf, &known) & 15) || !known) return 1;
    }
    BulkRowsParams b;
    b.pitched = pp.pitched;
    b.tight = pp.tight;
    b.n_frames = pp.n_frames;
    b.pad_zero = (!pp.to_tight && ((j->flags & JMC_JOB_PAD_ZERO) || jmc_env().pad_zero)) ? 1u : 0u;
    uint32_t per_row = 0, widest = 16;                 /* shared memory per staged row (worst part); surface bytes per row */
    for (int i = 0; i < 2; i++) {
        const Part &pt = pp.part[i];
        b.part[i] = pt;
        b.rstride[i] = b.ldbytes[i] = 16;
        if (pt.kind == PART_NONE) continue;
        const uint32_t row_bytes = pt.kind == PART_COPY ? pt.row_elems : 2 * pt.row_elems;
        if (((uint64_t)pt.p_off | (uint32_t)pt.p_pitch) & 15) return 1;
        b.ldbytes[i] = (row_bytes + 15) & ~15u;
        if ((uint32_t)pt.p_pitch < b.ldbytes[i]) return 1;
        b.rstride[i] = pt.kind == PART_COPY ? b.ldbytes[i] : ((row_bytes + 31) & ~31u);
        widest = std::max(widest, b.rstride[i]);
        /* decode: the staged surface rows (+ the two planar halves of a chroma row); encode: the tight run(s) */
        per_row = std::max(per_row, pp.to_tight ? (pt.kind == PART_COPY ? b.rstride[i] : 2 * b.rstride[i]) : row_bytes);
    }
    if (per_row == 0) return JMC_OK;
    /* ~12 KB of surface data per tile, a multiple of the four warps that store the rows */
    uint32_t rows = std::min<uint32_t>(32, std::max<uint32_t>(4, (12288 / widest) & ~3u));
    if (jmc_env().brows_rows > 0) rows = (uint32_t)jmc_env().brows_rows;
    const size_t slack = 192;                          /* spare chunks behind each staging area, run alignment */
    while (rows > 1 && (size_t)rows * per_row + slack > 96 * 1024) rows >>= 1;
    if ((size_t)rows * per_row + slack > 96 * 1024) return 1;
    b.rows_per_tile = rows;
    b.tiles[0] = pp.part[0].kind == PART_NONE ? 0 : (pp.part[0].rows + rows - 1) / rows;
    b.tiles[1] = pp.part[1].kind == PART_NONE ? 0 : (pp.part[1].rows + rows - 1) / rows;
    const uint64_t total = (uint64_t)(b.tiles[0] + b.tiles[1]) * b.n_frames;
    if (total == 0) return JMC_OK;
    if (total > 0x7fffffffull) return 1;
    const size_t smem = (size_t)rows * per_row + slack;
    const uint32_t grid = (uint32_t)total;
#define JMC_BROWS(KERNEL)                                                                                         \
    do {                                                                                                          \
        JMC_SMEM_ONCE(96 * 1024, KERNEL);                                                                         \
        KERNEL<<<grid, BROWS_THREADS, smem, stream>>>(b);                                                         \
    } while (0)
    if (pp.to_tight) {
        if (k1 == PART_SPLIT) JMC_BROWS(bulk_rows_kernel<PART_SPLIT>); else JMC_BROWS(bulk_rows_kernel<PART_COPY>);
    } else {
        if (k1 == PART_MERGE) JMC_BROWS(bulk_rows_pack_kernel<PART_MERGE>); else JMC_BROWS(bulk_rows_pack_kernel<PART_COPY>);
    }
#undef JMC_BROWS
    JMC_CUDA(cudaGetLastError());
    ctx->launches++;
    return JMC_OK;
}

/* Can the host prove that every access of this job is 16-byte aligned?  (1080p, 4K, 720p ... are.) */
static bool all_wide(const jmc_job *j, const PlaneParams &p)
{
    bool known = true;
    uint64_t bits = frames_bits(j, j->surf, &known) | frames_bits(j, j->tight, &known);
    if (!known) return false;
    for (int i = 0; i < 2; i++) {
        const Part &pt = p.part[i];
        if (pt.kind == PART_NONE) continue;
        bits |= (uint64_t)pt.p_off | (uint64_t)(uint32_t)pt.p_pitch | (uint64_t)pt.a_off | pt.row_elems;
        if (pt.kind != PART_COPY) bits |= (uint64_t)pt.b_off;
    }
    return (bits & 15) == 0;
}

static int launch_planes(jmc_ctx *ctx, const jmc_job *j, cudaStream_t stream)
{
    if (!frames_ok(j->surf) || !frames_ok(j->tight)) { jmc_set_error("jmc_convert: surf/tight frame set is empty"); return JMC_ERR_INVALID; }
    const uint32_t w = (uint32_t)j->width, h = (uint32_t)j->height;
    PlaneParams p;
    p.pitched = to_set(j, j->surf);
    p.tight = to_set(j, j->tight);
    p.n_frames = (uint32_t)j->n_frames;
    p.to_tight = (j->op == JMC_OP_NV12_TO_NV12 || j->op == JMC_OP_NV12_TO_I420) ? 1 : 0;
    /* luma: h rows of w bytes (nv_dec.cpp:787-790,801-804; intel_enc.cpp:291-295; nv_enc.cpp:1043-1051) */
    p.part[0] = make_part(PART_COPY, h, w, j->surf_y_off, j->pitch, 0, 0);
    if (j->op == JMC_OP_NV12_TO_NV12 || j->op == JMC_OP_NV12_TO_SURF) {
        /* chroma kept interleaved: h>>1 rows of w bytes at tight offset w*h (nv_dec.cpp:792-796) */
        p.part[1] = make_part(PART_COPY, h >> 1, w, j->surf_uv_off, j->pitch, (int64_t)w * h, 0);
    } else {
        /* (h>>1) x (w>>1) chroma pairs (nv_dec.cpp:807-818; intel_enc.cpp:366-380) */
        p.part[1] = make_part(j->op == JMC_OP_NV12_TO_I420 ? PART_SPLIT : PART_MERGE, h >> 1, w >> 1,
                              j->surf_uv_off, j->pitch, j->tight_u_off, j->tight_v_off);
    }
    p.tiles_per_frame = p.part[0].tiles + p.part[1].tiles;
    const uint64_t total = (uint64_t)p.tiles_per_frame * p.n_frames;
    if (total == 0) return JMC_OK;
    if (total > 0x7fffffffull) { jmc_set_error("jmc_convert: batch too large for one launch"); return JMC_ERR_INVALID; }
    p.total_tiles = (uint32_t)total;
    const uint32_t grid = p.total_tiles;             /* one CTA per 16 KB tile (see Cfg256x4) */
    const bool wide = all_wide(j, p);
    const int k1 = p.part[1].kind == PART_NONE ? PART_COPY : p.part[1].kind;
    if (wide && !jmc_env().no_bulk) {
        int r = launch_bulk(ctx, p, k1, stream);
        if (r != 1) return r;                        /* 1: geometry does not fit the bulk kernel, use LDG/STG */
    }
    if (!wide && !jmc_env().no_rows) {
        /* width not a multiple of 16 on an aligned surface: bulk-loaded row tiles, both directions */
        if (!jmc_env().no_bulk) {
            int r = launch_bulk_rows(ctx, j, p, k1, stream);
            if (r != 1) return r;                    /* 1: surface side not 16-byte friendly or tile too large */
        }
        if (narrow_vectors(j, p) || jmc_env().rows_always) {
            int r = launch_rows(ctx, j, p, k1, stream);
            if (r != 1) return r;                    /* 1: surface side not 16-byte friendly, use the any-alignment kernel */
        }
    }
#define JMC_LAUNCH(TT, K1)                                                                               \
    do {                                                                                                 \
        if (wide) planes_kernel<PlaneCfg, TT, K1, true><<<grid, PlaneCfg::THREADS, 0, stream>>>(p);      \
        else planes_kernel<PlaneCfg, TT, K1, false><<<grid, PlaneCfg::THREADS, 0, stream>>>(p);          \
    } while (0)
    if (p.to_tight) {
        if (k1 == PART_SPLIT) JMC_LAUNCH(true, PART_SPLIT); else JMC_LAUNCH(true, PART_COPY);
    } else {
        if (k1 == PART_MERGE) JMC_LAUNCH(false, PART_MERGE); else JMC_LAUNCH(false, PART_COPY);
    }
#undef JMC_LAUNCH
    JMC_CUDA(cudaGetLastError());
    ctx->launches++;
    return JMC_OK;
}

static int launch_rgb(jmc_ctx *ctx, const jmc_job *j, cudaStream_t stream)
{
    const bool fused = j->op == JMC_OP_NV12_TO_I420_RGB24;
    const bool argb = j->op == JMC_OP_NV12_TO_ARGB32;
    if (!frames_ok(j->surf) || !frames_ok(j->rgb) || (fused && !frames_ok(j->tight))) {
        jmc_set_error("jmc_convert: surf/rgb/tight frame set is empty");
        return JMC_ERR_INVALID;
    }
    if ((j->width >> 1) < 1 || (j->height >> 1) < 1) { jmc_set_error("jmc_convert: RGB needs width,height >= 2"); return JMC_ERR_INVALID; }
    if (j->rgb_pitch < (argb ? 4 : 3) * j->width) { jmc_set_error("jmc_convert: rgb_pitch < %d*width", argb ? 4 : 3); return JMC_ERR_INVALID; }
    RgbParams p;
    p.surf = to_set(j, j->surf);
    p.tight = to_set(j, j->tight);
    p.rgb = to_set(j, j->rgb);
    p.n_frames = (uint32_t)j->n_frames;
    p.width = j->width; p.height = j->height; p.pitch = j->pitch;
    p.y_off = j->surf_y_off; p.uv_off = j->surf_uv_off;
    p.u_off = j->tight_u_off; p.v_off = j->tight_v_off;
    p.rgb_pitch = j->rgb_pitch;
    p.fused = fused ? 1 : 0;
    p.argb = argb ? 1 : 0;
    p.segs_per_row = ((uint32_t)j->width + 511) / 512;
    p.row_pairs = ((uint32_t)j->height + 1) / 2;
    p.tasks_per_frame = p.segs_per_row * p.row_pairs;
    p.tpf_div = make_fastdiv(p.tasks_per_frame);
    p.seg_div = make_fastdiv(p.segs_per_row);
    const uint64_t total = (uint64_t)p.tasks_per_frame * p.n_frames;
    if (total == 0) return JMC_OK;
    if (total > 0x7fffffffull) { jmc_set_error("jmc_convert: batch too large for one launch"); return JMC_ERR_INVALID; }
    p.total_tasks = (uint32_t)total;
    /* bulk-copy-engine variants: 10 bytes of shared memory per pixel of a segment */
    {
        uint64_t bits = (uint64_t)j->surf_y_off | (uint64_t)j->surf_uv_off | (uint32_t)j->pitch | (uint32_t)j->rgb_pitch | (uint32_t)j->width;
        if (fused) bits |= (uint32_t)(j->width >> 1);            /* U / V rows are bulk-stored too */
        const jmc_frames *sets[3] = { &j->surf, &j->rgb, fused ? &j->tight : nullptr };
        const bool ok = !jmc_env().no_bulk && !argb;    /* ARGB32 runs on the vector kernel */
        bool known = true;                                       /* device-side pointer lists: aligned only if the caller says so */
        for (int i = 0; i < 3; i++)
            if (sets[i]) bits |= frames_bits(j, *sets[i], &known);
        if (fused) bits |= (uint64_t)j->tight_u_off | (uint64_t)j->tight_v_off;
        /* !aligned: only the surface has to be 16-byte friendly (any even width; rows over-readable to the next
         * multiple of 16 inside the pitch) - RGB / tight rows at any address are written with re-aligned stores */
        const bool aligned = known && (bits & 15) == 0;
        bool surf_ok = (j->width & 1) == 0 && j->pitch >= ((j->width + 15) & ~15) &&
                       (((uint64_t)j->surf_y_off | (uint64_t)j->surf_uv_off | (uint32_t)j->pitch) & 15) == 0;
        {
            bool sknown = true;
            const uint64_t sbits = frames_bits(j, j->surf, &sknown);
            surf_ok = surf_ok && sknown && (sbits & 15) == 0;
        }
        /* measured (profiles/r1_odd_sizes_rgb_bulk.txt): with unaligned rows the bulk-loaded variant wins for the
         * fused op (1366-wide: 0.88 vs 0.79 of peak) but not for RGB alone (0.77 vs 0.82; 1080-wide 0.68 vs 0.82) */
        /* ... and for RGB alone on narrow frames the warp-per-task kernel is ahead even when everything is aligned
         * (profiles/r1c_rgb_bulk_vs_vector.txt: 1536 wide 1.03 vs 0.99, 1376 wide 0.95 vs 0.92; from 1920 wide on the
         * bulk kernel wins, 1.02 vs 1.00) */
        /* row pairs per CTA of the re-aligning variant (!aligned).  Measured (profiles/r2_rgb_tile_shape_sweep.txt): two
         * pairs - one whole-row store job per warp - lift the fused op on 1366 / 1080 / 854-wide frames from 0.75 / 0.68 /
         * 0.59 to 0.88 / 0.85 / 0.75 of peak; three and more lose it again to occupancy.  RGB alone stays on the
         * warp-per-task kernel (0.85 / 0.91 / 0.78 vs 0.83 / 0.82 / 0.74) unless JMC_RGB_BULK_PAIRS=n asks (A/B). */
        const int want_pairs = jmc_env().rgb_bulk_pairs;
        const bool bulk_pays = fused || j->width >= 1664 || jmc_env().rgb_bulk_always;
        if (ok && ((aligned && bulk_pays) || (surf_ok && (fused || jmc_env().rgb_bulk_always || (!aligned && want_pairs > 0))))) {
            RgbBulkParams b;
            b.surf = p.surf; b.tight = p.tight; b.rgb = p.rgb;
            b.n_frames = p.n_frames; b.width = p.width; b.height = p.height; b.pitch = p.pitch;
            b.y_off = p.y_off; b.uv_off = p.uv_off; b.u_off = p.u_off; b.v_off = p.v_off;
            b.rgb_pitch = p.rgb_pitch; b.fused = p.fused; b.row_pairs = p.row_pairs;
            /* RGB only: column segments of <= 2048 pixels, a multiple of 32 (3840 -> 2 x 1920): more CTAs per
             * SM, +1.4 %.  Fused: whole rows while they fit (the luma store is then one 2*w-byte burst), +1.9 %
             * over segments (profiles/README.md). */
            const uint32_t seg_max = (fused && (size_t)j->width * 10 <= 96 * 1024) ? (uint32_t)j->width : 2048u;
            b.segs = ((uint32_t)j->width + seg_max - 1) / seg_max;
            b.seg_w = ((((uint32_t)j->width + b.segs - 1) / b.segs) + 31) & ~31u;
            b.segs = ((uint32_t)j->width + b.seg_w - 1) / b.seg_w;
            /* whole unaligned rows: two row pairs per CTA (rgb_bulk_pairs_kernel), while they fit ~56 KB (4 CTAs per SM) */
            b.pairs = 1;
            if (!aligned && b.segs == 1 && want_pairs != 1) {
                b.pairs = want_pairs > 1 ? (uint32_t)want_pairs : 2u;
                while (b.pairs > 1 && (size_t)b.pairs * b.seg_w * 10 + 64 > 56 * 1024) b.pairs--;
                if (b.pairs > b.row_pairs) b.pairs = b.row_pairs;
            }
            b.groups = (b.row_pairs + b.pairs - 1) / b.pairs;
            b.upr_div = make_fastdiv(((std::min<uint32_t>(b.seg_w, (uint32_t)j->width) + 15) & ~15u) >> 4);
            const uint64_t ctas = (uint64_t)b.groups * b.segs * b.n_frames;
            const size_t smem = (size_t)b.pairs * b.seg_w * 10 + 64; /* + spare chunks read by the re-aligning stores */
            if (ctas <= 0x7fffffffull && smem <= 100 * 1024) {
#define JMC_RGB_BULK(AL)                                                                                          \
    do {                                                                                                          \
        JMC_SMEM_ONCE(100 * 1024, rgb_bulk_kernel<AL>);                                                           \
        rgb_bulk_kernel<AL><<<(uint32_t)ctas, RGB_BULK_THREADS, smem, stream>>>(b);                               \
    } while (0)
                if (aligned) JMC_RGB_BULK(true);
                else if (b.pairs == 1) JMC_RGB_BULK(false);
                else {
                    JMC_SMEM_ONCE(100 * 1024, rgb_bulk_pairs_kernel);
                    rgb_bulk_pairs_kernel<<<(uint32_t)ctas, RGB_BULK_THREADS, smem, stream>>>(b);
                }
#undef JMC_RGB_BULK
                JMC_CUDA(cudaGetLastError());
                ctx->launches++;
                return JMC_OK;
            }
        }
    }
    constexpr uint32_t WARPS = RgbCfg::THREADS / 32;
    /* flattened variant: when a row pair leaves more than ~1/7 of the lane slots of its warps idle and the surface
     * allows 16-byte loads for every lane.  Measured (profiles/r1c_rgb_flat.txt): 1080 wide (71 % of the slots used)
     * RGB24 0.84 -> 0.92, ARGB32 0.93 -> 1.02; 1280 wide (83 %) 0.96 -> 0.98; but 1366 / 1376 wide (90 %) lose
     * 3-5 % to the extra runs per warp, and 4K (94 %) is a tie.  JMC_RGB_FLAT=0 / 1 forces it off / on (A/B). */
    {
        bool surf_ok = !fused && (j->width & 1) == 0 && j->pitch >= ((j->width + 15) & ~15) &&
                       (((uint64_t)j->surf_y_off | (uint64_t)j->surf_uv_off | (uint32_t)j->pitch) & 15) == 0;
        {
            bool sknown = true;
            const uint64_t sbits = frames_bits(j, j->surf, &sknown);
            surf_ok = surf_ok && sknown && (sbits & 15) == 0;
        }
        const int force = jmc_env().rgb_flat;
        const uint32_t units = ((uint32_t)j->width + 15) / 16, slots = 32 * ((units + 31) / 32);
        const bool want = force >= 0 ? force != 0 : units * 100 < slots * 86;
        if (surf_ok && want) {
            RgbFlatParams q;
            q.surf = p.surf; q.rgb = p.rgb; q.n_frames = p.n_frames;
            q.width = p.width; q.height = p.height; q.pitch = p.pitch;
            q.y_off = p.y_off; q.uv_off = p.uv_off; q.rgb_pitch = p.rgb_pitch;
            q.units_per_row = ((uint32_t)j->width + 15) / 16;
            q.units_per_frame = q.units_per_row * p.row_pairs;
            q.tasks_per_frame = (q.units_per_frame + 31) / 32;
            q.tpf_div = make_fastdiv(q.tasks_per_frame);
            q.unit_div = make_fastdiv(q.units_per_row);
            const uint64_t tasks = (uint64_t)q.tasks_per_frame * q.n_frames;
            if (tasks <= 0x7fffffffull) {
                q.total_tasks = (uint32_t)tasks;
                const uint32_t g2 = (q.total_tasks + WARPS - 1) / WARPS;
                if (argb) rgb_flat_kernel<RgbCfg, true><<<g2, RgbCfg::THREADS, 0, stream>>>(q);
                else rgb_flat_kernel<RgbCfg, false><<<g2, RgbCfg::THREADS, 0, stream>>>(q);
                JMC_CUDA(cudaGetLastError());
                ctx->launches++;
                return JMC_OK;
            }
        }
    }
    const uint32_t blocks_needed = (p.total_tasks + WARPS - 1) / WARPS;
    const uint32_t grid = blocks_needed;             /* one warp per (row pair, 512-pixel segment) task */
    if (argb) rgb_kernel<RgbCfg, true><<<grid, RgbCfg::THREADS, 0, stream>>>(p);
    else rgb_kernel<RgbCfg, false><<<grid, RgbCfg::THREADS, 0, stream>>>(p);
    JMC_CUDA(cudaGetLastError());
    ctx->launches++;
    return JMC_OK;
}

static int launch_rgb_to_nv12(jmc_ctx *ctx, const jmc_job *j, cudaStream_t stream)
{
    if (!frames_ok(j->surf) || !frames_ok(j->rgb)) { jmc_set_error("jmc_convert: surf/rgb frame set is empty"); return JMC_ERR_INVALID; }
    if (j->rgb_pitch < 3 * j->width) { jmc_set_error("jmc_convert: rgb_pitch %d < 3*width", j->rgb_pitch); return JMC_ERR_INVALID; }
    Rgb2Params p;
    p.rgb = to_set(j, j->rgb);
    p.surf = to_set(j, j->surf);
    p.n_frames = (uint32_t)j->n_frames;
    p.width = j->width; p.height = j->height; p.pitch = j->pitch; p.rgb_pitch = j->rgb_pitch;
    p.y_off = j->surf_y_off; p.uv_off = j->surf_uv_off;
    p.row_pairs = ((uint32_t)j->height + 1) / 2;
    p.segs_per_row = ((uint32_t)j->width + 511) / 512;
    p.tasks_per_frame = p.row_pairs * p.segs_per_row;
    p.tpf_div = make_fastdiv(p.tasks_per_frame);
    p.seg_div = make_fastdiv(p.segs_per_row);
    const uint64_t total = (uint64_t)p.tasks_per_frame * p.n_frames;
    if (total == 0) return JMC_OK;
    if (total > 0x7fffffffull) { jmc_set_error("jmc_convert: batch too large for one launch"); return JMC_ERR_INVALID; }
    p.total_tasks = (uint32_t)total;
    {
        /* JMC_JOB_PAD_ZERO is honoured when every surface row is 16-byte aligned and long enough for whole chunks */
        bool known = true;
        const uint64_t bits = frames_bits(j, j->surf, &known) | (uint64_t)j->surf_y_off | (uint64_t)j->surf_uv_off | (uint32_t)j->pitch;
        p.pad_zero = (((j->flags & JMC_JOB_PAD_ZERO) || jmc_env().pad_zero) && known && (bits & 15) == 0 && j->pitch >= ((j->width + 15) & ~15)) ? 1u : 0u;
    }
    constexpr uint32_t WARPS = RGB2_THREADS / 32;
    rgb_to_nv12_kernel<<<(p.total_tasks + WARPS - 1) / WARPS, RGB2_THREADS, 0, stream>>>(p);
    JMC_CUDA(cudaGetLastError());
    ctx->launches++;
    return JMC_OK;
}

int jmc_launch_job(jmc_ctx *ctx, const jmc_job *j, cudaStream_t stream)
{
    if (!ctx || !j) { jmc_set_error("jmc_convert: NULL ctx/job"); return JMC_ERR_INVALID; }
    if (j->n_frames < 0 || j->width < 0 || j->height < 0 || j->pitch < 0) { jmc_set_error("jmc_convert: negative geometry"); return JMC_ERR_INVALID; }
    if ((uint64_t)j->width * (uint64_t)j->height > 0x7fffffffull) { jmc_set_error("jmc_convert: frame too large"); return JMC_ERR_INVALID; }
    if (j->n_frames == 0 || j->width == 0 || j->height == 0) return JMC_OK;
    if (j->pitch < j->width) { jmc_set_error("jmc_convert: pitch %d < width %d", j->pitch, j->width); return JMC_ERR_INVALID; }
    if ((j->flags & JMC_JOB_LIST_ON_HOST) && j->n_frames > INLINE_LIST_MAX) {
        jmc_set_error("jmc_convert: JMC_JOB_LIST_ON_HOST carries at most %d frames per launch", INLINE_LIST_MAX);
        return JMC_ERR_INVALID;
    }
    switch (j->op) {
    case JMC_OP_NV12_TO_NV12:
    case JMC_OP_NV12_TO_I420:
    case JMC_OP_NV12_TO_SURF:
    case JMC_OP_I420_TO_SURF:
        return launch_planes(ctx, j, stream);
    case JMC_OP_NV12_TO_RGB24:
    case JMC_OP_NV12_TO_I420_RGB24:
    case JMC_OP_NV12_TO_ARGB32:
        return launch_rgb(ctx, j, stream);
    case JMC_OP_RGB24_TO_SURF:
        return launch_rgb_to_nv12(ctx, j, stream);
    default:
        jmc_set_error("jmc_convert: unknown op %d", j->op);
        return JMC_ERR_INVALID;
    }
}
