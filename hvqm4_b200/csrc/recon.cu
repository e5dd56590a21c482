/*
 * recon.cu -- pixel reconstruction kernels for sm_100a.
 *
 * One launch reconstructs a BATCH of independent pictures (one per stream; I, P and B
 * mixed freely).  The reference decodes macroblocks serially in raster order
 * (/root/reference/h4m_audio_decode.c:1487-1518 for I pictures, 1922-1967 for P/B), but
 * no block ever depends on another block's *pixels* of the same picture -- only on the
 * completed type/DC maps and on other frames -- so every 4x4 block is an independent
 * work item here.
 *
 * Mapping
 *   picture -> ctas_per_pic CTAs of kWarps warps; a CTA stages the picture's nest once
 *              (as the 38x64 "8 nibbles from x" table of recon_core.h) and its warps loop
 *              over kUnitsPerWarp SEGMENTS each;
 *   warp    -> one segment at a time: 16 macroblocks of one macroblock row = 96 blocks,
 *              walked in three passes (upper luma block row: 32 blocks, lower luma block
 *              row: 32 blocks, chroma: 16 U + 16 V), one block per lane and pass.
 * Per segment:
 *   1. every lane loads its type byte; a warp prefix sum over sym_side_words(type) on top
 *      of the segment-table entry gives the block's slot in the side-word array;
 *   2. cheap blocks (weighted DC, flat, raw, motion compensation only) are computed at
 *      once; blocks with an AOT basis loop are only QUEUED (ballot/popc compaction into
 *      two per-warp queues: intra AOT, predicted AOT);
 *   3. the queues are drained with all lanes busy on the same kind of work -- without
 *      this the basis loops run at the occupancy of the rarest block type in the warp;
 *   4. every result goes to a per-warp shared-memory tile (8 x 128 B luma, 2 x 4 x 64 B
 *      chroma); the tile is written out with 16-byte vector stores, 128 contiguous bytes
 *      per 8 lanes, whatever order the blocks were computed in.
 * Reference pixels (motion compensation, nest windows) are fetched straight from global
 * memory through the read-only path: with per-macroblock vectors the footprint of a
 * segment is scattered, and staging a superset in shared memory would read more.
 *
 * The block arithmetic itself lives in recon_core.h.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "recon.h"
#include "recon_core.h"

namespace {

constexpr int kTileLumaBytes = 8 * 128;
constexpr int kTileBytes = kTileLumaBytes + 2 * 4 * 64;   /* 1536 per segment */
constexpr int kQueues = 4;                                /* weighted, MC, intra AOT, predicted AOT */

/* kSegs = segments a warp classifies and drains together (1 or 2) */
template <int kSegs>
struct __align__(16) WarpScratchT
{
    uint32_t tile[kSegs][kTileBytes / 4];
    uint16_t queue[kQueues][kSegs * 96];                  /* entry: [7:0] slot, [15:8] type byte */
    uint16_t side_off[kSegs * 96];                        /* side-word offset of the slot relative to its pass base */
    uint32_t pass_base[kSegs * 3 + 2];
};

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t &total)
{
    const unsigned lane = threadIdx.x & 31;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (unsigned)d) inc += n;
    }
    total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    return inc - v;
}

/* slot = (segment-in-iteration * 3 + pass) * 32 + lane  ->  block coordinates and tile position */
struct Slot
{
    int plane, bx, by, tile_word, seg;
};

/* (row, first macroblock) of the segments of the current iteration, warp-uniform */
struct IterGeom
{
    int row[2], mx0[2];
};

__device__ __forceinline__ Slot decode_slot(int slot, const IterGeom &g)
{
    Slot s;
    const int grp = slot >> 5, lane = slot & 31;
    s.seg = grp >= 3;
    const int pass = grp - 3 * s.seg;
    const int row = s.seg ? g.row[1] : g.row[0], mx0 = s.seg ? g.mx0[1] : g.mx0[0];
    if (pass < 2)
    {
        s.plane = 0;
        s.bx = mx0 * 2 + lane;
        s.by = row * 2 + pass;
        s.tile_word = pass * 4 * 32 + lane;                       /* row stride 32 words */
    }
    else
    {
        s.plane = 1 + (lane >> 4);
        s.bx = mx0 + (lane & 15);
        s.by = row;
        s.tile_word = kTileLumaBytes / 4 + (lane >> 4) * 64 + (lane & 15);   /* row stride 16 words */
    }
    return s;
}

template <typename WS>
__device__ __forceinline__ void tile_put(WS &ws, const Slot &s, const uint32_t rows[4])
{
    const int stride = s.plane ? 16 : 32;
    uint32_t *t = ws.tile[s.seg] + s.tile_word;
#pragma unroll
    for (int r = 0; r < 4; ++r) t[r * stride] = rows[r];
}

template <int kWarps, int kItersPerWarp, int kMinBlocks, int kSegsPerIter>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
recon_pictures_kernel(const ReconJob *__restrict__ jobs, int units_per_pic, int ctas_per_pic)
{
    using WarpScratch = WarpScratchT<kSegsPerIter>;
    /* dynamic shared memory: [nest table | mcdiv | div] (fixed offsets, recon_core.h) then one scratch per warp */
    WarpScratch *s_warp = reinterpret_cast<WarpScratch *>(rc_smem + RC_SMEM_TABLE_BYTES);
    uint32_t *s_nest_tab = reinterpret_cast<uint32_t *>(rc_smem + RC_SMEM_NEST_OFF);
    int32_t *s_mcdiv = reinterpret_cast<int32_t *>(rc_smem + RC_SMEM_MCDIV_OFF);
    int32_t *s_div = reinterpret_cast<int32_t *>(rc_smem + RC_SMEM_DIV_OFF);

    const int job = blockIdx.x / ctas_per_pic;
    const int cta = blockIdx.x - job * ctas_per_pic;

    /* Picture parameters live in shared memory (one ReconView per CTA) instead of ~50 registers per
       thread: every field is a constant-offset LDS away and nothing stays live across the drains. */
    static_assert(sizeof(ReconView) <= 256, "ReconView must fit its shared-memory slot");
    ReconView &vw = *reinterpret_cast<ReconView *>(rc_smem + RC_SMEM_VIEW_OFF);
    if (threadIdx.x == 0)
    {
        const ReconJob J = jobs[job];
        SymHeader h;
        const uint4 *src = reinterpret_cast<const uint4 *>(J.blob);
        uint4 *dst = reinterpret_cast<uint4 *>(&h);
#pragma unroll
        for (int i = 0; i < 5; ++i) dst[i] = __ldg(src + i);     /* header fields end at byte 76 */
        rc_make_view(vw, J.blob, h, nullptr, nullptr, nullptr, J.past, J.future);
        vw.present = J.present;
    }
    /* constants of h4m:262-273 */
    for (int i = threadIdx.x; i < 256; i += kWarps * 32) s_mcdiv[i] = i ? 0x1000 / i : 0;
    if (threadIdx.x < 16) s_div[threadIdx.x] = threadIdx.x ? 0x1000 / (threadIdx.x * 16) * 16 : 0;
    __syncthreads();
    const ReconView &v = vw;
    if (v.has_nest)
    {
        /* stage the packed rows (35 B) at a 40-byte pitch, zero padded, in scratch that is free until the barrier */
        uint8_t *packed = reinterpret_cast<uint8_t *>(s_warp);
        const uint8_t *src = v.blob + v.off_nest;
        for (int i = threadIdx.x; i < SYM_NEST_H * 40; i += kWarps * 32)
        {
            const int y = i / 40, x = i - y * 40;
            packed[i] = x < SYM_NEST_ROW_BYTES ? __ldg(src + y * SYM_NEST_ROW_BYTES + x) : (uint8_t)0;
        }
        __syncthreads();
        /* entries (y, 2j) and (y, 2j+1) share the five bytes j..j+4 of row y */
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(packed);
        for (int i = threadIdx.x; i < SYM_NEST_H * 32; i += kWarps * 32)
        {
            const int y = i >> 5, j = i & 31;
            const int w = y * 10 + (j >> 2), sh = (j & 3) * 8;
            const uint32_t w0 = pw[w], w1 = pw[w + 1], w2 = (j & 3) ? pw[w + 2] : 0u;
            const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
            s_nest_tab[y * 64 + 2 * j] = lo;
            s_nest_tab[y * 64 + 2 * j + 1] = (lo >> 4) | (hi << 28);
        }
        __syncthreads();
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch &ws = s_warp[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
#define W (v.width)
#define H (v.height)
#define nseg (v.nseg)
#define mcb_w (v.mcb_w)
#define side (v.side)
#define segtab (v.segtab)

#pragma unroll 1
    for (int it = 0; it < kItersPerWarp; ++it)
    {
        const int unit0 = ((cta * kItersPerWarp + it) * kWarps + warp) * kSegsPerIter;
        if (unit0 >= units_per_pic) break;
        const int n_units = min(kSegsPerIter, units_per_pic - unit0);
        IterGeom geom;
#pragma unroll
        for (int u = 0; u < 2; ++u)
        {
            const int unit = min(unit0 + (u < kSegsPerIter ? u : 0), units_per_pic - 1);
            geom.row[u] = unit / nseg;
            geom.mx0[u] = (unit - geom.row[u] * nseg) * SYM_SEG_MCBS;
        }
        int qn[kQueues] = {0, 0, 0, 0};

        /* ---- phase 1: place every block's side words, do the trivial blocks, queue the rest ---- */
#pragma unroll 1
        for (int grp = 0; grp < n_units * 3; ++grp)
        {
            const int slot = grp * 32 + lane;
            const Slot s = decode_slot(slot, geom);
            const bool valid = s.plane ? s.bx < mcb_w : s.bx < mcb_w * 2;
            const int bstride = ((s.plane ? W >> 1 : W) >> 2) + 2;
            uint32_t t = 0;
            if (valid) t = __ldg(v.blob + rc_pick3(v.off_type, s.plane) + (s.by + 1) * bstride + s.bx + 1);
            const uint32_t nwords = valid ? sym_side_words(t, v.is_ipic) : 0u;
            uint32_t total;
            const uint32_t rel = warp_excl_scan(nwords, total);
            /* segment-table entry at the first pass of a segment, running sum afterwards (warp-uniform) */
            uint32_t base;
            if (grp == 0 || grp == 3) base = __ldg(segtab + unit0 + (grp == 3));
            else base = ws.pass_base[grp - 1] + ws.pass_base[kSegsPerIter * 3 + ((grp - 1) & 1)];
            if (lane == 0)
            {
                ws.pass_base[grp] = base;
                ws.pass_base[kSegsPerIter * 3 + (grp & 1)] = total;
            }
            ws.side_off[slot] = (uint16_t)rel;
            const int cls = valid ? rc_classify(t, v.is_ipic) : -1;
            const uint16_t entry = (uint16_t)(slot | t << 8);
#pragma unroll
            for (int q = 0; q < kQueues; ++q)
            {
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, cls == q + 1);
                if (cls == q + 1) ws.queue[q][qn[q] + __popc(m & lt_mask)] = entry;
                qn[q] += __popc(m);
            }
            if (cls == RC_DIRECT)
            {
                uint32_t rows[4];
                rc_direct_block(v, s.plane, s.bx, s.by, t, side + base + rel, rows);
                tile_put(ws, s, rows);
            }
            __syncwarp();
        }

        /* ---- phase 2: drain the queues, all lanes on one kind of work at a time ---- */
#pragma unroll 1
        for (int i = lane; i < qn[RC_WEIGHTED - 1]; i += 32)
        {
            const uint32_t e = ws.queue[RC_WEIGHTED - 1][i];
            const Slot s = decode_slot(e & 255, geom);
            uint32_t rows[4];
            rc_weighted_block(v, s.plane, s.bx, s.by, rows);
            tile_put(ws, s, rows);
        }
#pragma unroll 1
        for (int i = lane; i < qn[RC_MC - 1]; i += 32)
        {
            const uint32_t e = ws.queue[RC_MC - 1][i];
            const Slot s = decode_slot(e & 255, geom);
            uint32_t rows[4];
            rc_mc_block(v, s.plane, s.bx, s.by, e >> 8, rows);
            tile_put(ws, s, rows);
        }
#pragma unroll 1
        for (int i = lane; i < qn[RC_AOT_INTRA - 1]; i += 32)
        {
            const uint32_t e = ws.queue[RC_AOT_INTRA - 1][i];
            const int slot = e & 255;
            const Slot s = decode_slot(slot, geom);
            uint32_t rows[4];
            rc_aot_intra_block(v, s.plane, s.bx, s.by, e >> 8, side + ws.pass_base[slot >> 5] + ws.side_off[slot], rows);
            tile_put(ws, s, rows);
        }
#pragma unroll 1
        for (int i = lane; i < qn[RC_AOT_INTER - 1]; i += 32)
        {
            const uint32_t e = ws.queue[RC_AOT_INTER - 1][i];
            const int slot = e & 255;
            const Slot s = decode_slot(slot, geom);
            uint32_t rows[4];
            rc_aot_inter_block(v, s.plane, s.bx, s.by, e >> 8, side + ws.pass_base[slot >> 5] + ws.side_off[slot], rows);
            tile_put(ws, s, rows);
        }
        __syncwarp();

        /* ---- phase 3: write the tiles out ---- */
#pragma unroll 1
        for (int u = 0; u < n_units; ++u)
        {
            const int row = u ? geom.row[1] : geom.row[0], mx0 = u ? geom.mx0[1] : geom.mx0[0];
            const int valid_mcbs = min(SYM_SEG_MCBS, mcb_w - mx0);
            uint8_t *const y_dst = v.present + (size_t)(row * 8) * W + mx0 * 8;
            uint8_t *const u_dst = v.present + (size_t)W * H + (size_t)(row * 4) * (W >> 1) + mx0 * 4;
            uint8_t *const v_dst = u_dst + (size_t)(W >> 1) * (H >> 1);
            const uint32_t *tile = ws.tile[u];
            if ((W & 31) == 0)
            {   /* rows of every plane are 16-byte aligned and segments are whole vectors */
                const uint4 *tile4 = reinterpret_cast<const uint4 *>(tile);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                {
                    const int vec = lane + 32 * j, r = vec >> 3, c = vec & 7;
                    if (c * 2 < valid_mcbs) *reinterpret_cast<uint4 *>(y_dst + (size_t)r * W + c * 16) = tile4[vec];
                }
                {
                    const int p = lane >> 4, r = (lane >> 2) & 3, c = lane & 3;
                    uint8_t *dst = (p ? v_dst : u_dst) + (size_t)r * (W >> 1) + c * 16;
                    if (c * 4 < valid_mcbs) *reinterpret_cast<uint4 *>(dst) = tile4[kTileLumaBytes / 16 + lane];
                }
            }
            else
            {   /* odd widths: 4-byte stores, still 128 contiguous bytes per warp instruction */
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (lane < valid_mcbs * 2) *reinterpret_cast<uint32_t *>(y_dst + (size_t)r * W + lane * 4) = tile[r * 32 + lane];
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if ((lane & 15) < valid_mcbs)
                        *reinterpret_cast<uint32_t *>(((lane >> 4) ? v_dst : u_dst) + (size_t)r * (W >> 1) + (lane & 15) * 4) =
                            tile[kTileLumaBytes / 4 + (lane >> 4) * 64 + r * 16 + (lane & 15)];
            }
        }
        __syncwarp();
    }
#undef W
#undef H
#undef nseg
#undef mcb_w
#undef side
#undef segtab
}

template <int kWarps, int kItersPerWarp, int kMinBlocks, int kSegsPerIter = 2>
int launch(const ReconJob *d_jobs, int n_jobs, int units, cudaStream_t stream)
{
    using WarpScratch = WarpScratchT<kSegsPerIter>;
    const int per_cta = kWarps * kItersPerWarp * kSegsPerIter;
    const int ctas_per_pic = (units + per_cta - 1) / per_cta;
    const long long grid = (long long)ctas_per_pic * n_jobs;
    if (grid > 0x7FFFFFFFll) return (int)cudaErrorInvalidConfiguration;
    constexpr size_t smem = sizeof(WarpScratch) * kWarps + RC_SMEM_TABLE_BYTES;
    static const cudaError_t attr = cudaFuncSetAttribute(recon_pictures_kernel<kWarps, kItersPerWarp, kMinBlocks, kSegsPerIter>,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (attr != cudaSuccess) return (int)attr;
    recon_pictures_kernel<kWarps, kItersPerWarp, kMinBlocks, kSegsPerIter><<<(unsigned)grid, kWarps * 32, smem, stream>>>(d_jobs, units, ctas_per_pic);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" int hvqm4_recon_launch(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    const int nseg = (mcb_w + SYM_SEG_MCBS - 1) / SYM_SEG_MCBS;
    const int units = nseg * mcb_h;
    /* HVQM4_RECON_CFG=1..6 pins a configuration (tuning experiments); default: by batch size */
    static const int forced = [] {
        const char *e = getenv("HVQM4_RECON_CFG");
        return e ? atoi(e) : 0;
    }();
    switch (forced)
    {
    case 1: return launch<4, 1, 1>(d_jobs, n_jobs, units, stream);
    case 2: return launch<4, 2, 6>(d_jobs, n_jobs, units, stream);
    case 3: return launch<4, 2, 7>(d_jobs, n_jobs, units, stream);
    case 4: return launch<8, 2, 3>(d_jobs, n_jobs, units, stream);
    case 5: return launch<8, 2, 4>(d_jobs, n_jobs, units, stream);
    case 6: return launch<4, 4, 7>(d_jobs, n_jobs, units, stream);
    case 7: return launch<8, 4, 3, 1>(d_jobs, n_jobs, units, stream);
    case 8: return launch<8, 2, 3, 1>(d_jobs, n_jobs, units, stream);
    case 9: return launch<4, 4, 6, 1>(d_jobs, n_jobs, units, stream);
    case 10: return launch<16, 2, 1, 1>(d_jobs, n_jobs, units, stream);
    case 11: return launch<12, 2, 2, 1>(d_jobs, n_jobs, units, stream);
    default: break;
    }
    /* few pictures: many small CTAs (latency); large batches: amortise the per-CTA nest table */
    if ((long long)n_jobs * units >= 148ll * 8 * 16) return launch<4, 2, 6>(d_jobs, n_jobs, units, stream);
    return launch<4, 1, 1>(d_jobs, n_jobs, units, stream);
}
