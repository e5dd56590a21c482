/*
 * recon.cu -- pixel reconstruction kernels for sm_100a.
 *
 * One STEP reconstructs a batch of independent pictures (one per stream; I, P and B mixed
 * freely) with two launches.  The reference decodes macroblocks serially in raster order
 * (/root/reference/h4m_audio_decode.c:1487-1518 for I pictures, 1922-1967 for P/B), but no
 * block ever depends on another block's *pixels* of the same picture -- only on the
 * completed type/DC maps and on other frames -- so every 4x4 block is an independent work
 * item, and the host, which knows every block's type before it emits side data, can hand
 * the GPU its work already sorted (symbuf.h).
 *
 * recon_map_kernel     everything that follows from the type/DC maps and the reference
 *                      frames alone: weighted-DC fill, flat fill, half-sample motion
 *                      compensation (also the prediction of predicted-AOT blocks).
 *                      warp -> one segment of 16 macroblocks, three passes (upper luma block
 *                      row, lower luma block row, 16 U + 16 V); lane -> one 4x4 block; a warp
 *                      store instruction covers 128 contiguous bytes of a picture row.  No
 *                      prefix sums, no queues, few registers: occupancy hides the latency of
 *                      the scattered reference reads.
 * recon_record_kernel  blocks with side data, from the host-grouped record list: raw blocks
 *                      and AOT basis loops.  Chunks hold <= 32 records of one (class, band,
 *                      length) group, records are at first + index * length (no searching);
 *                      a warp takes up to four chunks and deals their records to its lanes
 *                      32 at a time (full chunks: one chunk per pass, all lanes in one class
 *                      with the same number of bases; the few-record chunks of sparse content:
 *                      one pass for all of them).  The per-picture nest is expanded once per
 *                      CTA into a shared-memory table in which a basis row is a single 32-bit
 *                      load.  Predicted-AOT blocks read back the prediction the map kernel
 *                      left in the picture (L2 resident).
 * Both are launched with programmatic stream serialization (pdl_release / pdl_wait below).
 *
 * The block arithmetic itself lives in recon_core.h.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "recon.h"
#include "recon_core.h"
#include "recon_dev.cuh"

namespace {

/* Map and record kernels are launched with programmatic stream serialization: a kernel may become resident
   while the kernel in front of it in the stream drains, reads what does not depend on it (job, picture header,
   segment heads / chunk descriptors, record headers, nest table -- symbol data uploaded earlier) and waits in
   pdl_wait() before it touches a frame surface.  Every CTA passes pdl_wait(), also on its early exits, so that
   a kernel's completion implies the completion of everything in front of it. */
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

/* ------------------------------------------------------------------------------------------
 * building blocks shared by the kernels
 * ------------------------------------------------------------------------------------------ */

/* pictures with more record chunks than this (dense content: random vectors, ~400 chunks at 640x480) do not
   prefetch reference rows: their gathers already keep the SM's outstanding requests full and the extra
   requests only delay them (measured: -30 % on dense content with the kernel pair, +2.4 % on sparse) */
constexpr uint32_t kPrefetchMaxChunks = 160;

/* what a lane needs before it can start on a segment: the type bytes of its three blocks (upper luma,
   lower luma, chroma) and the two vector words (its luma blocks share a macroblock) */
struct SegHead
{
    uint32_t t0, t1, t2, mv_l, mv_c;
};

__device__ __forceinline__ SegHead segment_head(const ReconView &v, int row, int mx0, int lane)
{
    const int lbx = mx0 * 2 + lane, cbx = mx0 + (lane & 15), cplane = 1 + (lane >> 4);
    const int lstride = (v.width >> 2) + 2, cstride = (v.width >> 3) + 2;
    SegHead h = {0, 0, 0, 0, 0};
    if (lbx < v.mcb_w * 2)
    {
        h.t0 = __ldg(v.blob + v.off_type[0] + (row * 2 + 1) * lstride + lbx + 1);
        h.t1 = __ldg(v.blob + v.off_type[0] + (row * 2 + 2) * lstride + lbx + 1);
        if (!v.is_ipic) h.mv_l = __ldg(reinterpret_cast<const uint32_t *>(v.blob + v.off_mv) + row * v.mcb_w + (lbx >> 1));
    }
    if (cbx < v.mcb_w)
    {
        h.t2 = __ldg(v.blob + (cplane == 1 ? v.off_type[1] : v.off_type[2]) + (row + 1) * cstride + cbx + 1);
        if (!v.is_ipic) h.mv_c = __ldg(reinterpret_cast<const uint32_t *>(v.blob + v.off_mv) + row * v.mcb_w + cbx);
    }
    return h;
}

/* MAP work of one segment (16 macroblocks of macroblock row `row` starting at macroblock mx0):
   three passes, one 4x4 block per lane and pass, 128 contiguous bytes per warp row store.  The head is
   fetched by the caller ahead of time (the next segment's while this one is reconstructed): three
   dependent memory latencies per pass (type -> vector -> reference rows) become the rows alone. */
__device__ __forceinline__ void map_segment(const ReconView &v, int row, int mx0, int lane, const SegHead &head)
{
    const int lbx = mx0 * 2 + lane, cbx = mx0 + (lane & 15), cplane = 1 + (lane >> 4);
    const bool l_ok = lbx < v.mcb_w * 2, c_ok = cbx < v.mcb_w;
    const uint32_t t0 = head.t0, t1 = head.t1, t2 = head.t2, mv_l = head.mv_l, mv_c = head.mv_c;
    /* motion of the three blocks, resolved once: the two luma blocks share the macroblock's vector and
       reference, the lower one starts four rows further down */
    uint32_t mp0 = 0, mp1 = 0, mpc = 0;
    if (!v.is_ipic)
    {
        if (l_ok && (t0 & 0x60))
        {
            mp0 = rc_motion_pack(v, 0, lbx, row * 2, t0, mv_l);
            mp1 = (mp0 & RC_MP_POISON) ? mp0 : mp0 + 4u * (uint32_t)v.width;
        }
        if (c_ok && (t2 & 0x60)) mpc = rc_motion_pack(v, cplane, cbx, row, t2, mv_c);
        /* ask L2 for the reference rows of the lower luma block and of the chroma block while the upper luma
           block is done (+2.4 % on realistic content; the same in the band kernel's classifying walk loses 7 %
           on dense content, see kPrefetchMaxChunks) */
        if (v.n_chunks < kPrefetchMaxChunks)
        {
            if (l_ok && (t0 & 0x60) && !(mp1 & RC_MP_POISON))
            {
                const uint8_t *src = ((mp1 & RC_MP_FUTURE) ? v.ref[1] : v.ref[0]) + RC_MP_OFFSET(mp1);
#pragma unroll
                for (int r = 0; r < 5; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + r * v.width));
            }
            if (c_ok && (t2 & 0x60) && !(mpc & RC_MP_POISON))
            {
                const uint8_t *src = ((mpc & RC_MP_FUTURE) ? v.ref[1] : v.ref[0]) + RC_MP_OFFSET(mpc);
#pragma unroll
                for (int r = 0; r < 5; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + r * (v.width >> 1)));
            }
        }
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass)
    {   /* luma: the segment's two block rows (plane is a compile-time 0 here) */
        const int bx = lbx, by = row * 2 + pass;
        if (!l_ok) continue;
        const int pw = v.width;
        uint32_t rows[4];
        if (!rc_map_block_mp(v, 0, bx, by, pass ? t1 : t0, pass ? mp1 : mp0, rows)) continue;
        uint8_t *dst = v.present + (by * 4) * pw + bx * 4;
#pragma unroll
        for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + r * pw) = rows[r];
    }
    {   /* chroma: lanes 0-15 U, 16-31 V */
        const int plane = cplane, bx = cbx, by = row;
        if (!c_ok) return;
        const int pw = v.width >> 1;
        uint32_t rows[4];
        if (!rc_map_block_mp(v, plane, bx, by, t2, mpc, rows)) return;
        const int plane_off = v.width * v.height + (plane == 2 ? pw * (v.height >> 1) : 0);
        uint8_t *dst = v.present + plane_off + (by * 4) * pw + bx * 4;
#pragma unroll
        for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + r * pw) = rows[r];
    }
}

struct BandOut;
template <bool kTile> __device__ __forceinline__ uint32_t block_off(const ReconView &v, const BandOut &o, int plane, int bx, int by, int &pw);
template <bool kTile> __device__ __forceinline__ void block_st(uint8_t *base, uint32_t off, uint32_t val);
template <bool kTile> __device__ __forceinline__ uint32_t block_ld(const uint8_t *base, uint32_t off);

/* What the band kernel stages in shared memory for the record phase of a band (cp.async.bulk, issued before the map phase,
   landed long before the record phase starts): the chunk descriptors and the records of the band's three classes and
   the band's rows of the vector table.  In the profile of the kernel that read them from global memory a third of all
   stall samples sat on the heads of the record phase's load chain (descriptor -> header -> vector word -> window rows,
   each a trip to L2 or HBM); staged, only the window rows are left.  A class whose records do not fit stays where it is:
   the pointers below are generic and point into shared or global memory. */
struct BandStage
{
    unsigned long long bar;
    int next;                                 /* record phase: chunks of the band not yet taken by a warp (counts down) */
    uint32_t pad0;
    const uint2 *desc[SYM_REC_CLASSES];       /* descriptor of chunk c of class k: desc[k][c - c0[k]] */
    const uint32_t *rec[SYM_REC_CLASSES];     /* record word w (counted from the picture's first record word): rec[k][w] */
    const uint32_t *mv;                       /* vector word of macroblock (mx, my): mv[my * mcb_w + mx] */
    uint32_t c0[SYM_REC_CLASSES], c1[SYM_REC_CLASSES];
};

/* RECORD work of one chunk of class cls: one record per lane */
template <bool kTile>
__device__ __forceinline__ void record_chunk(const ReconView &v, const BandOut &o, const BandStage &st, int cls, uint32_t c, int lane)
{
    const uint2 cd = st.desc[cls][c - st.c0[cls]];
    const uint32_t count = cd.y & 0xFF, len = ((cd.y >> 8) & 0xFF) + 1;
    if ((uint32_t)lane >= count) return;
    const uint32_t *rec = st.rec[cls] + cd.x + lane * len;
    const uint32_t hdr = *rec;
    uint32_t t;
    int plane, bx, by;
    rc_record_coords(hdr, t, plane, bx, by);
    uint32_t extra = 0;
    if (cls == SYM_REC_INTER) extra = st.mv[(plane ? by : by >> 1) * v.mcb_w + (plane ? bx : bx >> 1)];
    else if (cls == SYM_REC_INTRA) extra = rc_record_extra(v, cls, hdr);
    int pw;
    const uint32_t dst = block_off<kTile>(v, o, plane, bx, by, pw);
    uint8_t *const pic = v.present;
    uint32_t rows[4];
    if (cls == SYM_REC_INTER)
    {   /* the prediction left by the map work; L2-coherent loads (it may have been written by another warp) */
#pragma unroll
        for (int r = 0; r < 4; ++r) rows[r] = block_ld<kTile>(pic, dst + r * pw);
    }
    rc_record_block_pre<true>(v, cls, len, rec, hdr, extra, rows);
#pragma unroll
    for (int r = 0; r < 4; ++r) block_st<kTile>(pic, dst + r * pw, rows[r]);
}

/* the same with the chunk descriptor, the lane's header word and rc_record_extra() fetched by the caller ahead of time */
__device__ __forceinline__ void record_chunk_pre(const ReconView &v, uint2 cd, int lane, uint32_t hdr, uint32_t extra)
{
    const uint32_t count = cd.y & 0xFF, len = ((cd.y >> 8) & 0xFF) + 1;
    const int cls = (int)((cd.y >> 16) & 0xFF);
    if ((uint32_t)lane >= count) return;
    const uint32_t *rec = v.rec + cd.x + lane * len;
    uint32_t t;
    int plane, bx, by;
    rc_record_coords(hdr, t, plane, bx, by);
    const int pw = plane ? v.width >> 1 : v.width;
    const int plane_off = plane == 0 ? 0 : plane == 1 ? v.width * v.height : v.width * v.height + (v.width >> 1) * (v.height >> 1);
    uint8_t *dst = v.present + plane_off + (by * 4) * pw + bx * 4;
    uint32_t rows[4];
    if (cls == SYM_REC_INTER)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r) rows[r] = __ldcg(reinterpret_cast<const uint32_t *>(dst + r * pw));
    }
    rc_record_block_pre(v, cls, len, rec, hdr, extra, rows);
#pragma unroll
    for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + r * pw) = rows[r];
}

/* ------------------------------------------------------------------------------------------
 * map kernel (small batches: many CTAs per picture)
 * ------------------------------------------------------------------------------------------ */
template <int kWarps, int kUnitsPerWarp, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
recon_map_kernel(const ReconJob *__restrict__ jobs, int units_per_pic, int ctas_per_pic)
{
    ReconView &vw = *reinterpret_cast<ReconView *>(rc_smem + RC_SMEM_VIEW_OFF);
    const int job = blockIdx.x / ctas_per_pic;
    const int cta = blockIdx.x - job * ctas_per_pic;
    if (threadIdx.x == 0) load_view(vw, jobs[job]);
    __syncthreads();
    const ReconView &v = vw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int unit = cta * kUnitsPerWarp * kWarps + warp;
    pdl_release();
    if (unit >= units_per_pic)
    {
        pdl_wait();
        return;
    }
    const int nseg = v.nseg;
    int row = unit / nseg, seg = unit - row * nseg;       /* one division per warp; then stepped */
    SegHead head = segment_head(v, row, seg * SYM_SEG_MCBS, lane);
    pdl_wait();                                           /* the surfaces belong to the kernels in front until here */
#pragma unroll 1
    for (int it = 0; it < kUnitsPerWarp; ++it)
    {
        const int next = unit + kWarps;
        const bool more = it + 1 < kUnitsPerWarp && next < units_per_pic;
        int nrow = row, nseg_i = seg + kWarps;
        while (nseg_i >= nseg) { nseg_i -= nseg; ++nrow; }
        SegHead nhead = {0, 0, 0, 0, 0};
        if (more) nhead = segment_head(v, nrow, nseg_i * SYM_SEG_MCBS, lane);
        map_segment(v, row, seg * SYM_SEG_MCBS, lane, head);
        if (!more) break;
        unit = next; row = nrow; seg = nseg_i; head = nhead;
    }
}

/* ------------------------------------------------------------------------------------------
 * record kernel (small batches)
 * shared memory: [nest table | mcdiv | div | view] at the fixed offsets of recon_core.h, then
 * a scratch area used only while the nest table is being built
 * ------------------------------------------------------------------------------------------ */
constexpr int kRecWarps = 8;
constexpr int kRecSmem = RC_SMEM_TABLE_BYTES + SYM_NEST_H * 40 + 32;
/* band kernels: [tables + view | nest staging scratch | BandStage | queues that do not fit over the nest table | staged record
   data | output tile (kTile)] */
constexpr int kBandStageOff = (kRecSmem + 15) & ~15;
constexpr int kBandQueueOff = kBandStageOff + 128;
/* The warps' queues (16-bit entries, band_map_tile) live where the NEST TABLE will be: the table is spread only after the map
   phase, when the queues are dead (band_item); the warps whose queues do not fit there follow the BandStage.  Shared memory
   is L1 the gathers do not get: with 32-bit entries behind the tables the tile variant needed 96 KB per CTA -- two CTAs =
   the 196 KB carve-out, 60 KB of L1 -- now 75 KB, the 164 KB carve-out and 92 KB of L1 (profiles/r02_carve_ab.txt: the
   plain kernel loses 3 % from 124 to 92 KB of L1, 11 % more from 92 to 60 KB, 40 % more from 60 to 28 KB). */
constexpr int kBandQueueAlias = RC_SMEM_MCDIV_OFF - RC_SMEM_NEST_OFF;
__device__ __forceinline__ uint16_t *band_queue(int warp, int cap)
{
    const int per = cap * 2, n_alias = kBandQueueAlias / per;
    return reinterpret_cast<uint16_t *>(rc_smem + (warp < n_alias ? RC_SMEM_NEST_OFF + warp * per : kBandQueueOff + (warp - n_alias) * per));
}
/* bytes behind kBandQueueOff for the queues of `warps` warps of `cap` entries */
__host__ __device__ constexpr int band_queue_extra_dev(int warps, int cap)
{
    return warps > kBandQueueAlias / (cap * 2) ? (warps - kBandQueueAlias / (cap * 2)) * cap * 2 : 0;
}
static inline int band_queue_extra(int warps, int cap) { return band_queue_extra_dev(warps, cap); }

template <int kMinBlocks>
__global__ void __launch_bounds__(kRecWarps * 32, kMinBlocks)
recon_record_kernel(const ReconJob *__restrict__ jobs, int n_jobs, uint32_t cta_base)
{
    ReconView &vw = *reinterpret_cast<ReconView *>(rc_smem + RC_SMEM_VIEW_OFF);
    __shared__ uint32_t s_cta_in_pic;

    /* which picture does this CTA belong to: last job with rec_cta_begin <= blockIdx.x.  Warp 0 searches 32 ways
       (two rounds of independent loads for 1 024 pictures instead of ten dependent ones) */
    if (threadIdx.x < 32)
    {
        const uint32_t gcta = blockIdx.x + cta_base;     /* rec_cta_begin is a prefix over the whole step */
        const int l = (int)threadIdx.x;
        int lo = 0, hi = n_jobs;                         /* answer in [lo, hi); jobs[lo].rec_cta_begin <= gcta */
        while (hi - lo > 1)
        {
            const int step = (hi - lo + 31) >> 5, idx = lo + l * step;
            const bool ok = idx < hi && __ldg(&jobs[idx].rec_cta_begin) <= gcta;
            const int k = 31 - __clz((int)(__ballot_sync(0xFFFFFFFFu, ok) | 1u));
            lo += k * step;
            hi = min(lo + step, hi);
        }
        if (l == 0)
        {
            s_cta_in_pic = gcta - __ldg(&jobs[lo].rec_cta_begin);
            load_view(vw, jobs[lo]);
        }
    }
    build_div_tables<kRecWarps * 32>();
    pdl_release();
    __syncthreads();
    const ReconView &v = vw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t chunk0 = s_cta_in_pic * HVQM4_REC_CHUNKS_PER_CTA;
    const uint32_t chunk_end = min(chunk0 + HVQM4_REC_CHUNKS_PER_CTA, v.n_chunks);
    if (chunk0 < v.n_chunks_nest) nest_stage_begin<kRecWarps * 32>(v, rc_smem + RC_SMEM_TABLE_BYTES);
    /* A warp has up to kPerWarp chunks (chunk0 + warp + k * kRecWarps).  On the content this kernel runs on -- few
       records per picture -- a chunk holds only a few records (9 of 32 lanes were active) and every record is
       reached through a chain of dependent loads (chunk descriptor -> header word -> vector word or DC ->
       reference window rows) with an idle memory system around it.  So the warp's chunks are FLATTENED: lane i
       fetches the descriptor of the i-th chunk, the record counts are prefix-summed, and the records of all its
       chunks are dealt to the lanes 32 at a time (descriptor by shuffle from the owning lane): typically one walk
       of the chain per warp instead of one per chunk, classes diverging inside the pass.  Full chunks (dense content) still
       come out as one chunk per pass.  Descriptors and the first headers are requested before the nest table is
       built, the next pass's headers before the current pass is computed. */
    constexpr uint32_t kPerWarp = (HVQM4_REC_CHUNKS_PER_CTA + kRecWarps - 1) / kRecWarps;
    static_assert(kPerWarp >= 1 && kPerWarp <= 32, "one descriptor per lane");
    const uint32_t first = chunk0 + (uint32_t)warp;
    const uint32_t n = first < chunk_end ? (chunk_end - first + kRecWarps - 1) / kRecWarps : 0u;
    const uint2 my_cd = (uint32_t)lane < n ? __ldg(reinterpret_cast<const uint2 *>(v.chunks) + first + (uint32_t)lane * kRecWarps) : make_uint2(0u, 0u);
    const uint32_t my_count = my_cd.y & 0xFF;           /* 0 beyond the warp's chunks */
    uint32_t start[kPerWarp + 1];                       /* first record of chunk k; start[kPerWarp] = all of them */
    start[0] = 0;
#pragma unroll
    for (uint32_t k = 0; k < kPerWarp; ++k) start[k + 1] = start[k] + __shfl_sync(0xFFFFFFFFu, my_count, (int)k);
    const uint32_t total = start[kPerWarp];
    /* record g of the warp -> (descriptor, index inside its chunk, header word) */
    auto fetch = [&](uint32_t g, uint2 &cd, uint32_t &r, uint32_t &hdr) {
        uint32_t i = 0, s0 = 0;
#pragma unroll
        for (uint32_t k = 1; k < kPerWarp; ++k)
            if (g >= start[k]) { i = k; s0 = start[k]; }
        r = g - s0;
        cd = make_uint2(__shfl_sync(0xFFFFFFFFu, my_cd.x, (int)i), __shfl_sync(0xFFFFFFFFu, my_cd.y, (int)i));
        hdr = g < total ? __ldg(v.rec + cd.x + r * (((cd.y >> 8) & 0xFF) + 1)) : 0u;
    };
    uint2 cd;
    uint32_t r, hdr;
    fetch((uint32_t)lane, cd, r, hdr);
    if (chunk0 < v.n_chunks_nest)
    {
        nest_stage_wait();
        __syncthreads();
        nest_spread<kRecWarps * 32>(rc_smem + RC_SMEM_TABLE_BYTES, v.portrait != 0);
        __syncthreads();
    }
    pdl_wait();                                           /* the map kernel's predictions, and the surfaces at all */
#pragma unroll 1
    for (uint32_t base = 0; base < total; base += 32)
    {
        const bool active = base + (uint32_t)lane < total;
        const uint32_t extra = active ? rc_record_extra(v, (int)((cd.y >> 16) & 0xFF), hdr) : 0u;
        uint2 cd_n = make_uint2(0u, 0u);
        uint32_t r_n = 0, hdr_n = 0;
        if (base + 32 < total) fetch(base + 32 + (uint32_t)lane, cd_n, r_n, hdr_n);     /* warp-uniform condition */
        if (active) record_chunk_pre(v, cd, (int)r, hdr, extra);
        cd = cd_n;
        r = r_n;
        hdr = hdr_n;
    }
}

/* ------------------------------------------------------------------------------------------
 * band kernel (large batches): one CTA reconstructs one BAND (SYM_BAND_MCB_ROWS macroblock rows)
 * of one picture completely -- its map work, a block barrier, then exactly the records that lie
 * in the band (the host groups records per band, symbuf.h).  The sectors the record work
 * rewrites, and the reference rows both phases read, are then still in L2/L1 instead of making a
 * second round trip to HBM between two kernels.
 *
 * Map work is CLASS-SORTED: a warp that takes 32 neighbouring blocks as they come runs the
 * weighted-fill code, the motion-compensation code and the flat fill one after the other with
 * a third of its lanes each (measured: 15 of 32 lanes active, 63 % of the kernel's
 * instructions).  Instead every warp first walks the type map of its block rows (coalesced byte
 * loads, one block per lane), finishes flat blocks on the spot, and pushes the other blocks into
 * its private shared-memory queue -- weighted blocks from the front, motion-compensated blocks
 * from the back, positions from a ballot -- and then drains 32 entries of ONE class at a time.
 * Wide pictures are walked in column tiles of kTileMcbs macroblocks so that the queue has a
 * fixed upper size.
 *
 * Two variants (template parameter kTile).  Plain: finished blocks are stored straight into the picture.  Tile (the
 * default wherever two CTAs of it fit an SM): the band is assembled in shared memory -- same layout as its part of the
 * picture -- and leaves through three bulk stores; sixteen warps per CTA, 16-bit queue entries in the space of the nest
 * table, record chunks taken from a shared counter.  What the gathers of both phases get as L1 decides the speed
 * (last table), so every byte of shared memory here is accounted for.
 * ------------------------------------------------------------------------------------------ */
#ifndef HVQM4_BAND_WARPS
#define HVQM4_BAND_WARPS 8
#endif
constexpr int kBandWarps = HVQM4_BAND_WARPS;     /* plain band kernel: three or four CTAs of this many warps per SM */
#ifndef HVQM4_BAND_TILE_WARPS
#define HVQM4_BAND_TILE_WARPS 16
#endif
/* tile variant: two CTAs per SM by shared memory (75 KB + queues each), so sixteen warps per CTA at 64 registers (no spills)
   give the SM 32 warps where the plain kernel has 24, and the 32 row tasks of a band are two per warp.  Measured dense
   (profiles/r02_tile_warps2_ab.txt, after the queues moved over the nest table): 12 warps 1.307 M, 13 1.298 M, 14 1.305 M,
   16 warps 1.360 M frames/s (128 pictures: 1.08 -> 1.14 M); before that, at 96+ KB per CTA: 8 warps 1.09 M, 10 1.16 M,
   12 1.23 M, 14 and 16 one CTA per SM.  Plain kernel 1.20-1.22 M. */
constexpr int kBandTileWarps = HVQM4_BAND_TILE_WARPS;
/* warps per CTA of a band kernel instantiation: kBandTileWarps for the 8-row tile variant (two CTAs per SM), eight otherwise */
__host__ __device__ constexpr int band_warps(bool tile, int rows) { return tile && rows == 8 ? kBandTileWarps : kBandWarps; }
constexpr int kTileMcbs = 128;
constexpr int kBandRows = 8;   /* macroblock rows per CTA of the band kernel = kBandRows record bands of symbuf.h */
static_assert(SYM_BAND_MCB_ROWS == kBandRows, "a CTA of the band kernel takes one record band of 8 rows, or 8 bands of one row");

/* queue capacity of one warp in entries: its four block rows (two luma, one U, one V) of one column tile */
static inline int band_queue_entries(int mcb_w, int rows = 8)
{
    /* 8 rows per CTA: two luma and two chroma block rows per warp; 4 rows: one of each */
    return (mcb_w < kTileMcbs ? mcb_w : kTileMcbs) * (rows > 4 ? 6 : 3);
}

/* Where the band kernel's blocks go.  kTile = false: straight into the picture (4-byte stores scattered over the
   band: every 32-byte sector of a picture row is written 3-4 times, once per class, and the predictions of
   predicted-AOT blocks make a round trip through L2).  kTile = true: the band is assembled in shared memory -- same
   layout as its part of the picture: (row1 - row0) * 8 luma rows, then the U rows, then the V rows -- and leaves with
   three bulk stores of whole picture rows when the CTA is done. */
struct BandOut
{
    uint32_t tile_off;       /* shared-memory offset of the tile */
    uint32_t ybytes, cbytes; /* luma bytes, bytes of one chroma plane of the band */
    int row0;                /* first macroblock row of the band */
};

template <bool kTile>
__device__ __forceinline__ uint32_t block_off(const ReconView &v, const BandOut &o, int plane, int bx, int by, int &pw)
{
    pw = plane ? v.width >> 1 : v.width;
    if (kTile)
    {
        const uint32_t plane_off = plane == 0 ? 0u : plane == 1 ? o.ybytes : o.ybytes + o.cbytes;
        const int lby = by - (plane ? o.row0 : 2 * o.row0);
        return o.tile_off + plane_off + (uint32_t)((lby * 4) * pw + bx * 4);
    }
    const int plane_off = plane == 0 ? 0 : plane == 1 ? v.width * v.height : v.width * v.height + (v.width >> 1) * (v.height >> 1);
    return (uint32_t)(plane_off + (by * 4) * pw + bx * 4);
}
/* base = the picture (read ONCE by the caller: behind a store through a generic pointer the compiler reloads the view's
   `present` from shared memory for every row -- 6.6 % of the kernel's instructions); unused with kTile */
template <bool kTile>
__device__ __forceinline__ void block_st(uint8_t *base, uint32_t off, uint32_t val)
{
    if (kTile) *reinterpret_cast<uint32_t *>(rc_smem + off) = val;
    else *reinterpret_cast<uint32_t *>(base + off) = val;
}
/* the prediction the map phase left for a predicted-AOT block (another warp may have written it) */
template <bool kTile>
__device__ __forceinline__ uint32_t block_ld(const uint8_t *base, uint32_t off)
{
    if (kTile) return *reinterpret_cast<const uint32_t *>(rc_smem + off);
    return __ldcg(reinterpret_cast<const uint32_t *>(base + off));
}

/* map work of macroblock rows [row0, row1) x macroblock columns [mx0, mx1) of the CTA's picture.
   Every warp owns up to four block rows (two luma, one U, one V) and a private queue of `cap`
   entries: no atomics, no block barrier, counters in (warp-uniform) registers.  The classifying
   walk stays minimal and uniform (type and DC of the next 32 blocks are loaded while the current
   32 are classified); everything that costs instructions runs in the drains, where all lanes of
   a warp do the same thing, two queue entries per lane at a time so that twice as many
   reference rows are in flight. */
template <bool kTile, int kWarps>
__device__ __forceinline__ void band_map_tile(const ReconView &v, const BandOut &o, int row0, int row1, int mx0, int mx1, uint16_t *q, int cap)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const bool ipic = v.is_ipic != 0;
    uint8_t *const pic = v.present;
    uint32_t n_w = 0, n_mc = 0;
    /* classify: row tasks = 2 luma block rows per macroblock row, then the U rows, then the V rows.
       Queue entry (16 bits): [7:0] block x inside the column tile, [9:8] which of the warp's tasks, [10] the reference is
       the future picture -- everything else a drain needs follows from the task */
    static_assert(kTileMcbs * 2 <= 256, "a queue entry holds the block x inside the column tile in 8 bits");
    const int mrows = row1 - row0, n_tasks = mrows * 4;      /* at most 4 per warp: band_item asserts it */
    uint32_t slot = 0;
#pragma unroll 1
    for (int task = warp; task < n_tasks; task += kWarps, slot += 1u << 8)
    {
        const int plane = task < 2 * mrows ? 0 : task < 3 * mrows ? 1 : 2;
        const int by = plane == 0 ? row0 * 2 + task : row0 + (task - (plane + 1) * mrows);
        const int sh = plane ? 0 : 1;
        const int x0 = mx0 << sh, x1 = mx1 << sh;
        const int pw = v.width >> (plane ? 1 : 0), bstride = (pw >> 2) + 2;
        const uint8_t *trow = v.blob + (rc_pick3(v.off_type, plane) + (by + 1) * bstride + 1);
        const uint8_t *drow = v.blob + (rc_pick3(v.off_dc, plane) + (by + 1) * bstride + 1);
        int pw2;
        const uint32_t dst_row = block_off<kTile>(v, o, plane, 0, by, pw2);
        /* lanes past the row end get type 6 (raw: nothing to do here) */
        int bx = x0 + lane;
        uint32_t t = 6u, dc = 0u;
        if (bx < x1)
        {
            t = __ldg(trow + bx);
            dc = __ldg(drow + bx);
        }
#pragma unroll 1
        for (int xb = x0; xb < x1; xb += 32)
        {
            const int bx_n = bx + 32;
            uint32_t t_n = 6u, dc_n = 0u;
            if (bx_n < x1)
            {
                t_n = __ldg(trow + bx_n);
                dc_n = __ldg(drow + bx_n);
            }
            const bool inter = !ipic && (t & 0x60);
            const uint32_t nib = ipic ? t : (t & 0xF);
            const bool is_w = !inter && nib == 0;
            const bool is_mc = inter && ((t & 0x10) || nib != 6);
            const uint32_t entry = slot + (uint32_t)(bx - x0) + ((t & 0x40) << 4);
            const uint32_t m_w = __ballot_sync(0xFFFFFFFFu, is_w), m_mc = __ballot_sync(0xFFFFFFFFu, is_mc);
            if (is_w) q[n_w + __popc(m_w & lt)] = (uint16_t)entry;
            if (is_mc) q[cap - 1 - (int)(n_mc + __popc(m_mc & lt))] = (uint16_t)entry;
            n_w += __popc(m_w);
            n_mc += __popc(m_mc);
            if (!inter && nib == 8)
            {   /* flat fill, h4m:281 */
                const uint32_t V = dc * 0x01010101u;
                const uint32_t dst = dst_row + bx * 4;
#pragma unroll
                for (int r = 0; r < 4; ++r) block_st<kTile>(pic, dst + r * pw, V);
            }
            bx = bx_n; t = t_n; dc = dc_n;
        }
    }
    __syncwarp();
    /* entry -> plane, block coordinates, and the one type bit motion compensation looks at (past / future) */
    auto coords = [&](uint32_t e, uint32_t &t, int &plane, int &bx, int &by) {
        const int task = warp + (int)((e >> 8) & 3u) * kWarps;
        plane = task < 2 * mrows ? 0 : task < 3 * mrows ? 1 : 2;
        by = plane == 0 ? row0 * 2 + task : row0 + (task - (plane + 1) * mrows);
        bx = (int)(e & 0xFFu) + (plane ? mx0 : mx0 << 1);
        t = (e & 0x400u) ? 0x40u : 0x20u;
    };
#pragma unroll 1
    for (uint32_t i = lane; i < n_w; i += 32)
    {
        uint32_t t, rows[4];
        int plane, bx, by, pw;
        coords(q[i], t, plane, bx, by);
        rc_weighted_block(v, plane, bx, by, rows);
        const uint32_t dst = block_off<kTile>(v, o, plane, bx, by, pw);
#pragma unroll
        for (int r = 0; r < 4; ++r) block_st<kTile>(pic, dst + r * pw, rows[r]);
    }
    /* motion compensation: two entries per lane and round; a last round of at most 32 entries (with sixteen warps per CTA a
       warp holds ~96: every second round) takes one entry per lane instead of computing the same block twice */
    const uint16_t *q_mc = q + cap - 1;
#pragma unroll 1
    for (uint32_t i = lane; i - lane < n_mc; i += 64)
    {
        if (n_mc - (i - lane) <= 32u)
        {
            if (i < n_mc)
            {
                uint32_t t0, rows0[4];
                int plane0, bx0, by0, pw0;
                coords(q_mc[-(int)i], t0, plane0, bx0, by0);
                rc_mc_packed(v, plane0, rc_motion_pack(v, plane0, bx0, by0, t0, rc_mv_word(v, plane0, bx0, by0)), rows0);
                const uint32_t dst0 = block_off<kTile>(v, o, plane0, bx0, by0, pw0);
#pragma unroll
                for (int r = 0; r < 4; ++r) block_st<kTile>(pic, dst0 + r * pw0, rows0[r]);
            }
            break;
        }
        const bool two = i + 32 < n_mc;
        uint32_t t0, t1, rows0[4], rows1[4];
        int plane0, bx0, by0, plane1, bx1, by1, pw0, pw1;
        coords(q_mc[-(int)i], t0, plane0, bx0, by0);
        coords(q_mc[-(int)(two ? i + 32 : i)], t1, plane1, bx1, by1);
        const uint32_t mv0 = rc_mv_word(v, plane0, bx0, by0), mv1 = rc_mv_word(v, plane1, bx1, by1);
        const uint32_t mp0 = rc_motion_pack(v, plane0, bx0, by0, t0, mv0), mp1 = rc_motion_pack(v, plane1, bx1, by1, t1, mv1);
        rc_mc_packed2(v, plane0, mp0, rows0, plane1, mp1, rows1);
        const uint32_t dst0 = block_off<kTile>(v, o, plane0, bx0, by0, pw0);
#pragma unroll
        for (int r = 0; r < 4; ++r) block_st<kTile>(pic, dst0 + r * pw0, rows0[r]);
        if (two)
        {
            const uint32_t dst1 = block_off<kTile>(v, o, plane1, bx1, by1, pw1);
#pragma unroll
            for (int r = 0; r < 4; ++r) block_st<kTile>(pic, dst1 + r * pw1, rows1[r]);
        }
    }
    __syncwarp();
}

/* one band of kRows macroblock rows of one picture: the whole CTA.  tile_off: shared-memory offset of the output tile (kTile) */
/* stage_off / stage_cap: shared-memory area for the band's record data (0 bytes: nothing is staged); stage_phase: how often
   the CTA has used the staging barrier before (the walk kernel reuses it) */
template <bool kTile, int kRows, int kWarps>
__device__ __forceinline__ bool band_item(const ReconJob *__restrict__ jobs, int job, int band, uint16_t *queue, int queue_cap, uint32_t tile_off,
                                          uint32_t stage_off, uint32_t stage_cap, uint32_t stage_phase)
{
    static_assert(kRows * 4 <= 4 * kWarps, "a queue entry names one of at most four row tasks of its warp (band_map_tile)");
    ReconView &vw = *reinterpret_cast<ReconView *>(rc_smem + RC_SMEM_VIEW_OFF);
    BandStage &st = *reinterpret_cast<BandStage *>(rc_smem + kBandStageOff);
    if (threadIdx.x == 0)
    {
        load_view(vw, jobs[job]);
        if (stage_phase == 0)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&st.bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const ReconView &v = vw;
    if (!v.blob) return false;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (v.has_nest) nest_stage_begin<kWarps * 32>(v, rc_smem + RC_SMEM_TABLE_BYTES);   /* lands during the map phase */

    const int row0 = band * kRows, row1 = min(row0 + kRows, v.mcb_h);
    /* the chunks of a class are ordered by record band: bands of 8, 4 or 1 macroblock rows (h4e_set_band_rows; kRows is
       a multiple), of which rows row0..row1 are one range */
    const int nb = (int)v.n_bands;
    const int rpb = nb == v.mcb_h ? 1 : nb == (v.mcb_h + 3) / 4 ? 4 : 8;
    const int b0 = row0 / rpb, b1 = min((row1 + rpb - 1) / rpb, nb);
    if (warp == kWarps - 1)
    {   /* the last warp (it has the fewest block rows of the map phase when they do not divide) requests the band's record
           data: lane k < 3 looks up class k's chunk and record range (two dependent loads), lane 3 the vector rows */
        const uint32_t nb1 = v.n_bands + 1;
        uint32_t src = 0, bytes = 0, head = 0;         /* source offset in the blob, bytes to copy (16-byte units), misalignment */
        uint32_t c0 = 0, c1 = 0, r0 = 0, r1 = 0;
        if (lane < SYM_REC_CLASSES)
        {
            c0 = __ldg(v.bands + lane * nb1 + b0);
            c1 = __ldg(v.bands + lane * nb1 + b1);
            const uint32_t n_words = __ldg(reinterpret_cast<const uint32_t *>(v.blob + offsetof(SymHeader, n_rec_words)));
            r0 = c0 < v.n_chunks ? __ldg(v.chunks + 2 * c0) : n_words;
            r1 = c1 < v.n_chunks ? __ldg(v.chunks + 2 * c1) : n_words;
        }
        /* seven ranges of the blob: descriptors per class (lanes 0-2), vectors (lane 3), records per class (lanes 4-6: predicted
           AOT first -- its chain is the longest --, then intra AOT, then raw: what does not fit is the cheapest to leave) */
        const int k = lane < 4 ? lane & 3 : 6 - lane;
        const uint32_t kc0 = __shfl_sync(0xFFFFFFFFu, c0, k), kc1 = __shfl_sync(0xFFFFFFFFu, c1, k);
        const uint32_t kr0 = __shfl_sync(0xFFFFFFFFu, r0, k), kr1 = __shfl_sync(0xFFFFFFFFu, r1, k);
        uint32_t lo = 0, hi = 0;
        if (lane < 3) { lo = (uint32_t)((const uint8_t *)v.chunks - v.blob) + kc0 * 8u; hi = lo + (kc1 - kc0) * 8u; }
        else if (lane == 3 && !v.is_ipic) { lo = v.off_mv + (uint32_t)(row0 * v.mcb_w * 4); hi = lo + (uint32_t)((row1 - row0) * v.mcb_w * 4); }
        else if (lane >= 4 && lane < 7) { lo = (uint32_t)((const uint8_t *)v.rec - v.blob) + kr0 * 4u; hi = (uint32_t)((const uint8_t *)v.rec - v.blob) + kr1 * 4u; }
        if (hi > lo) { src = lo & ~15u; head = lo - src; bytes = ((hi + 15u) & ~15u) - src; }
        /* places in the staging area by prefix sum; a range that does not fit (and everything behind it) stays in global memory */
        uint32_t at = bytes;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1)
        {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, at, d);
            if ((lane & 7) >= d) at += up;
        }
        const bool fits = stage_cap != 0 && at <= stage_cap;
        const uint32_t dst = stage_off + at - bytes;
        const uint8_t *base = fits ? rc_smem + dst + head : v.blob + lo;      /* where byte `lo` of the blob is read from */
        if (lane < 3)
        {
            st.desc[lane] = reinterpret_cast<const uint2 *>(base);
            st.c0[lane] = kc0;
            st.c1[lane] = kc1;
        }
        else if (lane == 3) st.mv = reinterpret_cast<const uint32_t *>(base) - row0 * v.mcb_w;
        else if (lane < 7) st.rec[k] = reinterpret_cast<const uint32_t *>(base) - kr0;
        {   /* chunks of all three classes (lanes 0..2 hold one class each: k == lane there) */
            const uint32_t n = kc1 - kc0;
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, n, 0) + __shfl_sync(0xFFFFFFFFu, n, 1) + __shfl_sync(0xFFFFFFFFu, n, 2);
            if (lane == 0) st.next = (int)total;
        }
        uint32_t tx = fits && lane < 7 ? bytes : 0u;
#pragma unroll
        for (int d = 4; d; d >>= 1) tx += __shfl_xor_sync(0xFFFFFFFFu, tx, d);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&st.bar);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
        __syncwarp();
        if (fits && lane < 7 && bytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(rc_smem + dst)), "l"(v.blob + src), "r"(bytes), "r"(bar) : "memory");
    }

    /* map phase */
    const BandOut out = {tile_off, (uint32_t)((row1 - row0) * 8 * v.width), (uint32_t)((row1 - row0) * 4 * (v.width >> 1)), row0};
    for (int mx0 = 0; mx0 < v.mcb_w; mx0 += kTileMcbs)
        band_map_tile<kTile, kWarps>(v, out, row0, row1, mx0, min(mx0 + kTileMcbs, v.mcb_w), queue, queue_cap);
    /* record phase */
    if (v.has_nest) nest_stage_wait();
    __syncthreads();     /* map stores of the band and the staging pointers visible to the whole CTA */
    {   /* the band's record data has landed (it was requested before the map phase) */
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&st.bar);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar), "r"(stage_phase & 1u) : "memory");
    }
    const uint32_t raw0 = st.c0[SYM_REC_RAW], raw1 = st.c1[SYM_REC_RAW];
    const uint32_t intra0 = st.c0[SYM_REC_INTRA], intra1 = st.c1[SYM_REC_INTRA];
    const uint32_t inter0 = st.c0[SYM_REC_INTER];
    if (intra1 > intra0)
    {
        nest_spread<kWarps * 32>(rc_smem + RC_SMEM_TABLE_BYTES, v.portrait != 0);
        __syncthreads();     /* nest table complete */
    }
    /* The band's chunks are TAKEN by the warps one at a time from a shared counter, the expensive ones first (it counts
       down: predicted AOT, then intra AOT, then raw; inside a class the chunks are sorted by record length, longest last).
       Dealt out statically -- warp, warp + kWarps, ... per class -- a band's ~100 chunks of 1 to 6 words per record left the
       warps up to three chunks apart and 9 % of all warp samples sat at the barrier behind this loop. */
    const int n_raw = (int)(raw1 - raw0), n_intra = (int)(intra1 - intra0);
#pragma unroll 1
    for (;;)
    {
        int u = 0;
        if (lane == 0) u = atomicSub(&st.next, 1) - 1;
        u = __shfl_sync(0xFFFFFFFFu, u, 0);
        if (u < 0) break;
        if (u >= n_raw + n_intra) record_chunk<kTile>(v, out, st, SYM_REC_INTER, inter0 + (uint32_t)(u - n_raw - n_intra), lane);
        else if (u >= n_raw) record_chunk<kTile>(v, out, st, SYM_REC_INTRA, intra0 + (uint32_t)(u - n_raw), lane);
        else record_chunk<kTile>(v, out, st, SYM_REC_RAW, raw0 + (uint32_t)u, lane);
    }
    if (kTile)
    {   /* the band leaves as whole picture rows: three bulk stores (shared memory -> picture) */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     /* this thread's tile writes -> visible to the copies */
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const uint32_t tile = (uint32_t)__cvta_generic_to_shared(rc_smem + tile_off);
            uint8_t *py = v.present + (size_t)row0 * 8 * v.width;
            uint8_t *pu = v.present + (size_t)v.width * v.height + (size_t)row0 * 4 * (v.width >> 1);
            uint8_t *pv = pu + (size_t)(v.width >> 1) * (v.height >> 1);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(py), "r"(tile), "r"(out.ybytes) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(pu), "r"(tile + out.ybytes), "r"(out.cbytes) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(pv), "r"(tile + out.ybytes + out.cbytes), "r"(out.cbytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            /* the CTA's shared memory is released when it exits: the copies must have READ it by then (their writes to the
               picture complete by the end of the kernel like any store) */
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    return true;
}

/* one CTA per (picture, band) */
template <int kMinBlocks, bool kTile, int kRows>
__global__ void __launch_bounds__(band_warps(kTile, kRows) * 32, kMinBlocks)
recon_band_kernel(const ReconJob *__restrict__ jobs, int n_bands, int queue_cap, int stage_cap)
{
    constexpr int kWarps = band_warps(kTile, kRows);
    uint16_t *queue = band_queue(threadIdx.x >> 5, queue_cap);   /* the warp's own */
    build_div_tables<kWarps * 32>();
    const int job = blockIdx.x / n_bands;
    const uint32_t stage_off = (uint32_t)((kBandQueueOff + band_queue_extra_dev(kWarps, queue_cap) + 127) & ~127);
    band_item<kTile, kRows, kWarps>(jobs, job, blockIdx.x - job * n_bands, queue, queue_cap, stage_off + (uint32_t)stage_cap, stage_off, (uint32_t)stage_cap, 0u);
}

/* The fallback launch behind the sweep kernel (skip_handled = 1: pictures whose job says pad[0] = 1 are already
   reconstructed) or the row kernel (skip_handled = 2: only pictures whose job says pad[1] = 1 are left): a few CTAs per
   SM walk all items, because nearly all of them are skipped. */
__global__ void __launch_bounds__(kBandWarps * 32, 2)
recon_band_walk_kernel(const ReconJob *__restrict__ jobs, int n_bands, int queue_cap, int n_items, int skip_handled)
{
    uint16_t *queue = band_queue(threadIdx.x >> 5, queue_cap);
    build_div_tables<kBandWarps * 32>();
    uint32_t phase = 0;      /* uses of the staging barrier so far (nothing is staged here: the record data is read in place) */
#pragma unroll 1
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
    {
        const int job = item / n_bands;
        if (skip_handled == 1 && __ldg(&jobs[job].pad[0])) continue;
        if (skip_handled == 2 && !__ldg(&jobs[job].pad[1])) continue;
        __syncthreads();     /* the previous item is finished (view, tables, queues) */
        phase += band_item<false, kBandRows, kBandWarps>(jobs, job, item - job * n_bands, queue, queue_cap, 0u, 0u, 0u, phase) ? 1u : 0u;
    }
}

template <typename... KArgs, typename... Args>
int launch_overlapped(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <int kWarps, int kUnitsPerWarp, int kMinBlocks>
int launch_map(const ReconJob *d_jobs, int n_jobs, int units, cudaStream_t stream)
{
    const int per_cta = kWarps * kUnitsPerWarp;
    const int ctas_per_pic = (units + per_cta - 1) / per_cta;
    const long long grid = (long long)ctas_per_pic * n_jobs;
    if (grid > 0x7FFFFFFFll) return (int)cudaErrorInvalidConfiguration;
    return launch_overlapped(recon_map_kernel<kWarps, kUnitsPerWarp, kMinBlocks>, (unsigned)grid, kWarps * 32, RC_SMEM_TABLE_BYTES, stream, d_jobs, units,
                             ctas_per_pic);
}

template <int kMinBlocks>
int launch_record(const ReconJob *d_jobs, int n_jobs, uint32_t cta_base, uint32_t n_ctas, cudaStream_t stream)
{
    return launch_overlapped(recon_record_kernel<kMinBlocks>, n_ctas, kRecWarps * 32, kRecSmem, stream, d_jobs, n_jobs, cta_base);
}

int env_int(const char *name)
{
    const char *e = getenv(name);
    return e ? atoi(e) : 0;
}

}  // namespace

template <int kMinBlocks, bool kTile, int kRows>
int launch_band_plain(const ReconJob *d_jobs, long long items, int n_bands, int cap, int stage, int smem, cudaStream_t stream)
{
    if (smem > 48 * 1024)
    {   /* opt in (per device, so not cached) */
        const cudaError_t e = cudaFuncSetAttribute(recon_band_kernel<kMinBlocks, kTile, kRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    {   /* HVQM4_BAND_CARVEOUT=<percent of the SM's shared memory>: pins the L1 / shared memory split (experiments) */
        static const int carve = getenv("HVQM4_BAND_CARVEOUT") ? atoi(getenv("HVQM4_BAND_CARVEOUT")) : -1;
        if (carve >= 0) cudaFuncSetAttribute(recon_band_kernel<kMinBlocks, kTile, kRows>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    recon_band_kernel<kMinBlocks, kTile, kRows><<<(unsigned)items, band_warps(kTile, kRows) * 32, smem, stream>>>(d_jobs, n_bands, cap, stage);
    return (int)cudaGetLastError();
}

/* Bytes of shared memory for a band's staged record data (descriptors, vector rows, records; what does not fit is read in
   place).  Measured on B200, 1 024 streams (profiles/r02_band_stage_ab.txt): dense content 1.203 M frames/s with nothing
   staged, 1.154 M with 4 KB (descriptors + vectors), 1.171 M with 12 KB, 1.128 M with 20 KB, 1.017 M with 30 KB (everything):
   every staged kilobyte is taken from the L1 cache that the reference gathers live in (three CTAs per SM), and a chain that
   starts earlier only queues earlier on the same memory system; sparse content gains 5 % (2.07 -> 2.18 M) but runs the
   map + record kernels anyway (3.15 M).  So nothing is staged by default; HVQM4_BAND_STAGE=<bytes> (1 = eight bytes per
   block) turns it on for experiments. */
static inline int band_stage_bytes(int mcb_w, int rows)
{
    static const int env = getenv("HVQM4_BAND_STAGE") ? atoi(getenv("HVQM4_BAND_STAGE")) : 0;
    if (env <= 0) return 0;
    const int want = env > 1 ? env : rows * mcb_w * 6 * 8;
    return (want < 48 * 1024 ? want : 48 * 1024) & ~127;
}

/* bytes of the output tile of one band: `rows` macroblock rows of all three planes */
static inline int band_tile_bytes(int mcb_w, int rows) { return rows * 8 * (mcb_w * 8) * 3 / 2; }
static inline int band_plain_smem(int mcb_w, int rows, int warps = kBandWarps)
{
    return ((kBandQueueOff + band_queue_extra(warps, band_queue_entries(mcb_w, rows)) + 127) & ~127) + band_stage_bytes(mcb_w, rows);
}
static inline int band_tile_smem(int mcb_w, int rows) { return band_plain_smem(mcb_w, rows, band_warps(true, rows)) + band_tile_bytes(mcb_w, rows); }

/* tile = 0: blocks go straight into the picture, one CTA per band of 8 macroblock rows (n_bands of them per picture);
   tile = 8 / 4: the band (8 / 4 macroblock rows) is assembled in shared memory (two / four CTAs per SM) */
template <int kMinBlocks>
int launch_band(const ReconJob *d_jobs, int n_jobs, int n_bands, int mcb_w, cudaStream_t stream, int skip_handled = 0, int tile = 0, int mcb_h = 0)
{
    long long items = (long long)n_jobs * n_bands;
    if (items > 0x7FFFFFFFll) return (int)cudaErrorInvalidConfiguration;
    const int cap = band_queue_entries(mcb_w);
    const int smem = kBandQueueOff + band_queue_extra(kBandWarps, cap);
    if (skip_handled)
    {
        const long long grid = items > 148 * 2 ? 148 * 2 : items;
        if (smem > 48 * 1024)
        {
            const cudaError_t e = cudaFuncSetAttribute(recon_band_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
        }
        recon_band_walk_kernel<<<(unsigned)grid, kBandWarps * 32, smem, stream>>>(d_jobs, n_bands, cap, (int)items, skip_handled);
        return (int)cudaGetLastError();
    }
    if (tile == 8) return launch_band_plain<2, true, 8>(d_jobs, items, n_bands, cap, band_stage_bytes(mcb_w, 8), band_tile_smem(mcb_w, 8), stream);
    if (tile == 4)
    {
        const int nb4 = (mcb_h + 3) / 4;
        return launch_band_plain<4, true, 4>(d_jobs, (long long)n_jobs * nb4, nb4, band_queue_entries(mcb_w, 4), band_stage_bytes(mcb_w, 4),
                                             band_tile_smem(mcb_w, 4), stream);
    }
    return launch_band_plain<kMinBlocks, false, kBandRows>(d_jobs, items, n_bands, cap, band_stage_bytes(mcb_w, kBandRows), band_plain_smem(mcb_w, kBandRows), stream);
}

/* does the tile variant of the band kernel fit `ctas` times into an SM for pictures this wide -- and leave the gathers an L1
   worth having?  Up to the 196 KB carve-out (60 KB of L1) it is at least as fast as the plain kernel; in the 228 KB one
   (28 KB of L1) it loses a third (profiles/r02_queue16_ab.txt).  640-wide pictures: 2 x 80 KB, the 164 KB carve-out. */
static bool band_tile_fits(int mcb_w, int rows, int ctas)
{
    return (band_tile_smem(mcb_w, rows) + 1024) * ctas <= 196 * 1024;
}

extern "C" int hvqm4_sweep_supported(int mcb_w, int mcb_h);
extern "C" int hvqm4_sweep_launch(ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, cudaStream_t stream);
extern "C" int hvqm4_row_supported(int mcb_w, int mcb_h, const void *slab_base);
extern "C" int hvqm4_row_launch(ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const void *slab_base, cudaStream_t stream);
long long g_sweep_launches = 0, g_row_launches = 0;

/* The sweep kernel (sweep.cu) reconstructs the pictures its plan serves and marks the others; the band kernel
   behind it takes the marked ones (normally none: a few CTAs per SM walk the job list and leave). */
static int launch_sweep_then_band(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, cudaStream_t stream, int *launches);

/* 0 auto, 5 forced: does this step go to the sweep kernel?  One CTA per SM takes whole pictures, so it needs about
   as many pictures as SMs to fill the GPU. */
static bool use_sweep(int mode, int n_jobs, int mcb_w, int mcb_h)
{
    static const int env = getenv("HVQM4_SWEEP") ? atoi(getenv("HVQM4_SWEEP")) : -1;   /* 0: never in auto mode, 1: always */
    if (mode == 5 || (mode == 0 && env == 1)) return hvqm4_sweep_supported(mcb_w, mcb_h) != 0;
    (void)n_jobs;     /* measured (profiles/r02_*): the sweep kernel loses to the band kernel on every content; it stays selectable */
    return false;
}

/* The row kernel (row.cu) likewise, for pictures whose surfaces lie in a registered slab. */
static int launch_row_then_band(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const void *slab, cudaStream_t stream, int *launches);

/* 0 auto, 6 forced: does this step go to the row kernel?  One CTA per SM walks macroblock rows of many pictures. */
static bool use_row(int mode, int n_jobs, int mcb_w, int mcb_h, const void *slab)
{
    static const int env = getenv("HVQM4_ROW") ? atoi(getenv("HVQM4_ROW")) : -1;   /* 0: never in auto mode, 1: always */
    if (!slab) return false;
    if (mode == 6 || (mode == 0 && env == 1)) return hvqm4_row_supported(mcb_w, mcb_h, slab) != 0;
    (void)n_jobs;
    return false;
}

int g_band_mode = 0;
long long g_band_launches = 0;   /* steps issued as one fused band kernel (diagnostics) */   /* set by hvqm4_recon_set_mode: 0 auto, 1..4 force the band kernel, 5 force the sweep kernel, 6 force the row kernel, <0 force map+record kernels */

static int launch_map_cfg(int cfg, const ReconJob *d_jobs, int n_jobs, int units, bool record_heavy, cudaStream_t stream)
{
    switch (cfg)
    {
    case 1: return launch_map<4, 1, 1>(d_jobs, n_jobs, units, stream);
    case 2: return launch_map<4, 2, 8>(d_jobs, n_jobs, units, stream);
    case 3: return launch_map<8, 2, 4>(d_jobs, n_jobs, units, stream);
    case 4: return launch_map<8, 4, 4>(d_jobs, n_jobs, units, stream);
    case 5: return launch_map<4, 4, 10>(d_jobs, n_jobs, units, stream);
    default:
        /* few pictures: one segment per warp (latency); large batches: fewer, fatter CTAs -- four segments per
           warp with pipelined heads on sparse content (2.73 M vs 2.56 M frames/s realistic), two on record-heavy
           content, whose scattered gathers prefer more, shorter CTAs (950 k vs 672 k frames/s dense) */
        if ((long long)n_jobs * units < 148ll * 64) return launch_map<4, 1, 1>(d_jobs, n_jobs, units, stream);
        return record_heavy ? launch_map<4, 2, 8>(d_jobs, n_jobs, units, stream) : launch_map<4, 4, 10>(d_jobs, n_jobs, units, stream);
    }
}

static int launch_record_cfg(int cfg, const ReconJob *d_jobs, int n_jobs, uint32_t cta_base, uint32_t n_ctas, cudaStream_t stream)
{
    switch (cfg)
    {
    case 1: return launch_record<1>(d_jobs, n_jobs, cta_base, n_ctas, stream);
    case 2: return launch_record<2>(d_jobs, n_jobs, cta_base, n_ctas, stream);
    case 3: return launch_record<3>(d_jobs, n_jobs, cta_base, n_ctas, stream);
    default: return launch_record<4>(d_jobs, n_jobs, cta_base, n_ctas, stream);   /* 64 registers, 4 CTAs/SM: measured best */
    }
}

/*
 * h_rec_prefix[i] = record-kernel CTAs of pictures 0..i-1 (n_jobs + 1 entries, host memory).
 * The step can be issued in SUB-BATCHES of pictures (map kernel then record kernel per
 * sub-batch) so that the record kernel finds the sectors the map kernel has just written still
 * in L2; by default the sub-batch is the whole step (see below).
 */
extern "C" void hvqm4_recon_set_mode(int band_mode) { g_band_mode = band_mode; }
/* macroblock rows per record band the streams of a new batch / decoder should be created with: the sweep and row
   kernels need bands of one row, everything else is fastest with the band kernel's eight */
extern "C" int hvqm4_recon_band_rows(void)
{
    static const int env_sweep = getenv("HVQM4_SWEEP") ? atoi(getenv("HVQM4_SWEEP")) : -1, env_row = getenv("HVQM4_ROW") ? atoi(getenv("HVQM4_ROW")) : -1;
    static const int env_rows = getenv("HVQM4_BAND_ROWS") ? atoi(getenv("HVQM4_BAND_ROWS")) : 0;     /* 1, 4 or 8 pins it (experiments) */
    if (g_band_mode == 5 || g_band_mode == 6 || env_sweep == 1 || env_row == 1) return 1;
    if (env_rows == 1 || env_rows == 4 || env_rows == 8) return env_rows;
    return SYM_BAND_MCB_ROWS;     /* also for mode 7: tiles of 8 rows at two CTAs per SM beat tiles of 4 rows at four (1.07 M vs 0.79 M frames/s dense) */
}

/* the fused band kernel only (no host-side record prefix needed): used behind the GPU entropy stage, where it
   runs next to the parse kernels -- the smallest register footprint (end to end 96.4 k vs 94.4 k frames/s) */
extern "C" int hvqm4_recon_launch_band(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const void *slab, int band_rows, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    if (use_row(g_band_mode, n_jobs, mcb_w, mcb_h, slab)) return launch_row_then_band(d_jobs, n_jobs, mcb_w, mcb_h, slab, stream, nullptr);
    if (use_sweep(g_band_mode, n_jobs, mcb_w, mcb_h)) return launch_sweep_then_band(d_jobs, n_jobs, mcb_w, mcb_h, stream, nullptr);
    const int n_bands = (mcb_h + kBandRows - 1) / kBandRows;
    /* HVQM4_BAND_BEHIND_PARSER=7: the tile variant here too (experiments; its CTAs hold 30 k registers each) */
    static const int behind_env = getenv("HVQM4_BAND_BEHIND_PARSER") ? atoi(getenv("HVQM4_BAND_BEHIND_PARSER")) : 0;
    const bool tile = (behind_env == 7 || g_band_mode == 7) && band_rows > 4 && band_tile_fits(mcb_w, 8, 2);
    const int rc = tile ? launch_band<2>(d_jobs, n_jobs, n_bands, mcb_w, stream, 0, 8, mcb_h) : launch_band<4>(d_jobs, n_jobs, n_bands, mcb_w, stream);
    if (rc == 0) ++g_band_launches;
    return rc;
}
extern "C" long long hvqm4_recon_band_launches(void) { return g_band_launches; }
extern "C" long long hvqm4_recon_sweep_launches(void) { return g_sweep_launches; }
extern "C" long long hvqm4_recon_row_launches(void) { return g_row_launches; }

static int launch_row_then_band(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const void *slab, cudaStream_t stream, int *launches)
{
    int rc = hvqm4_row_launch(const_cast<ReconJob *>(d_jobs), n_jobs, mcb_w, mcb_h, slab, stream);
    if (rc != 0) return rc;
    ++g_row_launches;
    const int n_bands = (mcb_h + kBandRows - 1) / kBandRows;
    rc = launch_band<3>(d_jobs, n_jobs, n_bands, mcb_w, stream, 2);
    if (rc == 0 && launches) *launches += 2;
    return rc;
}

static int launch_sweep_then_band(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, cudaStream_t stream, int *launches)
{
    int rc = hvqm4_sweep_launch(const_cast<ReconJob *>(d_jobs), n_jobs, mcb_w, mcb_h, stream);
    if (rc != 0) return rc;
    ++g_sweep_launches;
    const int n_bands = (mcb_h + kBandRows - 1) / kBandRows;
    rc = launch_band<3>(d_jobs, n_jobs, n_bands, mcb_w, stream, 1);
    if (rc == 0 && launches) *launches += 2;
    return rc;
}

extern "C" int hvqm4_recon_launch(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const uint32_t *h_rec_prefix, const void *slab, int band_rows,
                                  cudaStream_t stream, int *launches)
{
    if (n_jobs <= 0) return 0;
    if (use_row(g_band_mode, n_jobs, mcb_w, mcb_h, slab)) return launch_row_then_band(d_jobs, n_jobs, mcb_w, mcb_h, slab, stream, launches);
    if (use_sweep(g_band_mode, n_jobs, mcb_w, mcb_h)) return launch_sweep_then_band(d_jobs, n_jobs, mcb_w, mcb_h, stream, launches);
    const int nseg = (mcb_w + SYM_SEG_MCBS - 1) / SYM_SEG_MCBS;
    const int units = nseg * mcb_h;
    /* HVQM4_MAP_CFG / HVQM4_REC_CFG / HVQM4_SUBBATCH pin a configuration (tuning experiments) */
    static const int map_cfg = env_int("HVQM4_MAP_CFG"), rec_cfg = env_int("HVQM4_REC_CFG"), sub_env = env_int("HVQM4_SUBBATCH");
    static const int band_env = env_int("HVQM4_BAND");   /* 1..4: force the band kernel (min blocks), -1: never */
    const int n_bands = (mcb_h + kBandRows - 1) / kBandRows;
    const int band_mode = g_band_mode != 0 ? g_band_mode : band_env;
    /* auto: the fused band kernel pays off when the pictures are record-heavy and there are enough bands
       (measured on dense 640x480 content: 16 pictures = 128 bands 444 k vs 424 k frames/s, 64 pictures 862 k vs
       729 k, 1024 pictures 1.14 M vs 0.94 M; 8 pictures 253 k vs 296 k).  Sparse content, where the map work
       dominates, keeps the kernel pair (2.73 M vs 2.3 M) */
    const bool record_heavy = h_rec_prefix[n_jobs] >= 8u * (uint32_t)n_jobs;
    if (band_mode > 0 || (band_mode == 0 && record_heavy && (long long)n_jobs * n_bands >= 128))
    {
        /* CTAs per SM (the kernel's register budget): 2, 3 or 4 pins one; otherwise by grid size.  Three 80-register
           CTAs per SM (no spills, a larger L1, fewer pictures in flight: -19 % DRAM bytes) finish a wave of 3 x 148
           bands in 0.72 of the time four 64-register CTAs need for 4 x 148 (8 192 bands: 1.21 M vs 1.17 M frames/s;
           two 92-register CTAs the same), so the choice is a matter of whole waves: 512 bands 750 k vs 892 k,
           768 bands 926 k vs 783 k, 1 024 bands 919 k vs 935 k, 1 536 bands 1.01 M vs 0.98 M. */
        const long long grid = (long long)n_jobs * n_bands;
        const long long waves4 = (grid + 4 * 148 - 1) / (4 * 148), waves3 = (grid + 3 * 148 - 1) / (3 * 148);
        const int per_sm = band_mode >= 2 && band_mode <= 4 ? band_mode : 1000 * waves4 <= 724 * waves3 ? 4 : 3;
        /* The band assembled in shared memory and written by bulk stores (two CTAs of twelve warps per SM) whenever it fits
           and no CTA count is pinned: no global store requests, DRAM bytes -29 %, and faster than the plain kernel at every
           grid size (profiles/r02_tile_grid_ab.txt, dense: 16 pictures 589 k vs 485 k frames/s, 32: 891 k vs 769 k,
           256: 1.15 M vs 1.08 M, 1 024: 1.235 M vs 1.204 M).  HVQM4_BAND_TILE=0: never; mode 7 forces it, modes 2..4 pin
           the plain kernel with that many CTAs per SM. */
        static const int tile_env = getenv("HVQM4_BAND_TILE") ? atoi(getenv("HVQM4_BAND_TILE")) : 1;
        const bool want_tile = band_mode == 7 || ((band_mode == 0 || band_mode == 1) && tile_env != 0);
        /* a CTA's rows are a multiple of the streams' record band: 8-row tiles (two CTAs per SM); mode 7 on streams with
           bands of 4 or 1 rows takes 4-row tiles (four CTAs per SM: measured slower, kept as a tested alternative) */
        const int tile = !want_tile ? 0 : (band_mode == 7 && band_rows <= 4 && band_tile_fits(mcb_w, 4, 4)) ? 4 : (band_tile_fits(mcb_w, 8, 2) ? 8 : 0);
        int rc;
        switch (tile ? 0 : per_sm)
        {
        case 0: rc = launch_band<2>(d_jobs, n_jobs, n_bands, mcb_w, stream, 0, tile, mcb_h); break;
        case 2: rc = launch_band<2>(d_jobs, n_jobs, n_bands, mcb_w, stream); break;
        case 3: rc = launch_band<3>(d_jobs, n_jobs, n_bands, mcb_w, stream); break;
        default: rc = launch_band<4>(d_jobs, n_jobs, n_bands, mcb_w, stream); break;
        }
        if (rc == 0 && launches) ++*launches;
        if (rc == 0) ++g_band_launches;
        return rc;
    }
    /* measured on B200 (1024 x 640x480 pictures): the whole step as ONE sub-batch is fastest -- grid
       tails cost more than the L2 misses of the record kernel's read-modify-writes; sub-batching
       stays available for experiments */
    const int sub = sub_env > 0 ? sub_env : n_jobs;
    for (int j0 = 0; j0 < n_jobs; j0 += sub)
    {
        const int nj = n_jobs - j0 < sub ? n_jobs - j0 : sub;
        int rc = launch_map_cfg(map_cfg, d_jobs + j0, nj, units, record_heavy, stream);
        if (rc != 0) return rc;
        if (launches) ++*launches;
        const uint32_t c0 = h_rec_prefix[j0], c1 = h_rec_prefix[j0 + nj];
        if (c1 > c0)
        {
            rc = launch_record_cfg(rec_cfg, d_jobs, n_jobs, c0, c1 - c0, stream);
            if (rc != 0) return rc;
            if (launches) ++*launches;
        }
    }
    return 0;
}
