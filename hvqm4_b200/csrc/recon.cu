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
 *                      and AOT basis loops.  warp -> one chunk of <= 32 records of one
 *                      (class, band, length) group; lane -> one record.  All lanes of a warp
 *                      run the same class with the same number of bases, records are at
 *                      first + lane * length (no searching), and the per-picture nest is
 *                      expanded once per CTA into a shared-memory table in which a basis row
 *                      is a single 32-bit load.  Predicted-AOT blocks read back the
 *                      prediction the map kernel left in the picture (L2 resident).
 *
 * The block arithmetic itself lives in recon_core.h.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "recon.h"
#include "recon_core.h"

namespace {

/* ------------------------------------------------------------------------------------------
 * picture parameters: one ReconView per CTA in shared memory (constant-offset LDS, no live
 * registers across the block functions)
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ void load_view(ReconView &vw, const ReconJob &J)
{
    SymHeader h;
    const uint4 *src = reinterpret_cast<const uint4 *>(J.blob);
    uint4 *dst = reinterpret_cast<uint4 *>(&h);
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i] = __ldg(src + i);     /* header fields end at byte 88 */
    rc_make_view(vw, J.blob, h, nullptr, nullptr, nullptr, J.past, J.future);
    vw.present = J.present;
}

/* ------------------------------------------------------------------------------------------
 * map kernel
 * ------------------------------------------------------------------------------------------ */
template <int kWarps, int kUnitsPerWarp, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
recon_map_kernel(const ReconJob *__restrict__ jobs, int units_per_pic, int ctas_per_pic)
{
    ReconView &vw = *reinterpret_cast<ReconView *>(rc_smem);
    const int job = blockIdx.x / ctas_per_pic;
    const int cta = blockIdx.x - job * ctas_per_pic;
    if (threadIdx.x == 0) load_view(vw, jobs[job]);
    __syncthreads();
    const ReconView &v = vw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

#pragma unroll 1
    for (int it = 0; it < kUnitsPerWarp; ++it)
    {
        const int unit = (cta * kUnitsPerWarp + it) * kWarps + warp;
        if (unit >= units_per_pic) break;
        const int row = unit / v.nseg, mx0 = (unit - row * v.nseg) * SYM_SEG_MCBS;
        /* luma: two passes over the segment's two block rows (plane is a compile-time 0 here) */
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass)
        {
            const int bx = mx0 * 2 + lane, by = row * 2 + pass;
            if (bx >= v.mcb_w * 2) continue;
            const int pw = v.width, bstride = (pw >> 2) + 2;
            const uint32_t t = __ldg(v.blob + v.off_type[0] + (by + 1) * bstride + bx + 1);
            uint32_t rows[4];
            if (!rc_map_block(v, 0, bx, by, t, rows)) continue;
            uint8_t *dst = v.present + (by * 4) * pw + bx * 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + r * pw) = rows[r];
        }
        /* chroma: lanes 0-15 U, 16-31 V */
        {
            const int plane = 1 + (lane >> 4), bx = mx0 + (lane & 15), by = row;
            if (bx >= v.mcb_w) continue;
            const int pw = v.width >> 1, bstride = (pw >> 2) + 2;
            const uint32_t t = __ldg(v.blob + (plane == 1 ? v.off_type[1] : v.off_type[2]) + (by + 1) * bstride + bx + 1);
            uint32_t rows[4];
            if (!rc_map_block(v, plane, bx, by, t, rows)) continue;
            const int plane_off = v.width * v.height + (plane == 2 ? pw * (v.height >> 1) : 0);
            uint8_t *dst = v.present + plane_off + (by * 4) * pw + bx * 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + r * pw) = rows[r];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * record kernel
 * shared memory: [nest table | mcdiv | div | view] at the fixed offsets of recon_core.h, then
 * a scratch area used only while the nest table is being built
 * ------------------------------------------------------------------------------------------ */
constexpr int kRecWarps = 8;
constexpr int kRecSmem = RC_SMEM_TABLE_BYTES + SYM_NEST_H * 40 + 16;

template <int kMinBlocks>
__global__ void __launch_bounds__(kRecWarps * 32, kMinBlocks)
recon_record_kernel(const ReconJob *__restrict__ jobs, int n_jobs, uint32_t cta_base)
{
    uint32_t *s_nest_tab = reinterpret_cast<uint32_t *>(rc_smem + RC_SMEM_NEST_OFF);
    int32_t *s_mcdiv = reinterpret_cast<int32_t *>(rc_smem + RC_SMEM_MCDIV_OFF);
    int32_t *s_div = reinterpret_cast<int32_t *>(rc_smem + RC_SMEM_DIV_OFF);
    ReconView &vw = *reinterpret_cast<ReconView *>(rc_smem + RC_SMEM_VIEW_OFF);
    __shared__ uint32_t s_cta_in_pic;

    /* which picture does this CTA belong to: last job with rec_cta_begin <= blockIdx.x */
    if (threadIdx.x == 0)
    {
        const uint32_t gcta = blockIdx.x + cta_base;     /* rec_cta_begin is a prefix over the whole step */
        int lo = 0, hi = n_jobs - 1;
        while (lo < hi)
        {
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(&jobs[mid].rec_cta_begin) <= gcta) lo = mid;
            else hi = mid - 1;
        }
        s_cta_in_pic = gcta - __ldg(&jobs[lo].rec_cta_begin);
        load_view(vw, jobs[lo]);
    }
    for (int i = threadIdx.x; i < 256; i += kRecWarps * 32) s_mcdiv[i] = i ? 0x1000 / i : 0;   /* h4m:272 */
    if (threadIdx.x < 16) s_div[threadIdx.x] = threadIdx.x ? 0x1000 / (threadIdx.x * 16) * 16 : 0;   /* h4m:270 */
    __syncthreads();
    const ReconView &v = vw;
    const uint32_t chunk0 = s_cta_in_pic * HVQM4_REC_CHUNKS_PER_CTA;
    const uint32_t chunk_end = min(chunk0 + HVQM4_REC_CHUNKS_PER_CTA, v.n_chunks);

    if (chunk0 < v.n_chunks_nest)
    {
        /* stage the packed nest rows (35 B) at a 40-byte pitch, zero padded */
        uint8_t *packed = rc_smem + RC_SMEM_TABLE_BYTES;
        const uint8_t *src = v.blob + v.off_nest;
        for (int i = threadIdx.x; i < SYM_NEST_H * 40; i += kRecWarps * 32)
        {
            const int y = i / 40, x = i - y * 40;
            packed[i] = x < SYM_NEST_ROW_BYTES ? __ldg(src + y * SYM_NEST_ROW_BYTES + x) : (uint8_t)0;
        }
        __syncthreads();
        /* table entry (y, x) = nibbles x..x+7 of row y; (y, 2j) and (y, 2j+1) share bytes j..j+4 */
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(packed);
        for (int i = threadIdx.x; i < SYM_NEST_H * 32; i += kRecWarps * 32)
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
#pragma unroll 1
    for (uint32_t c = chunk0 + warp; c < chunk_end; c += kRecWarps)
    {
        const uint2 cd = __ldg(reinterpret_cast<const uint2 *>(v.chunks) + c);
        const uint32_t count = cd.y & 0xFF, len = ((cd.y >> 8) & 0xFF) + 1;
        const int cls = (int)((cd.y >> 16) & 0xFF);
        if ((uint32_t)lane >= count) continue;
        const uint32_t *rec = v.rec + cd.x + lane * len;
        uint32_t t;
        int plane, bx, by;
        rc_record_coords(__ldg(rec), t, plane, bx, by);
        const int pw = plane ? v.width >> 1 : v.width;
        const int plane_off = plane == 0 ? 0 : plane == 1 ? v.width * v.height : v.width * v.height + (v.width >> 1) * (v.height >> 1);
        uint8_t *dst = v.present + plane_off + (by * 4) * pw + bx * 4;
        uint32_t rows[4];
        if (cls == SYM_REC_INTER)
        {   /* the prediction written by the map kernel (plain loads: written by the previous launch) */
#pragma unroll
            for (int r = 0; r < 4; ++r) rows[r] = *reinterpret_cast<const uint32_t *>(dst + r * pw);
        }
        rc_record_block(v, cls, len, rec, rows);
#pragma unroll
        for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + r * pw) = rows[r];
    }
}

template <int kWarps, int kUnitsPerWarp, int kMinBlocks>
int launch_map(const ReconJob *d_jobs, int n_jobs, int units, cudaStream_t stream)
{
    const int per_cta = kWarps * kUnitsPerWarp;
    const int ctas_per_pic = (units + per_cta - 1) / per_cta;
    const long long grid = (long long)ctas_per_pic * n_jobs;
    if (grid > 0x7FFFFFFFll) return (int)cudaErrorInvalidConfiguration;
    recon_map_kernel<kWarps, kUnitsPerWarp, kMinBlocks><<<(unsigned)grid, kWarps * 32, sizeof(ReconView), stream>>>(d_jobs, units, ctas_per_pic);
    return (int)cudaGetLastError();
}

template <int kMinBlocks>
int launch_record(const ReconJob *d_jobs, int n_jobs, uint32_t cta_base, uint32_t n_ctas, cudaStream_t stream)
{
    recon_record_kernel<kMinBlocks><<<n_ctas, kRecWarps * 32, kRecSmem, stream>>>(d_jobs, n_jobs, cta_base);
    return (int)cudaGetLastError();
}

int env_int(const char *name)
{
    const char *e = getenv(name);
    return e ? atoi(e) : 0;
}

}  // namespace

static int launch_map_cfg(int cfg, const ReconJob *d_jobs, int n_jobs, int units, cudaStream_t stream)
{
    switch (cfg)
    {
    case 1: return launch_map<4, 1, 1>(d_jobs, n_jobs, units, stream);
    case 2: return launch_map<4, 2, 8>(d_jobs, n_jobs, units, stream);
    case 3: return launch_map<8, 2, 4>(d_jobs, n_jobs, units, stream);
    case 4: return launch_map<8, 4, 4>(d_jobs, n_jobs, units, stream);
    case 5: return launch_map<4, 4, 10>(d_jobs, n_jobs, units, stream);
    default:
        /* few pictures: one segment per warp (latency); large batches: fewer, fatter CTAs */
        return (long long)n_jobs * units >= 148ll * 64 ? launch_map<4, 2, 8>(d_jobs, n_jobs, units, stream)
                                                       : launch_map<4, 1, 1>(d_jobs, n_jobs, units, stream);
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
extern "C" int hvqm4_recon_launch(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const uint32_t *h_rec_prefix,
                                  cudaStream_t stream, int *launches)
{
    if (n_jobs <= 0) return 0;
    const int nseg = (mcb_w + SYM_SEG_MCBS - 1) / SYM_SEG_MCBS;
    const int units = nseg * mcb_h;
    /* HVQM4_MAP_CFG / HVQM4_REC_CFG / HVQM4_SUBBATCH pin a configuration (tuning experiments) */
    static const int map_cfg = env_int("HVQM4_MAP_CFG"), rec_cfg = env_int("HVQM4_REC_CFG"), sub_env = env_int("HVQM4_SUBBATCH");
    /* measured on B200 (1024 x 640x480 pictures): the whole step as ONE sub-batch is fastest -- grid
       tails cost more than the L2 misses of the record kernel's read-modify-writes; sub-batching
       stays available for experiments */
    const int sub = sub_env > 0 ? sub_env : n_jobs;
    for (int j0 = 0; j0 < n_jobs; j0 += sub)
    {
        const int nj = n_jobs - j0 < sub ? n_jobs - j0 : sub;
        int rc = launch_map_cfg(map_cfg, d_jobs + j0, nj, units, stream);
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
