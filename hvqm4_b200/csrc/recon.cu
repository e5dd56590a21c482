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
 * Mapping:   picture -> ctas_per_pic CTAs of kWarps warps;
 *            warp    -> one SEGMENT: 16 macroblocks of one macroblock row, handled in
 *                       three passes (upper luma block row: 32 blocks, lower luma block
 *                       row: 32 blocks, chroma: 16 U + 16 V blocks);
 *            lane    -> one 4x4 block per pass.
 * A lane finds its variable-length side data with a warp prefix sum over
 * sym_side_words(type) added to the segment's base offset (symbuf.h).  Stores are one
 * 32-bit word per lane and row: a warp writes 128 contiguous bytes of a luma row
 * (64 + 64 for the two chroma planes) per store instruction.
 *
 * The block arithmetic itself lives in recon_core.h.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "recon.h"
#include "recon_core.h"

namespace {

constexpr int kWarps = 4;

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

__global__ void __launch_bounds__(kWarps * 32)
recon_pictures_kernel(const ReconJob *__restrict__ jobs, int units_per_pic, int ctas_per_pic)
{
    __shared__ int32_t s_div[16];
    __shared__ int32_t s_mcdiv[512];
    __shared__ __align__(16) uint8_t s_nest[(SYM_NEST_BYTES + 15) & ~15];

    const int job = blockIdx.x / ctas_per_pic;
    const int cta = blockIdx.x - job * ctas_per_pic;
    const ReconJob J = jobs[job];
    const SymHeader *__restrict__ hp = reinterpret_cast<const SymHeader *>(J.blob);
    SymHeader h;
    {
        /* 128-byte header, uniform across the CTA: eight 16-byte loads through the read-only path */
        const uint4 *src = reinterpret_cast<const uint4 *>(hp);
        uint4 *dst = reinterpret_cast<uint4 *>(&h);
#pragma unroll
        for (int i = 0; i < 5; ++i) dst[i] = __ldg(src + i);   /* fields end at byte 76 */
    }

    /* constants of h4m:262-273 */
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_mcdiv[i] = i ? 0x1000 / i : 0;
    if (threadIdx.x < 16) s_div[threadIdx.x] = threadIdx.x ? 0x1000 / (threadIdx.x * 16) * 16 : 0;
    if (h.has_nest)
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(J.blob + h.off_nest);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_nest);
        for (int i = threadIdx.x; i < (SYM_NEST_BYTES + 3) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = cta * kWarps + warp;
    if (unit >= units_per_pic) return;
    const int nseg = h.nseg;
    const int row = unit / nseg, sg = unit - row * nseg;

    ReconView v;
    rc_make_view(v, J.blob, h, s_nest, s_div, s_mcdiv, J.past, J.future);
    const uint32_t *__restrict__ side = reinterpret_cast<const uint32_t *>(J.blob + h.off_side);
    uint32_t word = __ldg(reinterpret_cast<const uint32_t *>(J.blob + h.off_seg) + unit);

    const int W = h.width, H = h.height;
    const int mx0 = sg * SYM_SEG_MCBS;

#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass)
    {
        int plane, bx, by;
        bool valid;
        if (pass < 2)
        {
            plane = 0;
            bx = mx0 * 2 + lane;
            by = row * 2 + pass;
            valid = bx < h.mcb_w * 2;
        }
        else
        {
            plane = 1 + (lane >> 4);
            bx = mx0 + (lane & 15);
            by = row;
            valid = bx < h.mcb_w;
        }
        const int pw = plane ? W >> 1 : W;
        const int bstride = (pw >> 2) + 2;
        uint32_t t = 0;
        if (valid) t = __ldg(J.blob + h.off_type[plane] + (by + 1) * bstride + bx + 1);
        const uint32_t nwords = valid ? sym_side_words(t, v.is_ipic) : 0u;
        uint32_t total;
        const uint32_t mine = word + warp_excl_scan(nwords, total);
        word += total;
        if (!valid) continue;

        uint32_t rows[4];
        rc_block(v, plane, bx, by, t, side + mine, rows);

        uint8_t *dst = J.present + (plane == 0 ? 0 : plane == 1 ? W * H : W * H + (W >> 1) * (H >> 1)) + (by * 4) * pw + bx * 4;
#pragma unroll
        for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + r * pw) = rows[r];
    }
}

}  // namespace

extern "C" int hvqm4_recon_launch(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    const int nseg = (mcb_w + SYM_SEG_MCBS - 1) / SYM_SEG_MCBS;
    const int units = nseg * mcb_h;
    const int ctas_per_pic = (units + kWarps - 1) / kWarps;
    const long long grid = (long long)ctas_per_pic * n_jobs;
    if (grid > 0x7FFFFFFFll) return (int)cudaErrorInvalidConfiguration;
    recon_pictures_kernel<<<(unsigned)grid, kWarps * 32, 0, stream>>>(d_jobs, units, ctas_per_pic);
    return (int)cudaGetLastError();
}

extern "C" int hvqm4_recon_launch_count(void) { return 1; }
