/*
 * recon_dev.cuh -- device helpers shared by the reconstruction kernels (recon.cu, sweep.cu): the picture's
 * ReconView in shared memory, the division tables (h4m:262-273) and the nest lookup table.
 */
#ifndef HVQM4_RECON_DEV_CUH
#define HVQM4_RECON_DEV_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "recon.h"
#include "recon_core.h"

namespace {

/* ------------------------------------------------------------------------------------------
 * picture parameters: one ReconView per CTA in shared memory (constant-offset LDS, no live
 * registers across the block functions)
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ void load_view(ReconView &vw, const ReconJob &J)
{
    if (!J.blob)
    {   /* the GPU entropy stage rejected this picture: nothing to reconstruct */
        vw.blob = nullptr;
        vw.mcb_h = 0; vw.mcb_w = 0; vw.nseg = 1; vw.n_bands = 0; vw.n_chunks = 0; vw.n_chunks_nest = 0; vw.has_nest = 0;
        return;
    }
    SymHeader h;
    const uint4 *src = reinterpret_cast<const uint4 *>(J.blob);
    uint4 *dst = reinterpret_cast<uint4 *>(&h);
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i] = __ldg(src + i);     /* header fields end at byte 88 */
    rc_make_view(vw, J.blob, h, nullptr, nullptr, nullptr, J.past, J.future);
    vw.present = J.present;
}

/* h4m:262-273 into shared memory */
template <int kThreads>
__device__ __forceinline__ void build_div_tables()
{
    int32_t *s_mcdiv = reinterpret_cast<int32_t *>(rc_smem + RC_SMEM_MCDIV_OFF);
    int32_t *s_div = reinterpret_cast<int32_t *>(rc_smem + RC_SMEM_DIV_OFF);
    for (int i = threadIdx.x; i < 256; i += kThreads) s_mcdiv[i] = i ? 0x1000 / i : 0;
    if (threadIdx.x < 16) s_div[threadIdx.x] = threadIdx.x ? 0x1000 / (threadIdx.x * 16) : 0;   /* divTable / 16 (recon_core.h) */
}

/* The packed nest of the CTA's picture (35-byte pitch, 16-byte aligned in the blob) is staged in shared memory
   by asynchronous copies -- no registers, nothing waits for it -- and expanded into the lookup tables later:
   nest_stage_begin() ... (other work) ... nest_stage_wait(); __syncthreads(); nest_spread(); __syncthreads().
   The caller provides 38 * 40 bytes of scratch. */
template <int kThreads>
__device__ __forceinline__ void nest_stage_begin(const ReconView &v, uint8_t *packed)
{
    const uint8_t *src = v.blob + v.off_nest;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(packed);
    for (int i = threadIdx.x; i < (SYM_NEST_BYTES + 15) / 16; i += kThreads)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 16 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void nest_stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int kThreads>
__device__ __forceinline__ void nest_spread(uint8_t *packed, bool portrait = false)
{
    /* landscape: 38 rows of 35 packed bytes -> 68 entries per row; portrait: 70 rows of 19 bytes -> 36 entries per row */
    const int rows = portrait ? SYM_NEST_W : SYM_NEST_H, row_bytes = portrait ? SYM_NEST_H / 2 : SYM_NEST_ROW_BYTES;
    const int half = portrait ? RC_NEST_PITCH_PORTRAIT / 2 : RC_NEST_PITCH / 2, pitch = 2 * half;
    uint32_t *s_nest_tab = reinterpret_cast<uint32_t *>(rc_smem + RC_SMEM_NEST_OFF);
    uint32_t *stage = reinterpret_cast<uint32_t *>(packed);
    /* nibbles x..x+7 of row y, spread into table entries x = 2j and 2j+1 (samples x..x+3, one per byte, times 16);
       both share bytes j..j+4 of the row.  Entries near the end of a row run into the next row: those nibbles lie
       beyond column 69, which no descriptor reaches (offset <= 63, largest pattern + 6). */
    const uint32_t *pw = stage;
    for (int i = threadIdx.x; i < rows * half; i += kThreads)
    {
        const int y = i / half, j = i - y * half;
        const int b = y * row_bytes + j, w = b >> 2, sh = (b & 3) * 8;
        const uint32_t w0 = pw[w], w1 = pw[w + 1], w2 = sh ? pw[w + 2] : 0u;
        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        const uint32_t odd = (lo >> 4) | (hi << 28);
        s_nest_tab[y * pitch + 2 * j] = rc_nest_spread_step1(lo);
        s_nest_tab[y * pitch + 2 * j + 1] = rc_nest_spread_step1(odd);
    }
}

}  // namespace

#endif
