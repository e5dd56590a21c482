/*
 * sweep.cu -- the SWEEP reconstruction kernel for sm_100a: one persistent CTA per SM walks a
 * picture top to bottom, band by band, out of shared memory (sweep_core.h has the design and all
 * of the logic; this file is the machinery around it).
 *
 * Roles inside the CTA (kConsumerWarps + 1 warps):
 *   producer warp   per band: one bulk asynchronous copy (cp.async.bulk, the TMA engine) per
 *                   symbol-buffer slice and per run of reference rows, all completing on the
 *                   band's `full` mbarrier; when they have landed, the band's task list
 *                   (sw_prep_band) and an arrive on `ready`; when the consumers are done with a
 *                   band, the bulk stores of its tile (shared memory -> picture) and the
 *                   retirement of its slot and of the ring rows no later band needs.  It runs as
 *                   far ahead as slots and ring capacity allow.
 *   consumer warps  per band: wait for `ready`, take tasks from the band's ticket counter until
 *                   none is left (map task = 32 blocks, record task = 32 records), make the tile
 *                   writes visible to the asynchronous proxy and arrive on `done`.  No CTA-wide
 *                   barrier inside a sweep: a warp that finishes a band early starts the next.
 *
 * A picture the plan cannot serve is marked in its job (pad[0] = 0) and reconstructed by the
 * band kernel, which the host launches behind this one (recon.cu).
 */
#define RC_PLAIN_LOADS 1
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "recon.h"
#include "sweep_core.h"
#include "recon_dev.cuh"

#ifndef HVQM4_SWEEP_WARPS
#define HVQM4_SWEEP_WARPS 16
#endif

namespace {

constexpr int kConsumerWarps = HVQM4_SWEEP_WARPS;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr long long kTimeoutCycles = 4000000000ll;    /* ~2 s: a wait that long is a bug; everybody leaves */

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
/* false: the CTA is aborting (some wait timed out) */
__device__ __forceinline__ bool mbar_wait(SweepCtl &ctl, unsigned long long *bar, uint32_t parity)
{
    if (mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    for (;;)
    {
        if (mbar_try(bar, parity)) return true;
        if (*reinterpret_cast<volatile uint32_t *>(&ctl.abort_flag)) return false;
        if (clock64() - t0 > kTimeoutCycles)
        {
            *reinterpret_cast<volatile uint32_t *>(&ctl.abort_flag) = 1;
            return false;
        }
    }
}

/* cp.async.bulk: global -> shared, completes on an mbarrier; shared -> global, completes in a bulk group */
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

/* per-role phase parities of the three barrier arrays, one bit per slot */
struct Phases
{
    uint32_t full, ready, done;
};

struct SweepArgs
{
    int mode, f;
    uint8_t *present;
    const uint8_t *ref;
    uint8_t *scratch;
};

/* ---- producer warp ---------------------------------------------------------------------------- */
__device__ bool producer(const SweepGeom &g, const ReconView &v, SweepCtl &ctl, const SweepArgs &a, Phases &ph)
{
    const int lane = threadIdx.x & 31;
    const int nb = g.n_bands;
    int ki = 0, kp = 0, ks = 0;            /* next band to request, to prepare, to retire */
    SwRingState st[2] = {{SW_EMPTY_HI}, {SW_EMPTY_HI}};
    long long idle_since = 0;
    while (ks < nb)
    {
        bool progressed = false;
        /* retire: store the tile of the oldest band once every consumer warp is done with it */
        if (ks < kp)
        {
            const int slot = ks % SW_NSLOTS;
            if (mbar_test(&ctl.bar_done[slot], (ph.done >> slot) & 1u))
            {
                ph.done ^= 1u << slot;
                if (lane == 0)
                {
                    const uint32_t slot_off = g.off_slot0 + (uint32_t)slot * g.slot_bytes;
                    const SweepSlotMeta &m = *reinterpret_cast<const SweepSlotMeta *>(rc_smem + slot_off + g.s_meta);
                    const uint32_t tile = smem_addr(rc_smem + slot_off + g.s_tile);
                    if (a.mode == SW_MODE_FUTURE)
                    {
                        if (m.n2) bulk_store(a.scratch + (size_t)m.side_off * SW_MCB_BYTES, tile, m.n2 * SW_MCB_BYTES);
                    }
                    else
                    {
                        const uint32_t wy = (uint32_t)g.width, wc = wy / 2;
                        const uint32_t ny = (uint32_t)m.rows * 8u * wy, nc = (uint32_t)m.rows * 4u * wc;
                        uint8_t *py = a.present + (size_t)ks * g.tile_y_bytes;
                        uint8_t *pu = a.present + (size_t)wy * g.height + (size_t)ks * g.tile_c_bytes;
                        uint8_t *pv = pu + (size_t)wc * (g.height / 2);
                        bulk_store(py, tile, ny);
                        bulk_store(pu, tile + g.tile_y_bytes, nc);
                        bulk_store(pv, tile + g.tile_y_bytes + g.tile_c_bytes, nc);
                    }
                    bulk_commit();
                }
                ++ks;
                progressed = true;
            }
        }
        /* request: the next band's symbol slices and the reference rows it adds */
        if (ki < nb && ki < ks + SW_NSLOTS && sw_ring_fits(ctl, a.f, ki, ks))
        {
            const int slot = ki % SW_NSLOTS;
            const uint32_t slot_off = g.off_slot0 + (uint32_t)slot * g.slot_bytes;
            SweepSlotMeta &m = *reinterpret_cast<SweepSlotMeta *>(rc_smem + slot_off + g.s_meta);
            if (ki >= SW_NSLOTS)
            {   /* the slot's previous tile must have left shared memory */
                if (lane == 0) bulk_wait_read();
                __syncwarp();
            }
            int r0[2], r1[2];
            sw_ring_new_rows(ctl, a.f, 0, ki, st[0], r0[0], r1[0]);
            sw_ring_new_rows(ctl, a.f, 1, ki, st[1], r0[1], r1[1]);
            SwCopy k = {SW_SRC_BLOB, 0, 0, 0};
            if (lane < SW_N_SYM_COPIES) k = sw_sym_copy(g, v, ctl, ki, lane, slot_off, m);
            else if (lane < SW_N_SYM_COPIES + 6)
            {
                const int j = lane - SW_N_SYM_COPIES, p = j >> 1, pc = p ? 1 : 0;
                k = sw_ring_copy(g, ctl, p, r0[pc], r1[pc], j & 1);
            }
            uint32_t total = k.bytes;
#pragma unroll
            for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, d);
            if (lane == 0) mbar_arrive_expect_tx(&ctl.bar_full[slot], total);
            __syncwarp();
            if (k.bytes)
            {
                const uint8_t *src = (k.src_kind == SW_SRC_BLOB ? v.blob : a.ref) + k.src_off;
                bulk_load(smem_addr(rc_smem + k.dst_off), src, k.bytes, &ctl.bar_full[slot]);
            }
            if (r1[0] > r0[0]) st[0].loaded_hi = r1[0];
            if (r1[1] > r0[1]) st[1].loaded_hi = r1[1];
            ++ki;
            progressed = true;
        }
        /* prepare: the task list of the oldest requested band whose data has landed */
        if (kp < ki)
        {
            const int slot = kp % SW_NSLOTS;
            if (mbar_test(&ctl.bar_full[slot], (ph.full >> slot) & 1u))
            {
                ph.full ^= 1u << slot;
                const uint32_t slot_off = g.off_slot0 + (uint32_t)slot * g.slot_bytes;
                SweepSlotMeta &m = *reinterpret_cast<SweepSlotMeta *>(rc_smem + slot_off + g.s_meta);
                sw_prep_band(g, v, ctl, kp, a.mode, a.f, slot_off, m, lane);
                __threadfence_block();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl.bar_ready[slot]);
                ++kp;
                progressed = true;
            }
        }
        if (progressed) idle_since = 0;
        else
        {
            if (*reinterpret_cast<volatile uint32_t *>(&ctl.abort_flag)) return false;
            const long long now = clock64();
            if (!idle_since) idle_since = now;
            else if (now - idle_since > kTimeoutCycles)
            {
                *reinterpret_cast<volatile uint32_t *>(&ctl.abort_flag) = 1;
                return false;
            }
            __nanosleep(40);
        }
    }
    /* the tiles have left shared memory and (scratch list) are visible to the loads of the next sweep */
    if (lane == 0)
    {
        bulk_wait_all();
        fence_async_all();
    }
    __syncwarp();
    return true;
}

/* ---- consumer warps --------------------------------------------------------------------------- */
__device__ bool consumer(const SweepGeom &g, const ReconView &v, SweepCtl &ctl, const SweepArgs &a, Phases &ph)
{
    const int lane = threadIdx.x & 31;
    const int nb = g.n_bands;
    for (int b = 0; b < nb; ++b)
    {
        const int slot = b % SW_NSLOTS;
        if (!mbar_wait(ctl, &ctl.bar_ready[slot], (ph.ready >> slot) & 1u)) return false;
        ph.ready ^= 1u << slot;
        const uint32_t slot_off = g.off_slot0 + (uint32_t)slot * g.slot_bytes;
        SweepSlotMeta &m = *reinterpret_cast<SweepSlotMeta *>(rc_smem + slot_off + g.s_meta);
        const SweepBand sb = {&g, &v, &ctl, &m, slot_off, b, a.mode, a.scratch};
        const uint32_t n_tasks = m.n_tasks;
        for (;;)
        {
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(&m.ticket, 1u);
            t = __shfl_sync(0xFFFFFFFFu, t, 0);
            if (t >= n_tasks) break;
            sw_run_task(sb, t, lane);
            __syncwarp();
        }
        fence_async_smem();      /* this lane's tile writes -> visible to the bulk store */
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl.bar_done[slot]);
    }
    return true;
}

/* one sweep; all threads.  false = abort */
__device__ bool run_sweep(const SweepGeom &g, const ReconView &v, SweepCtl &ctl, const SweepArgs &a, Phases &ph)
{
    if (threadIdx.x == 0) sw_plan_rings(g, ctl, a.f);
    __syncthreads();
    const bool ok = (threadIdx.x >> 5) == kConsumerWarps ? producer(g, v, ctl, a, ph) : consumer(g, v, ctl, a, ph);
    return ok;
}

__global__ void __launch_bounds__(kThreads, 1)
recon_sweep_kernel(ReconJob *__restrict__ jobs, int n_jobs, const __grid_constant__ SweepGeom g, uint8_t *scratch_base, size_t scratch_stride,
                   uint32_t *err)
{
    ReconView &vw = *reinterpret_cast<ReconView *>(rc_smem + RC_SMEM_VIEW_OFF);
    SweepCtl &ctl = *reinterpret_cast<SweepCtl *>(rc_smem + g.off_ctl);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0)
    {
        for (int s = 0; s < SW_NSLOTS; ++s)
        {
            mbar_init(&ctl.bar_full[s], 1);
            mbar_init(&ctl.bar_ready[s], 1);
            mbar_init(&ctl.bar_done[s], kConsumerWarps);
        }
        ctl.abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    build_div_tables<kThreads>();
    Phases ph = {0, 0, 0};
    uint8_t *scratch = scratch_base + (size_t)blockIdx.x * scratch_stride;
    bool ok = true;
    for (int job = blockIdx.x; job < n_jobs && ok; job += gridDim.x)
    {
        __syncthreads();        /* the previous picture is finished by every role */
        if (tid == 0)
        {
            load_view(vw, jobs[job]);
            ctl.unsupported = 0;
            ctl.n_future = 0;
        }
        __syncthreads();
        const ReconView &v = vw;
        if (!v.blob)
        {   /* the GPU entropy stage rejected this picture: nothing to reconstruct */
            if (tid == 0) jobs[job].pad[0] = 1;
            continue;
        }
        if (v.width != g.width || v.height != g.height || (int)v.n_bands != g.mcb_h)
        {
            if (tid == 0) jobs[job].pad[0] = 0;
            continue;
        }
        /* nest table (packed nest staged in the window area, which no sweep uses yet); band table */
        if (v.has_nest) nest_stage_begin<kThreads>(v, rc_smem + g.off_win);
        const int nr1 = g.mcb_h + 1;
        for (int i = tid; i < SYM_REC_CLASSES * nr1; i += kThreads)
        {
            const int cls = i / nr1, r = i - cls * nr1;
            ctl.bf[cls][r] = __ldg(v.bands + cls * nr1 + r);
        }
        if (v.has_nest) nest_stage_wait();
        __syncthreads();
        if (v.has_nest) nest_spread<kThreads>(rc_smem + g.off_win);
        for (int i = tid; i < SYM_REC_CLASSES * nr1; i += kThreads)
        {
            const int cls = i / nr1, r = i - cls * nr1;
            const uint32_t ci = ctl.bf[cls][r];
            ctl.rec_off[cls][r] = ci < v.n_chunks ? __ldg(v.chunks + 2 * ci) : __ldg(reinterpret_cast<const uint32_t *>(v.blob + offsetof(SymHeader, n_rec_words)));
        }
        /* which reference rows every band needs (warp per band, lanes over its macroblocks) */
        const bool is_bpic = __ldg(v.blob + offsetof(SymHeader, pic_type)) == SYM_PIC_B;
        for (int b = warp; b < g.n_bands; b += kConsumerWarps + 1)
        {
            int lo[4] = {SW_EMPTY_LO, SW_EMPTY_LO, SW_EMPTY_LO, SW_EMPTY_LO}, hi[4] = {SW_EMPTY_HI, SW_EMPTY_HI, SW_EMPTY_HI, SW_EMPTY_HI};   /* [f * 2 + plane class] */
            int n2 = 0, bad_any = 0;
            if (!v.is_ipic)
            {
                const int n_mcb = sw_band_rows(g, b) * g.mcb_w;
                for (int i = lane; i < n_mcb; i += 32)
                {
                    const int lmy = i / g.mcb_w, mx = i - lmy * g.mcb_w;
                    int ref, ext[4], bad;
                    sw_prescan_mcb(v, mx, b * g.h + lmy, ref, ext, bad);
                    if (ref == 2 && !is_bpic) bad = 1;     /* P picture (h4m:2058-2061): `future` is the picture itself */
                    bad_any |= bad;
                    if (!ref || bad) continue;
                    n2 += ref == 2;
                    const int q = (ref - 1) * 2;
                    lo[q] = min(lo[q], ext[0]); hi[q] = max(hi[q], ext[1]);
                    lo[q + 1] = min(lo[q + 1], ext[2]); hi[q + 1] = max(hi[q + 1], ext[3]);
                }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1)
            {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                {
                    lo[q] = min(lo[q], __shfl_xor_sync(0xFFFFFFFFu, lo[q], d));
                    hi[q] = max(hi[q], __shfl_xor_sync(0xFFFFFFFFu, hi[q], d));
                }
                n2 += __shfl_xor_sync(0xFFFFFFFFu, n2, d);
                bad_any |= __shfl_xor_sync(0xFFFFFFFFu, bad_any, d);
            }
            if (lane == 0)
            {
                for (int f = 0; f < 2; ++f)
                {
                    ctl.lo_y[f][b] = (int16_t)lo[2 * f]; ctl.hi_y[f][b] = (int16_t)hi[2 * f];
                    ctl.lo_c[f][b] = (int16_t)lo[2 * f + 1]; ctl.hi_c[f][b] = (int16_t)hi[2 * f + 1];
                }
                ctl.n2[b] = (uint16_t)n2;
                if (bad_any) ctl.unsupported = 1;
            }
        }
        __syncthreads();
        if (tid < 4) sw_plan_scan((tid & 1) ? ctl.lo_c[tid >> 1] : ctl.lo_y[tid >> 1], (tid & 1) ? ctl.hi_c[tid >> 1] : ctl.hi_y[tid >> 1], g.n_bands);
        else if (tid == 32)
        {
            uint32_t at = 0;
            for (int b = 0; b < g.n_bands; ++b)
            {
                ctl.side_off[b] = at;
                at += ctl.n2[b];
            }
            ctl.side_off[g.n_bands] = at;
            ctl.n_future = (int32_t)at;
        }
        else if (tid >= 64)
        {
            for (int b = tid - 64; b < g.n_bands; b += kThreads - 64)
                if (!sw_band_fits(g, ctl, b)) ctl.unsupported = 1;
        }
        __syncthreads();
        if (tid == 0 && !ctl.unsupported)
            if (!sw_plan_rings(g, ctl, 0) || (ctl.n_future > 0 && !sw_plan_rings(g, ctl, 1))) ctl.unsupported = 1;
        __syncthreads();
        if (ctl.unsupported)
        {
            if (tid == 0) jobs[job].pad[0] = 0;
            continue;
        }
        SweepArgs a;
        a.present = v.present;
        a.scratch = scratch;
        if (ctl.n_future > 0)
        {
            a.mode = SW_MODE_FUTURE; a.f = 1; a.ref = v.ref[1];
            ok = run_sweep(g, v, ctl, a, ph);
            if (ok)
            {
                __syncthreads();
                a.mode = SW_MODE_MERGE; a.f = 0; a.ref = v.ref[0];
                ok = run_sweep(g, v, ctl, a, ph);
            }
        }
        else
        {
            a.mode = SW_MODE_ALL; a.f = 0; a.ref = v.ref[0];
            ok = run_sweep(g, v, ctl, a, ph);
        }
        if (ok && tid == 0) jobs[job].pad[0] = 1;
    }
    if (!ok && lane == 0) atomicOr(err, 1u);
}

uint8_t *g_scratch;
size_t g_scratch_bytes;
uint32_t *g_err;
int g_sm_count, g_smem_optin;

}  // namespace

/* 1 if the sweep kernel serves pictures of this size on this device (shared-memory budget, width multiple of 32) */
extern "C" int hvqm4_sweep_supported(int mcb_w, int mcb_h)
{
    if (!g_sm_count)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    SweepGeom g;
    return g_sm_count > 0 && sw_make_geom(g, mcb_w * 8, mcb_h * 8, 1, (uint32_t)g_smem_optin);
}

/* Sweeps n_jobs pictures; every job's pad[0] says afterwards whether it was reconstructed (1) or is left to the
   band kernel (0).  Returns a cudaError_t. */
extern "C" int hvqm4_sweep_launch(ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    if (!hvqm4_sweep_supported(mcb_w, mcb_h)) return (int)cudaErrorInvalidConfiguration;
    static const int h_env = getenv("HVQM4_SWEEP_H") ? atoi(getenv("HVQM4_SWEEP_H")) : 1;
    SweepGeom g;
    if (!sw_make_geom(g, mcb_w * 8, mcb_h * 8, h_env, (uint32_t)g_smem_optin) && !sw_make_geom(g, mcb_w * 8, mcb_h * 8, 1, (uint32_t)g_smem_optin))
        return (int)cudaErrorInvalidConfiguration;
    const int grid = n_jobs < g_sm_count ? n_jobs : g_sm_count;
    const size_t stride = ((size_t)mcb_w * mcb_h * SW_MCB_BYTES + 255) & ~(size_t)255;
    if (g_scratch_bytes < stride * (size_t)g_sm_count)
    {
        if (g_scratch) cudaFree(g_scratch);
        g_scratch = nullptr;
        g_scratch_bytes = 0;
        cudaError_t e = cudaMalloc((void **)&g_scratch, stride * (size_t)g_sm_count);
        if (e != cudaSuccess) return (int)e;
        g_scratch_bytes = stride * (size_t)g_sm_count;
    }
    if (!g_err)
    {
        cudaError_t e = cudaMalloc((void **)&g_err, sizeof(uint32_t));
        if (e != cudaSuccess) return (int)e;
        cudaMemset(g_err, 0, sizeof(uint32_t));
    }
    cudaError_t e = cudaFuncSetAttribute(recon_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    recon_sweep_kernel<<<grid, kThreads, g.smem_bytes, stream>>>(d_jobs, n_jobs, g, g_scratch, stride, g_err);
    return (int)cudaGetLastError();
}

/* nonzero if a sweep CTA ever gave up waiting (diagnostics; synchronises the device) */
extern "C" int hvqm4_sweep_errors(void)
{
    uint32_t h = 0;
    if (g_err && cudaMemcpy(&h, g_err, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)h;
}
