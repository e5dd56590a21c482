/*
 * row.cu -- the ROW reconstruction kernel for sm_100a: one persistent CTA per SM walks a contiguous range of the
 * step's macroblock rows through a shared-memory pipeline (row_core.h has the design and all of the logic; this
 * file is the machinery around it: mbarriers, bulk copies, tensor copies, roles).
 *
 * Roles inside the CTA (2 + kWorkWarps warps):
 *   request   per row: bulk asynchronous copies (cp.async.bulk) of the row's symbol slices, completing on `sym`;
 *             when they have landed, ring space for the patches of its inter macroblocks, `go`.  Runs as far
 *             ahead as slots and ring allow.
 *   retire    per row: when the work warps are done with it (`done`), the bulk stores of its tile (shared memory
 *             -> picture); once the tile has left shared memory the slot and the ring space are free again.
 *   work      everything that grows with the row's content, as tasks from two ticket counters per row:
 *             FETCH tasks (one per 32 macroblocks, for this row and the rows ahead whose `go` has been given):
 *             sort the group's blocks into the row's class lists, give its inter macroblocks a patch place,
 *             announce the patch bytes on `ready` and issue two tensor copies (cp.async.bulk.tensor, luma box
 *             32 x 9 and chroma box 32 x 5 x 2) per inter macroblock -- `ready` completes when every fetch task
 *             has arrived and all patches have landed; then the row's WORK tasks (32 list entries or 32 records
 *             of one class) until none is left, the tile writes made visible to the asynchronous proxy, `done`.
 *             No CTA-wide barrier inside a picture: a warp that runs out of tasks in one row starts on the next,
 *             and a warp that waits for a row to become ready fetches for the rows ahead meanwhile.
 *
 * A picture the kernel cannot serve is marked in its job (pad[1] = 1) and reconstructed by the band kernel, which
 * the host launches behind this one (recon.cu).
 */
#define RC_PLAIN_LOADS 1
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "recon.h"
#include "row_core.h"
#include "recon_dev.cuh"

#ifndef HVQM4_ROW_WORK_WARPS
#define HVQM4_ROW_WORK_WARPS 22
#endif

namespace {

constexpr int kWorkWarps = HVQM4_ROW_WORK_WARPS;
constexpr int kThreads = (2 + kWorkWarps) * 32;
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
/* suspends the warp (no issue slots) until the phase completes or about `ns` nanoseconds have passed */
__device__ __forceinline__ bool mbar_try_for(unsigned long long *bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
/* false: the CTA is aborting (some wait timed out) */
__device__ __forceinline__ bool mbar_wait(RowCtl &ctl, unsigned long long *bar, uint32_t parity)
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
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* tensor copies: the box starts at a multiple of 16 bytes in x (any other x is an illegal instruction on B200) */
__device__ __forceinline__ void tma_box_3d(uint32_t dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_addr(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tma_box_4d(uint32_t dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z, int w)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_addr(bar)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

/* HVQM4_ROW_TRACE (compile-time, tuning builds only): cycles every role of CTA 0 spends waiting and working,
   summed per role into a global array that hvqm4_row_trace_dump() prints */
#ifdef HVQM4_ROW_TRACE
enum { TR_SEQ_LOOP, TR_SEQ_IDLE, TR_SEQ_RETIRE, TR_SEQ_RELEASE, TR_SEQ_REQUEST, TR_SEQ_RING_FULL, TR_FE_WAIT, TR_FE_CLASSIFY, TR_FE_ISSUE, TR_WK_WAIT, TR_WK_TASKS, TR_WK_TICKETS,
       TR_WK_NTASKS, TR_SETUP, TR_ROWS, TR_N };
__device__ unsigned long long g_trace[TR_N];
#define TR_T0() const long long tr_t0 = clock64()
#define TR_ADD(k) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_trace[k], (unsigned long long)(clock64() - tr_t0)); } while (0)
#define TR_COUNT(k, n) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_trace[k], (unsigned long long)(n)); } while (0)
#else
#define TR_T0() do { } while (0)
#define TR_ADD(k) do { } while (0)
#define TR_COUNT(k, n) do { } while (0)
#endif

struct RowArgs
{
    int r0, n;              /* first macroblock row of the segment inside the picture, number of rows */
    uint32_t seq0;          /* the CTA's row sequence number of the segment's first row (slot and phase of every barrier) */
    uint8_t *present;
};

__device__ __forceinline__ uint32_t slot_of(const RowGeom &g, uint32_t seq) { return seq % (uint32_t)g.n_slots; }
__device__ __forceinline__ uint32_t parity_of(const RowGeom &g, uint32_t seq) { return (seq / (uint32_t)g.n_slots) & 1u; }
__device__ __forceinline__ RowSlotMeta &meta_of(const RowGeom &g, uint32_t slot)
{
    return *reinterpret_cast<RowSlotMeta *>(rc_smem + g.off_slot0 + slot * g.slot_bytes + g.s_meta);
}

/* ---- request warp: symbol slices, ring space, `go` ------------------------------------------------- */
__device__ bool requester(const RowGeom &g, const ReconView &v, RowCtl &ctl, const RowArgs &a)
{
    const int lane = threadIdx.x & 31;
    const int n = a.n;
    int ki = 0, kc = 0;                    /* next row to request, to release to the fetch tasks */
    RwRing ring = {0};
    uint32_t tail = 0, freed = 0;          /* ring offset up to which patches are free again: the first `freed` rows are done */
    uint32_t *ring_ends = ctl.ring_ends;
    bool counted = false;
    uint32_t n_inter = 0;
    long long idle_since = 0;
    TR_T0();
    while (kc < n)
    {
        bool progressed = false;
        const uint32_t retired = ctl.retired;
        /* patch space is free as soon as the work warps are done with a row (its tile may still be on its way out) */
        while ((int)freed < kc)
        {
            const uint32_t dseq = a.seq0 + freed, dslot = slot_of(g, dseq);
            if (!mbar_test(&ctl.bar_done[dslot], parity_of(g, dseq))) break;
            tail = ring_ends[dslot];
            ++freed;
        }
        /* request: the next row's symbol slices, as soon as the row that used the slot before has been stored */
        if (ki < n && ki < (int)retired + g.n_slots)
        {
            TR_T0();
            const uint32_t seq = a.seq0 + (uint32_t)ki, slot = slot_of(g, seq);
            const uint32_t slot_off = g.off_slot0 + slot * g.slot_bytes;
            RowSlotMeta &m = meta_of(g, slot);
            SwCopy k = {SW_SRC_BLOB, 0, 0, 0};
            if (lane < RW_N_SYM_COPIES) k = rw_sym_copy(g, v, ctl, a.r0 + ki, lane, slot_off, m);
            uint32_t total = k.bytes;
#pragma unroll
            for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, d);
            if (lane == 0) mbar_arrive_expect_tx(&ctl.bar_sym[slot], total);
            __syncwarp();
            if (k.bytes) bulk_load(smem_addr(rc_smem + k.dst_off), v.blob + k.src_off, k.bytes, &ctl.bar_sym[slot]);
            ++ki;
            progressed = true;
            TR_ADD(TR_SEQ_REQUEST);
        }
        /* release: the oldest requested row whose symbol slices have landed gets ring space for its patches */
        if (kc < ki)
        {
            TR_T0();
            const uint32_t seq = a.seq0 + (uint32_t)kc, slot = slot_of(g, seq);
            if (counted || mbar_test(&ctl.bar_sym[slot], parity_of(g, seq)))
            {
                RowSlotMeta &m = meta_of(g, slot);
                if (!counted)
                {
                    n_inter = rw_count_inter(g, v, m, lane);
                    counted = true;
                }
                const uint32_t live = (uint32_t)kc - freed;
                if (!live) ring.head = tail = 0;
                const uint32_t pos = rw_ring_alloc(ring, g.ring_bytes, n_inter * RW_PATCH_BYTES, live, tail);
                if (pos != 0xFFFFFFFFu)
                {
                    if (lane == 0)
                    {
                        m.patch_base = g.off_ring + pos;
                        m.n_patch = n_inter;
                        m.ring_end = ring.head;
                        ring_ends[slot] = ring.head;
                        m.ticket = m.fticket = m.patch_count = 0;
                        m.n_list[0] = m.n_list[1] = m.n_list[2] = 0;
                    }
                    __syncwarp();
                    __threadfence_block();
                    if (lane == 0) mbar_arrive(&ctl.bar_go[slot]);
                    counted = false;
                    ++kc;
                    progressed = true;
                }
                else TR_COUNT(TR_SEQ_RING_FULL, 1);
            }
            TR_ADD(TR_SEQ_RELEASE);
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
            {
                TR_T0();
                __nanosleep(200);
                TR_ADD(TR_SEQ_IDLE);
            }
        }
    }
    TR_ADD(TR_SEQ_LOOP);
    TR_COUNT(TR_ROWS, n);
    return true;
}

/* ---- retire warp: finished tiles leave with three bulk stores of whole picture rows ------------------ */
__device__ bool retirer(const RowGeom &g, RowCtl &ctl, const RowArgs &a)
{
    const int lane = threadIdx.x & 31;
    for (int ks = 0; ks < a.n; ++ks)
    {
        const uint32_t seq = a.seq0 + (uint32_t)ks, slot = slot_of(g, seq);
        if (!mbar_wait(ctl, &ctl.bar_done[slot], parity_of(g, seq))) return false;
        TR_T0();
        if (lane == 0)
        {
            const uint32_t tile = smem_addr(rc_smem + g.off_slot0 + slot * g.slot_bytes + g.s_tile);
            const uint32_t wy = (uint32_t)g.width, wc = wy / 2;
            const int row = a.r0 + ks;
            uint8_t *py = a.present + (size_t)row * g.tile_y_bytes;
            uint8_t *pu = a.present + (size_t)wy * g.height + (size_t)row * g.tile_c_bytes;
            uint8_t *pv = pu + (size_t)wc * (g.height / 2);
            bulk_store(py, tile, g.tile_y_bytes);
            bulk_store(pu, tile + g.tile_y_bytes, g.tile_c_bytes);
            bulk_store(pv, tile + g.tile_y_bytes + g.tile_c_bytes, g.tile_c_bytes);
            bulk_commit();
            bulk_wait_read();                 /* the tile has left shared memory: the slot may be used again */
            __threadfence_block();
            ctl.retired = (uint32_t)ks + 1u;
        }
        __syncwarp();
        TR_ADD(TR_SEQ_RETIRE);
    }
    return true;
}

/* ---- work warps: fetch tasks of the rows ahead, then the row's own tasks --------------------------------- */
__device__ __forceinline__ void fetch_task(const RowGeom &g, const ReconView &v, RowCtl &ctl, int row, int grp, uint32_t slot, const CUtensorMap *map_y,
                                           const CUtensorMap *map_c)
{
    const int lane = threadIdx.x & 31;
    const uint32_t slot_off = g.off_slot0 + slot * g.slot_bytes;
    RowSlotMeta &m = meta_of(g, slot);
    RwBox box;
    uint32_t n;
    {
        TR_T0();
        n = rw_fetch_group(g, v, ctl, row, grp, slot_off, m, lane, box);
        __threadfence_block();
        __syncwarp();
        TR_ADD(TR_FE_CLASSIFY);
    }
    TR_T0();
    if (lane == 0) mbar_arrive_expect_tx(&ctl.bar_ready[slot], n * RW_PATCH_TX);
    __syncwarp();
    if (box.dst)
    {
        const uint32_t dst = smem_addr(rc_smem + box.dst);
        if (!box.chroma) tma_box_3d(dst, map_y, &ctl.bar_ready[slot], box.x, box.y, box.z);
        else tma_box_4d(dst, map_c, &ctl.bar_ready[slot], box.x, box.y, 0, box.z);
    }
    __syncwarp();
    TR_ADD(TR_FE_ISSUE);
}

__device__ bool worker(const RowGeom &g, const ReconView &v, RowCtl &ctl, const RowArgs &a, const CUtensorMap *map_y, const CUtensorMap *map_c)
{
    const int lane = threadIdx.x & 31;
    int f = 0;        /* first row that may still have fetch tasks to hand out */
    for (int k = 0; k < a.n; ++k)
    {
        const uint32_t seq = a.seq0 + (uint32_t)k, slot = slot_of(g, seq), slot_off = g.off_slot0 + slot * g.slot_bytes;
        RowSlotMeta &m = meta_of(g, slot);
        if (f < k) f = k;
        const int f_end = k + g.n_slots < a.n ? k + g.n_slots : a.n;
        /* Fetch tasks come first, whenever a warp looks for work: they are the head of the latency chain of a row
           (sort, tensor copies in flight, then its work tasks).  Rows [k, k + n_slots) can have been given `go`. */
        auto fetch_ahead = [&]() {
            while (f < f_end)
            {
                const uint32_t fseq = a.seq0 + (uint32_t)f, fslot = slot_of(g, fseq);
                if (!mbar_test(&ctl.bar_go[fslot], parity_of(g, fseq))) break;
                uint32_t t = 0;
                if (lane == 0) t = atomicAdd(&meta_of(g, fslot).fticket, 1u);
                t = __shfl_sync(0xFFFFFFFFu, t, 0);
                if (t >= (uint32_t)g.n_groups) { ++f; continue; }
                fetch_task(g, v, ctl, a.r0 + f, (int)t, fslot, map_y, map_c);
            }
        };
        {
            TR_T0();
            const long long t_start = clock64();
            for (;;)
            {
                fetch_ahead();
                /* asleep until the row is ready, with a look at the rows ahead now and then: a warp that polls takes
                   issue slots from the warps that work (22 polling warps: fetch tasks ran at 30 cycles per instruction) */
                if (mbar_try_for(&ctl.bar_ready[slot], parity_of(g, seq), 2000u)) break;
                if (*reinterpret_cast<volatile uint32_t *>(&ctl.abort_flag)) return false;
                if (clock64() - t_start > kTimeoutCycles)
                {
                    *reinterpret_cast<volatile uint32_t *>(&ctl.abort_flag) = 1;
                    return false;
                }
            }
            TR_ADD(TR_WK_WAIT);
        }
        const RowWork w = {&g, &v, &m, slot_off};
        uint32_t t_end[RW_TASK_CLASSES];
        rw_task_ends(m, t_end);
        const uint32_t n_tasks = t_end[RW_TASK_CLASSES - 1];
        for (;;)
        {
            fetch_ahead();
            uint32_t t = 0;
            {
                TR_T0();
                if (lane == 0) t = atomicAdd(&m.ticket, 1u);
                t = __shfl_sync(0xFFFFFFFFu, t, 0);
                TR_ADD(TR_WK_TICKETS);
            }
            if (t >= n_tasks) break;
            TR_T0();
            rw_run_task(w, t_end, t, lane);
            __syncwarp();
            TR_ADD(TR_WK_TASKS);
            TR_COUNT(TR_WK_NTASKS, 1);
        }
        fence_async_smem();      /* this lane's tile writes -> visible to the bulk store */
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl.bar_done[slot]);
    }
    return true;
}

__global__ void __launch_bounds__(kThreads, 1)
recon_row_kernel(ReconJob *__restrict__ jobs, int n_jobs, const __grid_constant__ RowGeom g, const __grid_constant__ CUtensorMap map_y,
                 const __grid_constant__ CUtensorMap map_c, const uint8_t *slab_base, unsigned long long slab_stride, int slab_count, uint32_t *err)
{
    ReconView &vw = *reinterpret_cast<ReconView *>(rc_smem + RC_SMEM_VIEW_OFF);
    RowCtl &ctl = *reinterpret_cast<RowCtl *>(rc_smem + g.off_ctl);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0)
    {
        for (int s = 0; s < g.n_slots; ++s)
        {
            mbar_init(&ctl.bar_sym[s], 1);
            mbar_init(&ctl.bar_go[s], 1);
            mbar_init(&ctl.bar_ready[s], (uint32_t)g.n_groups);      /* one arrival (with its patch bytes) per fetch task */
            mbar_init(&ctl.bar_done[s], kWorkWarps);
        }
        ctl.abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    build_div_tables<kThreads>();
    const long long total_rows = (long long)n_jobs * g.mcb_h;
    long long R = total_rows * blockIdx.x / gridDim.x;
    const long long R1 = total_rows * (blockIdx.x + 1) / gridDim.x;
    uint32_t seq = 0;
    bool ok = true;
    while (R < R1 && ok)
    {
        const int job = (int)(R / g.mcb_h), r0 = (int)(R - (long long)job * g.mcb_h);
        const int r1 = (int)((long long)g.mcb_h < r0 + (R1 - R) ? (long long)g.mcb_h : r0 + (R1 - R));
        R += r1 - r0;
        __syncthreads();        /* the previous segment is finished by every role */
        TR_T0();
        if (tid == 0)
        {
            load_view(vw, jobs[job]);
            ctl.unsupported = 0;
            ctl.retired = 0;
            ctl.pad = (int32_t)jobs[job].pad[1];        /* one read for the whole CTA: another CTA may be marking the picture */
        }
        __syncthreads();
        const ReconView &v = vw;
        if (!v.blob || ctl.pad) continue;               /* rejected by the GPU entropy stage / already left to the band kernel */
        if (v.width != g.width || v.height != g.height || (int)v.n_bands != g.mcb_h)
        {
            if (tid == 0) jobs[job].pad[1] = 1;
            continue;
        }
        /* band table rows r0 .. r1 of the three classes, then the record offsets they point at */
        const int nr = r1 - r0 + 1, nb1 = g.mcb_h + 1;
        for (int i = tid; i < SYM_REC_CLASSES * nr; i += kThreads)
        {
            const int cls = i / nr, r = r0 + i - cls * nr;
            ctl.bf[cls][r] = __ldg(v.bands + cls * nb1 + r);
        }
        if (v.has_nest) nest_stage_begin<kThreads>(v, rc_smem + g.off_ring);
        if (tid >= 32 && tid < 32 + 7) rw_sym_table(g, v, ctl, tid - 32);
        if (tid == 0)
        {
            ctl.is_bpic = __ldg(v.blob + offsetof(SymHeader, pic_type)) == SYM_PIC_B;
            for (int f = 0; f < 2; ++f)
            {
                long long z = -1;
                if (v.ref[f] && v.ref[f] >= slab_base)
                {
                    const unsigned long long d = (unsigned long long)(v.ref[f] - slab_base);
                    if (d % slab_stride == 0 && d / slab_stride < (unsigned long long)slab_count) z = (long long)(d / slab_stride);
                }
                ctl.z[f] = (int32_t)z;
            }
        }
        if (v.has_nest) nest_stage_wait();
        __syncthreads();
        if (v.has_nest) nest_spread<kThreads>(rc_smem + g.off_ring);
        for (int i = tid; i < SYM_REC_CLASSES * nr; i += kThreads)
        {
            const int cls = i / nr, r = r0 + i - cls * nr;
            const uint32_t ci = ctl.bf[cls][r];
            ctl.rec_off[cls][r] = ci < v.n_chunks ? __ldg(v.chunks + 2 * ci) : __ldg(reinterpret_cast<const uint32_t *>(v.blob + offsetof(SymHeader, n_rec_words)));
        }
        for (int r = r0 + tid; r < r1; r += kThreads)
            if (!rw_row_fits(ctl, r)) ctl.unsupported = 1;
        __syncthreads();
        if (tid == 0) TR_ADD(TR_SETUP);
        if (!ctl.unsupported)
        {
            const RowArgs a = {r0, r1 - r0, seq, v.present};
            if (warp == 0) ok = requester(g, v, ctl, a);
            else if (warp == 1) ok = retirer(g, ctl, a);
            else ok = worker(g, v, ctl, a, &map_y, &map_c);
            seq += (uint32_t)(r1 - r0);
            __syncthreads();
        }
        if (ctl.unsupported && tid == 0) jobs[job].pad[1] = 1;
    }
    if (warp == 1 && lane == 0) bulk_wait_all();      /* the last tiles have reached the picture */
    if (!ok && lane == 0) atomicOr(err, 1u);
}

/* ---- host side: registered surface slabs and their tensor maps ------------------------------------ */
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Slab
{
    const uint8_t *base;
    size_t stride;
    int count, width, height;
    CUtensorMap map_y, map_c;
};
constexpr int kMaxSlabs = 64;
Slab g_slabs[kMaxSlabs];
int g_n_slabs;
uint32_t *g_err;
int g_sm_count, g_smem_optin;

bool device_limits()
{
    if (!g_sm_count)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    return g_sm_count > 0;
}

}  // namespace

/* Announces a slab of `count` frame surfaces (planar Y|U|V of width x height at a pitch of `stride` bytes) whose
   surfaces the row kernel may read through tensor copies.  Returns 0, or nonzero if the slab cannot be described
   (the row kernel is then not used for it). */
extern "C" int hvqm4_row_register_slab(const void *base, size_t stride, int count, int width, int height)
{
    if (!base || count <= 0 || (width & 31) || (height & 7) || (stride & 15) || ((uintptr_t)base & 15) || g_n_slabs >= kMaxSlabs) return 1;
    static EncodeTiled encode = nullptr;
    if (!encode)
    {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) return 2;
    }
    Slab s;
    s.base = (const uint8_t *)base; s.stride = stride; s.count = count; s.width = width; s.height = height;
    {
        cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)height, (cuuint64_t)count}, strides[2] = {(cuuint64_t)width, (cuuint64_t)stride};
        cuuint32_t box[3] = {RW_BOX_W, 9, 1}, es[3] = {1, 1, 1};
        if (encode(&s.map_y, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)s.base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 3;
    }
    {
        const cuuint64_t cw = (cuuint64_t)width / 2, ch = (cuuint64_t)height / 2;
        cuuint64_t dims[4] = {cw, ch, 2, (cuuint64_t)count}, strides[3] = {cw, cw * ch, (cuuint64_t)stride};
        cuuint32_t box[4] = {RW_BOX_W, 5, 2, 1}, es[4] = {1, 1, 1, 1};
        if (encode(&s.map_c, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)(s.base + (size_t)width * height), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 3;
    }
    g_slabs[g_n_slabs++] = s;
    return 0;
}

extern "C" void hvqm4_row_unregister_slab(const void *base)
{
    for (int i = 0; i < g_n_slabs; ++i)
        if (g_slabs[i].base == base)
        {
            g_slabs[i] = g_slabs[--g_n_slabs];
            return;
        }
}

/* 1 if the row kernel serves pictures of this size out of this slab on this device */
extern "C" int hvqm4_row_supported(int mcb_w, int mcb_h, const void *slab_base)
{
    if (!slab_base || !device_limits()) return 0;
    const Slab *s = nullptr;
    for (int i = 0; i < g_n_slabs; ++i)
        if (g_slabs[i].base == slab_base) s = &g_slabs[i];
    if (!s || s->width != mcb_w * 8 || s->height != mcb_h * 8) return 0;
    RowGeom g;
    return rw_make_geom(g, mcb_w * 8, mcb_h * 8, (uint32_t)g_smem_optin);
}

/* Reconstructs n_jobs pictures; a job's pad[1] says afterwards whether it is left to the band kernel (1).
   Returns a cudaError_t. */
extern "C" int hvqm4_row_launch(ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const void *slab_base, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    if (!hvqm4_row_supported(mcb_w, mcb_h, slab_base)) return (int)cudaErrorInvalidConfiguration;
    const Slab *s = nullptr;
    for (int i = 0; i < g_n_slabs; ++i)
        if (g_slabs[i].base == slab_base) s = &g_slabs[i];
    RowGeom g;
    rw_make_geom(g, mcb_w * 8, mcb_h * 8, (uint32_t)g_smem_optin);
    const long long rows = (long long)n_jobs * mcb_h;
    /* a CTA should have a few rows to amortise the pipeline fill */
    long long grid = rows / 8 < g_sm_count ? rows / 8 : g_sm_count;
    if (grid < 1) grid = 1;
    if (!g_err)
    {
        cudaError_t e = cudaMalloc((void **)&g_err, sizeof(uint32_t));
        if (e != cudaSuccess) return (int)e;
        cudaMemset(g_err, 0, sizeof(uint32_t));
    }
    cudaError_t e = cudaFuncSetAttribute(recon_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    recon_row_kernel<<<(unsigned)grid, kThreads, g.smem_bytes, stream>>>(d_jobs, n_jobs, g, s->map_y, s->map_c, s->base, (unsigned long long)s->stride,
                                                                        s->count, g_err);
    return (int)cudaGetLastError();
}

#ifdef HVQM4_ROW_TRACE
extern "C" __attribute__((visibility("default"))) void hvqm4_row_trace_dump(void)
{
    unsigned long long h[TR_N];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_trace, sizeof h);
    static const char *names[TR_N] = {"seq loop", "seq idle", "seq retire", "seq release", "seq request", "seq ring-full events", "fetch wait", "fetch classify", "fetch issue",
                                      "work wait", "work tasks", "work tickets", "work n_tasks", "setup", "rows"};
    const double rows = (double)(h[TR_ROWS] ? h[TR_ROWS] : 1);
    for (int i = 0; i < TR_N; ++i) fprintf(stderr, "row trace: %-22s %14llu  %10.1f per row\n", names[i], h[i], (double)h[i] / rows);
    memset(h, 0, sizeof h);
    cudaMemcpyToSymbol(g_trace, h, sizeof h);
}
#endif

/* nonzero if a CTA of the row kernel ever gave up waiting (diagnostics; synchronises the device) */
extern "C" int hvqm4_row_errors(void)
{
    uint32_t h = 0;
    if (g_err && cudaMemcpy(&h, g_err, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)h;
}
