/*
 * row_core.h -- everything of the ROW kernel (row.cu) that is not a CUDA synchronisation primitive or a TMA
 * instruction: geometry and shared-memory layout, what is copied for a macroblock row and where, the patch ring,
 * the classification of a row into class-sorted block lists, and the per-lane work of every task.
 * tests/emul/row_emul.cpp drives the very same functions serially on the CPU (test infrastructure), so the slot
 * and ring arithmetic, the lists and the task decomposition are checked against the oracle without a GPU.
 *
 * The row pipeline ("h4m:N" = /root/reference/h4m_audio_decode.c line N):
 *   A persistent CTA (one per SM) owns a contiguous range of macroblock rows of the step's pictures and walks
 *   it top to bottom -- the reference's raster walk (h4m:1487-1518 / 1922-1967) one macroblock row at a time.
 *   Up to n_slots rows are in flight; everything a row needs is staged in shared memory:
 *     - its slice of the symbol buffer (type / DC map rows with the neighbour rows of the weighted fill,
 *       h4m:299-383; the vector row; the chunk descriptors of its records) by bulk asynchronous copies;
 *     - one reference PATCH per inter macroblock (h4m:1327-1355): the 9 luma rows and 2 x 5 chroma rows its
 *       half-sample prediction reads, fetched by two tensor copies (cp.async.bulk.tensor: boxes of 32 x 9 luma
 *       bytes and 32 x 5 x 2 chroma bytes; a box must start at a multiple of 16 bytes in x -- measured,
 *       tools/ubench/tma_dbg2.cu -- so the box is the aligned 32 bytes around the 9 / 5 the filter needs) into a
 *       ring of patches: two requests per macroblock instead of ~60 per-lane word loads, served by the TMA unit
 *       instead of the L1 load pipeline, from either reference of a B picture (h4m:2018-2056) in one pass;
 *     - the row's output tile (8 luma rows, 4 rows of U and V), assembled in shared memory and written by
 *       three bulk stores of whole picture rows.
 *   The blocks of a row are CLASS-SORTED into lists (weighted fill, flat fill, motion compensation) next to the
 *   host-grouped records (symbuf.h); a task is 32 entries of one list, so all lanes of a warp run one code path.
 *   Only the window gathers of predicted-AOT bases (h4m:734-773) still read the reference luma with per-lane
 *   loads: a 70 x 38 window per macroblock is 30 x the bytes its bases touch (tools/ubench/tma_box.cu).
 *
 *   Roles (row.cu): a REQUEST warp asks for symbol slices and allocates ring space, a RETIRE warp stores finished
 *   tiles; everything else is a task taken from a row's ticket counters by the WORK warps: FETCH tasks (one per 16
 *   macroblocks of a row: sort their blocks into the lists, issue their patch copies) for the rows ahead, then the
 *   row's work tasks.  A first version with one warp per role measured 13 000 cycles of SERIAL work per row (one
 *   warp issues an instruction every 5-8 cycles; a tensor copy costs its issuing warp ~86 cycles): nothing that
 *   grows with the row's content may sit in a single warp.
 *
 *   A macroblock whose prediction is not inside its plane (the reference addresses linearly, such vectors wrap
 *   around picture rows, h4m:1344,1897; a tensor copy would fill with zeros instead) marks the picture for the
 *   band kernel, which the host launches behind this one for marked pictures only.
 */
#ifndef HVQM4_ROW_CORE_H
#define HVQM4_ROW_CORE_H

#include "sweep_core.h"   /* SW_SMEM, SwCopy, lane helpers (ballot / sum with one lane on the CPU) */

#define RW_MAX_ROWS 128          /* macroblock rows per picture the per-picture tables hold */
#define RW_MAX_SLOTS 6           /* rows in flight */
#define RW_MIN_SLOTS 3
#define RW_DESC_CAP 32           /* chunk descriptors per (class, row) a slot holds */
#define RW_BOX_W 32              /* bytes per patch row */
#define RW_PATCH_Y (9 * RW_BOX_W)
#define RW_PATCH_C (5 * RW_BOX_W)            /* one chroma plane; U then V */
#define RW_PATCH_TX (RW_PATCH_Y + 2 * RW_PATCH_C)      /* 608 bytes land per patch */
#define RW_PATCH_C_OFF 384                   /* the destination of a tensor copy is 128-byte aligned: luma at 0, chroma at 384 */
#define RW_PATCH_BYTES 768                   /* ring space per patch */
#define RW_NO_PATCH 0xFFFFu

enum { RW_LIST_W = 0, RW_LIST_FLAT = 1, RW_LIST_MC = 2, RW_LISTS = 3 };
/* task classes in the order they are handed out (long tasks first) */
enum { RW_T_INTER = 0, RW_T_INTRA = 1, RW_T_MC = 2, RW_T_W = 3, RW_T_RAW = 4, RW_T_FLAT = 5, RW_TASK_CLASSES = 6 };

struct RowGeom
{
    int width, height, mcb_w, mcb_h;
    int bw[3], stride[3];        /* blocks per block row, bordered map pitch (plane 0, 1, 2) */
    int n_slots;
    int n_groups;                /* fetch tasks per row: groups of 16 macroblocks */
    uint32_t off_ctl, off_slot0, slot_bytes, off_ring, ring_bytes, smem_bytes;
    /* inside a slot */
    uint32_t s_meta, s_type[3], s_dc[3], s_mv, s_desc, s_poff, s_list[RW_LISTS], s_tile;
    uint32_t list_cap;           /* entries per list: every block of the row */
    uint32_t tile_y_bytes, tile_c_bytes;
};

/* per-picture tables and state of the CTA */
struct RowCtl
{
    unsigned long long bar_sym[RW_MAX_SLOTS], bar_go[RW_MAX_SLOTS], bar_ready[RW_MAX_SLOTS], bar_done[RW_MAX_SLOTS];
    uint32_t abort_flag;                    /* a wait timed out: everybody leaves (debug guard) */
    int32_t unsupported;                    /* the picture is left to the band kernel */
    int32_t z[2];                           /* surface index of past / future inside the registered slab; -1 = not in it */
    int32_t is_bpic, pad;                   /* pad: the job was already marked when the segment started */
    volatile uint32_t retired;              /* rows of the segment the retire warp has stored */
    uint32_t ring_ends[RW_MAX_SLOTS];       /* ring head after the allocation of the row in slot s (request warp) */
    uint32_t sym_base[7], sym_step[7], sym_span[7], sym_region[7];   /* symbol slices 0-6 of row r: blob bytes [base + r * step, + span) */
    uint32_t bf[SYM_REC_CLASSES][RW_MAX_ROWS + 1];        /* first chunk of (class, macroblock row) */
    uint32_t rec_off[SYM_REC_CLASSES][RW_MAX_ROWS + 1];   /* first record word of (class, macroblock row) */
};

/* what the pipeline publishes for a row (lives in the row's slot) */
struct RowSlotMeta
{
    uint32_t ticket;                         /* next work task (work warps fetch-and-add) */
    uint32_t fticket;                        /* next fetch task */
    uint32_t patch_count;                    /* patches handed out so far (fetch tasks fetch-and-add) */
    uint32_t n_rec[SYM_REC_CLASSES], rec_lo[SYM_REC_CLASSES];
    uint32_t n_list[RW_LISTS];               /* list lengths (fetch tasks fetch-and-add their share) */
    uint32_t p_type[3], p_dc[3];             /* shared-memory offset of the bordered map row above the row's first block row */
    uint32_t p_mv, p_desc[SYM_REC_CLASSES];
    uint32_t patch_base;                     /* shared-memory offset of the row's patches */
    uint32_t n_patch;                        /* ring space of the row in patches: its inter macroblocks */
    uint32_t ring_end;                       /* ring head after this row's allocation (the tail once it retires) */
    int32_t row;                             /* macroblock row inside the picture */
};

RC_HD int rw_make_geom(RowGeom &g, int width, int height, uint32_t smem_limit)
{
    if (width <= 0 || height <= 0 || (width & 31) || (height & 7) || width > 2048) return 0;
    if (width < height) return 0;      /* portrait pictures (38 x 70 nest, swapped window origin): the other kernels */
    g.width = width; g.height = height; g.mcb_w = width / 8; g.mcb_h = height / 8;
    g.n_groups = (g.mcb_w + 15) / 16;
    if (g.mcb_h > RW_MAX_ROWS) return 0;
    for (int p = 0; p < 3; ++p)
    {
        g.bw[p] = (width >> (p ? 1 : 0)) / 4;
        g.stride[p] = g.bw[p] + 2;
    }
    uint32_t at = RC_SMEM_TABLE_BYTES;
    g.off_ctl = at = sw_align16(at);
    at += (uint32_t)sizeof(RowCtl);
    g.off_slot0 = at = (at + 127u) & ~127u;
    uint32_t s = 0;
    g.s_meta = s; s += sw_align16((uint32_t)sizeof(RowSlotMeta));
    for (int p = 0; p < 3; ++p) { g.s_type[p] = s; s += sw_align16((uint32_t)(((p ? 1 : 2) + 2) * g.stride[p]) + 32); }
    for (int p = 0; p < 3; ++p) { g.s_dc[p] = s; s += sw_align16((uint32_t)(((p ? 1 : 2) + 2) * g.stride[p]) + 32); }
    g.s_mv = s; s += sw_align16((uint32_t)(g.mcb_w * 4) + 32);
    g.s_desc = s; s += SYM_REC_CLASSES * (RW_DESC_CAP * 8 + 16);
    g.s_poff = s; s += sw_align16((uint32_t)(g.mcb_w * 2));
    g.list_cap = (uint32_t)(g.mcb_w * 6);
    for (int l = 0; l < RW_LISTS; ++l) { g.s_list[l] = s; s += sw_align16(g.list_cap * 2); }
    g.tile_y_bytes = (uint32_t)(8 * width);
    g.tile_c_bytes = (uint32_t)(4 * (width / 2));
    g.s_tile = s = (s + 127u) & ~127u;
    s += g.tile_y_bytes + 2 * g.tile_c_bytes;
    g.slot_bytes = (s + 127u) & ~127u;
    /* the ring must hold the patches of one full row with room to spare; the rest goes to slots */
    const uint32_t ring_min = (uint32_t)(g.mcb_w * RW_PATCH_BYTES) * 3u / 2u;
    if (at + RW_MIN_SLOTS * g.slot_bytes + ring_min > smem_limit) return 0;
    int n = (int)((smem_limit - at - ring_min) / g.slot_bytes);
    g.n_slots = n > RW_MAX_SLOTS ? RW_MAX_SLOTS : n;
    at += (uint32_t)g.n_slots * g.slot_bytes;
    g.off_ring = at;
    g.ring_bytes = (smem_limit - at) & ~127u;
    g.smem_bytes = at + g.ring_bytes;
    return 1;
}

/* ---- symbol slices of a row ------------------------------------------------------------------ */
#define RW_N_SYM_COPIES 10

/* per segment: where slices 0-6 (0-2 type rows, 3-5 DC rows, 6 vectors) of row r lie in the blob and in a slot */
RC_HD void rw_sym_table(const RowGeom &g, const ReconView &v, RowCtl &c, int id)
{
    uint32_t base = 0, step = 0, span = 0, region = 0;
    if (id < 6)
    {
        const int p = id % 3;
        /* bordered row by0 = block row by0 - 1: the rows of the macroblock row plus one above and one below */
        base = id < 3 ? rc_pick3(v.off_type, p) : rc_pick3(v.off_dc, p);
        step = (uint32_t)((p ? 1 : 2) * g.stride[p]);
        span = (uint32_t)(((p ? 1 : 2) + 2) * g.stride[p]);
        region = id < 3 ? g.s_type[p] : g.s_dc[p];
    }
    else
    {
        region = g.s_mv;
        if (!v.is_ipic)
        {
            base = v.off_mv;
            step = span = (uint32_t)(g.mcb_w * 4);
        }
    }
    c.sym_base[id] = base; c.sym_step[id] = step; c.sym_span[id] = span; c.sym_region[id] = region;
}

/* copy `id` (0-6 as above, 7-9 chunk descriptors) of macroblock row `row`; also fills the matching pointer of the
   slot's meta (lane id does both, the fields are disjoint) */
RC_HD SwCopy rw_sym_copy(const RowGeom &g, const ReconView &v, const RowCtl &c, int row, int id, uint32_t slot_off, RowSlotMeta &m)
{
    SwCopy k = {SW_SRC_BLOB, 0, 0, 0};
    uint32_t lo, hi, region;
    if (id < 7)
    {
        lo = c.sym_base[id] + (uint32_t)row * c.sym_step[id];
        hi = lo + c.sym_span[id];
        region = c.sym_region[id];
    }
    else
    {
        const int cls = id - 7;
        lo = (uint32_t)((const uint8_t *)v.chunks - v.blob) + c.bf[cls][row] * 8u;
        hi = lo + (c.bf[cls][row + 1] - c.bf[cls][row]) * 8u;
        region = g.s_desc + (uint32_t)cls * (RW_DESC_CAP * 8 + 16);
    }
    const uint32_t a0 = lo & ~15u, a1 = sw_align16(hi);
    const uint32_t ptr = slot_off + region + (lo - a0);
    if (id < 3) m.p_type[id] = ptr;
    else if (id < 6) m.p_dc[id - 3] = ptr;
    else if (id == 6) m.p_mv = ptr;
    else m.p_desc[id - 7] = ptr;
    if (hi <= lo) return k;
    k.src_off = a0;
    k.dst_off = slot_off + region;
    k.bytes = a1 - a0;
    return k;
}

/* the chunk descriptors of a row must fit its slot */
RC_HD int rw_row_fits(const RowCtl &c, int row)
{
    for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
        if (c.bf[cls][row + 1] - c.bf[cls][row] > RW_DESC_CAP) return 0;
    return 1;
}

/* ---- patches ----------------------------------------------------------------------------------- */

/* Box coordinates of a macroblock's patch: luma box (xl, yl), chroma box (xc, yc), both in samples of their plane.
   ref = 0 none (intra or poisoned), 1 past, 2 future; bad != 0: the prediction is not inside its plane. */
RC_HD void rw_patch_box(const ReconView &v, uint32_t tag, uint32_t mvw, int &ref, int &xl, int &yl, int &xc, int &yc, int &bad)
{
    ref = (int)((tag >> 5) & 3);
    bad = 0;
    xl = yl = xc = yc = 0;
    if (!ref) return;
    if (ref == 3) { bad = 1; ref = 0; return; }
    const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
    if (rx == -32768) { ref = 0; return; }                      /* poisoned: painted grey, reads nothing */
    {   /* luma 8x8 (+1 with a half step), h4m:1327-1355 */
        const int px = rx >> 1, py = ry >> 1, hx = rx & 1, hy = ry & 1;
        if (px < 0 || py < 0 || px + 8 + hx > v.width || py + 8 + hy > v.height) bad = 1;
        xl = px & ~15;
        yl = py;
    }
    {   /* chroma 4x4 per plane; 1.3 reuses the luma phase (h4m:1337-1343) */
        const int pxc = rx >> 1, pyc = ry >> 1;
        const int hx = (v.version15 ? pxc : rx) & 1, hy = (v.version15 ? pyc : ry) & 1;
        const int cx = pxc >> 1, cy = pyc >> 1;
        if (cx < 0 || cy < 0 || cx + 4 + hx > (v.width >> 1) || cy + 4 + hy > (v.height >> 1)) bad = 1;
        xc = cx & ~15;
        yc = cy;
    }
    if (bad) ref = 0;
}

/* the same with the rules of the picture applied: what a fetch task really asks for */
RC_HD void rw_patch_of(const ReconView &v, const RowCtl &c, uint32_t tag, uint32_t mvw, int &ref, int &xl, int &yl, int &xc, int &yc, int &bad)
{
    rw_patch_box(v, tag, mvw, ref, xl, yl, xc, yc, bad);
    if (ref == 2 && !c.is_bpic) { bad = 1; ref = 0; }      /* P picture (h4m:2058-2061): `future` is the picture itself */
    if (ref && c.z[ref - 1] < 0) { bad = 1; ref = 0; }     /* the reference is not a surface of the registered slab */
}

/* Request warp, once a row's symbol slice has landed: the ring space the row may need = its inter macroblocks (a
   poisoned one wastes its place).  Warp-collective on the GPU. */
RC_HD uint32_t rw_count_inter(const RowGeom &g, const ReconView &v, const RowSlotMeta &m, int lane)
{
    if (v.is_ipic) return 0;
    const uint8_t *tags = SW_SMEM(m.p_type[0]) + g.stride[0] + 1;
    uint32_t n = 0;
    for (int mx0 = 0; mx0 < g.mcb_w; mx0 += SW_LANES)
    {
        const int mx = mx0 + lane;
        n += (uint32_t)SW_POPC(sw_ballot(mx < g.mcb_w && (tags[2 * mx] & 0x60) != 0));
    }
    return n;
}

/* Ring allocation.  The request warp owns the head, the retire warp publishes the tail and the number of retired
   rows; rows retire in order.  live = rows that hold an allocation, tail = offset up to which the ring is free. */
struct RwRing
{
    uint32_t head;
};

/* returns the offset inside the ring, or 0xFFFFFFFF if the bytes are not free yet */
RC_HD uint32_t rw_ring_alloc(RwRing &r, uint32_t cap, uint32_t bytes, uint32_t live, uint32_t tail)
{
    bytes = (bytes + 127u) & ~127u;
    uint32_t pos;
    if (!live)
    {   /* nothing in use: start over at the front */
        if (bytes > cap) return 0xFFFFFFFFu;
        pos = 0;
    }
    else if (r.head >= tail)
    {   /* in use: [tail, head); free: [head, cap) and [0, tail) */
        if (r.head + bytes <= cap) pos = r.head;
        else if (bytes < tail) pos = 0;
        else return 0xFFFFFFFFu;
    }
    else
    {   /* wrapped: free is [head, tail) */
        if (r.head + bytes < tail) pos = r.head;
        else return 0xFFFFFFFFu;
    }
    r.head = pos + bytes;
    return pos;
}

/* ---- fetch task: one group of 32 macroblocks of a row ------------------------------------------------ */

/* list entry: [8:0] block x, [9] local block row (luma), [11:10] plane */
RC_HD uint32_t rw_entry(int p, int lrow, int bx) { return (uint32_t)bx | (uint32_t)lrow << 9 | (uint32_t)p << 10; }

/* which list a block with type byte t belongs to; -1: none (it has a record, or nothing to do) */
RC_HD int rw_block_list(uint32_t t, bool ipic)
{
    const uint32_t nib = ipic ? t : (t & 0xF);
    if (!ipic && (t & 0x60)) return ((t & 0x10) || nib == 0) ? RW_LIST_MC : -1;
    return nib == 0 ? RW_LIST_W : nib == 8 ? RW_LIST_FLAT : -1;
}

/* what a lane of a fetch task asks the TMA unit for */
struct RwBox
{
    uint32_t dst;         /* shared-memory offset (128-byte aligned), 0 = nothing to fetch */
    int chroma;           /* 0: luma box 32 x 9 at (x, y, z); 1: chroma box 32 x 5 x 2 at (x, y, 0, z) */
    int x, y, z;
};

#define RW_GROUP_MCBS 16     /* macroblocks per fetch task */

/* Group `grp` (16 macroblocks) of row `row`: sorts the blocks of its macroblocks into the row's lists (ranges
   reserved with one fetch-and-add per list), gives every macroblock that needs one a patch place (another
   fetch-and-add), and returns in `box` what the calling lane must fetch: lanes 0-15 the luma box of macroblock
   16 * grp + lane, lanes 16-31 the chroma box of macroblock 16 * grp + lane - 16.  Group 0 also sums the record
   counts of the row.  Warp-collective on the GPU; on the CPU (tests/emul) the caller passes lane = 0..31 in turn
   and the reservations are plain additions, which gives a different but equally valid order inside the lists.
   Returns the number of patches of the group (GPU: the same in every lane; CPU: 1 or 0 for this lane). */
RC_HD uint32_t rw_fetch_group(const RowGeom &g, const ReconView &v, RowCtl &c, int row, int grp, uint32_t slot_off, RowSlotMeta &m, int lane, RwBox &box)
{
    const bool ipic = v.is_ipic != 0;
    uint16_t *lists[RW_LISTS];
    for (int l = 0; l < RW_LISTS; ++l) lists[l] = reinterpret_cast<uint16_t *>(SW_SMEM(slot_off + g.s_list[l]));
    /* the group's blocks as three rounds of 32: luma upper row, luma lower row, 16 U + 16 V */
    uint32_t t3[3], e3[3];
    int l3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
    {
        const int p = r < 2 ? 0 : 1 + (lane >> 4), lrow = r < 2 ? r : 0;
        const int bx = p ? RW_GROUP_MCBS * grp + (lane & 15) : 2 * RW_GROUP_MCBS * grp + lane;
        const bool in_row = bx < g.bw[p];
        t3[r] = in_row ? *(SW_SMEM(rc_pick3(m.p_type, p)) + (lrow + 1) * g.stride[p] + 1 + bx) : 6u;   /* past the row end: nothing to do */
        l3[r] = rw_block_list(t3[r], ipic);
        e3[r] = rw_entry(p, lrow, bx);
    }
#if defined(__CUDA_ARCH__)
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t bal[RW_LISTS][3], cnt[RW_LISTS] = {0, 0, 0};
#pragma unroll
    for (int l = 0; l < RW_LISTS; ++l)
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
            bal[l][r] = __ballot_sync(0xFFFFFFFFu, l3[r] == l);
            cnt[l] += (uint32_t)__popc(bal[l][r]);
        }
    uint32_t base = 0;
    if (lane < RW_LISTS) base = atomicAdd(&m.n_list[lane], lane == 0 ? cnt[0] : lane == 1 ? cnt[1] : cnt[2]);
#pragma unroll
    for (int l = 0; l < RW_LISTS; ++l)
    {
        uint32_t at = __shfl_sync(0xFFFFFFFFu, base, l);
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
            if (l3[r] == l) lists[l][at + (uint32_t)__popc(bal[l][r] & lt)] = (uint16_t)e3[r];
            at += (uint32_t)__popc(bal[l][r]);
        }
    }
#else
    for (int r = 0; r < 3; ++r)
        if (l3[r] >= 0) lists[l3[r]][m.n_list[l3[r]]++] = (uint16_t)e3[r];
#endif
    /* patches: both half warps look at the same 16 macroblocks */
    const int mx = RW_GROUP_MCBS * grp + (lane & 15), half = lane >> 4;
    int ref = 0, bad = 0, xl = 0, yl = 0, xc = 0, yc = 0;
    box.dst = 0;
    box.chroma = half;
    box.x = box.y = box.z = 0;
    if (mx < g.mcb_w && !ipic)
    {
        const uint32_t tag = *(SW_SMEM(m.p_type[0]) + g.stride[0] + 1 + 2 * mx);
        const uint32_t mvw = reinterpret_cast<const uint32_t *>(SW_SMEM(m.p_mv))[mx];
        rw_patch_of(v, c, tag, mvw, ref, xl, yl, xc, yc, bad);
    }
    uint16_t *poff = reinterpret_cast<uint16_t *>(SW_SMEM(slot_off + g.s_poff));
    uint32_t n, idx;
#if defined(__CUDA_ARCH__)
    const uint32_t balp = __ballot_sync(0xFFFFFFFFu, ref != 0) & 0xFFFFu;
    const bool bad_any = __any_sync(0xFFFFFFFFu, bad != 0);
    n = (uint32_t)__popc(balp);
    idx = 0;
    if (lane == 0 && n) idx = atomicAdd(&m.patch_count, n);
    idx = __shfl_sync(0xFFFFFFFFu, idx, 0) + (uint32_t)__popc(balp & ((1u << (lane & 15)) - 1u));
#else
    const bool bad_any = bad != 0;
    n = 0;
    if (half == 0)
    {
        n = ref != 0;
        idx = m.patch_count;
        m.patch_count += n;
    }
    else
        idx = ref ? poff[mx] / (RW_PATCH_BYTES / RW_BOX_W) : 0u;      /* the place lane - 16 has just given it */
#endif
    if (mx < g.mcb_w && half == 0) poff[mx] = ref ? (uint16_t)(idx * (RW_PATCH_BYTES / RW_BOX_W)) : (uint16_t)RW_NO_PATCH;
    if (ref)
    {
        box.dst = m.patch_base + idx * RW_PATCH_BYTES + (half ? (uint32_t)RW_PATCH_C_OFF : 0u);
        box.x = half ? xc : xl;
        box.y = half ? yc : yl;
        box.z = c.z[ref - 1];
    }
    if (bad_any && lane == 0) c.unsupported = 1;
    if (grp == 0)
    {   /* record counts of the row: the chunk descriptors of a class hold them */
        for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
        {
            const uint32_t nd = c.bf[cls][row + 1] - c.bf[cls][row];
            const uint2 *d = reinterpret_cast<const uint2 *>(SW_SMEM(m.p_desc[cls]));
            uint32_t sum = 0;
#if defined(__CUDA_ARCH__)
            for (uint32_t j = (uint32_t)lane; j < nd; j += 32) sum += d[j].y & 0xFF;
            sum = sw_lane_sum(sum);
#else
            if (lane == 0)
                for (uint32_t j = 0; j < nd; ++j) sum += d[j].y & 0xFF;
#endif
            if (lane == 0)
            {
                m.n_rec[cls] = sum;
                m.rec_lo[cls] = c.rec_off[cls][row];
            }
        }
        if (lane == 0) m.row = row;
    }
    return n;
}

/* task boundaries of a row once all of its fetch tasks are done: tasks [t_end[k-1], t_end[k]) belong to class k */
RC_HD void rw_task_ends(const RowSlotMeta &m, uint32_t t_end[RW_TASK_CLASSES])
{
    uint32_t t = 0;
    t += (m.n_rec[SYM_REC_INTER] + 31u) / 32u; t_end[RW_T_INTER] = t;
    t += (m.n_rec[SYM_REC_INTRA] + 31u) / 32u; t_end[RW_T_INTRA] = t;
    t += (m.n_list[RW_LIST_MC] + 31u) / 32u;   t_end[RW_T_MC] = t;
    t += (m.n_list[RW_LIST_W] + 31u) / 32u;    t_end[RW_T_W] = t;
    t += (m.n_rec[SYM_REC_RAW] + 31u) / 32u;   t_end[RW_T_RAW] = t;
    t += (m.n_list[RW_LIST_FLAT] + 31u) / 32u; t_end[RW_T_FLAT] = t;
}

/* ---- work: one list entry or one record per lane ------------------------------------------------ */

/* rows of a patch: aligned word k of patch row r */
struct RwPatchRows
{
    uint32_t off;     /* shared-memory offset of row 0's first aligned word */
    RC_HDM uint32_t ld(int r, int k) const { return *reinterpret_cast<const uint32_t *>(SW_SMEM(off + (uint32_t)r * RW_BOX_W + 4u * (uint32_t)k)); }
};

/* window gathers of predicted-AOT bases read the reference luma in global memory (read-only during the launch) */
struct RwGlobalRows
{
    const uint8_t *base;
    int pitch;
    RC_HDM uint32_t ld(int r, int k) const
    {
#if defined(__CUDA_ARCH__)
        return __ldg(reinterpret_cast<const uint32_t *>(base + r * pitch + 4 * k));
#else
        return *reinterpret_cast<const uint32_t *>(base + r * pitch + 4 * k);
#endif
    }
};
struct RwGlobalWindow
{
    const uint8_t *origin;
    int width;
    RC_HDM RwGlobalRows rows(int ox, int oy, int ys, uint32_t &a) const
    {
        const uint8_t *p = origin + oy * width + ox;
        a = (uint32_t)((uintptr_t)p & 3);
        return RwGlobalRows{p - a, ys * width};
    }
};

struct RowWork
{
    const RowGeom *g;
    const ReconView *v;
    const RowSlotMeta *m;
    uint32_t slot_off;
};

/* half-sample prediction of block (bx, local block row lrow) of plane p from the macroblock's patch (h4m:1327-1355) */
RC_HD void rw_predict_block(const RowWork &w, int p, int bx, int lrow, uint32_t mvw, uint32_t poff, uint32_t rows[4])
{
    const ReconView &v = *w.v;
    const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
    const bool none = poff == RW_NO_PATCH;           /* poisoned (SYM_ERR_MV_RANGE: grey), or left to the band kernel */
    const int sh = p ? 1 : 0;
    const int px = rx >> sh, py = ry >> sh;
    const int hx = none ? 0 : (v.version15 ? px : rx) & 1, hy = none ? 0 : (v.version15 ? py : ry) & 1;
#if defined(__CUDA_ARCH__)
    const bool any_diag = __any_sync(__activemask(), hx & hy);
#else
    const bool any_diag = hx & hy;
#endif
    if (none) { rows[0] = rows[1] = rows[2] = rows[3] = 0x80808080u; return; }
    /* column and row inside the patch: the box starts at the integer position, x rounded down to 16 */
    const uint32_t col = (uint32_t)((px >> 1) & 15) + (p ? 0u : (uint32_t)(bx & 1) * 4u);
    const uint32_t row = p ? 0u : (uint32_t)(lrow & 1) * 4u;
    const uint32_t base = w.m->patch_base + poff * RW_BOX_W + (p == 0 ? 0u : p == 1 ? (uint32_t)RW_PATCH_C_OFF : (uint32_t)(RW_PATCH_C_OFF + RW_PATCH_C));
    const uint32_t a = col & 3u;
    const RwPatchRows pr = {base + row * RW_BOX_W + (col - a)};
    uint32_t W[10];
    rc_predict_load_rows<false>(W, pr, a, hx, hy);
    rc_predict_filter(rows, W, a, hx, hy, any_diag);
}

/* a finished block goes into the row's tile (picture layout: 8 luma rows, then 4 rows of U, of V) */
RC_HD void rw_store_block(const RowWork &w, int p, int bx, int lrow, const uint32_t rows[4])
{
    const RowGeom &g = *w.g;
    const uint32_t pitch = (uint32_t)(g.width >> (p ? 1 : 0));
    const uint32_t plane_off = p == 0 ? 0u : g.tile_y_bytes + (p == 2 ? g.tile_c_bytes : 0u);
    uint8_t *dst = SW_SMEM(w.slot_off + g.s_tile) + plane_off + (uint32_t)(lrow * 4) * pitch + (uint32_t)bx * 4u;
#pragma unroll
    for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + (uint32_t)r * pitch) = rows[r];
}

/* entry idx of list l */
RC_HD void rw_list_lane(const RowWork &w, int l, uint32_t idx)
{
    const RowGeom &g = *w.g;
    const ReconView &v = *w.v;
    const RowSlotMeta &m = *w.m;
    const bool live = idx < m.n_list[l];
    const uint32_t e = live ? reinterpret_cast<const uint16_t *>(SW_SMEM(w.slot_off + g.s_list[l]))[idx] : 0u;
    const int bx = (int)(e & 0x1FF), lrow = (int)((e >> 9) & 1), p = (int)((e >> 10) & 3);
    uint32_t rows[4];
    if (l == RW_LIST_MC)
    {
        const int mx = p ? bx : bx >> 1;
        const uint32_t mvw = live ? reinterpret_cast<const uint32_t *>(SW_SMEM(m.p_mv))[mx] : 0u;
        const uint32_t poff = live ? reinterpret_cast<const uint16_t *>(SW_SMEM(w.slot_off + g.s_poff))[mx] : RW_NO_PATCH;
        rw_predict_block(w, p, bx, lrow, mvw, poff, rows);       /* every lane: the filter form is a warp vote */
        if (!live) return;
    }
    else
    {
        if (!live) return;
        const int stride = g.stride[p];
        const uint8_t *tcell = SW_SMEM(m.p_type[p]) + (lrow + 1) * stride + bx + 1;
        const uint8_t *dcell = SW_SMEM(m.p_dc[p]) + (lrow + 1) * stride + bx + 1;
        if (l == RW_LIST_W) rc_weighted_at(tcell, dcell, stride, v.is_ipic, rows);
        else rows[0] = rows[1] = rows[2] = rows[3] = (uint32_t)*dcell * 0x01010101u;
    }
    rw_store_block(w, p, bx, lrow, rows);
}

/* record idx of the (class, row) range: raw block h4m:543-549, intra AOT h4m:1358-1377, predicted AOT h4m:1379-1420
   including its motion-compensated prediction */
RC_HD void rw_record_lane(const RowWork &w, int cls, uint32_t idx)
{
    const RowGeom &g = *w.g;
    const ReconView &v = *w.v;
    const RowSlotMeta &m = *w.m;
    const bool live = idx < m.n_rec[cls];
    /* the chunks are ordered by length, each holds count records of one length */
    const uint2 *d = reinterpret_cast<const uint2 *>(SW_SMEM(m.p_desc[cls]));
    uint32_t i = idx, len = 1, first = m.rec_lo[cls];
    if (live)
    {
        uint2 cd = d[0];
        uint32_t j = 0;
        while (i >= (cd.y & 0xFF))
        {
            i -= cd.y & 0xFF;
            cd = d[++j];
        }
        len = ((cd.y >> 8) & 0xFF) + 1;
        first = cd.x;
    }
    const uint32_t *rec = v.rec + first + i * len;          /* global memory: every record is read once */
    uint32_t t = 0;
    int p = 0, bx = 0, by = 0;
#if defined(__CUDA_ARCH__)
    if (live) rc_record_coords(__ldg(rec), t, p, bx, by);
#else
    if (live) rc_record_coords(rec[0], t, p, bx, by);
#endif
    const int lrow = by - (p ? m.row : 2 * m.row);
    const int mx = p ? bx : bx >> 1;
    uint32_t rows[4];
    if (cls == SYM_REC_INTER)
    {
        const uint32_t mvw = live ? reinterpret_cast<const uint32_t *>(SW_SMEM(m.p_mv))[mx] : 0u;
        const uint32_t poff = live ? reinterpret_cast<const uint16_t *>(SW_SMEM(w.slot_off + g.s_poff))[mx] : RW_NO_PATCH;
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        rw_predict_block(w, p, bx, lrow, mvw, poff, rows);       /* every lane: the filter form is a warp vote */
        if (!live) return;
        const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
        if (rx != -32768)
        {   /* window origin, h4m:1864-1868; linear addressing like the reference */
            const uint8_t *ref = ((t >> 5) & 3) == 2 ? v.ref[1] : v.ref[0];
            const RwGlobalWindow win = {ref + rx / 2 + (ry / 2 - 16) * v.width - 32, v.width};
            rc_predicted_aot(v, rows, rec + 1, (int)len - 1, win);
        }
    }
    else
    {
        if (!live) return;
        if (cls == SYM_REC_RAW)
        {
#pragma unroll
            for (int r = 0; r < 4; ++r) rows[r] = RC_LD32(rec + 1 + r);
        }
        else
        {
            const int V = *(SW_SMEM(m.p_dc[p]) + (lrow + 1) * g.stride[p] + bx + 1);
            rc_intra_aot(v, rows, rec + 1, (int)len - 1, V);
        }
    }
    rw_store_block(w, p, bx, lrow, rows);
}

/* task t of a row -> what to do */
RC_HD void rw_run_task(const RowWork &w, const uint32_t t_end[RW_TASK_CLASSES], uint32_t t, int lane)
{
    if (t < t_end[RW_T_INTER]) rw_record_lane(w, SYM_REC_INTER, t * 32u + (uint32_t)lane);
    else if (t < t_end[RW_T_INTRA]) rw_record_lane(w, SYM_REC_INTRA, (t - t_end[RW_T_INTER]) * 32u + (uint32_t)lane);
    else if (t < t_end[RW_T_MC]) rw_list_lane(w, RW_LIST_MC, (t - t_end[RW_T_INTRA]) * 32u + (uint32_t)lane);
    else if (t < t_end[RW_T_W]) rw_list_lane(w, RW_LIST_W, (t - t_end[RW_T_MC]) * 32u + (uint32_t)lane);
    else if (t < t_end[RW_T_RAW]) rw_record_lane(w, SYM_REC_RAW, (t - t_end[RW_T_W]) * 32u + (uint32_t)lane);
    else rw_list_lane(w, RW_LIST_FLAT, (t - t_end[RW_T_RAW]) * 32u + (uint32_t)lane);
}

#endif
