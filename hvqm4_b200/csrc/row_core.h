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
 *   Roles (row.cu): a SEQUENCER warp requests symbol slices, allocates ring space and retires rows; FETCH warps
 *   classify a row and issue its patch copies; WORK warps take tasks from a row's ticket counter.
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
    uint32_t bf[SYM_REC_CLASSES][RW_MAX_ROWS + 1];        /* first chunk of (class, macroblock row) */
    uint32_t rec_off[SYM_REC_CLASSES][RW_MAX_ROWS + 1];   /* first record word of (class, macroblock row) */
};

/* what the pipeline publishes for a row (lives in the row's slot) */
struct RowSlotMeta
{
    uint32_t ticket;                         /* next task (work warps fetch-and-add) */
    uint32_t t_end[RW_TASK_CLASSES];         /* tasks [t_end[k-1], t_end[k]) belong to class k */
    uint32_t n_rec[SYM_REC_CLASSES], rec_lo[SYM_REC_CLASSES];
    uint32_t n_list[RW_LISTS];
    uint32_t p_type[3], p_dc[3];             /* shared-memory offset of the bordered map row above the row's first block row */
    uint32_t p_mv, p_desc[SYM_REC_CLASSES];
    uint32_t patch_base;                     /* shared-memory offset of the row's patches */
    uint32_t n_patch;
    uint32_t ring_end;                       /* ring head after this row's allocation (the tail once it retires) */
    int32_t row;                             /* macroblock row inside the picture */
};

RC_HD int rw_make_geom(RowGeom &g, int width, int height, uint32_t smem_limit)
{
    if (width <= 0 || height <= 0 || (width & 31) || (height & 7) || width > 2048) return 0;
    g.width = width; g.height = height; g.mcb_w = width / 8; g.mcb_h = height / 8;
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

/* copy `id` (0-2 type rows, 3-5 DC rows, 6 vectors, 7-9 chunk descriptors) of macroblock row `row`; also fills the
   matching pointer of the slot's meta (lane id does both, the fields are disjoint) */
RC_HD SwCopy rw_sym_copy(const RowGeom &g, const ReconView &v, const RowCtl &c, int row, int id, uint32_t slot_off, RowSlotMeta &m)
{
    SwCopy k = {SW_SRC_BLOB, 0, 0, 0};
    uint32_t lo = 0, hi = 0, region = 0;
    if (id < 6)
    {
        const int p = id % 3;
        const int by0 = p ? row : 2 * row, nrows = (p ? 1 : 2) + 2;
        lo = (id < 3 ? rc_pick3(v.off_type, p) : rc_pick3(v.off_dc, p)) + (uint32_t)(by0 * g.stride[p]);   /* bordered row by0 = block row by0 - 1 */
        hi = lo + (uint32_t)(nrows * g.stride[p]);
        region = id < 3 ? g.s_type[p] : g.s_dc[p];
    }
    else if (id == 6)
    {
        region = g.s_mv;
        if (!v.is_ipic)
        {
            lo = v.off_mv + (uint32_t)(row * g.mcb_w * 4);
            hi = lo + (uint32_t)(g.mcb_w * 4);
        }
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

/* Sequencer, once a row's symbol slice has landed: which macroblocks get a patch and where (offset in units of
   RW_BOX_W bytes from the row's patch base, RW_NO_PATCH = none).  Warp-collective on the GPU.  Returns the number
   of patches; *bad is set if some macroblock cannot be served. */
RC_HD uint32_t rw_plan_patches(const RowGeom &g, const ReconView &v, const RowCtl &c, uint32_t slot_off, const RowSlotMeta &m, int lane, int *bad_any)
{
    uint16_t *poff = reinterpret_cast<uint16_t *>(SW_SMEM(slot_off + g.s_poff));
    uint32_t n = 0;
    int bad_acc = 0;
    for (int mx0 = 0; mx0 < g.mcb_w; mx0 += SW_LANES)
    {
        const int mx = mx0 + lane;
        int ref = 0, xl, yl, xc, yc, bad = 0;
        if (mx < g.mcb_w && !v.is_ipic)
        {
            const uint32_t tag = *(SW_SMEM(m.p_type[0]) + g.stride[0] + 2 * mx + 1);
            const uint32_t mvw = reinterpret_cast<const uint32_t *>(SW_SMEM(m.p_mv))[mx];
            rw_patch_box(v, tag, mvw, ref, xl, yl, xc, yc, bad);
            if (ref == 2 && !c.is_bpic) { bad = 1; ref = 0; }      /* P picture (h4m:2058-2061): `future` is the picture itself */
            if (ref && c.z[ref - 1] < 0) { bad = 1; ref = 0; }     /* the reference is not a surface of the registered slab */
        }
        bad_acc |= bad;
        const uint32_t bal = sw_ballot(ref != 0);
        if (mx < g.mcb_w) poff[mx] = ref ? (uint16_t)((n + (uint32_t)SW_POPC(bal & ((1u << lane) - 1u))) * (RW_PATCH_BYTES / RW_BOX_W)) : (uint16_t)RW_NO_PATCH;
        n += (uint32_t)SW_POPC(bal);
    }
    *bad_any = (int)sw_ballot(bad_acc != 0);
    return n;
}

/* Ring allocation (sequencer registers): [tail, head) in ring order is in use; rows retire in order. */
struct RwRing
{
    uint32_t head, tail, live;    /* live: rows that hold an allocation */
};

/* returns the offset inside the ring, or 0xFFFFFFFF if the bytes are not free yet */
RC_HD uint32_t rw_ring_alloc(RwRing &r, uint32_t cap, uint32_t bytes)
{
    bytes = (bytes + 127u) & ~127u;
    if (!r.live) { r.head = r.tail = 0; }
    uint32_t pos;
    if (r.head >= r.tail && r.live)
    {   /* in use: [tail, head); free: [head, cap) and [0, tail) */
        if (r.head + bytes <= cap) pos = r.head;
        else if (bytes < r.tail) pos = 0;
        else return 0xFFFFFFFFu;
    }
    else if (!r.live)
    {
        if (bytes > cap) return 0xFFFFFFFFu;
        pos = 0;
    }
    else
    {   /* wrapped: free is [head, tail) */
        if (r.head + bytes < r.tail) pos = r.head;
        else return 0xFFFFFFFFu;
    }
    r.head = pos + bytes;
    ++r.live;
    return pos;
}
RC_HD void rw_ring_retire(RwRing &r, uint32_t ring_end)
{
    r.tail = ring_end;
    --r.live;
}

/* ---- classification of a row (fetch warp) ------------------------------------------------------ */

/* list entry: [8:0] block x, [9] local block row (luma), [11:10] plane */
RC_HD uint32_t rw_entry(int p, int lrow, int bx) { return (uint32_t)bx | (uint32_t)lrow << 9 | (uint32_t)p << 10; }

/* Builds the three block lists of the row, sums the record counts and sets the task boundaries.  Warp-collective. */
RC_HD void rw_classify_row(const RowGeom &g, const ReconView &v, const RowCtl &c, int row, uint32_t slot_off, RowSlotMeta &m, int lane)
{
    uint32_t n_list[RW_LISTS] = {0, 0, 0};
    uint16_t *lists[RW_LISTS];
    for (int l = 0; l < RW_LISTS; ++l) lists[l] = reinterpret_cast<uint16_t *>(SW_SMEM(slot_off + g.s_list[l]));
    const bool ipic = v.is_ipic != 0;
    for (int br = 0; br < 4; ++br)
    {   /* block rows: luma upper, luma lower, U, V */
        const int p = br < 2 ? 0 : br - 1, lrow = br < 2 ? br : 0;
        const uint8_t *trow = SW_SMEM(m.p_type[p]) + (lrow + 1) * g.stride[p] + 1;
        for (int x0 = 0; x0 < g.bw[p]; x0 += SW_LANES)
        {
            const int bx = x0 + lane;
            const uint32_t t = bx < g.bw[p] ? trow[bx] : 6u;       /* past the row end: a raw block is nothing to do here */
            const uint32_t nib = ipic ? t : (t & 0xF);
            const bool inter = !ipic && (t & 0x60);
            int l = -1;
            if (inter) { if ((t & 0x10) || nib == 0) l = RW_LIST_MC; }
            else l = nib == 0 ? RW_LIST_W : nib == 8 ? RW_LIST_FLAT : -1;
            const uint32_t e = rw_entry(p, lrow, bx);
#pragma unroll
            for (int k = 0; k < RW_LISTS; ++k)
            {
                const uint32_t bal = sw_ballot(l == k);
                if (l == k) lists[k][n_list[k] + (uint32_t)SW_POPC(bal & ((1u << lane) - 1u))] = (uint16_t)e;
                n_list[k] += (uint32_t)SW_POPC(bal);
            }
        }
    }
    uint32_t n_rec[SYM_REC_CLASSES];
    for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
    {
        const uint32_t nd = c.bf[cls][row + 1] - c.bf[cls][row];
        const uint2 *d = reinterpret_cast<const uint2 *>(SW_SMEM(m.p_desc[cls]));
        uint32_t sum = 0;
        for (uint32_t j = (uint32_t)lane; j < nd; j += SW_LANES) sum += d[j].y & 0xFF;
        n_rec[cls] = sw_lane_sum(sum);
    }
    if (lane == 0)
    {
        for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
        {
            m.n_rec[cls] = n_rec[cls];
            m.rec_lo[cls] = c.rec_off[cls][row];
        }
        for (int l = 0; l < RW_LISTS; ++l) m.n_list[l] = n_list[l];
        uint32_t t = 0;
        t += (n_rec[SYM_REC_INTER] + 31u) / 32u; m.t_end[RW_T_INTER] = t;
        t += (n_rec[SYM_REC_INTRA] + 31u) / 32u; m.t_end[RW_T_INTRA] = t;
        t += (n_list[RW_LIST_MC] + 31u) / 32u;   m.t_end[RW_T_MC] = t;
        t += (n_list[RW_LIST_W] + 31u) / 32u;    m.t_end[RW_T_W] = t;
        t += (n_rec[SYM_REC_RAW] + 31u) / 32u;   m.t_end[RW_T_RAW] = t;
        t += (n_list[RW_LIST_FLAT] + 31u) / 32u; m.t_end[RW_T_FLAT] = t;
        m.ticket = 0;
        m.row = row;
    }
    sw_lane_sync();
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
            rc_predicted_aot(v, rows, rec + 1, (int)len - 1, &win);
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
RC_HD void rw_run_task(const RowWork &w, uint32_t t, int lane)
{
    const RowSlotMeta &m = *w.m;
    if (t < m.t_end[RW_T_INTER]) rw_record_lane(w, SYM_REC_INTER, t * 32u + (uint32_t)lane);
    else if (t < m.t_end[RW_T_INTRA]) rw_record_lane(w, SYM_REC_INTRA, (t - m.t_end[RW_T_INTER]) * 32u + (uint32_t)lane);
    else if (t < m.t_end[RW_T_MC]) rw_list_lane(w, RW_LIST_MC, (t - m.t_end[RW_T_INTRA]) * 32u + (uint32_t)lane);
    else if (t < m.t_end[RW_T_W]) rw_list_lane(w, RW_LIST_W, (t - m.t_end[RW_T_MC]) * 32u + (uint32_t)lane);
    else if (t < m.t_end[RW_T_RAW]) rw_record_lane(w, SYM_REC_RAW, (t - m.t_end[RW_T_W]) * 32u + (uint32_t)lane);
    else rw_list_lane(w, RW_LIST_FLAT, (t - m.t_end[RW_T_RAW]) * 32u + (uint32_t)lane);
}

#endif
