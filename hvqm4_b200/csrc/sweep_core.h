/*
 * sweep_core.h -- everything of the SWEEP kernel (sweep.cu) that is not a CUDA synchronisation
 * primitive: geometry and shared-memory layout, the per-picture plan (which reference rows every
 * band of macroblock rows needs), what the producer copies for a band and where, the ring
 * addressing of the reference windows, and the per-lane work of a map task and of a record task.
 * tests/emul/sweep_emul.cpp drives the very same functions serially on the CPU (test
 * infrastructure), so the plan, the ring arithmetic and the task decomposition are checked
 * against the oracle without a GPU; sweep.cu adds the TMA copies and the mbarrier pipeline.
 *
 * The sweep ("h4m:N" = /root/reference/h4m_audio_decode.c line N):
 *   One CTA reconstructs one picture top to bottom in BANDS of h macroblock rows (the
 *   reference's raster walk, h4m:1487-1518 / 1922-1967, at band granularity).  Everything a
 *   band needs is staged in shared memory by bulk asynchronous copies issued bands ahead:
 *     - the band's slice of the symbol buffer: type / DC map rows (with the neighbour rows of the
 *       weighted fill, h4m:299-383), the vector row, the chunk descriptors and the records of the
 *       band (symbuf.h groups records per macroblock row);
 *     - the REFERENCE WINDOW: the rows of the reference frame that the band's motion vectors
 *       (h4m:1327-1355) and AOT windows (h4m:1864-1868) reach, kept in a ring of whole picture
 *       rows per plane.  The window slides down with the bands, so every reference byte is read
 *       from L2/HBM once per picture, by 128-byte-line bulk copies, instead of once per 4x4
 *       block through a 32-byte sector; all gathers (half-sample rows, basis rows) are
 *       shared-memory loads.
 *   The band's output tile (8h luma rows, 4h rows of U and V) is assembled in shared memory and
 *   leaves with three bulk stores.
 *   B pictures use two references (h4m:2018-2056).  Both windows do not fit next to each other, so
 *   a B picture is swept twice: first the macroblocks predicted from `future` alone, finished
 *   macroblocks going to a compact per-CTA scratch list (96 bytes per macroblock, L2 resident),
 *   then everything else with the `past` window, the first sweep's macroblocks copied in.
 *   A picture the plan cannot serve (vectors that wrap around picture rows or leave their plane,
 *   windows larger than shared memory, bands with more symbol data than a slot holds) is left to
 *   the band kernel: the sweep kernel marks it in its job entry.
 */
#ifndef HVQM4_SWEEP_CORE_H
#define HVQM4_SWEEP_CORE_H

#include "recon_core.h"

#if !defined(__CUDACC__)   /* tests/emul: the two vector types the lane functions use */
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
#endif

#define SW_MAX_BANDS 128        /* bands and macroblock rows per picture the per-band tables hold */
#define SW_NSLOTS 3             /* bands in flight: symbol data + tile per slot */
#define SW_DESC_CAP 32          /* chunk descriptors per (class, band) a slot holds */
#define SW_MCB_BYTES 96         /* one finished macroblock in the scratch list: 4 luma blocks, U, V, 16 bytes each */
#define SW_EMPTY_LO 0x3FFF
#define SW_EMPTY_HI (-0x3FFF)

enum { SW_MODE_ALL = 0, SW_MODE_FUTURE = 1, SW_MODE_MERGE = 2 };

#if defined(__CUDA_ARCH__)
#define SW_SMEM(off) (rc_smem + (off))
#else
extern uint8_t *sw_host_smem;   /* tests/emul: the CTA's shared memory */
#define SW_SMEM(off) (sw_host_smem + (off))
#endif

RC_HD uint32_t sw_align16(uint32_t x) { return (x + 15u) & ~15u; }

/* ---- geometry and shared-memory layout (one per launch; all pictures of a launch share it) ---- */
struct SweepGeom
{
    int width, height, mcb_w, mcb_h;
    int h;                       /* macroblock rows per band */
    int n_bands;
    int bw[3], stride[3];        /* blocks per block row, bordered map pitch (plane 0, 1, 2) */
    int seg[3];                  /* 32-block map tasks per block row */
    uint32_t off_ctl, off_slot0, slot_bytes, off_win, win_bytes, smem_bytes;
    /* inside a slot */
    uint32_t s_meta, s_type[3], s_dc[3], s_mv, s_desc, s_rec, s_rank, s_tile;
    uint32_t cap_rec;            /* bytes of records a slot holds (all three classes) */
    uint32_t tile_y_bytes, tile_c_bytes;   /* full band: 8h rows of luma, 4h rows of one chroma plane */
};

/* per-picture tables and state of the CTA */
struct SweepCtl
{
    unsigned long long bar_full[SW_NSLOTS], bar_ready[SW_NSLOTS], bar_done[SW_NSLOTS];
    int32_t unsupported;                    /* set by any thread while planning */
    int32_t n_future;                       /* macroblocks predicted from `future` in the picture */
    int32_t cap_y, cap_c;                   /* ring rows (luma, one chroma plane) of the current sweep */
    int32_t first_y, first_c;               /* first row ever loaded = ring slot 0 */
    uint32_t ring_off[3], ring_bytes[2];    /* ring_bytes: luma, chroma */
    uint32_t abort_flag;                    /* a wait timed out: everybody leaves (debug guard) */
    uint32_t pad;
    uint32_t bf[SYM_REC_CLASSES][SW_MAX_BANDS + 1];        /* band table of the picture (first chunk per class and macroblock row) */
    uint32_t rec_off[SYM_REC_CLASSES][SW_MAX_BANDS + 1];   /* first record word of (class, macroblock row) */
    /* reference rows needed per band and reference (0 past, 1 future): [lo, hi) in rows of the plane; after
       sw_plan_scan(): lo = suffix minimum (nothing below it is needed by this or any later band), hi = prefix maximum */
    int16_t lo_y[2][SW_MAX_BANDS], hi_y[2][SW_MAX_BANDS], lo_c[2][SW_MAX_BANDS], hi_c[2][SW_MAX_BANDS];
    uint16_t n2[SW_MAX_BANDS];              /* `future` macroblocks of the band */
    uint32_t side_off[SW_MAX_BANDS + 1];    /* their first entry in the scratch list */
};

/* what the producer publishes for a band (lives in the band's slot) */
struct SweepSlotMeta
{
    uint32_t ticket;                         /* next task (consumers fetch-and-add) */
    uint32_t n_tasks, t_inter, t_intra, t_map;   /* tasks [0, t_inter) predicted AOT, [t_inter, t_intra) intra AOT, [t_intra, t_map) map, [t_map, n_tasks) raw */
    uint32_t n_rec[SYM_REC_CLASSES], n_desc[SYM_REC_CLASSES];
    uint32_t rec_lo[SYM_REC_CLASSES];        /* first record word of the (class, band) range */
    uint32_t p_type[3], p_dc[3];             /* shared-memory offset of map cell (bx = -1, local block row = -1) */
    uint32_t p_mv, p_desc[SYM_REC_CLASSES], p_rec[SYM_REC_CLASSES];
    int32_t row0_y, slot0_y, row0_c, slot0_c;    /* ring mapping: row row0 of the plane sits in ring row slot0 */
    int32_t rows;                            /* macroblock rows of this band (the last band may be short) */
    uint32_t n2, side_off;
};

RC_HD int sw_band_rows(const SweepGeom &g, int band)
{
    const int r = g.mcb_h - band * g.h;
    return r < g.h ? r : g.h;
}

/* returns 0 if the geometry is not served by the sweep kernel */
RC_HD int sw_make_geom(SweepGeom &g, int width, int height, int h, uint32_t smem_limit)
{
    if (width <= 0 || height <= 0 || (width & 31) || (height & 7) || width > 2048 || h < 1 || h > 4) return 0;
    if (width < height) return 0;      /* portrait pictures (38 x 70 nest, swapped window origin): the other kernels */
    g.width = width; g.height = height; g.mcb_w = width / 8; g.mcb_h = height / 8;
    g.h = h;
    g.n_bands = (g.mcb_h + h - 1) / h;
    if (g.mcb_h > SW_MAX_BANDS) return 0;
    for (int p = 0; p < 3; ++p)
    {
        g.bw[p] = (width >> (p ? 1 : 0)) / 4;
        g.stride[p] = g.bw[p] + 2;
        g.seg[p] = (g.bw[p] + 31) / 32;
    }
    uint32_t at = RC_SMEM_TABLE_BYTES;
    g.off_ctl = at = sw_align16(at);
    at += (uint32_t)sizeof(SweepCtl);
    g.off_slot0 = at = (at + 127u) & ~127u;
    /* slot */
    uint32_t s = 0;
    g.s_meta = s; s += sw_align16((uint32_t)sizeof(SweepSlotMeta));
    for (int p = 0; p < 3; ++p) { g.s_type[p] = s; s += sw_align16((uint32_t)(((p ? h : 2 * h) + 2) * g.stride[p]) + 32); }
    for (int p = 0; p < 3; ++p) { g.s_dc[p] = s; s += sw_align16((uint32_t)(((p ? h : 2 * h) + 2) * g.stride[p]) + 32); }
    g.s_mv = s; s += sw_align16((uint32_t)(h * g.mcb_w * 4) + 32);
    g.s_desc = s; s += SYM_REC_CLASSES * (SW_DESC_CAP * 8 + 16);
    /* records: dense synthetic content carries ~9 bytes per block (SURVEY 8d distributions); twelve leaves headroom */
    g.cap_rec = sw_align16((uint32_t)(h * g.mcb_w * 6 * 12)) + 64;
    g.s_rec = s; s += g.cap_rec;
    g.s_rank = s; s += sw_align16((uint32_t)(h * g.mcb_w * 2));
    g.tile_y_bytes = (uint32_t)(8 * h * width);
    g.tile_c_bytes = (uint32_t)(4 * h * (width / 2));
    g.s_tile = s = (s + 127u) & ~127u;
    s += g.tile_y_bytes + 2 * g.tile_c_bytes;
    g.slot_bytes = (s + 127u) & ~127u;
    at += SW_NSLOTS * g.slot_bytes;
    g.off_win = at;
    if (at + 16u * (uint32_t)width > smem_limit) return 0;
    g.win_bytes = (smem_limit - at) & ~15u;
    g.smem_bytes = at + g.win_bytes;
    return 1;
}

/* ---- per-picture plan ---------------------------------------------------------------------- */

/* Reference rows one inter macroblock needs: ref = 0 none (intra), 1 past, 2 future; ext = {lo_y, hi_y, lo_c, hi_c}
   ([lo, hi) in rows of the plane, empty when the macroblock is poisoned); bad != 0: the sweep kernel does not serve this
   picture (a prediction or window that is not inside its plane rows and columns: the reference addresses linearly
   and such vectors wrap around rows, h4m:1344,1866,1897 -- the band kernel reproduces that). */
RC_HD void sw_prescan_mcb(const ReconView &v, int mx, int my, int &ref, int ext[4], int &bad)
{
    const int strideY = (v.width >> 2) + 2, strideC = (v.width >> 3) + 2;
    const uint8_t *ty = v.blob + v.off_type[0] + (2 * my + 1) * strideY + 2 * mx + 1;
    const uint32_t tag = RC_LD8(ty);
    ref = (int)((tag >> 5) & 3);
    bad = 0;
    ext[0] = ext[2] = SW_EMPTY_LO;
    ext[1] = ext[3] = SW_EMPTY_HI;
    if (!ref) return;
    if (ref == 3) { bad = 1; return; }
    const uint32_t mvw = RC_LD32(v.blob + v.off_mv + 4 * (my * v.mcb_w + mx));
    const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
    if (rx == -32768) return;                                  /* poisoned: painted grey, reads nothing */
    {   /* luma 8x8 (+1 with a half step), h4m:1327-1355 */
        const int px = rx >> 1, py = ry >> 1, hx = rx & 1, hy = ry & 1;
        if (px < 0 || py < 0 || px + 8 + hx > v.width || py + 8 + hy > v.height) { bad = 1; return; }
        ext[0] = py;
        ext[1] = py + 8 + hy;
    }
    {   /* chroma 4x4 per plane; 1.3 reuses the luma phase (h4m:1337-1343) */
        const int pxc = rx >> 1, pyc = ry >> 1;
        const int hx = (v.version15 ? pxc : rx) & 1, hy = (v.version15 ? pyc : ry) & 1;
        const int cx = pxc >> 1, cy = pyc >> 1;
        if (cx < 0 || cy < 0 || cx + 4 + hx > (v.width >> 1) || cy + 4 + hy > (v.height >> 1)) { bad = 1; return; }
        ext[2] = cy;
        ext[3] = cy + 4 + hy;
    }
    if (!(tag & 0x10))
    {   /* proc 0: predicted-AOT blocks gather from the 70x38 luma window (h4m:1864-1868) */
        const int wx0 = rx / 2 - 32, wy0 = ry / 2 - 16;
        if (wx0 >= 0 && wy0 >= 0 && wx0 + SYM_NEST_W <= v.width && wy0 + SYM_NEST_H <= v.height)
        {
            ext[0] = wy0 < ext[0] ? wy0 : ext[0];
            ext[1] = wy0 + SYM_NEST_H > ext[1] ? wy0 + SYM_NEST_H : ext[1];
        }
        else
        {   /* a window outside the plane is fine as long as no block of the macroblock has bases */
            const uint32_t t6[6] = {tag, RC_LD8(ty + 1), RC_LD8(ty + strideY), RC_LD8(ty + strideY + 1),
                                    RC_LD8(v.blob + v.off_type[1] + (my + 1) * strideC + mx + 1),
                                    RC_LD8(v.blob + v.off_type[2] + (my + 1) * strideC + mx + 1)};
#pragma unroll
            for (int k = 0; k < 6; ++k)
            {
                const uint32_t nib = t6[k] & 0xF;
                if (nib > 1 && nib != 6) bad = 1;
            }
        }
    }
}

/* After every band's [lo, hi) is known: lo -> suffix minimum, hi -> prefix maximum (one call per array pair),
   so that [lo[b], hi[b]) is what must be resident while band b is computed, both non-decreasing in b. */
RC_HD void sw_plan_scan(int16_t *lo, int16_t *hi, int n)
{
    int m = SW_EMPTY_LO;
    for (int b = n - 1; b >= 0; --b)
    {
        m = lo[b] < m ? lo[b] : m;
        lo[b] = (int16_t)m;
    }
    m = SW_EMPTY_HI;
    for (int b = 0; b < n; ++b)
    {
        m = hi[b] > m ? hi[b] : m;
        hi[b] = (int16_t)m;
    }
}

/* rows the ring must hold for reference f: max over bands of hi - lo (0 if the reference is unused) */
RC_HD int sw_plan_depth(const int16_t *lo, const int16_t *hi, int n)
{
    int d = 0;
    for (int b = 0; b < n; ++b)
        if (hi[b] > lo[b] && hi[b] - lo[b] > d) d = hi[b] - lo[b];
    return d;
}

/* Ring sizes for a sweep over reference f (after the scans).  Returns 0 if the windows do not fit. */
RC_HD int sw_plan_rings(const SweepGeom &g, SweepCtl &c, int f)
{
    const int dy = sw_plan_depth(c.lo_y[f], c.hi_y[f], g.n_bands), dc = sw_plan_depth(c.lo_c[f], c.hi_c[f], g.n_bands);
    const uint32_t need = (uint32_t)(dy + dc) * (uint32_t)g.width;    /* luma rows + rows of U and V (half width each) */
    if (need > g.win_bytes) return 0;
    /* spare rows let the producer load ahead: 8h luma + 4h chroma rows per band of look-ahead */
    const uint32_t step = (uint32_t)(12 * g.h * g.width);
    uint32_t extra = (g.win_bytes - need) / step;
    if (extra > SW_NSLOTS) extra = SW_NSLOTS;
    c.cap_y = dy + 8 * g.h * (int)extra;
    c.cap_c = dc + 4 * g.h * (int)extra;
    if (c.cap_y < 1) c.cap_y = 1;
    if (c.cap_c < 1) c.cap_c = 1;
    c.ring_bytes[0] = (uint32_t)c.cap_y * (uint32_t)g.width;
    c.ring_bytes[1] = (uint32_t)c.cap_c * (uint32_t)(g.width / 2);
    c.ring_off[0] = g.off_win;
    c.ring_off[1] = c.ring_off[0] + c.ring_bytes[0];
    c.ring_off[2] = c.ring_off[1] + c.ring_bytes[1];
    c.first_y = SW_EMPTY_LO;
    c.first_c = SW_EMPTY_LO;
    for (int b = g.n_bands - 1; b >= 0; --b)
    {
        if (c.hi_y[f][b] > c.lo_y[f][b]) c.first_y = c.lo_y[f][b];
        if (c.hi_c[f][b] > c.lo_c[f][b]) c.first_c = c.lo_c[f][b];
    }
    return 1;
}

/* symbol data of a band must fit its slot */
RC_HD int sw_band_fits(const SweepGeom &g, const SweepCtl &c, int band)
{
    const int r0 = band * g.h, r1 = r0 + sw_band_rows(g, band);
    uint32_t bytes = 0;
    for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
    {
        if (c.bf[cls][r1] - c.bf[cls][r0] > SW_DESC_CAP) return 0;
        const uint32_t lo = c.rec_off[cls][r0] * 4u, hi = c.rec_off[cls][r1] * 4u;
        bytes += sw_align16(hi + 15u - (lo & ~15u));
    }
    return bytes <= g.cap_rec;
}

/* ---- what the producer copies for a band ---------------------------------------------------- */
enum { SW_SRC_BLOB = 0, SW_SRC_REF = 1, SW_SRC_SCRATCH = 2 };
struct SwCopy
{
    uint32_t src_kind;
    uint32_t src_off;     /* bytes from the blob / reference surface (16-byte aligned) */
    uint32_t dst_off;     /* shared memory (16-byte aligned) */
    uint32_t bytes;       /* multiple of 16; 0 = nothing */
};
#define SW_N_SYM_COPIES 13

/* copy `id` (0-2 type rows, 3-5 DC rows, 6 vectors, 7-9 chunk descriptors, 10-12 records) of a band's symbol data;
   also fills the matching pointer of the slot's meta (lane id does both, the fields are disjoint) */
RC_HD SwCopy sw_sym_copy(const SweepGeom &g, const ReconView &v, const SweepCtl &c, int band, int id, uint32_t slot_off, SweepSlotMeta &m)
{
    SwCopy k = {SW_SRC_BLOB, 0, 0, 0};
    const int r0 = band * g.h, rows = sw_band_rows(g, band), r1 = r0 + rows;
    uint32_t lo = 0, hi = 0, region = 0;
    if (id < 6)
    {
        const int p = id % 3;
        const int by0 = p ? r0 : 2 * r0, nrows = (p ? rows : 2 * rows) + 2;
        lo = (id < 3 ? rc_pick3(v.off_type, p) : rc_pick3(v.off_dc, p)) + (uint32_t)(by0 * g.stride[p]);   /* bordered row by0 = block row by0 - 1 */
        hi = lo + (uint32_t)(nrows * g.stride[p]);
        region = id < 3 ? g.s_type[p] : g.s_dc[p];
    }
    else if (id == 6)
    {
        if (v.is_ipic) return k;
        lo = v.off_mv + (uint32_t)(r0 * g.mcb_w * 4);
        hi = lo + (uint32_t)(rows * g.mcb_w * 4);
        region = g.s_mv;
    }
    else if (id < 10)
    {
        const int cls = id - 7;
        lo = (uint32_t)((const uint8_t *)v.chunks - v.blob) + c.bf[cls][r0] * 8u;
        hi = lo + (c.bf[cls][r1] - c.bf[cls][r0]) * 8u;
        region = g.s_desc + (uint32_t)cls * (SW_DESC_CAP * 8 + 16);
    }
    else
    {
        const int cls = id - 10;
        region = g.s_rec;
        for (int q = 0; q <= cls; ++q)
        {
            lo = (uint32_t)((const uint8_t *)v.rec - v.blob) + c.rec_off[q][r0] * 4u;
            hi = (uint32_t)((const uint8_t *)v.rec - v.blob) + c.rec_off[q][r1] * 4u;
            if (q < cls) region += hi > lo ? sw_align16(hi + 15u - (lo & ~15u)) : 0u;
        }
    }
    const uint32_t a0 = lo & ~15u, a1 = sw_align16(hi);
    const uint32_t ptr = slot_off + region + (lo - a0);
    if (id < 3) m.p_type[id] = ptr;
    else if (id < 6) m.p_dc[id - 3] = ptr;
    else if (id == 6) m.p_mv = ptr;
    else if (id < 10) m.p_desc[id - 7] = ptr;
    else m.p_rec[id - 10] = ptr;
    if (hi <= lo) return k;
    k.src_off = a0;
    k.dst_off = slot_off + region;
    k.bytes = a1 - a0;
    return k;
}

/* Reference rows.  State of one plane class (luma, or U and V together) of the producer: */
struct SwRingState
{
    int loaded_hi;    /* rows [.., loaded_hi) have been requested */
};

/* Rows band `band` adds for plane class pc (0 luma, 1 chroma): [r0, r1), empty if none. */
RC_HD void sw_ring_new_rows(const SweepCtl &c, int f, int pc, int band, const SwRingState &st, int &r0, int &r1)
{
    const int lo = pc ? c.lo_c[f][band] : c.lo_y[f][band], hi = pc ? c.hi_c[f][band] : c.hi_y[f][band];
    r0 = r1 = 0;
    if (hi <= lo) return;
    r0 = st.loaded_hi > lo ? st.loaded_hi : lo;      /* rows below lo are needed by no band from here on */
    r1 = hi > r0 ? hi : r0;
}

/* may band `band` be requested while band `oldest` is the oldest one not yet retired? (its rows must not overwrite
   rows the unretired bands still read) */
RC_HD int sw_ring_fits(const SweepCtl &c, int f, int band, int oldest)
{
    if (c.hi_y[f][band] > c.lo_y[f][oldest] && c.hi_y[f][band] - c.lo_y[f][oldest] > c.cap_y) return 0;
    if (c.hi_c[f][band] > c.lo_c[f][oldest] && c.hi_c[f][band] - c.lo_c[f][oldest] > c.cap_c) return 0;
    return 1;
}

RC_HD int sw_ring_slot(int row, int first, int cap)
{
    const int d = row - first;
    return d % cap;      /* d >= 0: first is the smallest row of the sweep */
}

/* Copy `part` (0, 1: a run of rows splits in two where the ring wraps) of rows [r0, r1) of plane p into its ring. */
RC_HD SwCopy sw_ring_copy(const SweepGeom &g, const SweepCtl &c, int p, int r0, int r1, int part)
{
    SwCopy k = {SW_SRC_REF, 0, 0, 0};
    if (r1 <= r0) return k;
    const int pc = p ? 1 : 0;
    const int cap = pc ? c.cap_c : c.cap_y, first = pc ? c.first_c : c.first_y;
    const uint32_t pitch = (uint32_t)(g.width >> pc);
    const uint32_t plane_off = p == 0 ? 0u : p == 1 ? (uint32_t)(g.width * g.height) : (uint32_t)(g.width * g.height + (g.width >> 1) * (g.height >> 1));
    const int s0 = sw_ring_slot(r0, first, cap);
    const int n = r1 - r0, n0 = n < cap - s0 ? n : cap - s0;     /* rows before the wrap */
    if (part == 0)
    {
        k.src_off = plane_off + (uint32_t)r0 * pitch;
        k.dst_off = c.ring_off[p] + (uint32_t)s0 * pitch;
        k.bytes = (uint32_t)n0 * pitch;
    }
    else if (n > n0)
    {
        k.src_off = plane_off + (uint32_t)(r0 + n0) * pitch;
        k.dst_off = c.ring_off[p];
        k.bytes = (uint32_t)(n - n0) * pitch;
    }
    return k;
}

/* ---- ring addressing for the block functions (recon_core.h row / window policies) ------------- */
struct RcRingRows
{
    uint32_t ring;     /* shared-memory offset of the ring */
    uint32_t off0;     /* offset inside the ring of row 0's first aligned word (wrapped) */
    uint32_t pitch;    /* bytes between the patch's rows */
    uint32_t bytes;    /* ring size */
    RC_HDM uint32_t ld(int r, int k) const
    {
        const uint32_t o = off0 + (uint32_t)r * pitch, w = o - bytes;     /* o < 2 * bytes; w wraps around when o < bytes */
        return *reinterpret_cast<const uint32_t *>(SW_SMEM(ring + (w < o ? w : o) + 4u * (uint32_t)k));
    }
};
struct RcRingWindow
{
    uint32_t ring, origin, width, bytes;    /* origin: offset inside the ring of the window's first sample (wrapped) */
    RC_HDM RcRingRows rows(int ox, int oy, int ys, uint32_t &a) const
    {
        const uint32_t o = origin + (uint32_t)(oy * (int)width + ox), w = o - bytes;
        const uint32_t oo = w < o ? w : o;
        a = oo & 3u;      /* ring rows are multiples of 16 bytes: same misalignment as in the frame */
        return RcRingRows{ring, oo - a, (uint32_t)ys * width, bytes};
    }
};

/* everything a lane needs to know about the band it works on */
struct SweepBand
{
    const SweepGeom *g;
    const ReconView *v;
    const SweepCtl *c;
    const SweepSlotMeta *m;
    uint32_t slot_off;
    int band, mode;
    const uint8_t *scratch;     /* the CTA's list of finished `future` macroblocks (global memory) */
};

/* offset inside ring p of sample (x, row) */
RC_HD uint32_t sw_ring_at(const SweepBand &b, int p, int x, int row)
{
    const SweepSlotMeta &m = *b.m;
    const int pc = p ? 1 : 0;
    const int cap = pc ? b.c->cap_c : b.c->cap_y;
    int s = (pc ? m.slot0_c : m.slot0_y) + row - (pc ? m.row0_c : m.row0_y);
    s = s >= cap ? s - cap : s;
    return (uint32_t)s * (uint32_t)(b.g->width >> pc) + (uint32_t)x;
}

/* half-sample prediction of the 4x4 block at (bx, by) of plane p of a macroblock with vector word mvw (h4m:1327-1355) */
RC_HD void sw_predict_block(const SweepBand &b, int p, int bx, int by, uint32_t mvw, uint32_t rows[4])
{
    const ReconView &v = *b.v;
    const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
    const bool poison = rx == -32768;
    const int sh = p ? 1 : 0;
    const int px = rx >> sh, py = ry >> sh;
    const int hx = poison ? 0 : (v.version15 ? px : rx) & 1, hy = poison ? 0 : (v.version15 ? py : ry) & 1;
#if defined(__CUDA_ARCH__)
    const bool any_diag = __any_sync(__activemask(), hx & hy);
#else
    const bool any_diag = hx & hy;
#endif
    if (poison) { rows[0] = rows[1] = rows[2] = rows[3] = 0x80808080u; return; }   /* SYM_ERR_MV_RANGE: grey */
    const int x = (px >> 1) + (p ? 0 : (bx & 1) * 4), y = (py >> 1) + (p ? 0 : (by & 1) * 4);
    const int pc = p ? 1 : 0;
    const uint32_t o = sw_ring_at(b, p, x, y), a = o & 3u;
    const RcRingRows rr = {b.c->ring_off[p], o - a, (uint32_t)(v.width >> pc), b.c->ring_bytes[pc]};
    uint32_t W[10];
    rc_predict_load_rows<false>(W, rr, a, hx, hy);
    rc_predict_filter(rows, W, a, hx, hy, any_diag);
}

/* where a finished block goes: the band's tile (picture layout: 8h luma rows, then 4h rows of U, of V), or, in the
   `future` sweep, the macroblock's 96 bytes of the band's scratch list */
RC_HD void sw_store_block(const SweepBand &b, int p, int bx, int lrow, const uint32_t rows[4])
{
    const SweepGeom &g = *b.g;
    uint8_t *tile = SW_SMEM(b.slot_off + g.s_tile);
    if (b.mode == SW_MODE_FUTURE)
    {
        const int mx = p ? bx : bx >> 1, lmy = p ? lrow : lrow >> 1;
        const uint32_t rank = reinterpret_cast<const uint16_t *>(SW_SMEM(b.slot_off + g.s_rank))[lmy * g.mcb_w + mx];
        const uint32_t unit = p ? 3u + (uint32_t)p : (uint32_t)((lrow & 1) * 2 + (bx & 1));
        uint4 *dst = reinterpret_cast<uint4 *>(tile + rank * SW_MCB_BYTES + unit * 16u);
        *dst = make_uint4(rows[0], rows[1], rows[2], rows[3]);
        return;
    }
    const uint32_t pitch = (uint32_t)(g.width >> (p ? 1 : 0));
    const uint32_t plane_off = p == 0 ? 0u : g.tile_y_bytes + (p == 2 ? g.tile_c_bytes : 0u);
    uint8_t *dst = tile + plane_off + (uint32_t)(lrow * 4) * pitch + (uint32_t)bx * 4u;
#pragma unroll
    for (int r = 0; r < 4; ++r) *reinterpret_cast<uint32_t *>(dst + (uint32_t)r * pitch) = rows[r];
}

/* ---- map task: 32 consecutive blocks of one block row of the band ------------------------------
 * Every block that needs nothing but the maps and the reference window: weighted DC fill (h4m:299-383), flat fill
 * (h4m:281), motion compensation of proc-1 macroblocks and of nibble-0 blocks (h4m:1327-1355, 1886-1901); in the
 * merge sweep of a B picture also the copy of the macroblocks the `future` sweep finished. */
RC_HD void sw_map_lane(const SweepBand &b, int task, int lane)
{
    const SweepGeom &g = *b.g;
    const ReconView &v = *b.v;
    const SweepSlotMeta &m = *b.m;
    const int rows_mcb = m.rows;
    const int nY = 2 * rows_mcb * g.seg[0], nC = rows_mcb * g.seg[1];
    int p, lrow, sg;
    if (task < nY) { p = 0; lrow = task / g.seg[0]; sg = task - lrow * g.seg[0]; }
    else
    {
        int t = task - nY;
        p = 1;
        if (t >= nC) { p = 2; t -= nC; }
        lrow = t / g.seg[1];
        sg = t - lrow * g.seg[1];
    }
    const int bx = sg * 32 + lane;
    const bool in_row = bx < g.bw[p];
    const int stride = g.stride[p];
    const uint8_t *tcell = SW_SMEM(m.p_type[p]) + (lrow + 1) * stride + bx + 1;
    const uint8_t *dcell = SW_SMEM(m.p_dc[p]) + (lrow + 1) * stride + bx + 1;
    const uint32_t t = in_row ? *tcell : 6u;                  /* lanes past the row end: a raw block is nothing to do here */
    const bool ipic = v.is_ipic != 0;
    const uint32_t nib = ipic ? t : (t & 0xF);
    const uint32_t ref = ipic ? 0u : (t >> 5) & 3u;
    enum { kNone, kWeighted, kFlat, kMc, kCopy };
    int what = kNone;
    if (ref)
    {
        const bool mine = b.mode == SW_MODE_FUTURE ? ref == 2 : b.mode == SW_MODE_MERGE ? ref == 1 : true;
        if (!mine) what = b.mode == SW_MODE_MERGE ? kCopy : kNone;
        else if ((t & 0x10) || nib == 0) what = kMc;
    }
    else if (b.mode != SW_MODE_FUTURE)
        what = nib == 0 ? kWeighted : nib == 8 ? kFlat : kNone;
    const int lmy = p ? lrow : lrow >> 1, mx = p ? bx : bx >> 1;
    uint32_t rows[4];
    /* motion compensation first and by itself: the lanes of a warp vote on the filter form */
#if defined(__CUDA_ARCH__)
    if (__any_sync(0xFFFFFFFFu, what == kMc))
#endif
    {
        if (what == kMc)
        {
            const uint32_t mvw = reinterpret_cast<const uint32_t *>(SW_SMEM(m.p_mv))[lmy * g.mcb_w + mx];
            sw_predict_block(b, p, bx, p ? b.band * g.h + lrow : 2 * b.band * g.h + lrow, mvw, rows);
        }
    }
    if (what == kWeighted) rc_weighted_at(tcell, dcell, stride, ipic, rows);
    else if (what == kFlat) rows[0] = rows[1] = rows[2] = rows[3] = (uint32_t)*dcell * 0x01010101u;
    else if (what == kCopy)
    {
        const uint32_t rank = reinterpret_cast<const uint16_t *>(SW_SMEM(b.slot_off + g.s_rank))[lmy * g.mcb_w + mx];
        const uint32_t unit = p ? 3u + (uint32_t)p : (uint32_t)((lrow & 1) * 2 + (bx & 1));
        const uint8_t *src = b.scratch + (size_t)(m.side_off + rank) * SW_MCB_BYTES + unit * 16u;
#if defined(__CUDA_ARCH__)
        const uint4 q = __ldcg(reinterpret_cast<const uint4 *>(src));     /* written by this CTA's bulk stores: not through L1 */
#else
        const uint4 q = *reinterpret_cast<const uint4 *>(src);
#endif
        rows[0] = q.x; rows[1] = q.y; rows[2] = q.z; rows[3] = q.w;
    }
    if (what != kNone) sw_store_block(b, p, bx, lrow, rows);
}

/* ---- record task: one record per lane (raw block h4m:543-549, intra AOT h4m:1358-1377, predicted AOT
 * h4m:1379-1420 including its motion-compensated prediction) -------------------------------------------- */
RC_HD void sw_record_lane(const SweepBand &b, int cls, uint32_t idx)
{
    const SweepGeom &g = *b.g;
    const ReconView &v = *b.v;
    const SweepSlotMeta &m = *b.m;
    const bool live = idx < m.n_rec[cls];
    /* record idx of the (class, band) range: the chunks are ordered by length, each holds count records of one length */
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
    const uint32_t *rec = reinterpret_cast<const uint32_t *>(SW_SMEM(m.p_rec[cls])) + (first - m.rec_lo[cls]) + i * len;
    uint32_t t = 0;
    int p = 0, bx = 0, by = 0;
    if (live) rc_record_coords(rec[0], t, p, bx, by);
    const uint32_t ref = v.is_ipic ? 0u : (t >> 5) & 3u;
    const bool mine = live && (b.mode == SW_MODE_FUTURE ? ref == 2 : b.mode == SW_MODE_MERGE ? ref != 2 : true);
    const int lrow = by - (p ? b.band * g.h : 2 * b.band * g.h);
    const int lmy = p ? lrow : lrow >> 1, mx = p ? bx : bx >> 1;
    uint32_t rows[4];
    if (cls == SYM_REC_INTER)
    {
        const uint32_t mvw = mine ? reinterpret_cast<const uint32_t *>(SW_SMEM(m.p_mv))[lmy * g.mcb_w + mx] : 0x80008000u;
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        sw_predict_block(b, p, bx, by, mvw, rows);       /* every lane: the filter form is a warp vote */
        if (!mine) return;
        const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
        if (rx != -32768)
        {   /* window origin, h4m:1864-1868 */
            const uint32_t o = sw_ring_at(b, 0, rx / 2 - 32, ry / 2 - 16);
            const RcRingWindow win = {b.c->ring_off[0], o, (uint32_t)v.width, b.c->ring_bytes[0]};
            rc_predicted_aot(v, rows, rec + 1, (int)len - 1, win);
        }
    }
    else
    {
        if (!mine) return;
        if (cls == SYM_REC_RAW)
        {
#pragma unroll
            for (int r = 0; r < 4; ++r) rows[r] = rec[1 + r];
        }
        else
        {
            const int V = *(SW_SMEM(m.p_dc[p]) + (lrow + 1) * g.stride[p] + bx + 1);
            rc_intra_aot(v, rows, rec + 1, (int)len - 1, V);
        }
    }
    sw_store_block(b, p, bx, lrow, rows);
}


/* ---- what the producer publishes for a band once its symbol data has landed --------------------
 * Warp-collective on the GPU (32 lanes), one lane on the CPU.  f: reference of the sweep (0 past, 1 future). */
#if defined(__CUDA_ARCH__)
#define SW_POPC(x) __popc(x)
#define SW_LANES 32
RC_HD uint32_t sw_lane_sum(uint32_t v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}
RC_HD uint32_t sw_ballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
RC_HD void sw_lane_sync() { __syncwarp(); }
#else
#define SW_POPC(x) __builtin_popcount(x)
#define SW_LANES 1
RC_HD uint32_t sw_lane_sum(uint32_t v) { return v; }
RC_HD uint32_t sw_ballot(bool p) { return p ? 1u : 0u; }
RC_HD void sw_lane_sync() { }
#endif

RC_HD void sw_prep_band(const SweepGeom &g, const ReconView &v, const SweepCtl &c, int band, int mode, int f, uint32_t slot_off,
                        SweepSlotMeta &m, int lane)
{
    (void)v;
    const int r0 = band * g.h, rows = sw_band_rows(g, band), r1 = r0 + rows;
    uint32_t n_rec[SYM_REC_CLASSES];
    for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
    {
        const uint32_t nd = c.bf[cls][r1] - c.bf[cls][r0];
        const uint2 *d = reinterpret_cast<const uint2 *>(SW_SMEM(m.p_desc[cls]));
        uint32_t sum = 0;
        for (uint32_t j = (uint32_t)lane; j < nd; j += SW_LANES) sum += d[j].y & 0xFF;
        n_rec[cls] = sw_lane_sum(sum);
        if (lane == 0)
        {
            m.n_rec[cls] = n_rec[cls];
            m.n_desc[cls] = nd;
            m.rec_lo[cls] = c.rec_off[cls][r0];
        }
    }
    /* rank of every `future` macroblock of the band among them (B pictures swept twice) */
    uint32_t n2 = 0;
    if (mode != SW_MODE_ALL)
    {
        uint16_t *rank = reinterpret_cast<uint16_t *>(SW_SMEM(slot_off + g.s_rank));
        for (int lmy = 0; lmy < rows; ++lmy)
            for (int mx0 = 0; mx0 < g.mcb_w; mx0 += SW_LANES)
            {
                const int mx = mx0 + lane;
                const bool fut = mx < g.mcb_w && ((*(SW_SMEM(m.p_type[0]) + (2 * lmy + 1) * g.stride[0] + 2 * mx + 1) >> 5) & 3) == 2;
                const uint32_t bal = sw_ballot(fut);
                if (fut) rank[lmy * g.mcb_w + mx] = (uint16_t)(n2 + (uint32_t)SW_POPC(bal & ((1u << lane) - 1u)));
                n2 += (uint32_t)SW_POPC(bal);
            }
    }
    if (lane == 0)
    {
        const bool idle = mode == SW_MODE_FUTURE && n2 == 0;      /* no macroblock of this band belongs to the sweep */
        const uint32_t t_inter = idle ? 0u : (n_rec[SYM_REC_INTER] + 31u) / 32u;
        const uint32_t t_intra = idle || mode == SW_MODE_FUTURE ? 0u : (n_rec[SYM_REC_INTRA] + 31u) / 32u;
        const uint32_t t_map = idle ? 0u : (uint32_t)(rows * 2 * (g.seg[0] + g.seg[1]));
        const uint32_t t_raw = idle ? 0u : (n_rec[SYM_REC_RAW] + 31u) / 32u;
        m.t_inter = t_inter;
        m.t_intra = t_inter + t_intra;
        m.t_map = m.t_intra + t_map;
        m.n_tasks = m.t_map + t_raw;
        m.ticket = 0;
        m.rows = rows;
        m.n2 = n2;
        m.side_off = c.side_off[band];
        m.row0_y = c.lo_y[f][band];
        m.row0_c = c.lo_c[f][band];
        m.slot0_y = c.hi_y[f][band] > c.lo_y[f][band] ? sw_ring_slot(m.row0_y, c.first_y, c.cap_y) : 0;
        m.slot0_c = c.hi_c[f][band] > c.lo_c[f][band] ? sw_ring_slot(m.row0_c, c.first_c, c.cap_c) : 0;
    }
    sw_lane_sync();
}

/* task t of a band -> what to do (the order puts the long tasks first) */
RC_HD void sw_run_task(const SweepBand &b, uint32_t t, int lane)
{
    const SweepSlotMeta &m = *b.m;
    if (t < m.t_inter) sw_record_lane(b, SYM_REC_INTER, t * 32u + (uint32_t)lane);
    else if (t < m.t_intra) sw_record_lane(b, SYM_REC_INTRA, (t - m.t_inter) * 32u + (uint32_t)lane);
    else if (t < m.t_map) sw_map_lane(b, (int)(t - m.t_intra), lane);
    else sw_record_lane(b, SYM_REC_RAW, (t - m.t_map) * 32u + (uint32_t)lane);
}

#endif
