/*
 * tests/emul/row_emul.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Serial CPU driver for hvqm4_b200/csrc/row_core.h: the row kernel's slot layout, the symbol slices of a row,
 * the ring allocation and the fetch tasks (rows requested as far ahead as slots and ring allow, so that a patch or a
 * slot overwritten too early shows up as wrong pixels), the class lists and every task of every row, lane by
 * lane, with memcpy in place of the bulk and tensor copies (boxes filled with zeros outside the plane, like the
 * TMA unit does).  row.cu adds only the copies and the mbarrier pipeline around the same functions.  Never built
 * into, or reachable from, the product library.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../hvqm4_b200/csrc/row_core.h"

extern uint8_t *sw_host_smem;

static int32_t g_div[16], g_mcdiv[512];

namespace {

/* what cp.async.bulk.tensor does: a box of bw x bh bytes at (x, y) of a plane, zeros outside */
void box_copy(uint8_t *dst, const uint8_t *plane, int pw, int ph, int x, int y, int bw, int bh)
{
    for (int r = 0; r < bh; ++r)
        for (int c = 0; c < bw; ++c)
        {
            const int xx = x + c, yy = y + r;
            dst[r * bw + c] = (xx >= 0 && xx < pw && yy >= 0 && yy < ph) ? plane[yy * pw + xx] : 0;
        }
}

}  // namespace

/* rows [r0, r1) of one picture; returns 0 = reconstructed, 1 = left to the band kernel, < 0 = error */
extern "C" __attribute__((visibility("default")))
int emul_row_picture(const uint8_t *blob, uint8_t *present, const uint8_t *past, const uint8_t *future, int smem_limit, int lookahead_limit)
{
    for (int i = 1; i < 16; ++i) g_div[i] = 0x1000 / (i * 16) * 16;
    for (int i = 1; i < 512; ++i) g_mcdiv[i] = 0x1000 / i;
    SymHeader hd;
    memcpy(&hd, blob, sizeof hd);
    if (hd.magic != SYM_MAGIC) return -1;
    RowGeom g;
    if (!rw_make_geom(g, hd.width, hd.height, (uint32_t)smem_limit)) return 1;
    std::vector<uint8_t> smem(g.smem_bytes + 64, 0xCD);
    sw_host_smem = smem.data();
    static uint32_t nest_tab[RC_NEST_TABLE_WORDS];
    if (hd.has_nest)
        for (int y = 0; y < SYM_NEST_H; ++y)
            for (int x = 0; x < RC_NEST_PITCH; ++x)
                nest_tab[y * RC_NEST_PITCH + x] = rc_nest_spread_step1(rc_nest_table_entry(blob + hd.off_nest, y, x));
    ReconView v;
    rc_make_view(v, blob, hd, nest_tab, g_div, g_mcdiv, past, future);
    RowCtl &c = *reinterpret_cast<RowCtl *>(sw_host_smem + g.off_ctl);
    memset(&c, 0, sizeof c);
    if ((int)hd.n_bands != g.mcb_h) return -2;
    for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
        for (int r = 0; r <= g.mcb_h; ++r)
        {
            c.bf[cls][r] = v.bands[cls * (hd.n_bands + 1) + r];
            const uint32_t ci = c.bf[cls][r];
            c.rec_off[cls][r] = ci < hd.n_chunks ? v.chunks[2 * ci] : hd.n_rec_words;
        }
    c.is_bpic = hd.pic_type == SYM_PIC_B;
    c.z[0] = 0;
    c.z[1] = 1;
    for (int id = 0; id < 7; ++id) rw_sym_table(g, v, c, id);
    for (int r = 0; r < g.mcb_h; ++r)
        if (!rw_row_fits(c, r)) return 1;
    const int n = g.mcb_h;
    const int pw[3] = {g.width, g.width / 2, g.width / 2}, ph[3] = {g.height, g.height / 2, g.height / 2};
    const size_t plane_off[3] = {0, (size_t)g.width * g.height, (size_t)g.width * g.height + (size_t)pw[1] * ph[1]};
    int ki = 0, ks = 0;
    RwRing ring = {0};
    uint32_t tail = 0;
    while (ks < n)
    {
        /* request warp + fetch tasks: rows as far ahead as slots, ring and the look-ahead limit allow */
        while (ki < n && ki < ks + g.n_slots && ki < ks + lookahead_limit)
        {
            const uint32_t slot_off = g.off_slot0 + (uint32_t)(ki % g.n_slots) * g.slot_bytes;
            RowSlotMeta &m = *reinterpret_cast<RowSlotMeta *>(sw_host_smem + slot_off + g.s_meta);
            for (int id = 0; id < RW_N_SYM_COPIES; ++id)
            {
                const SwCopy k = rw_sym_copy(g, v, c, ki, id, slot_off, m);
                if (k.bytes) memcpy(sw_host_smem + k.dst_off, blob + k.src_off, k.bytes);
            }
            const uint32_t n_inter = rw_count_inter(g, v, m, 0);
            const uint32_t live = (uint32_t)(ki - ks);
            RwRing trial = ring;
            uint32_t ttail = tail;
            if (!live) trial.head = ttail = 0;
            const uint32_t pos = rw_ring_alloc(trial, g.ring_bytes, n_inter * RW_PATCH_BYTES, live, ttail);
            if (pos == 0xFFFFFFFFu)
            {
                if (ki == ks) return -22;     /* an empty ring must take any row */
                break;
            }
            ring = trial;
            tail = ttail;
            m.patch_base = g.off_ring + pos;
            m.n_patch = n_inter;
            m.ring_end = ring.head;
            m.ticket = m.fticket = m.patch_count = 0;
            m.n_list[0] = m.n_list[1] = m.n_list[2] = 0;
            uint32_t issued = 0;
            for (int grp = g.n_groups - 1; grp >= 0; --grp)       /* any order of the fetch tasks is valid */
                for (int lane = 0; lane < 32; ++lane)
                {
                    RwBox box;
                    issued += rw_fetch_group(g, v, c, ki, grp, slot_off, m, lane, box);
                    if (!box.dst) continue;
                    if (box.x & 15) return -24;                        /* the TMA unit faults on such a box */
                    if (box.dst & 127u) return -28;                    /* ... and on such a destination */
                    if (box.dst < m.patch_base || box.dst + RW_PATCH_C_OFF > m.patch_base + m.n_patch * RW_PATCH_BYTES) return -25;
                    const uint8_t *surf = box.z == 1 ? future : past;
                    uint8_t *dst = sw_host_smem + box.dst;
                    if (!box.chroma) box_copy(dst, surf + plane_off[0], pw[0], ph[0], box.x, box.y, RW_BOX_W, 9);
                    else
                    {
                        box_copy(dst, surf + plane_off[1], pw[1], ph[1], box.x, box.y, RW_BOX_W, 5);
                        box_copy(dst + RW_PATCH_C, surf + plane_off[2], pw[2], ph[2], box.x, box.y, RW_BOX_W, 5);
                    }
                }
            if (issued != m.patch_count || issued > n_inter) return -26;
            ++ki;
        }
        if (ki == ks) return -27;
        /* work warps: every task of row ks, lane by lane */
        const uint32_t slot_off = g.off_slot0 + (uint32_t)(ks % g.n_slots) * g.slot_bytes;
        const RowSlotMeta &m = *reinterpret_cast<const RowSlotMeta *>(sw_host_smem + slot_off + g.s_meta);
        const RowWork w = {&g, &v, &m, slot_off};
        uint32_t t_end[RW_TASK_CLASSES];
        rw_task_ends(m, t_end);
        for (uint32_t t = 0; t < t_end[RW_TASK_CLASSES - 1]; ++t)
            for (int lane = 0; lane < 32; ++lane) rw_run_task(w, t_end, t, lane);
        /* retire */
        const uint8_t *tile = sw_host_smem + slot_off + g.s_tile;
        memcpy(present + (size_t)ks * g.tile_y_bytes, tile, g.tile_y_bytes);
        memcpy(present + plane_off[1] + (size_t)ks * g.tile_c_bytes, tile + g.tile_y_bytes, g.tile_c_bytes);
        memcpy(present + plane_off[2] + (size_t)ks * g.tile_c_bytes, tile + g.tile_y_bytes + g.tile_c_bytes, g.tile_c_bytes);
        /* poison what the row held so that stale data cannot go unnoticed */
        if (m.n_patch) memset(sw_host_smem + m.patch_base, 0xCD, m.n_patch * RW_PATCH_BYTES);
        tail = m.ring_end;
        memset(sw_host_smem + slot_off, 0xCD, g.slot_bytes);
        ++ks;
    }
    const int bad_any = c.unsupported;
    sw_host_smem = nullptr;
    return bad_any ? 1 : 0;
}
