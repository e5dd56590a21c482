/*
 * tests/emul/sweep_emul.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Serial CPU driver for hvqm4_b200/csrc/sweep_core.h: the sweep kernel's per-picture plan, the
 * producer's copies (symbol slices into slots, reference rows into the rings -- requested as far
 * ahead as the kernel's rules allow, so that a ring row overwritten too early shows up as wrong
 * pixels), the per-band preparation and every task of every band, lane by lane, with memcpy in
 * place of the bulk asynchronous copies.  sweep.cu adds only the copies and the mbarrier
 * pipeline around the same functions.  Never built into, or reachable from, the product library.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../hvqm4_b200/csrc/sweep_core.h"

uint8_t *sw_host_smem;

static int32_t g_div[16], g_mcdiv[512];

namespace {

struct Picture
{
    const uint8_t *blob;
    uint8_t *present;
    const uint8_t *ref[2];
};

void do_copy(const SwCopy &k, const Picture &pic, int f, const uint8_t *scratch)
{
    if (!k.bytes) return;
    const uint8_t *src = k.src_kind == SW_SRC_BLOB ? pic.blob : k.src_kind == SW_SRC_REF ? pic.ref[f] : scratch;
    memcpy(sw_host_smem + k.dst_off, src + k.src_off, k.bytes);
}

/* one sweep over the picture; returns 0 or a negative code */
int run_sweep(const SweepGeom &g, const ReconView &v, SweepCtl &c, const Picture &pic, int mode, int f, uint8_t *scratch, int lookahead_limit)
{
    if (!sw_plan_rings(g, c, f)) return -20;
    int ki = 0, ks = 0;
    SwRingState st[2] = {{SW_EMPTY_HI}, {SW_EMPTY_HI}};
    const int nb = g.n_bands;
    while (ks < nb)
    {
        /* producer: request bands as far ahead as slots and rings allow */
        while (ki < nb && ki < ks + SW_NSLOTS && ki < ks + lookahead_limit && sw_ring_fits(c, f, ki, ks))
        {
            const uint32_t slot_off = g.off_slot0 + (uint32_t)(ki % SW_NSLOTS) * g.slot_bytes;
            SweepSlotMeta &m = *reinterpret_cast<SweepSlotMeta *>(sw_host_smem + slot_off + g.s_meta);
            for (int id = 0; id < SW_N_SYM_COPIES; ++id) do_copy(sw_sym_copy(g, v, c, ki, id, slot_off, m), pic, f, scratch);
            for (int pc = 0; pc < 2; ++pc)
            {
                int r0, r1;
                sw_ring_new_rows(c, f, pc, ki, st[pc], r0, r1);
                for (int p = pc ? 1 : 0; p <= (pc ? 2 : 0); ++p)
                    for (int part = 0; part < 2; ++part) do_copy(sw_ring_copy(g, c, p, r0, r1, part), pic, f, scratch);
                if (r1 > r0) st[pc].loaded_hi = r1;
            }
            sw_prep_band(g, v, c, ki, mode, f, slot_off, m, 0);
            if (mode != SW_MODE_ALL && m.n2 != c.n2[ki]) return -21;
            ++ki;
        }
        if (ki == ks) return -22;     /* the plan promised that an empty pipeline can always take the next band */
        /* consumers: every task of band ks, lane by lane */
        const uint32_t slot_off = g.off_slot0 + (uint32_t)(ks % SW_NSLOTS) * g.slot_bytes;
        const SweepSlotMeta &m = *reinterpret_cast<const SweepSlotMeta *>(sw_host_smem + slot_off + g.s_meta);
        const SweepBand sb = {&g, &v, &c, &m, slot_off, ks, mode, scratch};
        for (uint32_t t = 0; t < m.n_tasks; ++t)
            for (int lane = 0; lane < 32; ++lane) sw_run_task(sb, t, lane);
        /* store */
        const uint8_t *tile = sw_host_smem + slot_off + g.s_tile;
        if (mode == SW_MODE_FUTURE)
            memcpy(scratch + (size_t)m.side_off * SW_MCB_BYTES, tile, (size_t)m.n2 * SW_MCB_BYTES);
        else
        {
            const int rows = m.rows;
            const size_t wy = (size_t)g.width, wc = (size_t)g.width / 2;
            memcpy(pic.present + (size_t)ks * g.h * 8 * wy, tile, (size_t)rows * 8 * wy);
            uint8_t *u = pic.present + wy * g.height, *vv = u + wc * (g.height / 2);
            memcpy(u + (size_t)ks * g.h * 4 * wc, tile + g.tile_y_bytes, (size_t)rows * 4 * wc);
            memcpy(vv + (size_t)ks * g.h * 4 * wc, tile + g.tile_y_bytes + g.tile_c_bytes, (size_t)rows * 4 * wc);
        }
        /* poison the slot so that stale data cannot go unnoticed */
        memset(sw_host_smem + slot_off, 0xCD, g.slot_bytes);
        ++ks;
    }
    return 0;
}

}  // namespace

/* returns 0 = reconstructed, 1 = the plan leaves the picture to the band kernel, < 0 = error */
extern "C" __attribute__((visibility("default")))
int emul_sweep_picture(const uint8_t *blob, uint8_t *present, const uint8_t *past, const uint8_t *future, int h, int smem_limit, int lookahead_limit)
{
    for (int i = 1; i < 16; ++i) g_div[i] = 0x1000 / (i * 16) * 16;
    for (int i = 1; i < 512; ++i) g_mcdiv[i] = 0x1000 / i;
    SymHeader hd;
    memcpy(&hd, blob, sizeof hd);
    if (hd.magic != SYM_MAGIC) return -1;
    SweepGeom g;
    if (!sw_make_geom(g, hd.width, hd.height, h, (uint32_t)smem_limit)) return 1;
    std::vector<uint8_t> smem(g.smem_bytes + 64, 0xCD);
    sw_host_smem = smem.data();
    static uint32_t nest_tab[RC_NEST_TABLE_WORDS];
    if (hd.has_nest)
        for (int y = 0; y < SYM_NEST_H; ++y)
            for (int x = 0; x < RC_NEST_PITCH; ++x)
                nest_tab[y * RC_NEST_PITCH + x] = rc_nest_spread_step1(rc_nest_table_entry(blob + hd.off_nest, y, x));
    ReconView v;
    rc_make_view(v, blob, hd, nest_tab, g_div, g_mcdiv, past, future);
    SweepCtl &c = *reinterpret_cast<SweepCtl *>(sw_host_smem + g.off_ctl);
    memset(&c, 0, sizeof c);
    if ((int)hd.n_bands != g.mcb_h) return -2;
    /* band table, record offsets */
    for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
        for (int r = 0; r <= g.mcb_h; ++r) c.bf[cls][r] = v.bands[cls * (hd.n_bands + 1) + r];
    for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
        for (int r = 0; r <= g.mcb_h; ++r)
        {
            const uint32_t ci = c.bf[cls][r];
            c.rec_off[cls][r] = ci < hd.n_chunks ? v.chunks[2 * ci] : hd.n_rec_words;
        }
    /* prescan */
    int bad_any = 0;
    for (int b = 0; b < g.n_bands; ++b)
    {
        int lo[2][2], hi[2][2];
        for (int f = 0; f < 2; ++f)
            for (int pc = 0; pc < 2; ++pc) { lo[f][pc] = SW_EMPTY_LO; hi[f][pc] = SW_EMPTY_HI; }
        int n2 = 0;
        if (!v.is_ipic)
            for (int my = b * g.h; my < b * g.h + sw_band_rows(g, b); ++my)
                for (int mx = 0; mx < g.mcb_w; ++mx)
                {
                    int ref, ext[4], bad;
                    sw_prescan_mcb(v, mx, my, ref, ext, bad);
                    if (ref == 2 && hd.pic_type != SYM_PIC_B) bad = 1;
                    bad_any |= bad;
                    if (!ref || bad) continue;
                    n2 += ref == 2;
                    const int f = ref - 1;
                    if (ext[0] < lo[f][0]) lo[f][0] = ext[0];
                    if (ext[1] > hi[f][0]) hi[f][0] = ext[1];
                    if (ext[2] < lo[f][1]) lo[f][1] = ext[2];
                    if (ext[3] > hi[f][1]) hi[f][1] = ext[3];
                }
        for (int f = 0; f < 2; ++f)
        {
            c.lo_y[f][b] = (int16_t)lo[f][0]; c.hi_y[f][b] = (int16_t)hi[f][0];
            c.lo_c[f][b] = (int16_t)lo[f][1]; c.hi_c[f][b] = (int16_t)hi[f][1];
        }
        c.n2[b] = (uint16_t)n2;
    }
    if (bad_any) return 1;
    for (int f = 0; f < 2; ++f)
    {
        sw_plan_scan(c.lo_y[f], c.hi_y[f], g.n_bands);
        sw_plan_scan(c.lo_c[f], c.hi_c[f], g.n_bands);
    }
    c.side_off[0] = 0;
    for (int b = 0; b < g.n_bands; ++b) c.side_off[b + 1] = c.side_off[b] + c.n2[b];
    c.n_future = (int32_t)c.side_off[g.n_bands];
    for (int b = 0; b < g.n_bands; ++b)
        if (!sw_band_fits(g, c, b)) return 1;
    const bool two = c.n_future > 0;
    if (!sw_plan_rings(g, c, 0) || (two && !sw_plan_rings(g, c, 1))) return 1;
    const Picture pic = {blob, present, {past, future}};
    std::vector<uint8_t> scratch((size_t)g.mcb_w * g.mcb_h * SW_MCB_BYTES + 16, 0xEE);
    int rc = 0;
    if (two)
    {
        rc = run_sweep(g, v, c, pic, SW_MODE_FUTURE, 1, scratch.data(), lookahead_limit);
        if (rc == 0) rc = run_sweep(g, v, c, pic, SW_MODE_MERGE, 0, scratch.data(), lookahead_limit);
    }
    else
        rc = run_sweep(g, v, c, pic, SW_MODE_ALL, 0, scratch.data(), lookahead_limit);
    sw_host_smem = nullptr;
    return rc;
}
