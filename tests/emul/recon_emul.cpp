/*
 * tests/emul/recon_emul.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Serial CPU driver for hvqm4_b200/csrc/recon_core.h: walks a symbol buffer in exactly
 * the work order of the CUDA kernel (segments of 16 macroblocks; upper luma block row,
 * lower luma block row, U, V; running prefix sum of sym_side_words) and calls the same
 * __host__ __device__ block functions.  It exists so that the host stage (entropy.c)
 * and the block arithmetic can be checked against the oracle in a container without a
 * GPU.  It is never built into, or reachable from, the product library.
 */
#include <cstdint>
#include <cstring>
#include "../../hvqm4_b200/csrc/recon_core.h"

static int32_t g_div[16], g_mcdiv[512];
static bool g_init;

static void store_rows(uint8_t *plane, int pw, int bx, int by, const uint32_t rows[4])
{
    for (int r = 0; r < 4; ++r) memcpy(plane + (by * 4 + r) * pw + bx * 4, &rows[r], 4);
}

extern "C" __attribute__((visibility("default")))
int emul_recon_picture(const uint8_t *blob, uint8_t *present, const uint8_t *past, const uint8_t *future)
{
    if (!g_init)
    {
        for (int i = 1; i < 16; ++i) g_div[i] = 0x1000 / (i * 16) * 16;
        for (int i = 1; i < 512; ++i) g_mcdiv[i] = 0x1000 / i;
        g_init = true;
    }
    SymHeader h;
    memcpy(&h, blob, sizeof h);
    if (h.magic != SYM_MAGIC) return -1;
    static uint32_t nest_tab[RC_NEST_TABLE_WORDS];
    if (h.has_nest)
        for (int y = 0; y < SYM_NEST_H; ++y)
            for (int x = 0; x < 64; ++x) nest_tab[y * 64 + x] = rc_nest_table_entry(blob + h.off_nest, y, x);
    ReconView v;
    rc_make_view(v, blob, h, nest_tab, g_div, g_mcdiv, past, future);
    const uint32_t *seg = (const uint32_t *)(blob + h.off_seg);
    const uint32_t *side = (const uint32_t *)(blob + h.off_side);
    uint8_t *planes[3] = {present, present + h.width * h.height, present + h.width * h.height * 5 / 4};
    const int is_i = h.pic_type == SYM_PIC_I;
    for (int row = 0; row < h.mcb_h; ++row)
        for (int sg = 0; sg < h.nseg; ++sg)
        {
            uint32_t word = seg[row * h.nseg + sg];
            const int mx0 = sg * SYM_SEG_MCBS;
            for (int pass = 0; pass < 3; ++pass)
                for (int lane = 0; lane < 32; ++lane)
                {
                    int plane, bx, by;
                    bool valid;
                    if (pass < 2) { plane = 0; bx = mx0 * 2 + lane; by = row * 2 + pass; valid = bx < h.mcb_w * 2; }
                    else { plane = 1 + (lane >> 4); bx = mx0 + (lane & 15); by = row; valid = bx < h.mcb_w; }
                    if (!valid) continue;
                    const int pw = h.width >> (plane ? 1 : 0);
                    const int bstride = (pw >> 2) + 2;
                    const uint32_t t = blob[h.off_type[plane] + (by + 1) * bstride + bx + 1];
                    uint32_t rows[4];
                    rc_block(v, plane, bx, by, t, side + word, rows);
                    word += sym_side_words(t, is_i);
                    store_rows(planes[plane], pw, bx, by, rows);
                }
            if (word != seg[row * h.nseg + sg + 1]) return -2;   /* segment table and prefix sum must agree */
        }
    return 0;
}
