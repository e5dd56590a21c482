/*
 * tests/emul/recon_emul.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Serial CPU driver for hvqm4_b200/csrc/recon_core.h: runs the work of the two CUDA kernels
 * in the same order of phases -- MAP phase (every block from the type/DC maps: weighted, flat,
 * motion compensation), then RECORD phase (chunk table -> grouped records: raw, intra AOT,
 * predicted AOT on top of the prediction written by the map phase) -- calling the same
 * __host__ __device__ block functions.  It exists so that the host stage (entropy.c) and the
 * block arithmetic can be checked against the oracle in a container without a GPU.  It is
 * never built into, or reachable from, the product library.
 */
#include <cstdint>
#include <cstring>
#include "../../hvqm4_b200/csrc/recon_core.h"

static int32_t g_div[16], g_mcdiv[512];
static bool g_init;

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
    uint8_t *planes[3] = {present, present + h.width * h.height, present + h.width * h.height * 5 / 4};

    /* MAP phase */
    for (int plane = 0; plane < 3; ++plane)
    {
        const int pw = h.width >> (plane ? 1 : 0), ph = h.height >> (plane ? 1 : 0);
        const int bstride = (pw >> 2) + 2;
        for (int by = 0; by < ph / 4; ++by)
            for (int bx = 0; bx < pw / 4; ++bx)
            {
                const uint32_t t = blob[h.off_type[plane] + (by + 1) * bstride + bx + 1];
                uint32_t rows[4];
                if (!rc_map_block(v, plane, bx, by, t, rows)) continue;
                for (int r = 0; r < 4; ++r) memcpy(planes[plane] + (by * 4 + r) * pw + bx * 4, &rows[r], 4);
            }
    }
    /* RECORD phase */
    uint32_t seen = 0;
    for (uint32_t c = 0; c < h.n_chunks; ++c)
    {
        const uint32_t first = v.chunks[2 * c], desc = v.chunks[2 * c + 1];
        const uint32_t count = desc & 0xFF, len = ((desc >> 8) & 0xFF) + 1;
        const int cls = (int)((desc >> 16) & 0xFF);
        if (count == 0 || count > SYM_CHUNK) return -3;
        if ((c < h.n_chunks_nest) != (cls != SYM_REC_INTER)) return -4;
        for (uint32_t i = 0; i < count; ++i)
        {
            const uint32_t *rec = v.rec + first + i * len;
            if (first + (i + 1) * len > h.n_rec_words) return -5;
            uint32_t t;
            int plane, bx, by;
            rc_record_coords(rec[0], t, plane, bx, by);
            const int pw = h.width >> (plane ? 1 : 0);
            uint8_t *dst = planes[plane] + (by * 4) * pw + bx * 4;
            uint32_t rows[4];
            for (int r = 0; r < 4; ++r) memcpy(&rows[r], dst + r * pw, 4);
            rc_record_block(v, cls, len, rec, rows);
            for (int r = 0; r < 4; ++r) memcpy(dst + r * pw, &rows[r], 4);
            ++seen;
        }
    }
    if (seen != h.n_records) return -2;   /* chunk table must cover every record exactly once */
    return 0;
}
