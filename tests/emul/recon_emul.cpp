/*
 * tests/emul/recon_emul.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Serial CPU driver for hvqm4_b200/csrc/recon_core.h: runs the work of the CUDA kernels band by
 * band -- MAP phase (every block of the band from the type/DC maps: weighted, flat, motion
 * compensation), then RECORD phase (band table -> chunk table -> grouped records: raw, intra AOT,
 * predicted AOT on top of the prediction written by the map phase) -- calling the same
 * __host__ __device__ block functions, and checks that the band and chunk tables cover every
 * record exactly once.  It exists so that the host stage (entropy.c) and the
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
    {   /* landscape: 38 rows of 35 packed bytes, 68 entries per row; portrait: 70 rows of 19 bytes, 36 entries per row */
        const int rows = h.portrait ? SYM_NEST_W : SYM_NEST_H, row_bytes = h.portrait ? SYM_NEST_H / 2 : SYM_NEST_ROW_BYTES;
        const int pitch = h.portrait ? RC_NEST_PITCH_PORTRAIT : RC_NEST_PITCH;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < pitch; ++x)
                nest_tab[y * pitch + x] = rc_nest_spread_step1(rc_nest_table_entry(blob + h.off_nest, y, x, row_bytes));
    }
    ReconView v;
    rc_make_view(v, blob, h, nest_tab, g_div, g_mcdiv, past, future);
    uint8_t *planes[3] = {present, present + h.width * h.height, present + h.width * h.height * 5 / 4};

    /* band by band, like the fused CUDA kernel: the band's map blocks, then the records of the band
       (the chunks of every (class, band) pair are contiguous; the band table gives their start) */
    uint32_t seen = 0, chunks_seen = 0;
    /* bands of 8 macroblock rows, or of one (streams made for the sweep and row kernels) */
    const int band_rows = (h.n_bands == h.mcb_h && h.mcb_h > 1) ? 1 : SYM_BAND_MCB_ROWS;
    if (h.n_bands != (uint32_t)((h.mcb_h + band_rows - 1) / band_rows)) return -6;
    for (uint32_t band = 0; band < h.n_bands; ++band)
    {
        const int my0 = (int)band * band_rows, my1 = my0 + band_rows < h.mcb_h ? my0 + band_rows : h.mcb_h;
        for (int plane = 0; plane < 3; ++plane)
        {
            const int pw = h.width >> (plane ? 1 : 0);
            const int bstride = (pw >> 2) + 2;
            const int by0 = plane ? my0 : my0 * 2, by1 = plane ? my1 : my1 * 2;
            for (int by = by0; by < by1; ++by)
                for (int bx = 0; bx < pw / 4; ++bx)
                {
                    const uint32_t t = blob[h.off_type[plane] + (by + 1) * bstride + bx + 1];
                    uint32_t rows[4];
                    if (!rc_map_block(v, plane, bx, by, t, rows)) continue;
                    for (int r = 0; r < 4; ++r) memcpy(planes[plane] + (by * 4 + r) * pw + bx * 4, &rows[r], 4);
                }
        }
        for (int cls = 0; cls < SYM_REC_CLASSES; ++cls)
        {
            const uint32_t c0 = v.bands[cls * (h.n_bands + 1) + band], c1 = v.bands[cls * (h.n_bands + 1) + band + 1];
            if (c0 > c1 || c1 > h.n_chunks) return -7;
            for (uint32_t c = c0; c < c1; ++c, ++chunks_seen)
            {
                const uint32_t first = v.chunks[2 * c], desc = v.chunks[2 * c + 1];
                const uint32_t count = desc & 0xFF, len = ((desc >> 8) & 0xFF) + 1;
                if ((int)((desc >> 16) & 0xFF) != cls) return -8;
                if (count == 0 || count > SYM_CHUNK) return -3;
                if ((c < h.n_chunks_nest) != (cls != SYM_REC_INTER)) return -4;
                for (uint32_t i = 0; i < count; ++i)
                {
                    const uint32_t *rec = v.rec + first + i * len;
                    if (first + (i + 1) * len > h.n_rec_words) return -5;
                    uint32_t t;
                    int plane, bx, by;
                    rc_record_coords(rec[0], t, plane, bx, by);
                    if ((uint32_t)((plane ? by : by >> 1) / band_rows) != band) return -9;
                    const int pw = h.width >> (plane ? 1 : 0);
                    uint8_t *dst = planes[plane] + (by * 4) * pw + bx * 4;
                    uint32_t rows[4];
                    for (int r = 0; r < 4; ++r) memcpy(&rows[r], dst + r * pw, 4);
                    rc_record_block(v, cls, len, rec, rows);
                    for (int r = 0; r < 4; ++r) memcpy(dst + r * pw, &rows[r], 4);
                    ++seen;
                }
            }
        }
    }
    if (chunks_seen != h.n_chunks) return -10;
    if (seen != h.n_records) return -2;   /* chunk table must cover every record exactly once */
    return 0;
}

/* ---- leaf operators of recon_core.h, exported so that tests can sweep them against the
   reference's leaf functions (oracle/ref_wrap.c: ref_WeightImBlock, ref_MotionComp4x4) ---- */
extern "C" __attribute__((visibility("default")))
void emul_weighted(uint8_t *dst16, int V, int T, int B, int L, int R)
{
    uint32_t rows[4];
    rc_weighted(rows, V, T, B, L, R);
    memcpy(dst16, rows, 16);
}

extern "C" __attribute__((visibility("default")))
void emul_predict(uint8_t *dst16, const uint8_t *src, int stride, int hx, int hy)
{
    uint32_t rows[4];
    rc_predict(rows, src, stride, hx, hy);
    memcpy(dst16, rows, 16);
}

/* the warp-uniform "some lane needs both half steps" formulation, forced for any phase */
extern "C" __attribute__((visibility("default")))
void emul_predict_diag(uint8_t *dst16, const uint8_t *src, int stride, int hx, int hy)
{
    uint32_t rows[4], W[10];
    rc_predict_load<true>(W, src, stride, hx, hy);
    rc_predict_filter(rows, W, (uint32_t)((uintptr_t)src & 3), hx, hy, true);
    memcpy(dst16, rows, 16);
}
