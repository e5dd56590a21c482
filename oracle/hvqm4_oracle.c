/*
 * oracle/hvqm4_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C CPU restatement ("port") of the HVQM4 1.3/1.5 picture-decode path of the
 * reference decoder, written from scratch for this repository.  It is the checker
 * that travels with the repo (the GPU box has no /root/reference): the CUDA path
 * must reproduce its output bit for bit.
 *
 * Parity pin: this file is validated (tests/test_oracle.py) against
 *   (1) oracle/_ref/libhvqm4_ref.so -- the unmodified reference compiled from
 *       /root/reference/h4m_audio_decode.c by oracle/Makefile, frame by frame
 *       (planar YUV bytes, block maps and nest) on generated streams, and
 *   (2) tests/golden/ (JSON files) -- per-frame MD5s produced by that reference build
 *       (tests/golden/make_golden.py), so the pin also holds where the reference
 *       tree is absent.
 * The reference ships no golden vectors, fixtures or tests of its own
 * (SURVEY.md section 4), so running it on generated streams is the only pin there is.
 *
 * Every function cites the reference lines it restates ("h4m:N" =
 * /root/reference/h4m_audio_decode.c line N).
 *
 * Scope: 4:2:0 (h_samp = v_samp = 2) like every known stream, landscape and portrait (h4m:700-711, 743-754, 965-975,
 * 1865-1868; upstream calls portrait untested, README:23: pinned here against the reference build like everything else); other
 * sampling factors are rejected at open().
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>

#define PORT_API __attribute__((visibility("default")))

enum { PIC_I = 0x10, PIC_P = 0x20, PIC_B = 0x30 };   /* h4m:2065-2070 */
enum { NEST_W = 70, NEST_H = 38 };                    /* h4m:488, 966-970 */

/* ------------------------------------------------------------------ big-endian helpers (h4m:58-89) */

static uint32_t be16(const uint8_t *p) { return (uint32_t)p[0] << 8 | p[1]; }
static uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

/* ------------------------------------------------------------------ bit sections (h4m:552-602, 1061-1071)
 * The reference pulls MSB-first bits out of big-endian 32-bit words (1.5 ABI) or
 * bytes (1.3 ABI); both are the same plain MSB-first bit sequence, read here one
 * bit at a time.  A section is [BE32 size][bytes]; size 0 means "absent". */

typedef struct
{
    const uint8_t *base;   /* first payload byte, NULL if absent */
    uint32_t size;
    uint64_t bit;          /* next bit index (bit sections) */
    uint32_t byte;         /* next byte index (fixvl sections) */
} Section;

static void section_open(Section *s, const uint8_t *at)
{
    s->size = be32(at);
    s->base = s->size ? at + 4 : NULL;
    s->bit = 0;
    s->byte = 0;
}

static uint32_t take_bit(Section *s)
{
    uint32_t b = (s->base[s->bit >> 3] >> (7 - (s->bit & 7))) & 1;
    s->bit++;
    return b;
}

static uint32_t take_bits(Section *s, int n)
{
    uint32_t v = 0;
    while (n-- > 0) v = v << 1 | take_bit(s);
    return v;
}

/* ------------------------------------------------------------------ Huffman trees (h4m:385-394, 607-651)
 * Serialised pre-order: bit 1 = internal node (0-side subtree first), bit 0 + 8
 * bits = leaf.  Leaf value = byte, sign-extended when the tree is signed, shifted
 * left by the tree's scale.  The leaf table persists across pictures exactly like
 * the reference's Tree.array[0][0..255]; an empty leader section gives root 0,
 * i.e. every symbol decodes (without consuming bits) to the stale leaf[0]. */

typedef struct
{
    int32_t leaf[256];
    int16_t kid[2][256];   /* children of internal node n (stored as n - 256); < 256 = leaf byte */
    int32_t root;          /* < 256: leaf byte, >= 256: internal node */
    int32_t used;
} Tree;

static int tree_parse(Tree *t, Section *s, int is_signed, int scale)
{
    if (take_bit(s) == 0)
    {
        uint32_t byte = take_bits(s, 8);
        int32_t v = (is_signed && byte > 0x7F) ? (int32_t)byte - 256 : (int32_t)byte;
        /* int16_t symbol <<= scale in the reference (h4m:613-617); values stay in range */
        t->leaf[byte] = (int32_t)(int16_t)(v * (1 << scale));
        return (int)byte;
    }
    int node = t->used++;
    int a = tree_parse(t, s, is_signed, scale);
    int b = tree_parse(t, s, is_signed, scale);
    t->kid[0][node] = (int16_t)a;
    t->kid[1][node] = (int16_t)b;
    return node + 256;
}

static void tree_read(Tree *t, Section *leader, int is_signed, int scale)
{
    t->used = 0;
    t->root = leader->size ? tree_parse(t, leader, is_signed, scale) : 0;
}

static int32_t huff(const Tree *t, Section *s)
{
    int n = t->root;
    while (n >= 256) n = t->kid[take_bit(s)][n - 256];
    return t->leaf[n];
}

/* escape-extended symbols (h4m:654-677) */
static int32_t sym_signed_ovf(const Tree *t, Section *s, int32_t lo, int32_t hi)
{
    int32_t sum = 0, v;
    do { v = huff(t, s); sum += v; } while (v <= lo || v >= hi);
    return sum;
}
static int32_t sym_unsigned_ovf(const Tree *t, Section *s)
{
    int32_t sum = 0, v;
    do { v = huff(t, s); sum += v; } while (v >= 255);
    return sum;
}

/* ------------------------------------------------------------------ decoder state */

typedef struct { uint8_t dc, type; } Cell;   /* h4m:432-436 */

typedef struct
{
    int w, h;            /* samples */
    int bw, bh;          /* 4x4 blocks */
    int stride;          /* bw + 2 (bordered map, h4m:859-860) */
    int shift;           /* 1 for chroma (h4m:846-849) */
    Cell *map;           /* (bw+2)*(bh+2), border cells {0x7F,0xFF} (h4m:951-955) */
} Plane;

typedef struct
{
    int width, height, version15;
    Plane pl[3];
    Tree tree[6];        /* sharing as in h4m:977-999 */
    uint8_t nest[NEST_H * NEST_W];
    int32_t div_tab[16], mcdiv_tab[512];   /* h4m:262-273 */
    /* per picture */
    Section bn[2], bnr[2], dcv[3], sc[3], fix[3], rle[3], mvh, mvv, mcbt, mcbp;
    int dc_shift, unk_shift, rb[2][2];
    int32_t dc_lo, dc_hi;
    int last_type;
} Dec;

static Cell *cell(const Plane *p, int bx, int by) { return &p->map[(by + 1) * p->stride + bx + 1]; }

static void dec_init(Dec *d, int width, int height, int version15)
{
    memset(d, 0, sizeof *d);
    d->width = width; d->height = height; d->version15 = version15;
    for (int i = 0; i < 3; ++i)
    {
        Plane *p = &d->pl[i];
        p->shift = i ? 1 : 0;
        p->w = width >> p->shift; p->h = height >> p->shift;
        p->bw = p->w / 4; p->bh = p->h / 4;
        p->stride = p->bw + 2;
        size_t n = (size_t)p->stride * (p->bh + 2);
        p->map = calloc(n, sizeof(Cell));
        for (int y = 0; y < p->bh + 2; ++y)
            for (int x = 0; x < p->stride; ++x)
                if (y == 0 || y == p->bh + 1 || x == 0 || x == p->stride - 1)
                {
                    p->map[y * p->stride + x].dc = 0x7F;
                    p->map[y * p->stride + x].type = 0xFF;
                }
    }
    for (int i = 1; i < 16; ++i) d->div_tab[i] = 0x1000 / (i * 16) * 16;
    for (int i = 1; i < 512; ++i) d->mcdiv_tab[i] = 0x1000 / i;
}

static void dec_free(Dec *d)
{
    for (int i = 0; i < 3; ++i) free(d->pl[i].map);
}

static uint8_t clamp255(int32_t x) { return x < 0 ? 0 : x > 255 ? 255 : (uint8_t)x; }   /* h4m:288 */

/* ------------------------------------------------------------------ block reconstruction */

/* h4m:281 */
static void fill_flat(uint8_t *dst, int stride, uint8_t v)
{
    for (int y = 0; y < 4; ++y) memset(dst + y * stride, v, 4);
}

/* h4m:293-383.  The sixteen expressions of the reference collapse to a separable form:
 *   out(r,c) = sat_mean8(8V + rowterm[r] + colterm[c])
 *   rowterm = {2T-B-V, V-B, V-T, 2B-T-V}     colterm = {2L-R-V, V-R, V-L, 2R-L-V}
 * (e.g. r=0,c=0: 8V + vph + tpl = 6V+2T-B+2L-R, h4m:344,349; r=1,c=1: 8V - bpr = 10V-B-R,
 * h4m:357,362).  sat_mean8 divides (sum+4) by 8 as an UNSIGNED 32-bit number before
 * clamping (h4m:293-296): sums <= -5 wrap to a huge value -> 255, sums -4..-1 -> 0. */
static uint8_t sat_mean8_u32(int32_t sum)
{
    uint32_t q = ((uint32_t)sum + 4u) / 8u;
    return clamp255((int32_t)q);
}

static void fill_weighted(uint8_t *dst, int stride, int V, int T, int B, int L, int R)
{
    const int rowterm[4] = {2 * T - B - V, V - B, V - T, 2 * B - T - V};
    const int colterm[4] = {2 * L - R - V, V - R, V - L, 2 * R - L - V};
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            dst[r * stride + c] = sat_mean8_u32(8 * V + rowterm[r] + colterm[c]);
}

/* h4m:543-549: sixteen raw bytes, row-major, from the plane's fixvl byte stream */
static void fill_raw(Dec *d, int plane, uint8_t *dst, int stride)
{
    Section *s = &d->fix[plane];
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x)
            dst[y * stride + x] = s->base[s->byte++];
}

/* One AOT basis (h4m:679-732 intra, 734-773 inter).  `src` points at the origin of the
 * 70x38 nest: the I-picture nest (already 4-bit, src_shift 0) or a window of the
 * reference frame's luma (src_shift 4: sample = (pixel >> 4) & 0xF).  Descriptor bits:
 * [5:0] x offset, [10:6] y offset, [11] x step 2, [12] y step 2, [14:13] scale offset,
 * [15] negate.  Returns the 16 weighted samples added into acc[]. */
static void aot_add_basis(Dec *d, int plane, const uint8_t *src, int src_stride, int src_shift,
                          int32_t *scale_sum, int32_t acc[16])
{
    Section *fx = &d->fix[plane];
    uint32_t desc = be16(fx->base + fx->byte);
    fx->byte += 2;
    /* portrait pictures swap the axes: the 6-bit field is the row, the 5-bit field the column (h4m:700-711, 743-754) */
    const int portrait = d->width < d->height;
    const uint32_t f6 = desc & 0x3F, f5 = (desc >> 6) & 0x1F, s11 = (desc >> 11) & 1, s12 = (desc >> 12) & 1;
    const uint8_t *org = portrait ? src + src_stride * f6 + f5 : src + src_stride * f5 + f6;
    int xs = 1 << (portrait ? s12 : s11);
    int ys = src_stride << (portrait ? s11 : s12);
    uint8_t b[16];
    int lo = 255, hi = 0;
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x)
        {
            int v = (org[y * ys + x * xs] >> src_shift) & 0xF;
            b[y * 4 + x] = (uint8_t)v;
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
    *scale_sum += huff(&d->tree[2], &d->sc[plane]);          /* cumulative within the block, h4m:726,781 */
    int32_t inv = d->div_tab[hi - lo];
    if (desc & 0x8000) inv = -inv;
    uint32_t factor = (uint32_t)(*scale_sum + (int32_t)((desc >> 13) & 3)) * (uint32_t)inv;
    for (int i = 0; i < 16; ++i)
        acc[i] = (int32_t)((uint32_t)acc[i] + factor * b[i]);     /* mod 2^32, h4m:784-787 */
}

/* h4m:775-817: sum of n bases; returns the arithmetic-shift mean */
static int32_t aot_sum(Dec *d, int plane, int n, const uint8_t *src, int src_stride, int src_shift, int32_t acc[16])
{
    int32_t scale_sum = 0;
    memset(acc, 0, 16 * sizeof(int32_t));
    for (int k = 0; k < n; ++k) aot_add_basis(d, plane, src, src_stride, src_shift, &scale_sum, acc);
    uint32_t total = 0;
    for (int i = 0; i < 16; ++i) total += (uint32_t)acc[i];
    return (int32_t)total >> 4;
}

/* h4m:1358-1377 */
static void fill_intra_aot(Dec *d, int plane, uint8_t *dst, int stride, int dc, int type)
{
    if (type == 6) { fill_raw(d, plane, dst, stride); return; }
    int32_t acc[16];
    int32_t mean = aot_sum(d, plane, type, d->nest, d->width < d->height ? NEST_H : NEST_W, 0, acc);   /* nest 38 wide in portrait, h4m:965-975 */
    int32_t delta = (int32_t)((uint32_t)dc << d->unk_shift) - mean;
    for (int i = 0; i < 16; ++i)
        dst[(i >> 2) * stride + (i & 3)] = clamp255((acc[i] + delta) >> d->unk_shift);
}

/* h4m:1242-1294: 4x4 prediction at half-sample phase (hx,hy) */
static void predict4x4(uint8_t out[16], const uint8_t *src, int stride, int hx, int hy)
{
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x)
        {
            const uint8_t *p = src + y * stride + x;
            int v;
            if (!hx && !hy) v = p[0];
            else if (hx && !hy) v = (p[0] + p[1] + 1) / 2;
            else if (!hx && hy) v = (p[0] + p[stride] + 1) / 2;
            else v = (p[0] + p[1] + p[stride] + p[stride + 1] + 2) >> 2;
            out[y * 4 + x] = (uint8_t)v;
        }
}

/* h4m:1379-1420 */
static void fill_predicted_aot(Dec *d, int plane, uint8_t *dst, const uint8_t *src, int stride, int nibble,
                               const uint8_t *window, int window_stride, int hx, int hy)
{
    int32_t acc[16];
    uint32_t aot_mean = (uint32_t)aot_sum(d, plane, nibble - 1, window, window_stride, 4, acc);
    uint8_t m[16];
    predict4x4(m, src, stride, hx, hy);
    int32_t mean = 8;
    for (int i = 0; i < 16; ++i) mean += m[i];
    mean /= 16;
    int32_t diff[16], lo, hi;
    lo = hi = m[0] - mean;
    for (int i = 0; i < 16; ++i)
    {
        diff[i] = m[i] - mean;
        if (diff[i] < lo) lo = diff[i];
        if (diff[i] > hi) hi = diff[i];
    }
    int32_t s1 = sym_signed_ovf(&d->tree[0], &d->dcv[plane], d->dc_lo, d->dc_hi);
    int32_t s2 = sym_signed_ovf(&d->tree[0], &d->dcv[plane], d->dc_lo, d->dc_hi);
    uint32_t addend = ((uint32_t)(s1 >> d->dc_shift) << d->unk_shift) - aot_mean;
    uint32_t factor = (uint32_t)(s2 >> d->dc_shift) * (uint32_t)d->mcdiv_tab[hi - lo];
    for (int i = 0; i < 16; ++i)
    {
        int32_t r = (int32_t)((uint32_t)acc[i] + addend + (uint32_t)diff[i] * factor);
        dst[(i >> 2) * stride + (i & 3)] = clamp255((r >> d->unk_shift) + m[i]);
    }
}

/* neighbour rule shared by I pictures and intra macroblocks of P/B pictures
 * (h4m:1437-1441, 1811-1814): a neighbour's DC is used only if (type & 0x77) == 0. */
static int nbr_dc(const Cell *n, int own) { return (n->type & 0x77) ? own : n->dc; }

/* ------------------------------------------------------------------ I picture */

/* h4m:1073-1130 */
static void ipic_types(Dec *d)
{
    uint32_t run = 0;
    Plane *y = &d->pl[0];
    for (int by = 0; by < y->bh; ++by)
        for (int bx = 0; bx < y->bw; ++bx)
        {
            Cell *c = cell(y, bx, by);
            if (run) { c->type = 0; --run; continue; }
            int32_t n = huff(&d->tree[3], &d->bn[0]);
            if ((int16_t)n == 0) run = (uint32_t)huff(&d->tree[1], &d->bnr[0]);
            c->type = (uint8_t)n;
        }
    run = 0;
    Plane *u = &d->pl[1], *v = &d->pl[2];
    for (int by = 0; by < u->bh; ++by)
        for (int bx = 0; bx < u->bw; ++bx)
        {
            Cell *cu = cell(u, bx, by), *cv = cell(v, bx, by);
            if (run) { cu->type = cv->type = 0; --run; continue; }
            int32_t n = huff(&d->tree[3], &d->bn[1]);
            if ((int16_t)n == 0) run = (uint32_t)huff(&d->tree[1], &d->bnr[1]);
            cu->type = n & 0xF;
            cv->type = (n >> 4) & 0xF;
        }
}

/* h4m:1043-1058, 1132-1164: DPCM over the bordered DC map, 8-bit wraparound */
static void ipic_dcs(Dec *d)
{
    for (int p = 0; p < 3; ++p)
    {
        Plane *pl = &d->pl[p];
        uint32_t run = 0;
        for (int by = 0; by < pl->bh; ++by)
        {
            uint8_t v = cell(pl, 0, by - 1)->dc;
            for (int bx = 0; bx < pl->bw; ++bx)
            {
                uint32_t delta = 0;
                if (run) --run;
                else
                {
                    delta = (uint32_t)sym_signed_ovf(&d->tree[0], &d->dcv[p], d->dc_lo, d->dc_hi);
                    if (delta == 0) run = (uint32_t)huff(&d->tree[1], &d->rle[p]);
                }
                v = (uint8_t)(v + delta);
                cell(pl, bx, by)->dc = v;
                v = (uint8_t)((v + cell(pl, bx + 1, by - 1)->dc + 1) / 2);
            }
        }
    }
}

/* h4m:1166-1239 incl. the mirror / zero-fill path for pictures smaller than the nest */
static void make_nest(Dec *d, int nx, int ny)
{
    Plane *y = &d->pl[0];
    /* 70 x 38 in landscape, 38 x 70 in portrait pictures (h4m:965-975) */
    const int NW = d->width < d->height ? NEST_H : NEST_W, NH = d->width < d->height ? NEST_W : NEST_H;
    int cols = y->bw < NW ? y->bw : NW, rows = y->bh < NH ? y->bh : NH;
    int mcols = y->bw < NW ? (NW - y->bw < y->bw ? NW - y->bw : y->bw) : 0;
    int mrows = y->bh < NH ? (NH - y->bh < y->bh ? NH - y->bh : y->bh) : 0;
    memset(d->nest, 0, sizeof d->nest);
    for (int i = 0; i < rows; ++i)
    {
        uint8_t *row = d->nest + i * NW;
        for (int j = 0; j < cols; ++j) row[j] = (cell(y, nx + j, ny + i)->dc >> 4) & 0xF;
        for (int j = 0; j < mcols; ++j) row[cols + j] = (cell(y, nx + cols - 1 - j, ny + i)->dc >> 4) & 0xF;
    }
    for (int i = 0; i < mrows; ++i)
        memcpy(d->nest + (rows + i) * NW, d->nest + (rows - 1 - i) * NW, (size_t)NW);
}

/* h4m:1433-1518.  Top/bottom/right follow the 0x77 rule (the first/last line and the
 * last column alias the block itself, which gives "own DC" just like a border cell);
 * the left value is tracked by the reference as "previous block's DC if that block
 * was type 0 or 8, else own DC" (h4m:1441-1454). */
static void ipic_plane(Dec *d, int p, uint8_t *dst)
{
    Plane *pl = &d->pl[p];
    for (int by = 0; by < pl->bh; ++by)
        for (int bx = 0; bx < pl->bw; ++bx)
        {
            const Cell *c = cell(pl, bx, by);
            uint8_t *o = dst + (by * 4) * pl->w + bx * 4;
            if (c->type == 0)
            {
                int T = by == 0 ? c->dc : nbr_dc(cell(pl, bx, by - 1), c->dc);
                int B = by == pl->bh - 1 ? c->dc : nbr_dc(cell(pl, bx, by + 1), c->dc);
                int R = bx == pl->bw - 1 ? c->dc : nbr_dc(cell(pl, bx + 1, by), c->dc);
                int L = c->dc;
                if (bx > 0)
                {
                    const Cell *l = cell(pl, bx - 1, by);
                    if (l->type == 0 || l->type == 8) L = l->dc;
                }
                fill_weighted(o, pl->w, c->dc, T, B, L, R);
            }
            else if (c->type == 8) fill_flat(o, pl->w, c->dc);
            else fill_intra_aot(d, p, o, pl->w, c->dc, c->type);
        }
}

/* h4m:1970-2016 */
static void decode_ipic(Dec *d, const uint8_t *pic, uint8_t *present)
{
    int dc_shift = pic[0];
    d->unk_shift = pic[1];
    int nx = (int)be16(pic + 4), ny = (int)be16(pic + 6);
    const uint8_t *tab = pic + 8, *data = tab + 0x40;
    section_open(&d->bn[0], data + be32(tab + 0));
    section_open(&d->bnr[0], data + be32(tab + 4));
    section_open(&d->bn[1], data + be32(tab + 8));
    section_open(&d->bnr[1], data + be32(tab + 12));
    for (int p = 0; p < 3; ++p)
    {
        section_open(&d->dcv[p], data + be32(tab + 16 + 12 * p));
        section_open(&d->sc[p], data + be32(tab + 20 + 12 * p));
        section_open(&d->fix[p], data + be32(tab + 24 + 12 * p));
        section_open(&d->rle[p], data + be32(tab + 52 + 4 * p));
    }
    tree_read(&d->tree[3], &d->bn[0], 0, 0);
    tree_read(&d->tree[1], &d->bnr[0], 0, 0);
    tree_read(&d->tree[0], &d->dcv[0], 1, dc_shift);
    tree_read(&d->tree[2], &d->sc[0], 0, 2);
    d->dc_hi = 0x7F * (1 << dc_shift);
    d->dc_lo = -0x80 * (1 << dc_shift);
    ipic_types(d);
    ipic_dcs(d);
    make_nest(d, nx, ny);
    for (int p = 0; p < 3; ++p)
    {
        ipic_plane(d, p, present);
        present += d->pl[p].w * d->pl[p].h;
    }
}

/* ------------------------------------------------------------------ P/B picture */

/* order of the 4x4 blocks inside a macroblock: TL, BL, BR, TR (h4m:862-869) */
static const int SUBX[4] = {0, 0, 1, 1}, SUBY[4] = {0, 1, 1, 0};

static int blocks_in_mcb(int plane) { return plane == 0 ? 4 : 1; }

typedef struct { uint32_t value, count; } RunLen;

/* pass 1, h4m:1742-1776 with helpers 1551-1740: fills type (and, for intra
 * macroblocks, DC) of every block; symbols only, no pixels. */
static void pb_pass1(Dec *d)
{
    static const uint32_t next_type[2][3] = {{1, 2, 0}, {2, 0, 1}};   /* h4m:1591-1594 */
    RunLen proc = {0, 0}, type = {0, 0};
    if (d->mcbp.base)
    {
        proc.value = take_bit(&d->mcbp);
        proc.count = (uint32_t)sym_unsigned_ovf(&d->tree[5], &d->mcbp);
    }
    if (d->mcbt.base)
    {
        type.value = take_bits(&d->mcbt, 2);
        type.count = (uint32_t)sym_unsigned_ovf(&d->tree[5], &d->mcbt);
    }
    uint32_t run_y = 0, run_c = 0;
    uint32_t acc_dc[3] = {0x7F, 0x7F, 0x7F};
    for (int my = 0; my < d->height / 8; ++my)
        for (int mx = 0; mx < d->width / 8; ++mx)
        {
            if (type.count == 0)
            {
                type.value = next_type[take_bit(&d->mcbt)][type.value];
                type.count = (uint32_t)sym_unsigned_ovf(&d->tree[5], &d->mcbt);
            }
            --type.count;
            uint32_t pr = 0;
            if (type.value == 0)
            {
                for (int p = 0; p < 3; ++p)
                    for (int k = 0; k < blocks_in_mcb(p); ++k)
                    {
                        acc_dc[p] += (uint32_t)sym_signed_ovf(&d->tree[0], &d->dcv[p], d->dc_lo, d->dc_hi);
                        Cell *c = p == 0 ? cell(&d->pl[0], mx * 2 + SUBX[k], my * 2 + SUBY[k]) : cell(&d->pl[p], mx, my);
                        c->dc = (uint8_t)acc_dc[p];
                    }
            }
            else
            {
                acc_dc[0] = acc_dc[1] = acc_dc[2] = 0x7F;
                if (proc.count == 0)
                {
                    proc.value ^= 1;
                    proc.count = (uint32_t)sym_unsigned_ovf(&d->tree[5], &d->mcbp);
                }
                --proc.count;
                pr = proc.value;
            }
            uint8_t tag = (uint8_t)(type.value << 5 | pr << 4);
            if (pr == 1)
            {
                for (int k = 0; k < 4; ++k) cell(&d->pl[0], mx * 2 + SUBX[k], my * 2 + SUBY[k])->type = tag;
                cell(&d->pl[1], mx, my)->type = tag;
                cell(&d->pl[2], mx, my)->type = tag;
                continue;
            }
            for (int k = 0; k < 4; ++k)
            {
                Cell *c = cell(&d->pl[0], mx * 2 + SUBX[k], my * 2 + SUBY[k]);
                if (run_y) { c->type = tag; --run_y; continue; }
                int16_t n = (int16_t)huff(&d->tree[3], &d->bn[0]);
                if (n) c->type = (uint8_t)(tag | n);
                else { c->type = tag; run_y = (uint32_t)huff(&d->tree[1], &d->bnr[0]); }
            }
            Cell *cu = cell(&d->pl[1], mx, my), *cv = cell(&d->pl[2], mx, my);
            if (run_c) { cu->type = cv->type = tag; --run_c; }
            else
            {
                int16_t n = (int16_t)huff(&d->tree[3], &d->bn[1]);
                if (n) { cu->type = (uint8_t)(tag | (n & 0xF)); cv->type = (uint8_t)(tag | ((n >> 4) & 0xF)); }
                else { cu->type = cv->type = tag; run_c = (uint32_t)huff(&d->tree[1], &d->bnr[1]); }
            }
        }
}

/* h4m:1846-1860 */
static void read_mv(Dec *d, Section *s, int32_t *mv, int rbits)
{
    int32_t lim = 1 << (rbits + 5);
    int32_t v = huff(&d->tree[4], s) * (1 << rbits);
    for (int i = rbits - 1; i >= 0; --i) v += (int32_t)(take_bit(s) << i);
    *mv += v;
    if (*mv >= lim) *mv -= lim << 1;
    else if (*mv < -lim) *mv += lim << 1;
}

/* pass 2, h4m:1912-1968 with 1327-1355 (whole-MCB MC), 1789-1827 (intra MCB),
 * 1862-1910 (MC + residual MCB). */
static void pb_pass2(Dec *d, uint8_t *present, const uint8_t *past, const uint8_t *future)
{
    uint8_t *out[3];
    const uint8_t *refs[2][3];
    {
        size_t off = 0;
        for (int p = 0; p < 3; ++p)
        {
            out[p] = present + off; refs[0][p] = past + off; refs[1][p] = future + off;
            off += (size_t)d->pl[p].w * d->pl[p].h;
        }
    }
    int32_t mvx = 0, mvy = 0;
    int cur_ref = -1;
    for (int my = 0; my < d->height / 8; ++my)
        for (int mx = 0; mx < d->width / 8; ++mx)
        {
            uint8_t tag = cell(&d->pl[0], mx * 2, my * 2)->type;
            int mtype = (tag >> 5) & 3;
            if (mtype == 0)
            {
                for (int p = 0; p < 3; ++p)
                {
                    Plane *pl = &d->pl[p];
                    for (int k = 0; k < blocks_in_mcb(p); ++k)
                    {
                        int bx = p == 0 ? mx * 2 + SUBX[k] : mx, by = p == 0 ? my * 2 + SUBY[k] : my;
                        const Cell *c = cell(pl, bx, by);
                        uint8_t *o = out[p] + (by * 4) * pl->w + bx * 4;
                        int t = c->type & 0xF;
                        if (t == 0)
                            fill_weighted(o, pl->w, c->dc,
                                          nbr_dc(cell(pl, bx, by - 1), c->dc), nbr_dc(cell(pl, bx, by + 1), c->dc),
                                          nbr_dc(cell(pl, bx - 1, by), c->dc), nbr_dc(cell(pl, bx + 1, by), c->dc));
                        else if (t == 8) fill_flat(o, pl->w, c->dc);
                        else fill_intra_aot(d, p, o, pl->w, c->dc, t);
                    }
                }
                continue;
            }
            int ref = mtype - 1;
            if (ref != cur_ref) { cur_ref = ref; mvx = mvy = 0; }
            read_mv(d, &d->mvh, &mvx, d->rb[ref][0]);
            read_mv(d, &d->mvv, &mvy, d->rb[ref][1]);
            int32_t rx = mx * 16 + mvx, ry = my * 16 + mvy;      /* half-sample luma coordinates */
            int whole = (tag >> 4) & 1;
            /* 70x38 luma window of the reference frame used as the nest of this MCB (h4m:1864-1868) */
            const uint8_t *window = d->width < d->height ? refs[ref][0] + rx / 2 + (ry / 2 - 32) * d->pl[0].w - 16
                                                         : refs[ref][0] + rx / 2 + (ry / 2 - 16) * d->pl[0].w - 32;
            int hx = rx & 1, hy = ry & 1;                            /* 1.3: luma phase for every plane */
            for (int p = 0; p < 3; ++p)
            {
                Plane *pl = &d->pl[p];
                int px = rx >> pl->shift, py = ry >> pl->shift;
                if (d->version15) { hx = px & 1; hy = py & 1; }      /* 1.5: per-plane phase, h4m:1337-1343,1890-1896 */
                for (int k = 0; k < blocks_in_mcb(p); ++k)
                {
                    int bx = p == 0 ? mx * 2 + SUBX[k] : mx, by = p == 0 ? my * 2 + SUBY[k] : my;
                    int sub = p == 0 ? SUBY[k] * 4 * pl->w + SUBX[k] * 4 : 0;
                    uint8_t *o = out[p] + (by * 4) * pl->w + bx * 4;
                    /* linear addressing, no edge clamp (h4m:1344,1897) */
                    const uint8_t *src = refs[ref][p] + (py >> 1) * pl->w + (px >> 1) + sub;
                    int t = whole ? 0 : (cell(pl, bx, by)->type & 0xF);
                    if (t == 6) fill_raw(d, p, o, pl->w);
                    else if (t == 0)
                    {
                        uint8_t m[16];
                        predict4x4(m, src, pl->w, hx, hy);
                        for (int i = 0; i < 16; ++i) o[(i >> 2) * pl->w + (i & 3)] = m[i];
                    }
                    else fill_predicted_aot(d, p, o, src, pl->w, t, window, d->pl[0].w, hx, hy);
                }
            }
        }
}

/* h4m:2018-2061 */
static void decode_pbpic(Dec *d, const uint8_t *pic, uint8_t *present, const uint8_t *past, const uint8_t *future)
{
    d->dc_shift = pic[0];
    d->unk_shift = pic[1];
    d->rb[0][0] = pic[2]; d->rb[0][1] = pic[3];
    d->rb[1][0] = pic[4]; d->rb[1][1] = pic[5];
    const uint8_t *tab = pic + 8, *data = tab + 0x44;
    section_open(&d->bn[0], data + be32(tab + 0));
    section_open(&d->bnr[0], data + be32(tab + 4));
    section_open(&d->bn[1], data + be32(tab + 8));
    section_open(&d->bnr[1], data + be32(tab + 12));
    for (int p = 0; p < 3; ++p)
    {
        section_open(&d->dcv[p], data + be32(tab + 16 + 12 * p));
        section_open(&d->sc[p], data + be32(tab + 20 + 12 * p));
        section_open(&d->fix[p], data + be32(tab + 24 + 12 * p));
        memset(&d->rle[p], 0, sizeof(Section));
    }
    section_open(&d->mvh, data + be32(tab + 52));
    section_open(&d->mvv, data + be32(tab + 56));
    section_open(&d->mcbt, data + be32(tab + 60));
    section_open(&d->mcbp, data + be32(tab + 64));
    tree_read(&d->tree[3], &d->bn[0], 0, 0);
    tree_read(&d->tree[1], &d->bnr[0], 0, 0);
    tree_read(&d->tree[0], &d->dcv[0], 1, d->dc_shift);
    tree_read(&d->tree[2], &d->sc[0], 0, 2);
    tree_read(&d->tree[4], &d->mvh, 1, 0);
    tree_read(&d->tree[5], &d->mcbt, 0, 0);
    d->dc_hi = 0x7F * (1 << d->dc_shift);
    d->dc_lo = -0x80 * (1 << d->dc_shift);
    pb_pass1(d);
    pb_pass2(d, present, past, future);
}

/* ================================================================== stream walker (same C API shape as ref_wrap.c) */

typedef struct PortStream
{
    const uint8_t *data;
    size_t len, pos;
    int width, height, version, n_frames, n_gops;
    uint32_t picsize;
    uint32_t gops_left, vid_left, aud_left, gop_start, vid_in_gop;
    Dec dec;
    uint8_t *past, *present, *future;
} PortStream;

static int parse_header(const uint8_t *h, size_t len, int out[7])
{
    if (len < 0x44) return -1;
    if (!memcmp(h, "HVQM4 1.3\0\0\0\0\0\0\0", 16)) out[3] = 13;
    else if (!memcmp(h, "HVQM4 1.5\0\0\0\0\0\0\0", 16)) out[3] = 15;
    else return -1;
    if (be32(h + 0x10) != 0x44) return -1;
    out[5] = (int)be32(h + 0x18);
    out[2] = (int)be32(h + 0x1C);
    out[0] = (int)be16(h + 0x34);
    out[1] = (int)be16(h + 0x36);
    if (h[0x38] != 2 || h[0x39] != 2) return -2;
    return 0;
}

PORT_API PortStream *port_open(const uint8_t *data, size_t len)
{
    int hd[7];
    if (parse_header(data, len, hd)) return NULL;
    PortStream *s = calloc(1, sizeof *s);
    s->data = data; s->len = len; s->pos = 0x44;
    s->width = hd[0]; s->height = hd[1]; s->n_frames = hd[2]; s->version = hd[3]; s->n_gops = hd[5];
    s->gops_left = (uint32_t)s->n_gops;
    s->picsize = (uint32_t)(s->width * s->height * 3 / 2);
    dec_init(&s->dec, s->width, s->height, s->version == 15);
    s->past = calloc(1, s->picsize + 64);
    s->present = calloc(1, s->picsize + 64);
    s->future = calloc(1, s->picsize + 64);
    return s;
}

PORT_API void port_close(PortStream *s)
{
    if (!s) return;
    dec_free(&s->dec);
    free(s->past); free(s->present); free(s->future);
    free(s);
}

PORT_API void port_info(PortStream *s, int32_t out[6])
{
    out[0] = s->width; out[1] = s->height; out[2] = s->n_frames;
    out[3] = s->version; out[4] = (int32_t)s->picsize; out[5] = s->n_gops;
}

static void swap_ptr(uint8_t **a, uint8_t **b) { uint8_t *t = *a; *a = *b; *b = t; }

static void decode_record(PortStream *s, int type, const uint8_t *rec)
{
    /* reference-window rotation, h4m:2087-2093 and 2131-2137 */
    if (type != PIC_B) swap_ptr(&s->past, &s->future);
    if (type == PIC_I) decode_ipic(&s->dec, rec + 4, s->present);
    else if (type == PIC_P) decode_pbpic(&s->dec, rec + 4, s->present, s->past, s->present);
    else decode_pbpic(&s->dec, rec + 4, s->present, s->past, s->future);
    s->dec.last_type = type;
}

PORT_API int port_decode_next(PortStream *s, uint8_t *out, int32_t meta[4])
{
    for (;;)
    {
        if (s->vid_left == 0 && s->aud_left == 0)
        {
            if (s->gops_left == 0) return 0;
            if (s->pos + 20 > s->len) return -1;
            s->vid_left = be32(s->data + s->pos + 8);
            s->aud_left = be32(s->data + s->pos + 12);
            if (be32(s->data + s->pos + 16) != 0x01000000) return -2;
            s->pos += 20;
            s->gops_left--;
            s->gop_start += s->vid_in_gop;
            s->vid_in_gop = 0;
            continue;
        }
        if (s->pos + 8 > s->len) return -1;
        uint32_t id1 = be16(s->data + s->pos), id2 = be16(s->data + s->pos + 2), size = be32(s->data + s->pos + 4);
        s->pos += 8;
        if (s->pos + size > s->len) return -1;
        const uint8_t *rec = s->data + s->pos;
        s->pos += size;
        if (id1 == 0) { s->aud_left--; continue; }
        if (id1 != 1) return -3;
        if (id2 != PIC_I && id2 != PIC_P && id2 != PIC_B) return -4;
        s->vid_left--; s->vid_in_gop++;
        decode_record(s, (int)id2, rec);
        if (out) memcpy(out, s->present, s->picsize);
        if (meta)
        {
            meta[0] = (int32_t)id2; meta[1] = (int32_t)be32(rec);
            meta[2] = (int32_t)(s->gop_start + be32(rec)); meta[3] = (int32_t)size;
        }
        if (id2 != PIC_B) swap_ptr(&s->present, &s->future);
        return 1;
    }
}

PORT_API void port_section_usage(PortStream *s, int32_t consumed[17], int32_t size[17])
{
    Dec *d = &s->dec;
    Section *all[17] = {&d->bn[0], &d->bnr[0], &d->bn[1], &d->bnr[1],
                        &d->dcv[0], &d->sc[0], &d->fix[0], &d->dcv[1], &d->sc[1], &d->fix[1],
                        &d->dcv[2], &d->sc[2], &d->fix[2], NULL, NULL, NULL, NULL};
    if (d->last_type == PIC_I) { all[13] = &d->rle[0]; all[14] = &d->rle[1]; all[15] = &d->rle[2]; }
    else { all[13] = &d->mvh; all[14] = &d->mvv; all[15] = &d->mcbt; all[16] = &d->mcbp; }
    for (int i = 0; i < 17; ++i)
    {
        consumed[i] = size[i] = 0;
        if (!all[i] || !all[i]->base) continue;
        size[i] = (int32_t)all[i]->size;
        int is_fix = i == 6 || i == 9 || i == 12;
        consumed[i] = is_fix ? (int32_t)all[i]->byte : (int32_t)((all[i]->bit + 31) / 32 * 4);
    }
}

PORT_API int port_get_map(PortStream *s, int plane, uint8_t *out, int32_t dims[2])
{
    Plane *p = &s->dec.pl[plane];
    dims[0] = p->stride; dims[1] = p->bh + 2;
    if (out) memcpy(out, p->map, (size_t)p->stride * (p->bh + 2) * sizeof(Cell));
    return 0;
}

PORT_API void port_get_nest(PortStream *s, uint8_t out[NEST_W * NEST_H]) { memcpy(out, s->dec.nest, NEST_W * NEST_H); }

/* ------------------------------------------------------------------ timing (decode calls only) */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static double bench_once(const uint8_t *data, size_t len, int reps, long *frames_out)
{
    double t = 0;
    long frames = 0;
    PortStream *s = port_open(data, len);
    if (!s) { *frames_out = 0; return 0; }
    for (int r = 0; r < reps; ++r)
    {
        s->pos = 0x44; s->gops_left = (uint32_t)s->n_gops; s->vid_left = s->aud_left = 0;
        for (;;)
        {
            /* locate the next record off the clock, decode it on the clock */
            if (s->vid_left == 0 && s->aud_left == 0)
            {
                if (s->gops_left == 0) break;
                s->vid_left = be32(s->data + s->pos + 8);
                s->aud_left = be32(s->data + s->pos + 12);
                s->pos += 20; s->gops_left--;
                continue;
            }
            uint32_t id1 = be16(s->data + s->pos), id2 = be16(s->data + s->pos + 2), size = be32(s->data + s->pos + 4);
            const uint8_t *rec = s->data + s->pos + 8;
            s->pos += 8 + size;
            if (id1 != 1) { s->aud_left--; continue; }
            s->vid_left--;
            double t0 = now_s();
            decode_record(s, (int)id2, rec);
            t += now_s() - t0;
            if (id2 != PIC_B) swap_ptr(&s->present, &s->future);
            ++frames;
        }
    }
    port_close(s);
    *frames_out = frames;
    return t;
}

PORT_API double port_bench(const uint8_t *data, size_t len, int reps, int64_t *frames)
{
    long f;
    double t = bench_once(data, len, reps, &f);
    *frames = f;
    return t;
}

PORT_API int port_bench_mp(const uint8_t *data, size_t len, int nproc, int reps, double out[3])
{
    int (*pipes)[2] = malloc(sizeof(int[2]) * (size_t)nproc);
    pid_t *pids = malloc(sizeof(pid_t) * (size_t)nproc);
    double t0 = now_s();
    for (int i = 0; i < nproc; ++i)
    {
        if (pipe(pipes[i])) return -1;
        pids[i] = fork();
        if (pids[i] == 0)
        {
            close(pipes[i][0]);
            long f;
            double t = bench_once(data, len, reps, &f);
            double msg[2] = {t, (double)f};
            if (write(pipes[i][1], msg, sizeof msg) != (ssize_t)sizeof msg) _exit(1);
            _exit(0);
        }
        close(pipes[i][1]);
    }
    double tsum = 0, fsum = 0;
    int rc = 0;
    for (int i = 0; i < nproc; ++i)
    {
        double msg[2] = {0, 0};
        if (read(pipes[i][0], msg, sizeof msg) != (ssize_t)sizeof msg) rc = -2;
        close(pipes[i][0]);
        int st;
        waitpid(pids[i], &st, 0);
        tsum += msg[0]; fsum += msg[1];
    }
    out[0] = now_s() - t0; out[1] = tsum; out[2] = fsum;
    free(pipes); free(pids);
    return rc;
}

/*
 * The same over a SET of streams: worker i decodes streams i, i + nproc, i + 2 nproc, ... (cyclically) once each,
 * `reps` GOP decodes in all -- the bounded CPU sample of bench.py cycles the same distinct bitstreams as the GPU arm.
 */
PORT_API int port_bench_mp_streams(const uint8_t *const *datas, const size_t *lens, int n_streams, int nproc, int reps, double out[3])
{
    int (*pipes)[2] = malloc(sizeof(int[2]) * nproc);
    pid_t *pids = malloc(sizeof(pid_t) * nproc);
    double t0 = now_s();
    for (int i = 0; i < nproc; ++i)
    {
        if (pipe(pipes[i])) return -1;
        pids[i] = fork();
        if (pids[i] == 0)
        {
            close(pipes[i][0]);
            double msg[2] = {0, 0};
            for (int r = 0; r < reps; ++r)
            {
                const int k = (int)(((long)i + (long)r * nproc) % n_streams);
                long f;
                msg[0] += bench_once(datas[k], lens[k], 1, &f);
                msg[1] += (double)f;
            }
            if (write(pipes[i][1], msg, sizeof msg) != sizeof msg) _exit(1);
            _exit(0);
        }
        close(pipes[i][1]);
    }
    double tsum = 0, fsum = 0;
    int rc = 0;
    for (int i = 0; i < nproc; ++i)
    {
        double msg[2] = {0, 0};
        if (read(pipes[i][0], msg, sizeof msg) != sizeof msg) rc = -2;
        close(pipes[i][0]);
        int st;
        waitpid(pids[i], &st, 0);
        tsum += msg[0];
        fsum += msg[1];
    }
    out[0] = now_s() - t0;
    out[1] = tsum;
    out[2] = fsum;
    free(pipes);
    free(pids);
    return rc;
}

/* ---- planar 4:2:0 -> interleaved RGB as the reference's dumpRGB does it (h4m:895-926): JPEG
   matrix in single-precision float, chroma replicated 2x2 (no interpolation), results truncated
   towards zero and clamped to 0..255.  Returns 0 (same contract as ref_yuv_to_rgb). ---- */
static uint8_t rgb_clamp(float f) { return f < 0 ? 0 : f > 255 ? 255 : (uint8_t)f; }   /* h4m:896-899 */

PORT_API int port_yuv_to_rgb(const uint8_t *yuv, int w, int h, uint8_t *rgb)
{
    const uint8_t *yp = yuv, *up = yp + (size_t)w * h, *vp = up + (size_t)w * h / 4;
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j)
        {
            const float y = yp[(size_t)i * w + j];
            const float u = up[(size_t)(i / 2) * (w / 2) + j / 2], v = vp[(size_t)(i / 2) * (w / 2) + j / 2];
            *rgb++ = rgb_clamp(y + 1.402f * (v - 128.f));                                  /* h4m:918 */
            *rgb++ = rgb_clamp(y - 0.34414f * (u - 128.f) - 0.71414f * (v - 128.f));       /* h4m:919 */
            *rgb++ = rgb_clamp(y + 1.772f * (u - 128.f));                                  /* h4m:920 */
        }
    return 0;
}

/* ---- IMA-ADPCM audio track as the reference's decode_audio does it (h4m:185-258): same contract
   as ref_decode_audio in ref_wrap.c.  Seed (first record of a GOP block): per channel, DESCENDING,
   predictor high byte, then bit 7 = predictor bit 7 and bits 6:0 = step index; codes are 4 bits,
   high nibble first, channels descending inside a sample, fresh byte at every record. ---- */
static const int16_t ima_steps[89] = {
    7, 8, 9, 10, 11, 12, 13, 14, 16, 17, 19, 21, 23, 25, 28, 31, 34, 37, 41, 45, 50, 55, 60, 66, 73, 80, 88, 97, 107, 118, 130, 143,
    157, 173, 190, 209, 230, 253, 279, 307, 337, 371, 408, 449, 494, 544, 598, 658, 724, 796, 876, 963, 1060, 1166, 1282, 1411,
    1552, 1707, 1878, 2066, 2272, 2499, 2749, 3024, 3327, 3660, 4026, 4428, 4871, 5358, 5894, 6484, 7132, 7845, 8630, 9493, 10442,
    11487, 12635, 13899, 15289, 16818, 18500, 20350, 22385, 24623, 27086, 29794, 32767};
static const int8_t ima_index[16] = {-1, -1, -1, -1, 2, 4, 6, 8, -1, -1, -1, -1, 2, 4, 6, 8};   /* h4m:170-176 */

PORT_API int port_decode_audio(int32_t *state, int channels, int first, uint32_t sample_count, const uint8_t *data, size_t len, int16_t *pcm)
{
    size_t pos = 0;
    uint32_t i = 0, t = 0;
    if (first)
    {   /* h4m:190-212 */
        for (int c = channels - 1; c >= 0; --c)
        {
            if (pos + 2 > len) return -1;
            const uint8_t hi = data[pos++], b = data[pos++];
            state[2 * c] = (int16_t)((uint16_t)hi << 8 | (b & 0x80));
            state[2 * c + 1] = b & 0x7F;
            if (state[2 * c + 1] > 88) return -2;       /* the reference exits */
        }
        for (int c = 0; c < channels; ++c) pcm[t++] = (int16_t)state[2 * c];
        ++i;
    }
    uint8_t b = 0;
    int bitsleft = 0;
    for (; i < sample_count; ++i)
    {   /* h4m:216-246 */
        for (int c = channels - 1; c >= 0; --c)
        {
            if (bitsleft == 0)
            {
                if (pos >= len) return -1;
                b = data[pos++];
                bitsleft = 8;
            }
            const int32_t step = ima_steps[state[2 * c + 1]];
            int32_t delta = step >> 3;
            if (b & 0x10) delta += step >> 2;
            if (b & 0x20) delta += step >> 1;
            if (b & 0x40) delta += step;
            int32_t h = (b & 0x80) ? state[2 * c] - delta : state[2 * c] + delta;
            state[2 * c] = h > 32767 ? 32767 : h < -32768 ? -32768 : h;
            int32_t idx = state[2 * c + 1] + ima_index[(b & 0xF0) >> 4];
            state[2 * c + 1] = idx > 88 ? 88 : idx < 0 ? 0 : idx;
            b = (uint8_t)(b << 4);
            bitsleft -= 4;
        }
        for (int c = 0; c < channels; ++c) pcm[t++] = (int16_t)state[2 * c];
    }
    return (int)t;
}
