/*
 * tools/h4mgen.c -- seeded synthetic HVQM4 1.3/1.5 bitstream generator.
 *
 * No .h4m game assets are available offline, so tests and bench.py decode streams
 * produced here.  The generator is an *encoder of syntax*, not of images: it draws
 * random (seeded) block types, DC deltas, AOT basis descriptors, macroblock types
 * and motion vectors and serialises them exactly the way the reference decoder
 * consumes them:
 *
 *   container        /root/reference/h4m_audio_decode.c:2175-2247 (0x44 header),
 *                    2429-2438 (GOP block), 2456-2458 (frame record), 2085 (disp_id)
 *   I picture        h4m_audio_decode.c:1970-1999  (8-byte header, 16 sections, 4 trees)
 *   P/B picture      h4m_audio_decode.c:2018-2050  (8-byte header, 17 sections, 6 trees)
 *   section          h4m_audio_decode.c:1061-1071  (BE32 size + bytes)
 *   tree             h4m_audio_decode.c:607-651    (pre-order, 1=node, 0+8 bits=leaf)
 *   escapes          h4m_audio_decode.c:654-677    (signed 0x7F/0x80, unsigned 255)
 *   types / DCs      h4m_audio_decode.c:1073-1164  (I), 1649-1740 (P/B pass 1)
 *   MCB type/proc    h4m_audio_decode.c:1551-1622
 *   motion vectors   h4m_audio_decode.c:1846-1860, 1943-1955
 *   pass-2 order     h4m_audio_decode.c:1789-1910
 *
 * It obeys every validity rule in SURVEY.md section 9 (first frame I, nest window
 * inside the DC map, legal nibbles, non-empty leader sections, RL counts >= 1,
 * no future references in P pictures, in-bounds half-pel taps, in-frame 70x38
 * window for predicted-AOT macroblocks), because the reference has no input
 * validation at all and reads out of bounds otherwise.
 *
 * Build: gcc -O2 -shared -fPIC tools/h4mgen.c -o tools/libh4mgen.so
 *        gcc -O2 -DH4MGEN_MAIN tools/h4mgen.c -o tools/h4mgen
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define GEN_API __attribute__((visibility("default")))

typedef struct
{
    int32_t width, height;   /* multiples of 8, >= 16; below 280x152 the nest is mirrored / zero-filled (h4m:1173-1203) */
    int32_t version;         /* 13 or 15 */
    int32_t n_gops;
    int32_t profile;         /* 0 = dense (worst case), 1 = realistic (sparse, coherent motion), 2 = stress: dense plus everything the
                                reference accepts and an encoder rarely emits (block types 7 and 9..255 = that many bases in I-picture
                                luma, nibbles 7 and 9..15 elsewhere, scale symbols to 255, dc_shift 0..3, unk_shift 6..12, escape
                                chains, run lengths >= 255, rb 0..3), 3 / 4 = every luma block of an I picture carries 16 / 17 bases
                                (at / beyond the symbol capacity of the GPU entropy stage) */
    int32_t usec_per_frame;
    uint64_t seed;
    const char *gop;         /* decode-order pattern, e.g. "IPPP" or "IPBBPBB"; must start with I */
} H4MGenParams;

/* ------------------------------------------------------------------ rng */

typedef struct { uint64_t s; } Rng;

static uint64_t rng_next(Rng *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* uniform in [0, n) */
static uint32_t rnd(Rng *r, uint32_t n) { return n ? (uint32_t)((rng_next(r) >> 32) % n) : 0; }
/* uniform in [lo, hi] */
static int32_t rnd_range(Rng *r, int32_t lo, int32_t hi) { return lo + (int32_t)rnd(r, (uint32_t)(hi - lo + 1)); }
static int chance(Rng *r, uint32_t percent) { return rnd(r, 100) < percent; }

/* ------------------------------------------------------------------ growable buffers */

typedef struct { uint8_t *p; size_t n, cap; } Bytes;

static void bytes_reserve(Bytes *b, size_t extra)
{
    if (b->n + extra > b->cap)
    {
        size_t c = b->cap ? b->cap * 2 : 4096;
        while (c < b->n + extra) c *= 2;
        b->p = realloc(b->p, c);
        b->cap = c;
    }
}
static void put8(Bytes *b, uint32_t v) { bytes_reserve(b, 1); b->p[b->n++] = (uint8_t)v; }
static void put16(Bytes *b, uint32_t v) { put8(b, v >> 8); put8(b, v); }
static void put32(Bytes *b, uint32_t v) { put16(b, v >> 16); put16(b, v); }
static void putn(Bytes *b, const void *src, size_t n) { bytes_reserve(b, n); memcpy(b->p + b->n, src, n); b->n += n; }
static void poke32(Bytes *b, size_t at, uint32_t v)
{
    b->p[at] = v >> 24; b->p[at + 1] = v >> 16; b->p[at + 2] = v >> 8; b->p[at + 3] = v;
}

/* A coded section is a list of ops in consumption order: a Huffman symbol
   (value = leaf byte) or `nbits` raw bits. */
typedef struct { uint32_t *v; size_t n, cap; } Ops;
#define OP_RAW 0x80000000u

static void ops_push(Ops *o, uint32_t x)
{
    if (o->n == o->cap)
    {
        o->cap = o->cap ? o->cap * 2 : 1024;
        o->v = realloc(o->v, o->cap * sizeof(uint32_t));
    }
    o->v[o->n++] = x;
}
static void op_sym(Ops *o, uint32_t byte) { ops_push(o, byte & 0xFF); }
static void op_raw(Ops *o, uint32_t nbits, uint32_t value)
{
    if (nbits) ops_push(o, OP_RAW | (nbits << 16) | (value & ((1u << nbits) - 1)));
}
static void ops_append(Ops *dst, const Ops *src)
{
    for (size_t i = 0; i < src->n; ++i) ops_push(dst, src->v[i]);
}

/* signed escape-extended value in units of (1 << dc_shift); decodeSOvfSym, h4m:654-664 */
static void op_sovf(Ops *o, int32_t v)
{
    while (v >= 127) { op_sym(o, 0x7F); v -= 127; }
    while (v <= -128) { op_sym(o, 0x80); v += 128; }
    op_sym(o, (uint8_t)(int8_t)v);
}
/* unsigned escape-extended count; decodeUOvfSym, h4m:667-677 */
static void op_uovf(Ops *o, uint32_t v)
{
    while (v >= 255) { op_sym(o, 255); v -= 255; }
    op_sym(o, v);
}

/* ------------------------------------------------------------------ bit writer */

typedef struct { Bytes out; uint32_t acc; int nacc; } BitW;

static void bw_put(BitW *w, uint32_t nbits, uint32_t value)
{
    for (int i = (int)nbits - 1; i >= 0; --i)
    {
        w->acc = (w->acc << 1) | ((value >> i) & 1);
        if (++w->nacc == 8) { put8(&w->out, w->acc); w->acc = 0; w->nacc = 0; }
    }
}
static void bw_finish(BitW *w)
{
    if (w->nacc) { put8(&w->out, w->acc << (8 - w->nacc)); w->acc = 0; w->nacc = 0; }
    while (w->out.n & 3) put8(&w->out, 0);   /* sections are read as BE32 words */
}

/* ------------------------------------------------------------------ Huffman */

typedef struct
{
    int left[512], right[512];   /* children of internal nodes (index >= 256) */
    int root;
    int n_internal;
    uint8_t len[256];
    uint8_t code[256][32];       /* up to 255 bits, MSB first, packed */
} Huff;

static void huff_assign(Huff *h, int node, uint8_t *path, int depth)
{
    if (node < 256)
    {
        h->len[node] = (uint8_t)depth;
        memset(h->code[node], 0, 32);
        for (int i = 0; i < depth; ++i)
            if (path[i]) h->code[node][i >> 3] |= 0x80 >> (i & 7);
        return;
    }
    path[depth] = 0; huff_assign(h, h->left[node], path, depth + 1);
    path[depth] = 1; huff_assign(h, h->right[node], path, depth + 1);
}

/* Builds a Huffman tree over the symbols with freq > 0.  With no symbol at all a
   single leaf 0 is emitted (rule: every leader section carries a >= 1 leaf tree). */
static void huff_build(Huff *h, const uint64_t freq[256], Rng *r)
{
    int nodes[256];
    uint64_t w[512];
    int n = 0;
    memset(h, 0, sizeof *h);
    for (int s = 0; s < 256; ++s)
        if (freq[s]) { nodes[n++] = s; w[s] = freq[s]; }
    if (n == 0) { nodes[n++] = 0; w[0] = 1; }
    int next = 256;
    while (n > 1)
    {
        /* pick the two lightest (O(n^2) is fine for <= 256 symbols) */
        int a = 0;
        for (int i = 1; i < n; ++i) if (w[nodes[i]] < w[nodes[a]]) a = i;
        int na = nodes[a]; nodes[a] = nodes[--n];
        int b = 0;
        for (int i = 1; i < n; ++i) if (w[nodes[i]] < w[nodes[b]]) b = i;
        int nb = nodes[b];
        int in = next++;
        /* random child order so both bit polarities get exercised */
        if (rnd(r, 2)) { h->left[in] = na; h->right[in] = nb; }
        else           { h->left[in] = nb; h->right[in] = na; }
        w[in] = w[na] + w[nb];
        nodes[b] = in;
    }
    h->root = nodes[0];
    h->n_internal = next - 256;
    uint8_t path[256];
    huff_assign(h, h->root, path, 0);
}

static void huff_write_tree(const Huff *h, int node, BitW *w)
{
    if (node < 256) { bw_put(w, 1, 0); bw_put(w, 8, (uint32_t)node); return; }
    bw_put(w, 1, 1);
    huff_write_tree(h, h->left[node], w);    /* the 0 side first, h4m:624-627 */
    huff_write_tree(h, h->right[node], w);
}

static void huff_put(const Huff *h, BitW *w, uint32_t sym)
{
    for (int i = 0; i < h->len[sym]; ++i)
        bw_put(w, 1, (h->code[sym][i >> 3] >> (7 - (i & 7))) & 1);
}

/* ------------------------------------------------------------------ picture model */

enum
{
    S_BN_Y, S_BNR_Y, S_BN_C, S_BNR_C,
    S_DC_Y, S_SC_Y, S_FIX_Y, S_DC_U, S_SC_U, S_FIX_U, S_DC_V, S_SC_V, S_FIX_V,
    S_X0, S_X1, S_X2, S_X3,     /* I: dc_rle[Y,U,V], -    P/B: mv_h, mv_v, mcb_type, mcb_proc */
    S_COUNT
};
static const int S_DC[3] = {S_DC_Y, S_DC_U, S_DC_V};
static const int S_SC[3] = {S_SC_Y, S_SC_U, S_SC_V};
static const int S_FIX[3] = {S_FIX_Y, S_FIX_U, S_FIX_V};

typedef struct
{
    Ops ops[S_COUNT];
    Bytes fix[3];
    Ops dc_pass2[3];     /* P/B: PrediAot (S1,S2) pairs, appended after the pass-1 DC deltas */
} Pic;

static void pic_free(Pic *p)
{
    for (int i = 0; i < S_COUNT; ++i) free(p->ops[i].v);
    for (int i = 0; i < 3; ++i) { free(p->fix[i].p); free(p->dc_pass2[i].v); }
}

typedef struct
{
    const H4MGenParams *prm;
    Rng rng;
    int w, h, mbw, mbh;
    int bw[3], bh[3];     /* 4x4 blocks per plane */
} Gen;

/* -- distributions -------------------------------------------------- */

static int prof_dense(const Gen *g) { return g->prm->profile != 1; }
static int prof_stress(const Gen *g) { return g->prm->profile >= 2; }

/* luma_ipic: the block type is a whole byte there (h4m:1433-1459: type n = n bases), a nibble everywhere else */
static int pick_intra_type_ex(Gen *g, int luma_ipic)
{
    if (g->prm->profile == 3 || g->prm->profile == 4) return luma_ipic ? 13 + g->prm->profile : 0;
    if (prof_stress(g) && chance(&g->rng, 12))
    {
        if (luma_ipic) return chance(&g->rng, 8) ? rnd_range(&g->rng, 41, 255) : chance(&g->rng, 50) ? 7 : rnd_range(&g->rng, 9, 40);
        return chance(&g->rng, 40) ? 7 : rnd_range(&g->rng, 9, 15);
    }
    if (prof_dense(g))
    {
        static const uint8_t t[11] = {0, 0, 0, 8, 8, 1, 2, 3, 4, 5, 6};
        return t[rnd(&g->rng, 11)];
    }
    uint32_t x = rnd(&g->rng, 100);
    return x < 70 ? 0 : x < 80 ? 8 : x < 90 ? 1 : x < 95 ? 2 : x < 98 ? 3 : 6;
}
static int pick_intra_type(Gen *g) { return pick_intra_type_ex(g, 0); }

static int pick_inter_nibble(Gen *g, int window_ok)
{
    if (prof_dense(g))
    {
        if (window_ok)
        {
            /* k = k - 1 bases (h4m:1383); 8 is not a flat block here but 7 bases */
            if (prof_stress(g) && chance(&g->rng, 12)) return rnd_range(&g->rng, 7, 15);
            static const uint8_t t[10] = {0, 0, 0, 6, 1, 2, 3, 4, 5, 2};
            return t[rnd(&g->rng, 10)];
        }
        return chance(&g->rng, 25) ? 6 : 0;
    }
    uint32_t x = rnd(&g->rng, 100);
    if (!window_ok) return x < 96 ? 0 : 6;
    return x < 80 ? 0 : x < 88 ? 1 : x < 94 ? 2 : x < 97 ? 3 : 6;
}

static int pick_zero_run(Gen *g)
{
    if (prof_stress(g) && chance(&g->rng, 3)) return rnd_range(&g->rng, 200, 255);      /* the run length is one symbol: at most 255 */
    if (prof_dense(g)) return chance(&g->rng, 30) ? rnd_range(&g->rng, 1, 5) : 0;
    return chance(&g->rng, 60) ? rnd_range(&g->rng, 1, 24) : 0;
}

static int32_t pick_dc_delta(Gen *g)
{
    if (prof_stress(g) && chance(&g->rng, 3))
    {   /* a chain of escapes (h4m:654-664), both signs, also ending exactly on a multiple of the escape value */
        int32_t m = chance(&g->rng, 30) ? 127 * rnd_range(&g->rng, 1, 6) : rnd_range(&g->rng, 300, 1500);
        return chance(&g->rng, 50) ? m : -m - 1;
    }
    if (prof_dense(g))
    {
        if (chance(&g->rng, 6))
        {   /* needs one or more escape symbols */
            int32_t m = rnd_range(&g->rng, 127, 300);
            return chance(&g->rng, 50) ? m : -m - 1;
        }
        return rnd_range(&g->rng, -12, 12);
    }
    return rnd_range(&g->rng, -4, 4);
}

/* side data of an AOT-coded block with n bases: n x (16-bit descriptor, scale symbol) */
static void emit_bases(Gen *g, Pic *p, int plane, int n)
{
    for (int k = 0; k < n; ++k)
    {
        put16(&p->fix[plane], rnd(&g->rng, 0x10000));
        if (prof_stress(g) && chance(&g->rng, 10)) op_sym(&p->ops[S_SC[plane]], rnd(&g->rng, 256));   /* scale_sum wraps mod 2^32 in the products */
        else op_sym(&p->ops[S_SC[plane]], prof_dense(g) ? rnd(&g->rng, 6) : rnd(&g->rng, 4));
    }
}
static void emit_raw(Gen *g, Pic *p, int plane)
{
    for (int i = 0; i < 16; ++i) put8(&p->fix[plane], rnd(&g->rng, 256));
}

/* -- I picture ------------------------------------------------------- */

static void gen_ipic(Gen *g, Pic *p, uint8_t hdr[8])
{
    Rng *r = &g->rng;
    int dc_shift = prof_stress(g) ? (int)rnd(r, 4) : (int)rnd(r, 2);
    int unk_shift = prof_stress(g) ? rnd_range(r, 6, 12) : prof_dense(g) ? rnd_range(r, 8, 10) : 10;
    /* pictures narrower / lower than the nest take MakeNest's mirror + zero-fill path (h4m:1173-1203): origin 0 */
    /* the nest is 70 x 38 blocks, 38 x 70 in portrait pictures (h4m:965-975) */
    const int nw = g->w < g->h ? 38 : 70, nh = g->w < g->h ? 70 : 38;
    int nest_x = g->bw[0] >= nw ? (int)rnd(r, g->bw[0] - nw + 1) : 0, nest_y = g->bh[0] >= nh ? (int)rnd(r, g->bh[0] - nh + 1) : 0;
    hdr[0] = dc_shift; hdr[1] = unk_shift; hdr[2] = 0; hdr[3] = 0;
    hdr[4] = nest_x >> 8; hdr[5] = nest_x; hdr[6] = nest_y >> 8; hdr[7] = nest_y;

    /* block types (Ipic_BasisNumDec, h4m:1073-1130) */
    uint8_t *type[3];
    for (int pl = 0; pl < 3; ++pl) type[pl] = calloc((size_t)g->bw[pl] * g->bh[pl], 1);
    {
        int run = 0;
        for (int i = 0; i < g->bw[0] * g->bh[0]; ++i)
        {
            if (run) { type[0][i] = 0; --run; continue; }
            int t = pick_intra_type_ex(g, 1);
            type[0][i] = t;
            op_sym(&p->ops[S_BN_Y], t);
            if (t == 0) { run = pick_zero_run(g); op_sym(&p->ops[S_BNR_Y], run); }
        }
        run = 0;
        for (int i = 0; i < g->bw[1] * g->bh[1]; ++i)
        {
            if (run) { type[1][i] = type[2][i] = 0; --run; continue; }
            int u = pick_intra_type(g), v = pick_intra_type(g);
            /* 8 does not fit next to a second nibble only if >15; it does (0x88) */
            type[1][i] = u; type[2][i] = v;
            int sym = u | (v << 4);
            op_sym(&p->ops[S_BN_C], sym);
            if (sym == 0) { run = pick_zero_run(g); op_sym(&p->ops[S_BNR_C], run); }
        }
    }
    /* DC deltas with zero-run RLE (IpicDcvDec/getDeltaDC, h4m:1043-1058,1132-1164) */
    for (int pl = 0; pl < 3; ++pl)
    {
        int run = 0;
        for (int i = 0; i < g->bw[pl] * g->bh[pl]; ++i)
        {
            if (run) { --run; continue; }
            int32_t d = chance(r, prof_dense(g) ? 30 : 50) ? 0 : pick_dc_delta(g);
            op_sovf(&p->ops[S_DC[pl]], d);
            if (d == 0)
            {
                run = prof_stress(g) && chance(r, 3) ? rnd_range(r, 200, 255) : prof_dense(g) ? rnd(r, 6) : rnd(r, 12);
                op_sym(&p->ops[S_X0 + pl], run);
            }
        }
    }
    /* per-block side data, plane by plane in raster order (IpicPlaneDec, h4m:1487) */
    for (int pl = 0; pl < 3; ++pl)
        for (int i = 0; i < g->bw[pl] * g->bh[pl]; ++i)
        {
            int t = type[pl][i];
            if (t == 6) emit_raw(g, p, pl);
            else if (t != 0 && t != 8) emit_bases(g, p, pl, t);
        }
    for (int pl = 0; pl < 3; ++pl) free(type[pl]);
}

/* -- P/B picture ------------------------------------------------------ */

typedef struct { int lo, hi; } Span;

/* legal half-pel vectors for a macroblock at pixel `pos` of a `size`-pixel axis:
   0 <= (2pos+mv)>>1 and ((2pos+mv)>>1)+9 <= size, |mv| < M.  If `lo_pad`/`hi_pad`
   are non-zero the integer position must additionally keep lo_pad pixels before
   and hi_pad pixels after it inside the axis (70x38 window of PrediAot MCBs). */
static Span mv_span(int pos, int size, int M, int lo_pad, int hi_pad)
{
    int lo_i = lo_pad, hi_i = size - (hi_pad > 9 ? hi_pad : 9);   /* integer position range */
    Span s;
    s.lo = 2 * lo_i - 2 * pos;
    s.hi = 2 * hi_i + 1 - 2 * pos;
    if (s.lo < -M) s.lo = -M;
    if (s.hi > M - 1) s.hi = M - 1;
    return s;
}

static void emit_mv(Ops *o, int *pred, int target, int rb)
{
    int M = 1 << (rb + 5);
    int d = target - *pred;
    /* representative of d modulo 2M inside [-M, M) */
    d = ((d + M) % (2 * M) + 2 * M) % (2 * M) - M;
    int sym = d >> rb;                      /* floor; in [-32, 31] */
    int res = d & ((1 << rb) - 1);
    op_sym(o, (uint8_t)(int8_t)sym);
    op_raw(o, rb, res);
    *pred = target;
}

static void gen_pbpic(Gen *g, Pic *p, uint8_t hdr[8], int is_b)
{
    Rng *r = &g->rng;
    const int dense = prof_dense(g), stress = prof_stress(g);
    int dc_shift = stress ? (int)rnd(r, 4) : (int)rnd(r, 2);
    int unk_shift = stress ? rnd_range(r, 6, 12) : dense ? rnd_range(r, 8, 10) : 10;
    int rb[2][2];   /* [ref][h/v] */
    for (int f = 0; f < 2; ++f)
        for (int a = 0; a < 2; ++a) rb[f][a] = stress ? rnd_range(r, 0, 3) : rnd_range(r, 1, 2);
    hdr[0] = dc_shift; hdr[1] = unk_shift;
    hdr[2] = rb[0][0]; hdr[3] = rb[0][1]; hdr[4] = rb[1][0]; hdr[5] = rb[1][1];
    hdr[6] = 0; hdr[7] = 0;

    int nmb = g->mbw * g->mbh;
    uint8_t *mtype = malloc(nmb), *mproc = malloc(nmb);
    /* per-MCB nibbles: 4 luma (order TL,BL,BR,TR = mcb_offset, h4m:862-865), U, V */
    uint8_t (*nib)[6] = calloc(nmb, 6);
    int16_t (*mv)[2] = calloc(nmb, 4);
    uint8_t *winok = calloc(nmb, 1);

    /* macroblock types: random walk; P pictures never use type 2 (h4m:2060) */
    {
        int t = dense ? (int)rnd(r, is_b ? 3 : 2) : 1;
        /* stress: some pictures change type / proc so rarely that the run counts need escapes (>= 255, h4m:667-677) */
        const int long_runs = stress && chance(r, 40);
        int change = long_runs ? 1 : dense ? 15 : 4;
        for (int i = 0; i < nmb; ++i)
        {
            if (i && chance(r, change))
            {
                if (is_b) t = (t + 1 + (int)rnd(r, 2)) % 3;
                else t ^= 1;
                if (!dense && t == 0 && chance(r, 70)) t = 1;   /* intra is rare in real content */
            }
            mtype[i] = t;
        }
        int pr = rnd(r, 2);
        for (int i = 0; i < nmb; ++i)
        {
            if (mtype[i] == 0) { mproc[i] = 0; continue; }
            if (chance(r, long_runs ? 1 : dense ? 20 : 10)) pr ^= 1;
            if (!dense && pr == 0 && chance(r, 50)) pr = 1;
            mproc[i] = pr;
        }
    }
    /* motion vectors (targets), honouring the in-bounds rules */
    {
        int gm[2] = {rnd_range(r, -12, 12), rnd_range(r, -12, 12)};
        int cur[2] = {gm[0], gm[1]};
        for (int my = 0; my < g->mbh; ++my)
            for (int mx = 0; mx < g->mbw; ++mx)
            {
                int i = my * g->mbw + mx;
                if (mtype[i] == 0) continue;
                int f = mtype[i] - 1;
                int Mh = 1 << (rb[f][0] + 5), Mv = 1 << (rb[f][1] + 5);
                Span sh = mv_span(mx * 8, g->w, Mh, 0, 0), sv = mv_span(my * 8, g->h, Mv, 0, 0);
                /* window of predicted-AOT macroblocks: 70 x 38 at (-32, -16), in portrait 38 x 70 at (-16, -32), h4m:1864-1868 */
                const int portrait = g->w < g->h;
                Span wh = portrait ? mv_span(mx * 8, g->w, Mh, 16, 22) : mv_span(mx * 8, g->w, Mh, 32, 38);
                Span wv = portrait ? mv_span(my * 8, g->h, Mv, 32, 38) : mv_span(my * 8, g->h, Mv, 16, 22);
                int can_win = wh.lo <= wh.hi && wv.lo <= wv.hi;
                int want_win = mproc[i] == 0 && can_win && chance(r, dense ? 60 : 90);
                Span uh = want_win ? wh : sh, uv = want_win ? wv : sv;
                int th, tv;
                if (dense)
                {
                    th = rnd_range(r, uh.lo, uh.hi);
                    tv = rnd_range(r, uv.lo, uv.hi);
                }
                else
                {
                    if (chance(r, 20)) { cur[0] = gm[0] + rnd_range(r, -2, 2); cur[1] = gm[1] + rnd_range(r, -2, 2); }
                    th = cur[0] < uh.lo ? uh.lo : cur[0] > uh.hi ? uh.hi : cur[0];
                    tv = cur[1] < uv.lo ? uv.lo : cur[1] > uv.hi ? uv.hi : cur[1];
                }
                mv[i][0] = th; mv[i][1] = tv;
                int rx = (2 * mx * 8 + th) >> 1, ry = (2 * my * 8 + tv) >> 1;
                winok[i] = portrait ? rx >= 16 && rx + 22 <= g->w && ry >= 32 && ry + 38 <= g->h
                                    : rx >= 32 && rx + 38 <= g->w && ry >= 16 && ry + 22 <= g->h;
            }
    }

    /* ---- pass 1: spread_PB_descMap (h4m:1742-1776) ---- */
    Ops *ty = &p->ops[S_X2], *pc = &p->ops[S_X3];
    {
        /* type run-length stream */
        int i = 0, first = 1, prev = 0;
        while (i < nmb)
        {
            int t = mtype[i], j = i;
            while (j < nmb && mtype[j] == t) ++j;
            if (first) { op_raw(ty, 2, t); first = 0; }
            else op_raw(ty, 1, t == (prev + 1) % 3 ? 0 : 1);   /* mcbtypetrans, h4m:1591-1594 */
            op_uovf(ty, j - i);
            prev = t; i = j;
        }
        /* proc run-length stream over inter MCBs only */
        int have = 0, cur = 0; uint32_t cnt = 0;
        for (i = 0; i < nmb; ++i)
        {
            if (mtype[i] == 0) continue;
            if (!have) { have = 1; cur = mproc[i]; cnt = 1; op_raw(pc, 1, cur); continue; }
            if (mproc[i] == cur) { ++cnt; continue; }
            op_uovf(pc, cnt); cur = mproc[i]; cnt = 1;
        }
        if (have) op_uovf(pc, cnt);
    }
    {
        int runY = 0, runC = 0;
        for (int i = 0; i < nmb; ++i)
        {
            int intra = mtype[i] == 0;
            if (intra)
            {   /* decode_PB_dc, h4m:1649-1662: Y x4, U, V */
                for (int k = 0; k < 4; ++k) op_sovf(&p->ops[S_DC_Y], pick_dc_delta(g));
                op_sovf(&p->ops[S_DC_U], pick_dc_delta(g));
                op_sovf(&p->ops[S_DC_V], pick_dc_delta(g));
            }
            else if (mproc[i] == 1)
                continue;   /* decode_PB_cc with proc==1 reads nothing, h4m:1673-1683 */
            /* decode_PB_cc, h4m:1686-1738 */
            for (int k = 0; k < 4; ++k)
            {
                if (runY) { nib[i][k] = 0; --runY; continue; }
                int t = intra ? pick_intra_type(g) : pick_inter_nibble(g, winok[i]);
                nib[i][k] = t;
                op_sym(&p->ops[S_BN_Y], t);
                if (t == 0) { runY = pick_zero_run(g); op_sym(&p->ops[S_BNR_Y], runY); }
            }
            if (runC) { nib[i][4] = nib[i][5] = 0; --runC; }
            else
            {
                int u = intra ? pick_intra_type(g) : pick_inter_nibble(g, winok[i]);
                int v = intra ? pick_intra_type(g) : pick_inter_nibble(g, winok[i]);
                nib[i][4] = u; nib[i][5] = v;
                int sym = u | (v << 4);
                op_sym(&p->ops[S_BN_C], sym);
                if (sym == 0) { runC = pick_zero_run(g); op_sym(&p->ops[S_BNR_C], runC); }
            }
        }
    }
    /* ---- pass 2: BpicPlaneDec (h4m:1922-1967) ---- */
    {
        int ref = -1, ph = 0, pv = 0;
        for (int i = 0; i < nmb; ++i)
        {
            if (mtype[i] == 0)
            {   /* MCBlockDecDCNest, h4m:1789-1827 */
                for (int k = 0; k < 6; ++k)
                {
                    int pl = k < 4 ? 0 : k - 3, t = nib[i][k];
                    if (t == 6) emit_raw(g, p, pl);
                    else if (t != 0 && t != 8) emit_bases(g, p, pl, t);
                }
                continue;
            }
            int f = mtype[i] - 1;
            if (f != ref) { ref = f; ph = pv = 0; }             /* h4m:1943-1949 */
            emit_mv(&p->ops[S_X0], &ph, mv[i][0], rb[f][0]);
            emit_mv(&p->ops[S_X1], &pv, mv[i][1], rb[f][1]);
            if (mproc[i] == 1) continue;
            for (int k = 0; k < 6; ++k)
            {   /* MCBlockDecMCNest, h4m:1871-1909 */
                int pl = k < 4 ? 0 : k - 3, t = nib[i][k];
                if (t == 6) emit_raw(g, p, pl);
                else if (t != 0)
                {
                    emit_bases(g, p, pl, t - 1);
                    /* S1 (DC offset) and S2 (prediction gain), h4m:1405-1406 */
                    int32_t s1 = dense ? rnd_range(r, -40, 40) : rnd_range(r, -10, 10);
                    int32_t s2 = dense ? rnd_range(r, -48, 48) : rnd_range(r, -8, 8);
                    if (dense && chance(r, 4)) s1 = chance(r, 50) ? rnd_range(r, 127, 260) : -rnd_range(r, 128, 260);
                    if (stress && chance(r, 4)) s2 = chance(r, 50) ? rnd_range(r, 127, 2000) : -rnd_range(r, 128, 2000);
                    op_sovf(&p->dc_pass2[pl], s1);
                    op_sovf(&p->dc_pass2[pl], s2);
                }
            }
        }
    }
    for (int pl = 0; pl < 3; ++pl) ops_append(&p->ops[S_DC[pl]], &p->dc_pass2[pl]);
    free(mtype); free(mproc); free(nib); free(mv); free(winok);
}

/* -- serialisation ---------------------------------------------------- */

static void write_group(Gen *g, Pic *p, Bytes sec[S_COUNT], const int *members, int n)
{
    uint64_t freq[256] = {0};
    for (int m = 0; m < n; ++m)
    {
        const Ops *o = &p->ops[members[m]];
        for (size_t i = 0; i < o->n; ++i)
            if (!(o->v[i] & OP_RAW)) freq[o->v[i] & 0xFF]++;
    }
    Huff *h = malloc(sizeof *h);
    huff_build(h, freq, &g->rng);
    for (int m = 0; m < n; ++m)
    {
        BitW w = {{0}};
        if (m == 0) huff_write_tree(h, h->root, &w);   /* the leader carries the tree, h4m:1994-1999 */
        const Ops *o = &p->ops[members[m]];
        for (size_t i = 0; i < o->n; ++i)
        {
            uint32_t x = o->v[i];
            if (x & OP_RAW) bw_put(&w, (x >> 16) & 0x7F, x & 0xFFFF);
            else huff_put(h, &w, x & 0xFF);
        }
        bw_finish(&w);
        sec[members[m]] = w.out;
    }
    free(h);
}

static void write_picture(Gen *g, Bytes *out, int type)
{
    Pic pic;
    memset(&pic, 0, sizeof pic);
    uint8_t hdr[8];
    int is_i = type == 'I';
    if (is_i) gen_ipic(g, &pic, hdr);
    else gen_pbpic(g, &pic, hdr, type == 'B');

    Bytes sec[S_COUNT];
    memset(sec, 0, sizeof sec);
    static const int g_dc[3] = {S_DC_Y, S_DC_U, S_DC_V};
    static const int g_sc[3] = {S_SC_Y, S_SC_U, S_SC_V};
    static const int g_bn[2] = {S_BN_Y, S_BN_C};
    static const int g_run_i[5] = {S_BNR_Y, S_BNR_C, S_X0, S_X1, S_X2};
    static const int g_run_pb[2] = {S_BNR_Y, S_BNR_C};
    static const int g_mv[2] = {S_X0, S_X1};
    static const int g_mcb[2] = {S_X2, S_X3};
    write_group(g, &pic, sec, g_dc, 3);
    write_group(g, &pic, sec, g_sc, 3);
    write_group(g, &pic, sec, g_bn, 2);
    if (is_i) write_group(g, &pic, sec, g_run_i, 5);
    else
    {
        write_group(g, &pic, sec, g_run_pb, 2);
        write_group(g, &pic, sec, g_mv, 2);
        write_group(g, &pic, sec, g_mcb, 2);
    }
    for (int pl = 0; pl < 3; ++pl) { sec[S_FIX[pl]] = pic.fix[pl]; memset(&pic.fix[pl], 0, sizeof(Bytes)); }

    int nsec = is_i ? 16 : 17;
    putn(out, hdr, 8);
    size_t table = out->n;
    for (int i = 0; i < nsec; ++i) put32(out, 0);
    size_t data = out->n;
    for (int i = 0; i < nsec; ++i)
    {
        poke32(out, table + 4 * i, (uint32_t)(out->n - data));
        put32(out, (uint32_t)sec[i].n);
        if (sec[i].n) putn(out, sec[i].p, sec[i].n);
        while (out->n & 3) put8(out, 0);
    }
    for (int i = 0; i < S_COUNT; ++i) free(sec[i].p);
    pic_free(&pic);
}

GEN_API int h4mgen_generate(const H4MGenParams *prm, uint8_t **out_data, uint64_t *out_len)
{
    if (!prm || !prm->gop || prm->gop[0] != 'I') return -1;
    if (prm->width % 8 || prm->height % 8 || prm->width < 16 || prm->height < 16) return -2;
    if (prm->version != 13 && prm->version != 15) return -4;
    Gen g;
    memset(&g, 0, sizeof g);
    g.prm = prm;
    g.rng.s = prm->seed * 0x9E3779B97F4A7C15ull + 0x1234567;
    g.w = prm->width; g.h = prm->height;
    g.mbw = g.w / 8; g.mbh = g.h / 8;
    g.bw[0] = g.w / 4; g.bh[0] = g.h / 4;
    g.bw[1] = g.bw[2] = g.w / 8; g.bh[1] = g.bh[2] = g.h / 8;

    int gop_len = (int)strlen(prm->gop);
    Bytes f = {0};
    /* file header, h4m:2192-2211 */
    char magic[16] = {0};
    strcpy(magic, prm->version == 13 ? "HVQM4 1.3" : "HVQM4 1.5");
    putn(&f, magic, 16);
    put32(&f, 0x44);
    put32(&f, 0);                                  /* body size, patched below */
    put32(&f, prm->n_gops);
    put32(&f, prm->n_gops * gop_len);
    put32(&f, 0);                                  /* audio frames */
    put32(&f, prm->usec_per_frame ? prm->usec_per_frame : 33367);
    put32(&f, 0);                                  /* max frame size, patched below */
    put32(&f, 0);
    put32(&f, 0);                                  /* audio frame size */
    put16(&f, g.w); put16(&f, g.h);
    put8(&f, 2); put8(&f, 2); put8(&f, 0); put8(&f, 0);
    put8(&f, 0); put8(&f, 0); put8(&f, 0); put8(&f, 0);
    put32(&f, 0);                                  /* audio sample rate */

    uint32_t max_frame = 0;
    for (int gi = 0; gi < prm->n_gops; ++gi)
    {
        size_t gop_hdr = f.n;
        put32(&f, 0); put32(&f, 0); put32(&f, gop_len); put32(&f, 0); put32(&f, 0x01000000);
        size_t gop_data = f.n;
        /* display ids: anchors (I/P) are displayed after the B pictures that follow them in decode order */
        int *disp = malloc(sizeof(int) * gop_len);
        {
            int next = 0, pending = -1;
            for (int i = 0; i < gop_len; ++i)
            {
                if (prm->gop[i] == 'B') disp[i] = next++;
                else
                {
                    if (pending >= 0) disp[pending] = next++;
                    pending = i;
                }
            }
            if (pending >= 0) disp[pending] = next++;
        }
        for (int i = 0; i < gop_len; ++i)
        {
            int t = prm->gop[i];
            if (t != 'I' && t != 'P' && t != 'B') { free(disp); free(f.p); return -5; }
            put16(&f, 1);
            put16(&f, t == 'I' ? 0x10 : t == 'P' ? 0x20 : 0x30);
            size_t size_at = f.n;
            put32(&f, 0);
            size_t start = f.n;
            put32(&f, disp[i]);
            write_picture(&g, &f, t);
            uint32_t sz = (uint32_t)(f.n - start);
            poke32(&f, size_at, sz);
            if (sz > max_frame) max_frame = sz;
        }
        free(disp);
        poke32(&f, gop_hdr + 4, (uint32_t)(f.n - gop_data));
    }
    poke32(&f, 0x14, (uint32_t)(f.n - 0x44));
    poke32(&f, 0x28, max_frame);
    /* slack so that word-wise readers may overread the last record (h4m:2080-2082) */
    bytes_reserve(&f, 8);
    memset(f.p + f.n, 0, 8);
    *out_data = f.p;
    *out_len = f.n;
    return 0;
}

GEN_API void h4mgen_free(uint8_t *p) { free(p); }

#ifdef H4MGEN_MAIN
int main(int argc, char **argv)
{
    if (argc < 9)
    {
        fprintf(stderr, "usage: %s out.h4m W H version(13|15) gop n_gops seed profile(0 dense|1 realistic|2 stress|3,4 capacity)\n", argv[0]);
        return 2;
    }
    H4MGenParams p = {atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[6]), atoi(argv[8]), 0,
                      strtoull(argv[7], 0, 0), argv[5]};
    uint8_t *d; uint64_t n;
    int rc = h4mgen_generate(&p, &d, &n);
    if (rc) { fprintf(stderr, "h4mgen: error %d\n", rc); return 1; }
    FILE *f = fopen(argv[1], "wb");
    if (!f) { perror(argv[1]); return 1; }
    fwrite(d, 1, n, f);
    fclose(f);
    h4mgen_free(d);
    return 0;
}
#endif
