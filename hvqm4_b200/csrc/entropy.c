/*
 * entropy.c -- host serial stage: HVQM4 picture bitstream -> symbol buffer (symbuf.h).
 *
 * Written from scratch; the reference lines each routine corresponds to are cited as
 * "h4m:N" (= /root/reference/h4m_audio_decode.c line N).  Differences from the
 * reference by design:
 *   - the bit reader is a 64-bit left-aligned window refilled 8 bytes at a time, and
 *     is bounded by the declared section size (the reference reads BE32 words with no
 *     bounds, h4m:552-602); both consume the same MSB-first bit sequence;
 *   - Huffman codes are decoded through a 2^10-entry prefix table rebuilt per tree,
 *     falling back to a node walk for longer codes (the reference walks the tree bit
 *     by bit, h4m:644-651);
 *   - all parsing runs to completion before any pixel work: pass 2 of a P/B picture
 *     (h4m:1922-1967) only resolves motion vectors and copies each block's side data
 *     into its slot of the work-ordered symbol buffer;
 *   - malformed input raises SYM_ERR_* bits instead of reading out of bounds.
 */
#if defined(H4E_DEVICE)
#include <stddef.h>
#include <stdint.h>
#include "symbuf.h"
typedef struct H4Seq H4Seq;     /* the device build defines the entry points as static __device__ functions */
#else
#include "entropy.h"
#endif

#include <stdlib.h>
#include <string.h>

/*
 * This file is compiled twice: by gcc as the host serial stage, and -- #include'd by
 * entropy_dev.cu with H4E_DEVICE defined -- by nvcc as __device__ code, so that the very same
 * parser can also run on the GPU (one picture per warp; see entropy_dev.cu).  The macros below
 * are the only difference: function qualifiers, table storage, and where scratch memory comes
 * from (the device build never allocates: every buffer is carved from a per-stream arena).
 */
#if defined(H4E_DEVICE)
/* on the GPU the entry points are WARP-COLLECTIVE: all 32 lanes call them; serial parts run on
   lane 0, the data-parallel parts (flat section decode, record fill, copies) on all lanes */
#define H4E_LANE ((int)(threadIdx.x & 31))
#define H4E_LANES 32
#define H4E_SYNC() __syncwarp()
#define H4E_ERR(s, bits) atomicOr(&(s)->err, (uint32_t)(bits))
#define H4E_FETCH_ADD(ptr, v) atomicAdd((ptr), (uint32_t)(v))
#define H4E_FN static __device__
#define H4E_INL static __device__ __forceinline__
#define H4E_TABLE static __device__ const
#define H4E_API static __device__
#define H4E_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define H4E_LANE 0
#define H4E_LANES 1
#define H4E_SYNC() ((void)0)
#define H4E_ERR(s, bits) ((s)->err |= (uint32_t)(bits))
#define H4E_FETCH_ADD(ptr, v) h4e_fetch_add((ptr), (uint32_t)(v))
#define H4E_FN static
#define H4E_INL static inline
#define H4E_TABLE static const
#define H4E_API
#define H4E_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

#if !defined(H4E_DEVICE)
static inline uint32_t h4e_fetch_add(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
#endif


#if defined(H4E_DEVICE)   /* GPU build: per-phase clock64() totals, read back by tools/profile_e2e.py */
__device__ unsigned long long h4e_dev_prof[8];
#define PROF_T0() long long prof_t = clock64()
#define PROF_ADD(i) do { long long n_ = clock64(); if (H4E_LANE == 0) atomicAdd(&h4e_dev_prof[i], (unsigned long long)(n_ - prof_t)); prof_t = n_; } while (0)
#elif defined(H4E_PROFILE)   /* developer aid: cycle counters per phase (tools only, never defined in the product build) */
#include <x86intrin.h>
unsigned long long h4e_prof[8];
#define PROF_T0() unsigned long long prof_t = __rdtsc()
#define PROF_ADD(i) do { unsigned long long n_ = __rdtsc(); h4e_prof[i] += n_ - prof_t; prof_t = n_; } while (0)
#else
#define PROF_T0() do { } while (0)
#define PROF_ADD(i) do { } while (0)
#endif

/* ------------------------------------------------------------------ bit reader */

typedef struct
{
    const uint8_t *base, *p, *end;
    uint64_t buf;   /* unread bits, left aligned */
    int n;          /* number of valid bits in buf */
} BR;

H4E_INL uint32_t rd_be32(const uint8_t *p)
{
    return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
}

H4E_INL void br_open(BR *b, const uint8_t *base, uint32_t size)
{
    b->base = b->p = base;
    b->end = base ? base + size : base;
    b->buf = 0;
    b->n = 0;
}

H4E_INL void br_refill(BR *b)
{
    if (b->end - b->p >= 8)
    {
        uint64_t w;
#if defined(H4E_DEVICE)
        {   /* two aligned 8-byte loads around the (arbitrarily aligned) pointer, funnel-shifted and
               byte-swapped; this may touch up to 15 bytes past p + 8, which the batch runtime's
               16 bytes of padding behind every picture cover */
            const uintptr_t a = (uintptr_t)b->p;
            const unsigned long long *q = (const unsigned long long *)(a & ~(uintptr_t)7);
            const unsigned sh = (unsigned)(a & 7) * 8;
            if (!((uintptr_t)q & 127)) asm volatile("prefetch.global.L1 [%0];" ::"l"(q + 32));   /* two lines ahead */
            const unsigned long long lo = q[0], hi = q[1];
            const unsigned long long x = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
            w = (uint64_t)__byte_perm((unsigned)x, 0, 0x0123) << 32 | __byte_perm((unsigned)(x >> 32), 0, 0x0123);
        }
#else
        memcpy(&w, b->p, 8);
        w = __builtin_bswap64(w);
#endif
        b->buf |= w >> b->n;
        int adv = (63 - b->n) >> 3;
        b->p += adv;
        b->n += adv * 8;
    }
    else
    {
        while (b->n <= 56)
        {
            uint64_t byte = b->p < b->end ? *b->p : 0;   /* zeros past the end; checked by br_overrun() */
            b->p++;
            b->buf |= byte << (56 - b->n);
            b->n += 8;
        }
    }
}

H4E_INL uint32_t br_bit(BR *b)
{
    if (b->n < 1) br_refill(b);
    uint32_t v = (uint32_t)(b->buf >> 63);
    b->buf <<= 1;
    b->n -= 1;
    return v;
}

H4E_INL uint32_t br_bits(BR *b, int k)   /* 0 <= k <= 16 */
{
    if (k == 0) return 0;
    if (b->n < k) br_refill(b);
    uint32_t v = (uint32_t)(b->buf >> (64 - k));
    b->buf <<= k;
    b->n -= k;
    return v;
}

/* bits consumed so far, and whether the reader went past the section */
H4E_INL int64_t br_pos(const BR *b) { return (int64_t)(b->p - b->base) * 8 - b->n; }
H4E_INL int br_overrun(const BR *b) { return br_pos(b) > (int64_t)(b->end - b->base) * 8; }

/* ------------------------------------------------------------------ Huffman (h4m:385-394, 607-651) */

#define HT_BITS 10

typedef struct { int16_t val; uint8_t len; uint8_t walk; } HEnt;

typedef struct
{
    int32_t leaf[256];       /* value per leaf byte; persists across pictures like Tree.array[0][0..255] */
    int16_t kid[2][256];     /* children of internal node i (node id 256+i) */
    int root;
    int used;
    int bad;
    HEnt tab[1 << HT_BITS];
} HTab;

/* _readTree, h4m:607-630, without recursion (the same code runs as a GPU thread): `stk` holds
   the internal nodes whose subtrees are still open; bit 15 marks "0-side done". */
H4E_FN int ht_parse(HTab *t, BR *b, int is_signed, int scale)
{
    uint16_t stk[260];
    int sp = 0;
    for (;;)
    {
        int id;
        if (br_bit(b) == 0)
        {
            const uint32_t byte = br_bits(b, 8);
            const int32_t v = (is_signed && byte > 0x7F) ? (int32_t)byte - 256 : (int32_t)byte;
            t->leaf[byte] = (int32_t)(int16_t)((uint32_t)v << scale);   /* int16_t symbol <<= scale, h4m:613-617 */
            id = (int)byte;
        }
        else
        {
            if (t->used >= 256 || sp >= 258 || br_overrun(b))
            {
                t->bad = 1;
                return 0;
            }
            stk[sp++] = (uint16_t)t->used++;
            continue;                                   /* read the 0 side of the new node */
        }
        /* a subtree is complete: hang it under the innermost open node, closing nodes as they fill */
        for (;;)
        {
            if (sp == 0) return id;
            const int node = stk[sp - 1] & 0x7FFF;
            if (!(stk[sp - 1] & 0x8000))
            {
                t->kid[0][node] = (int16_t)id;
                stk[sp - 1] |= 0x8000;
                break;                                  /* now read the 1 side */
            }
            t->kid[1][node] = (int16_t)id;
            --sp;
            id = node + 256;
        }
    }
}

/* prefix table: every leaf at depth <= HT_BITS fills its 2^(HT_BITS-depth) entries; deeper
   subtrees get one "walk from this node" entry */
H4E_FN void ht_fill(HTab *t)
{
    struct { int16_t node; uint8_t depth; uint32_t code; } stk[2 * HT_BITS + 4];
    int sp = 0;
    stk[sp].node = (int16_t)t->root; stk[sp].depth = 0; stk[sp].code = 0; ++sp;
    while (sp)
    {
        --sp;
        const int node = stk[sp].node, depth = stk[sp].depth;
        const uint32_t code = stk[sp].code;
        if (node < 256)
        {
            HEnt e;
            e.val = (int16_t)t->leaf[node]; e.len = (uint8_t)depth; e.walk = 0;
            const uint32_t lo = code << (HT_BITS - depth), hi = (code + 1) << (HT_BITS - depth);
            for (uint32_t i = lo; i < hi; ++i) t->tab[i] = e;
        }
        else if (depth == HT_BITS)
        {
            HEnt e;
            e.val = (int16_t)node; e.len = HT_BITS; e.walk = 1;
            t->tab[code] = e;
        }
        else
        {
            stk[sp].node = t->kid[1][node - 256]; stk[sp].depth = (uint8_t)(depth + 1); stk[sp].code = code << 1 | 1; ++sp;
            stk[sp].node = t->kid[0][node - 256]; stk[sp].depth = (uint8_t)(depth + 1); stk[sp].code = code << 1; ++sp;
        }
    }
}

/* readTree, h4m:632-642: an empty leader section leaves root = 0, i.e. every symbol
   decodes to the stale leaf[0] without consuming bits. */
H4E_FN void ht_read(HTab *t, BR *leader, uint32_t leader_size, int is_signed, int scale)
{
    t->used = 0;
    t->bad = 0;
    t->root = leader_size ? ht_parse(t, leader, is_signed, scale) : 0;
    if (t->bad) t->root = 0;
    ht_fill(t);
}

H4E_INL int32_t ht_get(const HTab *t, BR *b)
{
    if (b->n < 32) br_refill(b);
    HEnt e;
    memcpy(&e, &t->tab[b->buf >> (64 - HT_BITS)], sizeof e);   /* one 32-bit load */
    b->buf <<= e.len;
    b->n -= e.len;
    if (!e.walk) return e.val;
    int node = e.val;
    while (node >= 256) node = t->kid[br_bit(b)][node - 256];
    return t->leaf[node];
}

/* decodeSOvfSym, h4m:654-664 */
H4E_INL int32_t ht_get_sovf(const HTab *t, BR *b, int32_t lo, int32_t hi)
{
    int32_t sum = 0, v;
    do
    {
        v = ht_get(t, b);
        sum += v;
    } while ((v <= lo || v >= hi) && !br_overrun(b));
    return sum;
}

/* decodeUOvfSym, h4m:667-677 */
H4E_INL int32_t ht_get_uovf(const HTab *t, BR *b)
{
    int32_t sum = 0, v;
    do
    {
        v = ht_get(t, b);
        sum += v;
    } while (v >= 255 && !br_overrun(b));
    return sum;
}

/* ------------------------------------------------------------------ flat symbol streams
 * Sections that hold nothing but Huffman symbols are decoded in ONE tight loop each, straight
 * after the trees are known, into arrays; the syntax loops below then consume plain array
 * elements.  Decoding a section needs no syntax knowledge (one tree per section), and taking
 * the serial bit-reader dependency chain out of the branchy syntax code is worth ~1.6x on the
 * host stage.  Sections that mix raw bits with symbols (mcb_type, mcb_proc, mv_h, mv_v) stay
 * interleaved; they are small. */

typedef struct
{
    int32_t *v;          /* decoded values (escape-extended sums for the DC sections) */
    uint32_t pos, n, cap;
    int32_t cval;        /* value of a constant stream */
    uint8_t is_const;    /* single-leaf or absent tree: every symbol costs 0 bits (h4m:638-650) */
    uint8_t over;        /* consumer ran past the end */
} SymStream;

H4E_INL int32_t ss_get(SymStream *q)
{
    if (q->pos < q->n) return q->v[q->pos++];
    if (q->is_const) return q->cval;
    q->over = 1;
    return 0;
}

H4E_FN int ss_reserve(SymStream *q, uint32_t need)
{
    if (need <= q->cap) return 1;
#if defined(H4E_DEVICE)
    return 0;                                            /* fixed capacity on the GPU */
#else
    uint32_t cap = q->cap ? q->cap : 1024;
    while (cap < need) cap *= 2;
    int32_t *nv = (int32_t *)realloc(q->v, (size_t)cap * sizeof(int32_t));
    if (!nv) return 0;
    q->v = nv;
    q->cap = cap;
    return 1;
#endif
}

/* plain symbols until the section's bits are exhausted (trailing pad bits may yield a few extra
   symbols that nobody consumes) */
H4E_FN void ss_decode(SymStream *q, const HTab *t, BR *b)
{
    q->pos = q->n = 0;
    q->over = 0;
    q->is_const = !t->tab[0].walk && t->tab[0].len == 0;
    q->cval = q->is_const ? t->tab[0].val : 0;
    if (q->is_const || !b->base) return;
    const int64_t end_bits = (int64_t)(b->end - b->base) * 8;
    uint32_t n = 0;
    while (br_pos(b) < end_bits)
    {
        if (n + 64 > q->cap && !ss_reserve(q, n + 64)) break;
        const uint32_t stop = q->cap - 8 < n + 4096 ? q->cap - 8 : n + 4096;
        int32_t *v = q->v;
        /* inner loop without capacity checks */
        while (n < stop && br_pos(b) < end_bits) v[n++] = ht_get(t, b);
    }
    q->n = n;
}

/* Two sections decoded in lock step: each section's decode is one serial dependency chain
   (window -> table entry -> length -> window); running two chains in the same loop roughly
   doubles the symbols per cycle.  sovf0/sovf1 select escape-summed value decoding. */
H4E_FN void ss_decode2(SymStream *q0, const HTab *t0, BR *b0, int sovf0,
                       SymStream *q1, const HTab *t1, BR *b1, int sovf1, int32_t lo, int32_t hi);

/* signed escape-extended values (decodeSOvfSym, h4m:654-664): a value ends at the first symbol
   strictly inside (lo, hi); a trailing unfinished escape run is dropped */
H4E_FN void ss_decode_sovf(SymStream *q, const HTab *t, BR *b, int32_t lo, int32_t hi)
{
    q->pos = q->n = 0;
    q->over = 0;
    q->is_const = !t->tab[0].walk && t->tab[0].len == 0;
    q->cval = q->is_const ? t->tab[0].val : 0;
    if (q->is_const)
    {   /* a constant escape symbol would never terminate in the reference either */
        if (q->cval <= lo || q->cval >= hi) q->cval = 0;
        return;
    }
    if (!b->base) return;
    const int64_t end_bits = (int64_t)(b->end - b->base) * 8;
    uint32_t n = 0;
    int32_t sum = 0;
    while (br_pos(b) < end_bits)
    {
        if (n + 64 > q->cap && !ss_reserve(q, n + 64)) break;
        const uint32_t stop = q->cap - 8 < n + 4096 ? q->cap - 8 : n + 4096;
        int32_t *v = q->v;
        while (n < stop && br_pos(b) < end_bits)
        {
            const int32_t x = ht_get(t, b);
            sum += x;
            if (x > lo && x < hi)
            {
                v[n++] = sum;
                sum = 0;
            }
        }
    }
    q->n = n;
}

#if defined(H4E_DEVICE)
/* GPU: every lane of the warp decodes its own section (q == NULL: idle lane) through one common
   instruction stream, so that the lanes run in SIMT lock step instead of one after the other */
H4E_FN void ss_decode_lane(SymStream *q, const HTab *t, BR *b, int sovf, int32_t lo, int32_t hi)
{
    int live = 0;
    int64_t end_bits = 0;
    uint32_t n = 0, cap = 0;
    int32_t sum = 0;
    int32_t *v = 0;
    BR x;
    memset(&x, 0, sizeof x);
    if (q)
    {
        q->pos = q->n = 0;
        q->over = 0;
        q->is_const = !t->tab[0].walk && t->tab[0].len == 0;
        q->cval = q->is_const ? t->tab[0].val : 0;
        if (q->is_const)
        {
            if (sovf && (q->cval <= lo || q->cval >= hi)) q->cval = 0;
        }
        else if (b->base && q->cap > 8)
        {
            x = *b;
            end_bits = (int64_t)(b->end - b->base) * 8;
            v = q->v;
            cap = q->cap - 8;
            live = br_pos(&x) < end_bits;
        }
    }
    /* Fast part: a 2 x 32-bit window fed by aligned words and a plain count of the bits left,
       while the section still holds more bits than any symbol can take (a code is at most 256
       deep); the generic reader below finishes the tail, where bits past the end must read as 0. */
    {
        const int64_t left64 = end_bits - br_pos(&x);
        int fast = live && left64 > 600 && left64 < (1ll << 30);
        if (fast)
        {   /* word-align the read pointer: hand whole bytes of the window back (the sections that
               lead with their tree arrive with a full window), or feed single bytes forward */
            const int k = (int)((uintptr_t)x.p & 3);
            if (k && x.n >= 8 * k)
            {
                x.p -= k;
                x.n -= 8 * k;
            }
            while (((uintptr_t)x.p & 3) && x.n <= 56)
            {
                x.buf |= (uint64_t)*x.p++ << (56 - x.n);
                x.n += 8;
            }
            fast = !((uintptr_t)x.p & 3);
        }
        uint32_t whi = (uint32_t)(x.buf >> 32), wlo = (uint32_t)x.buf;
        int nb = x.n, left = fast ? (int)left64 : 0;
        const uint32_t *wp = (const uint32_t *)x.p;
        const uint32_t *tab32 = t ? (const uint32_t *)t->tab : 0;
        fast = fast && left > 600 && n < cap;
        if (fast)
        {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(wp));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + 32));
        }
        while (fast)
        {
            if (nb <= 32)
            {
                /* the lanes run in lock step, so every lane's cache miss stalls all of them:
                   fetch each 128-byte line of the section two lines ahead of its first use */
                if (!((uintptr_t)wp & 127)) asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + 64));
                const uint32_t w = __byte_perm(*wp++, 0, 0x0123);
                whi |= __funnelshift_rc(w, 0, nb);
                wlo = __funnelshift_rc(0, w, nb);
                nb += 32;
            }
            const uint32_t e = tab32[whi >> (32 - HT_BITS)];
            const uint32_t len = (e >> 16) & 0xFF;
            int32_t a = (int16_t)(e & 0xFFFF);
            whi = __funnelshift_l(wlo, whi, len);
            wlo <<= len;
            nb -= (int)len;
            left -= (int)len;
            if (e >> 24)
            {   /* code longer than the table: walk the tree bit by bit */
                int node = a;
                while (node >= 256)
                {
                    if (nb == 0)
                    {
                        whi = __byte_perm(*wp++, 0, 0x0123);
                        wlo = 0;
                        nb = 32;
                    }
                    const uint32_t bit = whi >> 31;
                    whi = __funnelshift_l(wlo, whi, 1);
                    wlo <<= 1;
                    --nb;
                    --left;
                    node = t->kid[bit][node - 256];
                }
                a = t->leaf[node];
            }
            sum += a;
            const int done = !sovf || (a > lo && a < hi);
            v[n] = sum;
            n += (uint32_t)done;
            sum = done ? 0 : sum;
            fast = left > 600 && n < cap;
        }
        if (v)
        {
            x.buf = (uint64_t)whi << 32 | wlo;
            x.n = nb;
            x.p = (const uint8_t *)wp;
            live = live && br_pos(&x) < end_bits && n < cap;
        }
    }
    while (live)
    {
        const int32_t a = ht_get(t, &x);
        sum += a;
        const int done = !sovf || (a > lo && a < hi);
        v[n] = sum;
        n += (uint32_t)done;
        sum = done ? 0 : sum;
        live = br_pos(&x) < end_bits && n < cap;
    }
    if (v)
    {
        q->n = n;
        *b = x;
    }
}
#endif

H4E_FN void ss_decode2(SymStream *q0, const HTab *t0, BR *b0, int sovf0,
                       SymStream *q1, const HTab *t1, BR *b1, int sovf1, int32_t lo, int32_t hi)
{
    const int c0 = !t0->tab[0].walk && t0->tab[0].len == 0, c1 = !t1->tab[0].walk && t1->tab[0].len == 0;
    if (!c0 && !c1 && b0->base && b1->base)
    {
        q0->pos = q0->n = 0; q0->over = 0; q0->is_const = 0; q0->cval = 0;
        q1->pos = q1->n = 0; q1->over = 0; q1->is_const = 0; q1->cval = 0;
        const int64_t e0 = (int64_t)(b0->end - b0->base) * 8, e1 = (int64_t)(b1->end - b1->base) * 8;
        /* a symbol costs at least one bit: capacity for the worst case of the lock-step part */
        BR x = *b0, y = *b1;
        uint32_t n0 = 0, n1 = 0;
        int32_t s0 = 0, s1 = 0;
        for (;;)
        {
            if ((n0 + 4100 > q0->cap && !ss_reserve(q0, n0 + 8200)) || (n1 + 4100 > q1->cap && !ss_reserve(q1, n1 + 8200))) break;
            int32_t *v0 = q0->v, *v1 = q1->v;
            int k = 0;
            for (; k < 4096 && br_pos(&x) < e0 && br_pos(&y) < e1; ++k)
            {
                const int32_t a = ht_get(t0, &x), c = ht_get(t1, &y);
                if (sovf0)
                {
                    s0 += a;
                    v0[n0] = s0;
                    const int done = a > lo && a < hi;
                    n0 += (uint32_t)done;
                    s0 = done ? 0 : s0;
                }
                else
                    v0[n0++] = a;
                if (sovf1)
                {
                    s1 += c;
                    v1[n1] = s1;
                    const int done = c > lo && c < hi;
                    n1 += (uint32_t)done;
                    s1 = done ? 0 : s1;
                }
                else
                    v1[n1++] = c;
            }
            if (k < 4096) break;
        }
        /* the longer section finishes alone (an unfinished escape run carries over in s0/s1) */
        while (br_pos(&x) < e0)
        {
            if (n0 + 8 > q0->cap && !ss_reserve(q0, n0 + 4096)) break;
            const int32_t a = ht_get(t0, &x);
            if (sovf0) { s0 += a; if (a > lo && a < hi) { q0->v[n0++] = s0; s0 = 0; } }
            else q0->v[n0++] = a;
        }
        while (br_pos(&y) < e1)
        {
            if (n1 + 8 > q1->cap && !ss_reserve(q1, n1 + 4096)) break;
            const int32_t c = ht_get(t1, &y);
            if (sovf1) { s1 += c; if (c > lo && c < hi) { q1->v[n1++] = s1; s1 = 0; } }
            else q1->v[n1++] = c;
        }
        q0->n = n0; q1->n = n1;
        *b0 = x; *b1 = y;
        return;
    }
    if (sovf0) ss_decode_sovf(q0, t0, b0, lo, hi); else ss_decode(q0, t0, b0);
    if (sovf1) ss_decode_sovf(q1, t1, b1, lo, hi); else ss_decode(q1, t1, b1);
}

/* ------------------------------------------------------------------ stream state */

enum { T_DC = 0, T_RUN = 1, T_SCALE = 2, T_BNUM = 3, T_MV = 4, T_MCB = 5 };   /* tree sharing, h4m:977-999 */

typedef struct { const uint8_t *base; uint32_t size, pos; } ByteSec;   /* fixvl: plain byte stream */

/* one scheduled record (see schedule_record / fill_records below) */
typedef struct Work
{
    uint32_t at;                       /* record position (words from rec_base) */
    uint32_t fix_off, sc_off, dcv_off; /* consumption offsets in fixvl[p] (bytes), sc[p], dcv[p] (values) */
    uint16_t len;
    uint8_t cls, plane;
} Work;

struct H4Seq
{
    int width, height, version15;
    int bw[3], bh[3], stride[3];
    int mbw, mbh, nseg;
    size_t map_cells[3];
    uint8_t *type[3], *dc[3];            /* bordered, persistent (h4m:1001-1040) */
    uint8_t nest[SYM_NEST_BYTES];        /* packed nibbles of the last I picture's nest */
    HTab tree[6];
    int nbands, ngroups;                 /* record groups: [class][band][length bucket] */
    int bshift;                          /* log2 of the macroblock rows per record band (0 or 3) */
    uint32_t *grp_count, *grp_base, *grp_next, *grp_chunk, *grp_ord;
    uint8_t *mcb_tag;                    /* P/B pass 1, split form: type/proc tag of every macroblock ... */
    uint32_t *mcb_list, n_list;          /* ... and the macroblocks that carry block types, in bitstream order */
    int32_t *mv_raw;                     /* split pass 2: accumulated vector (h, v) of every macroblock */
    uint32_t *chunks;                    /* chunk table under construction (2 words per chunk) */
    uint32_t chunks_cap;
    uint32_t *band_first;                /* [SYM_REC_CLASSES][nbands + 1] */
    uint32_t errors_total;

    /* per picture, between parse_begin and parse_finish */
    int pic_type;
    uint32_t err;
    uint32_t n_rec_words, n_chunks, n_chunks_nest, n_records;
    uint32_t *rec_base;                  /* records region inside the blob being written */
    uint32_t n_inter_mcb;
    int need_nest;
    int dc_shift, unk_shift, rb[2][2];
    int32_t dc_lo, dc_hi;
    BR bn[2], bnr[2], dcv[3], sc[3], rle[3], mvh, mvv, mcbt, mcbp;
    SymStream q_bn[2], q_bnr[2], q_dcv[3], q_sc[3], q_rle[3];   /* flat-decoded sections (buffers persist) */
    ByteSec fix[3];
    size_t blob_bytes;
    SymHeader hdr;
    struct Work *dev_work;               /* GPU build only: per-stream record schedule scratch */
    uint32_t dev_work_cap;
    struct Work *cur_work;               /* schedule of the picture being finished */
    int setup_ok, nest_x, nest_y;
    int split_schedule;                  /* host only: run the GPU's pb_mvs + schedule_rows split (tests) */
};

H4E_INL size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

/* record class/length of every possible type byte: [is_ipic][type] = len | cls << 16 (0 = none) */
#if defined(H4E_DEVICE)
H4E_INL uint32_t rec_lut(int is_ipic, uint32_t t)
{
    int cls = 0;
    const uint32_t len = sym_record_len(t & 0xFF, is_ipic, &cls);
    return len ? (len | (uint32_t)cls << 16) : 0u;
}
#else
static uint32_t g_rec_lut[2][256];
static void init_rec_lut(void)
{
    for (int ip = 0; ip < 2; ++ip)
        for (int t = 0; t < 256; ++t)
        {
            int cls = 0;
            const uint32_t len = sym_record_len((uint32_t)t, ip, &cls);
            g_rec_lut[ip][t] = len ? (len | (uint32_t)cls << 16) : 0;
        }
}
H4E_INL uint32_t rec_lut(int is_ipic, uint32_t t) { return g_rec_lut[is_ipic][t & 0xFF]; }
#endif

/* geometry checks shared by both builds; 4:2:0 only: the only sampling HVQM4 content uses (h4m:872,896).  Portrait
   pictures (width < height: the nest is 38 x 70 and the axes of the basis descriptors swap, h4m:700-711, 743-754,
   965-975, 1865-1868) are decoded like the reference does; upstream calls that orientation untested (README:23). */
H4E_FN int seq_geometry_ok(int width, int height, int h_samp, int v_samp)
{
    if (width <= 0 || height <= 0 || (width & 7) || (height & 7) || h_samp != 2 || v_samp != 2) return 0;
    return width <= 8192 && height <= 8192;
}

/* Macroblock rows per record band of the streams created from now on: 8 (default: the band kernel's unit, few and
   full chunks) or 1 (a kernel can then take any number of consecutive macroblock rows as its unit: sweep and row
   kernels).  Process-wide; the device build keeps its own copy (entropy_dev.cu sets it before it creates streams). */
#if defined(H4E_DEVICE)
__device__ int g_h4e_band_shift = 3;
#else
static int g_h4e_band_shift = 3;
void h4e_set_band_rows(int rows) { g_h4e_band_shift = rows <= 1 ? 0 : rows <= 4 ? 2 : 3; }
#endif

H4E_FN void seq_set_dims(H4Seq *s, int width, int height, int version15)
{
    s->width = width;
    s->height = height;
    s->version15 = version15 ? 1 : 0;
    s->mbw = width / 8;
    s->mbh = height / 8;
    s->nseg = (s->mbw + SYM_SEG_MCBS - 1) / SYM_SEG_MCBS;
    for (int p = 0; p < 3; ++p)
    {
        const int sh = p ? 1 : 0;
        s->bw[p] = (width >> sh) / 4;
        s->bh[p] = (height >> sh) / 4;
        s->stride[p] = s->bw[p] + 2;
        s->map_cells[p] = (size_t)s->stride[p] * (s->bh[p] + 2);
    }
    s->bshift = g_h4e_band_shift;
    s->nbands = (s->mbh + (1 << s->bshift) - 1) >> s->bshift;
    s->ngroups = SYM_REC_CLASSES * s->nbands * SYM_LEN_BUCKETS;
    /* upper bound on chunks: one per group plus one per 32 blocks, or one per block in the long bucket */
    s->chunks_cap = (uint32_t)s->ngroups + (uint32_t)(s->bw[0] * s->bh[0] + 2 * s->bw[1] * s->bh[1]);
}

/* Assigns every per-stream array a slice of one allocation (mem == NULL: only measures).
   sym_cap / work_cap: fixed capacities of the flat symbol streams (symbols per block of the
   section's plane) and of the record schedule (GPU build; the host build grows those on demand
   and passes 0).  A section with more symbols than its capacity decodes as truncated. */
H4E_FN size_t seq_carve(H4Seq *s, uint8_t *mem, uint32_t sym_cap, uint32_t work_cap)
{
    size_t at = 0;
#define CARVE(ptr, type, count) do { if (mem) (ptr) = (type *)(mem + at); at = align16(at + (size_t)(count) * sizeof(type)); } while (0)
    for (int p = 0; p < 3; ++p)
    {
        CARVE(s->type[p], uint8_t, s->map_cells[p]);
        CARVE(s->dc[p], uint8_t, s->map_cells[p]);
    }
    CARVE(s->grp_count, uint32_t, s->ngroups);
    CARVE(s->grp_base, uint32_t, s->ngroups);
    CARVE(s->grp_next, uint32_t, s->ngroups);
    CARVE(s->grp_chunk, uint32_t, s->ngroups);
    CARVE(s->grp_ord, uint32_t, s->ngroups);
    CARVE(s->chunks, uint32_t, (size_t)s->chunks_cap * 2);
    CARVE(s->band_first, uint32_t, (size_t)SYM_REC_CLASSES * (s->nbands + 1));
    CARVE(s->mcb_tag, uint8_t, (size_t)s->mbw * s->mbh);
    CARVE(s->mcb_list, uint32_t, (size_t)s->mbw * s->mbh);
    CARVE(s->mv_raw, int32_t, (size_t)s->mbw * s->mbh * 2);
    if (sym_cap)
    {
        SymStream *all[13] = {&s->q_bn[0], &s->q_bn[1], &s->q_bnr[0], &s->q_bnr[1], &s->q_dcv[0], &s->q_dcv[1], &s->q_dcv[2],
                              &s->q_sc[0], &s->q_sc[1], &s->q_sc[2], &s->q_rle[0], &s->q_rle[1], &s->q_rle[2]};
        /* plane of every section above: bn/bnr 1 carry the U and V nibbles of the chroma blocks */
        static const uint8_t plane_of[13] = {0, 1, 0, 1, 0, 1, 2, 0, 1, 2, 0, 1, 2};
        for (int i = 0; i < 13; ++i)
        {
            const int p = plane_of[i];
            const uint32_t cap = (uint32_t)(s->bw[p] * s->bh[p]) * sym_cap + 1024;
            CARVE(all[i]->v, int32_t, cap);
            if (mem) all[i]->cap = cap;
        }
    }
    if (work_cap)
    {
        CARVE(s->dev_work, struct Work, work_cap);
        if (mem) s->dev_work_cap = work_cap;
    }
#undef CARVE
    return at;
}

/* border cells {0x7F, 0xFF} (h4m:951-955); everything else starts zeroed */
H4E_FN void seq_init_maps(H4Seq *s)
{
    for (int p = 0; p < 3; ++p)
    {
        memset(s->type[p], 0, s->map_cells[p]);
        memset(s->dc[p], 0, s->map_cells[p]);
        for (int y = 0; y < s->bh[p] + 2; ++y)
            for (int x = 0; x < s->stride[p]; ++x)
                if (y == 0 || y == s->bh[p] + 1 || x == 0 || x == s->stride[p] - 1)
                {
                    s->type[p][y * s->stride[p] + x] = 0xFF;
                    s->dc[p][y * s->stride[p] + x] = 0x7F;
                }
    }
}

#if !defined(H4E_DEVICE)
H4Seq *h4e_seq_create(int width, int height, int h_samp, int v_samp, int version15)
{
    if (!g_rec_lut[0][1]) init_rec_lut();   /* idempotent; identical values from every thread */
    if (!seq_geometry_ok(width, height, h_samp, v_samp)) return NULL;
    H4Seq tmp;
    memset(&tmp, 0, sizeof tmp);
    seq_set_dims(&tmp, width, height, version15);
    const size_t bytes = seq_carve(&tmp, NULL, 0, 0);
    uint8_t *mem = (uint8_t *)calloc(1, align16(sizeof(H4Seq)) + bytes);
    if (!mem) return NULL;
    H4Seq *s = (H4Seq *)mem;
    seq_set_dims(s, width, height, version15);
    seq_carve(s, mem + align16(sizeof(H4Seq)), 0, 0);
    seq_init_maps(s);
    return s;
}

void h4e_seq_destroy(H4Seq *s)
{
    if (!s) return;
    for (int i = 0; i < 2; ++i) { free(s->q_bn[i].v); free(s->q_bnr[i].v); }
    for (int i = 0; i < 3; ++i) { free(s->q_dcv[i].v); free(s->q_sc[i].v); free(s->q_rle[i].v); }
    free(s);
}
#endif

H4E_API void h4e_seq_set_version(H4Seq *s, int version15) { s->version15 = version15 ? 1 : 0; }
H4E_API uint32_t h4e_seq_errors(const H4Seq *s) { return s->errors_total; }
H4E_API size_t h4e_frame_bytes(const H4Seq *s) { return (size_t)s->width * s->height * 3 / 2; }

H4E_API void h4e_seq_set_split_schedule(H4Seq *s, int on) { s->split_schedule = on; }
H4E_API uint32_t h4e_last_inter_mcbs(const H4Seq *s) { return s->n_inter_mcb; }
H4E_API uint32_t h4e_last_chunks(const H4Seq *s) { return s->n_chunks; }

H4E_API void h4e_seq_dims(const H4Seq *s, int out[6])
{
    out[0] = s->width; out[1] = s->height; out[2] = s->mbw; out[3] = s->mbh; out[4] = s->nseg; out[5] = s->version15;
}

H4E_INL size_t cell_at(const H4Seq *s, int p, int bx, int by) { return (size_t)(by + 1) * s->stride[p] + bx + 1; }

/* setCode, h4m:1061-1071, with bounds */
H4E_FN void open_section(H4Seq *s, const uint8_t *data, size_t data_len, uint32_t off, const uint8_t **base, uint32_t *size)
{
    *base = NULL;
    *size = 0;
    if ((size_t)off + 4 > data_len)
    {
        s->err |= SYM_ERR_TRUNCATED;
        return;
    }
    uint32_t sz = rd_be32(data + off);
    if (sz == 0) return;
    if ((size_t)off + 4 + sz > data_len)
    {
        s->err |= SYM_ERR_TRUNCATED;
        sz = (uint32_t)(data_len - off - 4);
        if (sz == 0) return;
    }
    *base = data + off + 4;
    *size = sz;
}

H4E_FN void open_bits(H4Seq *s, BR *b, const uint8_t *data, size_t len, uint32_t off)
{
    const uint8_t *base;
    uint32_t size;
    open_section(s, data, len, off, &base, &size);
    br_open(b, base, size);
}

H4E_FN void open_bytes(H4Seq *s, ByteSec *b, const uint8_t *data, size_t len, uint32_t off)
{
    open_section(s, data, len, off, &b->base, &b->size);
    b->pos = 0;
}

/* ------------------------------------------------------------------ record groups */

/* block rows per band, as shifts */
#define SYM_BAND_SHIFT_CHROMA (s->bshift)
#define SYM_BAND_SHIFT_LUMA (s->bshift + 1)

H4E_INL int len_bucket(uint32_t len) { return len < SYM_LEN_BUCKETS ? (int)len - 1 : SYM_LEN_BUCKETS - 1; }
H4E_INL int group_of(const H4Seq *s, int cls, int band, uint32_t len)
{
    return (cls * s->nbands + band) * SYM_LEN_BUCKETS + len_bucket(len);
}

/* Called wherever a block's final type byte is written (ipic_types, pb_pass1): counts the
   record the block will own in its (class, band, length) group.  band_shift: log2 of block rows
   per band (luma 1, chroma 0 for bands of one macroblock row). */
H4E_INL void count_record(H4Seq *s, uint32_t t, int is_ipic, int by, int band_shift)
{
    const uint32_t lut = rec_lut(is_ipic, t);
    if (!lut) return;
    const int cls = (int)(lut >> 16);
    const uint32_t len = lut & 0xFFFF;
    const int g = group_of(s, cls, by >> band_shift, len);
    H4E_FETCH_ADD(&s->grp_count[g], 1);
    if (len >= SYM_LEN_BUCKETS) H4E_FETCH_ADD(&s->grp_base[g], len);   /* long-bucket word total, see plan_records */
}

/* warp-collective on the GPU */
H4E_FN void reset_record_counts(H4Seq *s)
{
    for (int g = H4E_LANE; g < s->ngroups; g += H4E_LANES) s->grp_count[g] = s->grp_next[g] = s->grp_base[g] = 0;
    if (H4E_LANE == 0) s->n_records = 0;
}

/* exclusive prefix sum of `v` over the lanes (total in *sum); the identity on a host thread */
H4E_INL uint32_t lane_scan(uint32_t v, uint32_t *sum)
{
#if defined(H4E_DEVICE)
    uint32_t inc = v;
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (H4E_LANE >= d) inc += o;
    }
    *sum = __shfl_sync(0xFFFFFFFFu, inc, 31);
    return inc - v;
#else
    *sum = v;
    return 0;
#endif
}

/* number of lanes whose predicate holds */
H4E_INL uint32_t lane_count(int pred)
{
#if defined(H4E_DEVICE)
    return (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, pred));
#else
    return pred ? 1u : 0u;
#endif
}

/*
 * Lays the (class, band, length) groups counted by count_record() out back to back
 * (class-major, so that raw/intra chunks precede inter chunks) and builds the chunk table.
 * Records of the last bucket ("long", >= SYM_LEN_BUCKETS words: only I-picture luma types
 * > 16, which no encoder emits) have individual lengths; they are packed in emission order
 * and get one chunk each.
 *
 * The layout order is the group index itself ((class * nbands + band) * SYM_LEN_BUCKETS + bucket),
 * so the plan is an exclusive prefix sum over the groups: every lane sums a contiguous slice,
 * the slice totals are scanned across the lanes, and a second walk writes the plan (a host thread
 * is one lane and one slice).  With one record band per macroblock row a 640x480 picture has
 * 3 240 groups; a single lane walking them spent 0.7 ms per picture on the GPU.
 */
typedef struct { uint32_t words, chunks, ord; } GroupSize;

H4E_INL GroupSize group_size(const H4Seq *s, int g)
{
    GroupSize z;
    const uint32_t n = s->grp_count[g];
    const int lb = g % SYM_LEN_BUCKETS;
    z.ord = n;
    if (lb < SYM_LEN_BUCKETS - 1)
    {
        z.words = n * ((uint32_t)lb + 1);
        z.chunks = (n + SYM_CHUNK - 1) / SYM_CHUNK;
    }
    else
    {   /* one chunk per long record; count_record() totalled their words in grp_base */
        z.words = s->grp_base[g];
        z.chunks = n;
    }
    return z;
}

/* warp-collective on the GPU */
H4E_FN void plan_records(H4Seq *s, int is_ipic)
{
    const int per = (s->ngroups + H4E_LANES - 1) / H4E_LANES;
    const int g0 = H4E_LANE * per < s->ngroups ? H4E_LANE * per : s->ngroups;
    const int g1 = g0 + per < s->ngroups ? g0 + per : s->ngroups;
    const int per_class = s->nbands * SYM_LEN_BUCKETS;
    uint32_t words = 0, chunks = 0, ords = 0;
    int nest = 0;
    for (int g = g0; g < g1; ++g)
    {
        const GroupSize z = group_size(s, g);
        words += z.words;
        chunks += z.chunks;
        ords += z.ord;
        nest |= z.ord != 0 && g / per_class == SYM_REC_INTRA;
    }
    uint32_t n_words, n_chunks, n_ords;
    uint32_t word = lane_scan(words, &n_words), chunk = lane_scan(chunks, &n_chunks), ord = lane_scan(ords, &n_ords);
    const int need_nest = lane_count(nest) != 0;
    for (int g = g0; g < g1; ++g)
    {
        const int cls = g / per_class, lb = g % SYM_LEN_BUCKETS;
        if (lb == 0) s->band_first[cls * (s->nbands + 1) + (g - cls * per_class) / SYM_LEN_BUCKETS] = chunk;
        const GroupSize z = group_size(s, g);
        s->grp_base[g] = word;
        s->grp_chunk[g] = chunk;
        s->grp_ord[g] = ord;
        if (lb < SYM_LEN_BUCKETS - 1)
        {
            const uint32_t len = (uint32_t)lb + 1;
            for (uint32_t i = 0; i < z.ord; i += SYM_CHUNK)
            {
                const uint32_t cnt = z.ord - i < SYM_CHUNK ? z.ord - i : SYM_CHUNK;
                s->chunks[2 * (chunk + i / SYM_CHUNK)] = word + i * len;
                s->chunks[2 * (chunk + i / SYM_CHUNK) + 1] = cnt | (len - 1) << 8 | (uint32_t)cls << 16;
            }
        }   /* else: one chunk per long record, filled in when the record is placed */
        word += z.words;
        chunk += z.chunks;
        ord += z.ord;
        /* first chunk past the class = chunk count after its last group */
        if (g + 1 - cls * per_class == per_class) s->band_first[cls * (s->nbands + 1) + s->nbands] = chunk;
    }
    H4E_SYNC();
    if (H4E_LANE == 0)
    {
        s->n_chunks_nest = s->band_first[SYM_REC_INTRA * (s->nbands + 1) + s->nbands];
        s->n_rec_words = n_words;
        s->n_records = n_ords;
        s->n_chunks = n_chunks;
        s->need_nest = is_ipic ? 1 : need_nest;
    }
}

/*
 * Record emission is split in two so that the expensive part runs over records sorted by
 * (class, length) -- uniform loop trip counts, predictable branches -- instead of in bitstream
 * order, where every block takes a different path:
 *   schedule_record()  bitstream order: reserves the record's slot in its group, writes the header
 *                      word and notes WHERE the block's side data starts in the flat-decoded
 *                      sections (the amounts consumed follow from the type byte alone);
 *   fill_records()     group order: copies descriptors / scale symbols / pairs / raw bytes.
 */
typedef struct { uint32_t fix[3], sc[3], dcv[3]; } Cursors;

H4E_TABLE int SUBX[4] = {0, 0, 1, 1}, SUBY[4] = {0, 1, 1, 0};   /* TL, BL, BR, TR: mcb_offset, h4m:862-865 */

#if defined(H4E_DEVICE)
H4E_FN Work *work_scratch(H4Seq *s, uint32_t n) { return n <= s->dev_work_cap ? s->dev_work : NULL; }
#else
static __thread Work *tl_work;         /* per host thread, grows to the largest picture seen */
static __thread uint32_t tl_work_cap;

static Work *work_scratch(H4Seq *s, uint32_t n)
{
    (void)s;
    if (n > tl_work_cap)
    {
        free(tl_work);
        tl_work_cap = n + n / 4 + 256;
        tl_work = (Work *)malloc((size_t)tl_work_cap * sizeof(Work));
        if (!tl_work) tl_work_cap = 0;
    }
    return tl_work;
}
#endif

/* what a record's block consumes from its plane's sections: raw 16 bytes | n x (2 descriptor
   bytes + 1 scale symbol) [+ 2 pair values] */
H4E_INL void cursors_advance(Cursors *c, int p, uint32_t lut)
{
    const int cls = (int)(lut >> 16);
    const uint32_t len = lut & 0xFFFF;
    const uint32_t nb = cls == SYM_REC_RAW ? 0 : len - 1 - (cls == SYM_REC_INTER);
    c->fix[p] += cls == SYM_REC_RAW ? 16 : 2 * nb;
    c->sc[p] += nb;
    c->dcv[p] += cls == SYM_REC_INTER ? 2 : 0;
}

/* slots inside a group are handed out with fetch-and-add: on the GPU the rows of a picture are
   scheduled by different lanes, and the order of records inside a group is free */
H4E_INL void schedule_record(H4Seq *s, Work *work, Cursors *c, uint32_t t, int is_ipic, int p, int bx, int by)
{
    const uint32_t lut = rec_lut(is_ipic, t);
    const int cls = (int)(lut >> 16);
    const uint32_t len = lut & 0xFFFF;
    const int band = by >> (p ? SYM_BAND_SHIFT_CHROMA : SYM_BAND_SHIFT_LUMA);
    const int g = group_of(s, cls, band, len);
    const uint32_t idx = H4E_FETCH_ADD(&s->grp_next[g], 1);
    uint32_t at;
    if (len < SYM_LEN_BUCKETS)
        at = s->grp_base[g] + idx * len;
    else
    {
        const uint32_t ch = H4E_FETCH_ADD(&s->grp_chunk[g], 1);
        at = H4E_FETCH_ADD(&s->grp_base[g], len);
        s->chunks[2 * ch] = at;
        s->chunks[2 * ch + 1] = 1u | ((len - 1) & 0xFF) << 8 | (uint32_t)cls << 16;
    }
    s->rec_base[at] = sym_record_header(t, p, bx, by);
    Work *w = &work[s->grp_ord[g] + idx];
    w->at = at;
    w->fix_off = c->fix[p];
    w->sc_off = c->sc[p];
    w->dcv_off = c->dcv[p];
    w->len = (uint16_t)len;
    w->cls = (uint8_t)cls;
    w->plane = (uint8_t)p;
    cursors_advance(c, p, lut);
}

/* Records of one row in bitstream order: one macroblock row of a P/B picture (all three planes,
   MCBlockDecDCNest h4m:1789-1827 / MCBlockDecMCNest h4m:1871-1909), or one block row of plane p of
   an I picture (IpicPlaneDec, h4m:1487-1518).  work == NULL: only advance the cursors. */
H4E_FN void row_records(H4Seq *s, int is_ipic, int p, int row, Work *work, Cursors *c)
{
    if (is_ipic)
    {
        const uint8_t *ty = s->type[p] + cell_at(s, p, 0, row);
        for (int bx = 0; bx < s->bw[p]; ++bx)
        {
            const uint32_t t = ty[bx];
            if (t == 0 || t == 8) continue;
            if (work) schedule_record(s, work, c, t, 1, p, bx, row);
            else cursors_advance(c, p, rec_lut(1, t));
        }
        return;
    }
    const int st0 = s->stride[0], my = row;
    const uint8_t *ty0 = s->type[0] + cell_at(s, 0, 0, my * 2);
    const uint8_t *ty1 = s->type[1] + cell_at(s, 1, 0, my), *ty2 = s->type[2] + cell_at(s, 2, 0, my);
    for (int mx = 0; mx < s->mbw; ++mx)
    {
        const int lx = mx * 2;
        const uint8_t tag = ty0[lx];
        const int intra = ((tag >> 5) & 3) == 0;
        if (!intra && (tag & 0x10)) continue;                 /* plain motion compensation: no records */
        for (int k = 0; k < 6; ++k)
        {
            const int pl = k < 4 ? 0 : k - 3;
            const int bx = k < 4 ? lx + SUBX[k] : mx, by = k < 4 ? my * 2 + SUBY[k] : my;
            const uint32_t t = k < 4 ? ty0[SUBY[k] * st0 + lx + SUBX[k]] : k == 4 ? ty1[mx] : ty2[mx];
            const uint32_t nib = t & 0xF;
            if (nib == 0 || (intra && nib == 8)) continue;
            if (work) schedule_record(s, work, c, t, 0, pl, bx, by);
            else cursors_advance(c, pl, rec_lut(0, t));
        }
    }
}

/* Schedules every record of the picture, rows dealt to the lanes: a row first adds up what it
   consumes, a prefix sum over the rows turns that into each row's starting cursors, then the row
   is walked again for real.  On a host thread (one lane) this degenerates to two serial walks. */
H4E_FN void schedule_rows(H4Seq *s, int is_ipic, Work *work, const Cursors *start)
{
    for (int p = 0; p < (is_ipic ? 3 : 1); ++p)
    {
        const int rows = is_ipic ? s->bh[p] : s->mbh;
#if !defined(H4E_DEVICE)
        Cursors one = *start;
        for (int row = 0; row < rows; ++row) row_records(s, is_ipic, p, row, work, &one);
        continue;
#endif
        Cursors base = *start;
        for (int r0 = 0; r0 < rows; r0 += H4E_LANES)
        {
            const int row = r0 + H4E_LANE;
            Cursors mine, at;
            memset(&mine, 0, sizeof mine);
            if (row < rows) row_records(s, is_ipic, p, row, NULL, &mine);
            for (int q = 0; q < 3; ++q)
            {
                uint32_t sum;
                at.fix[q] = base.fix[q] + lane_scan(mine.fix[q], &sum); base.fix[q] += sum;
                at.sc[q] = base.sc[q] + lane_scan(mine.sc[q], &sum); base.sc[q] += sum;
                at.dcv[q] = base.dcv[q] + lane_scan(mine.dcv[q], &sum); base.dcv[q] += sum;
            }
            if (row < rows) row_records(s, is_ipic, p, row, work, &at);
        }
    }
}

H4E_INL int32_t ss_at(SymStream *q, uint32_t i)
{
    if (i < q->n) return q->v[i];
    if (q->is_const) return q->cval;
    q->over = 1;
    return 0;
}

H4E_FN void fill_records(H4Seq *s, const Work *work, uint32_t n)
{
    for (uint32_t i = (uint32_t)H4E_LANE; i < n; i += H4E_LANES)
    {
        const Work w = work[i];
        uint32_t *rec = s->rec_base + w.at + 1;
        const int p = w.plane;
        const ByteSec *fx = &s->fix[p];
        if (w.cls == SYM_REC_RAW)
        {   /* OrgBlock, h4m:543-549 */
            if (fx->base && w.fix_off + 16 <= fx->size) memcpy(rec, fx->base + w.fix_off, 16);
            else
            {
                memset(rec, 0x80, 16);
                H4E_ERR(s, SYM_ERR_TRUNCATED);
            }
            continue;
        }
        /* n x (descriptor, scale symbol): read16(fixvl) + decodeHuff(bufTree0), h4m:691,726 / 738,767 */
        const uint32_t nb = (uint32_t)w.len - 1 - (w.cls == SYM_REC_INTER);
        const int fix_ok = fx->base && w.fix_off + 2 * nb <= fx->size;
        if (!fix_ok && nb) H4E_ERR(s, SYM_ERR_TRUNCATED);
        const uint8_t *f = fix_ok ? fx->base + w.fix_off : NULL;
        SymStream *sc = &s->q_sc[p];
        for (uint32_t k = 0; k < nb; ++k)
        {
            const uint32_t desc = f ? (uint32_t)f[2 * k] << 8 | f[2 * k + 1] : 0u;
            const uint32_t sym = (uint32_t)ss_at(sc, w.sc_off + k);
            rec[k] = desc | ((sym >> 2) & 0xFF) << 16;
        }
        if (w.cls == SYM_REC_INTER)
        {   /* the two decodeSOvfSym reads of PrediAotBlock, h4m:1405-1406, pre-shifted by dc_shift */
            int32_t a = ss_at(&s->q_dcv[p], w.dcv_off) >> s->dc_shift;
            int32_t g = ss_at(&s->q_dcv[p], w.dcv_off + 1) >> s->dc_shift;
            if (a < -32768 || a > 32767 || g < -32768 || g > 32767)
            {
                H4E_ERR(s, SYM_ERR_PAIR_RANGE);
                a = a < -32768 ? -32768 : a > 32767 ? 32767 : a;
                g = g < -32768 ? -32768 : g > 32767 ? 32767 : g;
            }
            rec[nb] = ((uint32_t)a & 0xFFFF) | (uint32_t)g << 16;
        }
    }
}

H4E_FN void plan_blob(H4Seq *s)
{
    SymHeader *h = &s->hdr;
    memset(h, 0, sizeof *h);
    h->magic = SYM_MAGIC;
    h->width = (uint16_t)s->width;
    h->height = (uint16_t)s->height;
    h->pic_type = (uint8_t)s->pic_type;
    h->version15 = (uint8_t)s->version15;
    h->dc_shift = (uint8_t)s->dc_shift;
    h->unk_shift = (uint8_t)s->unk_shift;
    h->has_nest = (uint8_t)s->need_nest;
    h->portrait = (uint8_t)(s->width < s->height);
    h->mcb_w = (uint16_t)s->mbw;
    h->mcb_h = (uint16_t)s->mbh;
    h->nseg = (uint16_t)s->nseg;
    size_t at = sizeof(SymHeader);
    h->off_chunks = (uint32_t)at;
    at = align16(at + (size_t)s->n_chunks * 8);
    h->off_bands = (uint32_t)at;
    h->n_bands = (uint32_t)s->nbands;
    at = align16(at + (size_t)SYM_REC_CLASSES * (s->nbands + 1) * 4);
    if (s->pic_type != SYM_PIC_I)
    {
        h->off_mv = (uint32_t)at;
        at = align16(at + (size_t)s->mbw * s->mbh * 4);
    }
    for (int p = 0; p < 3; ++p)
    {
        h->off_type[p] = (uint32_t)at;
        at = align16(at + s->map_cells[p]);
    }
    for (int p = 0; p < 3; ++p)
    {
        h->off_dc[p] = (uint32_t)at;
        at = align16(at + s->map_cells[p]);
    }
    if (s->need_nest)
    {
        h->off_nest = (uint32_t)at;
        at = align16(at + SYM_NEST_BYTES);
    }
    h->off_rec = (uint32_t)at;
    h->n_rec_words = s->n_rec_words;
    h->n_chunks = s->n_chunks;
    h->n_chunks_nest = s->n_chunks_nest;
    h->n_records = s->n_records;
    at = align16(at + (size_t)s->n_rec_words * 4);
    h->total_bytes = (uint32_t)at;
    s->blob_bytes = at;
}

/* ------------------------------------------------------------------ I picture, symbol part */

/* Ipic_BasisNumDec, h4m:1073-1130 */
H4E_FN void ipic_types(H4Seq *s)
{
    uint32_t run = 0;
    /* readers are copied to locals in every hot loop: the byte stores into the maps may alias
       anything, which would otherwise force the reader state back to memory after every symbol */
    SymStream bn = s->q_bn[0], bnr = s->q_bnr[0];
    for (int by = 0; by < s->bh[0]; ++by)
    {
        uint8_t *row = s->type[0] + cell_at(s, 0, 0, by);
        for (int bx = 0; bx < s->bw[0]; ++bx)
        {
            if (run)
            {
                row[bx] = 0;
                --run;
                continue;
            }
            int32_t n = ss_get(&bn);
            if ((int16_t)n == 0) run = (uint32_t)ss_get(&bnr);
            else count_record(s, (uint8_t)n, 1, by, SYM_BAND_SHIFT_LUMA);
            row[bx] = (uint8_t)n;
        }
    }
    s->q_bn[0] = bn;
    s->q_bnr[0] = bnr;
    bn = s->q_bn[1];
    bnr = s->q_bnr[1];
    run = 0;
    for (int by = 0; by < s->bh[1]; ++by)
    {
        uint8_t *ru = s->type[1] + cell_at(s, 1, 0, by), *rv = s->type[2] + cell_at(s, 2, 0, by);
        for (int bx = 0; bx < s->bw[1]; ++bx)
        {
            if (run)
            {
                ru[bx] = rv[bx] = 0;
                --run;
                continue;
            }
            int32_t n = ss_get(&bn);
            if ((int16_t)n == 0) run = (uint32_t)ss_get(&bnr);
            ru[bx] = n & 0xF;
            rv[bx] = (n >> 4) & 0xF;
            count_record(s, n & 0xF, 1, by, SYM_BAND_SHIFT_CHROMA);
            count_record(s, (n >> 4) & 0xF, 1, by, SYM_BAND_SHIFT_CHROMA);
        }
    }
    s->q_bn[1] = bn;
    s->q_bnr[1] = bnr;
}

/* IpicDcvDec + getDeltaDC, h4m:1043-1058, 1132-1164 */
H4E_FN void ipic_dcs(H4Seq *s)
{
    for (int p = 0; p < 3; ++p)
    {
        uint32_t run = 0;
        SymStream dcv = s->q_dcv[p], rle = s->q_rle[p];
        for (int by = 0; by < s->bh[p]; ++by)
        {
            uint8_t *cur = s->dc[p] + cell_at(s, p, 0, by);
            const uint8_t *up = cur - s->stride[p];
            uint8_t v = up[0];
            for (int bx = 0; bx < s->bw[p]; ++bx)
            {
                if (run) --run;
                else
                {
                    uint32_t delta = (uint32_t)ss_get(&dcv);
                    if (delta == 0) run = (uint32_t)ss_get(&rle);
                    v = (uint8_t)(v + delta);
                }
                cur[bx] = v;
                v = (uint8_t)((v + up[bx + 1] + 1) >> 1);
            }
        }
        s->q_dcv[p] = dcv;
        s->q_rle[p] = rle;
    }
}

/* MakeNest, h4m:1166-1239 (including the mirror / zero-fill path), packed to nibbles.  The nest is 70 x 38 in landscape
   and 38 x 70 in portrait pictures (h4m:965-975); packed rows are NW / 2 bytes either way (35 or 19), 1 330 bytes in all */
H4E_FN void make_nest(H4Seq *s, int nx, int ny)
{
    uint8_t full[SYM_NEST_H * SYM_NEST_W];
    const int NW = s->width < s->height ? SYM_NEST_H : SYM_NEST_W, NH = s->width < s->height ? SYM_NEST_W : SYM_NEST_H;
    int bw = s->bw[0], bh = s->bh[0];
    int cols = bw < NW ? bw : NW, rows = bh < NH ? bh : NH;
    int mcols = bw < NW ? (NW - bw < bw ? NW - bw : bw) : 0;
    int mrows = bh < NH ? (NH - bh < bh ? NH - bh : bh) : 0;
    /* keep the window inside the map even for a hostile header */
    if (nx < 0 || nx + cols > bw) { nx = 0; s->err |= SYM_ERR_MV_RANGE; }
    if (ny < 0 || ny + rows > bh) { ny = 0; s->err |= SYM_ERR_MV_RANGE; }
    memset(full, 0, sizeof full);
    for (int i = 0; i < rows; ++i)
    {
        const uint8_t *src = s->dc[0] + cell_at(s, 0, nx, ny + i);
        for (int j = 0; j < cols; ++j) full[i * NW + j] = (src[j] >> 4) & 0xF;
        for (int j = 0; j < mcols; ++j) full[i * NW + cols + j] = (src[cols - 1 - j] >> 4) & 0xF;
    }
    for (int i = 0; i < mrows; ++i) memcpy(full + (rows + i) * NW, full + (rows - 1 - i) * NW, (size_t)NW);
    for (int i = 0; i < NH; ++i)
        for (int j = 0; j < NW / 2; ++j)
            s->nest[i * (NW / 2) + j] = (uint8_t)(full[i * NW + 2 * j] | full[i * NW + 2 * j + 1] << 4);
}

/* ------------------------------------------------------------------ P/B picture, pass 1 */

typedef struct { uint32_t value, count; } RunLen;

H4E_TABLE uint8_t next_type[2][4] = {{1, 2, 0, 0}, {2, 0, 1, 0}};   /* mcbtypetrans, h4m:1591-1594 */

/* spread_PB_descMap, h4m:1742-1776, with decode_PB_dc (1649), decode_PB_cc (1670),
   getMCBtype (1596), getMCBproc (1613), initMCBtype/proc (1551-1569) */
H4E_FN void pb_pass1(H4Seq *s)
{
    const HTab *tm = &s->tree[T_MCB];
    RunLen proc = {0, 0}, type = {0, 0};
    BR mcbp = s->mcbp, mcbt = s->mcbt;
    SymStream bn0 = s->q_bn[0], bn1 = s->q_bn[1], bnr0 = s->q_bnr[0], bnr1 = s->q_bnr[1];
    SymStream dcv0 = s->q_dcv[0], dcv1 = s->q_dcv[1], dcv2 = s->q_dcv[2];
    if (mcbp.base)
    {
        proc.value = br_bit(&mcbp);
        proc.count = (uint32_t)ht_get_uovf(tm, &mcbp);
    }
    if (mcbt.base)
    {
        type.value = br_bits(&mcbt, 2);
        type.count = (uint32_t)ht_get_uovf(tm, &mcbt);
    }
    else
        s->err |= SYM_ERR_TRUNCATED;   /* the reference would use an uninitialised type here */
    uint32_t run_y = 0, run_c = 0;
    uint32_t acc[3] = {0x7F, 0x7F, 0x7F};
    const int st0 = s->stride[0];
    for (int my = 0; my < s->mbh; ++my)
    {
        uint8_t *ty0 = s->type[0] + cell_at(s, 0, 0, my * 2), *dc0 = s->dc[0] + cell_at(s, 0, 0, my * 2);
        uint8_t *ty1 = s->type[1] + cell_at(s, 1, 0, my), *dc1 = s->dc[1] + cell_at(s, 1, 0, my);
        uint8_t *ty2 = s->type[2] + cell_at(s, 2, 0, my), *dc2 = s->dc[2] + cell_at(s, 2, 0, my);
        for (int mx = 0; mx < s->mbw; ++mx)
        {
            if (type.count == 0 && mcbt.base)
            {
                type.value = next_type[br_bit(&mcbt)][type.value & 3];
                type.count = (uint32_t)ht_get_uovf(tm, &mcbt);
            }
            --type.count;
            uint32_t mt = type.value;
            if (mt == 3 || (mt == 2 && s->pic_type == SYM_PIC_P))
            {
                /* type 3 indexes outside mcbtypetrans in the reference; type 2 in a P picture
                   would predict from the picture being written (h4m:2060) */
                s->err |= SYM_ERR_MCB_TYPE;
                mt = 1;
            }
            uint32_t pr = 0;
            const int lx = mx * 2;
            if (mt == 0)
            {
                for (int k = 0; k < 4; ++k)
                {
                    acc[0] += (uint32_t)ss_get(&dcv0);
                    dc0[SUBY[k] * st0 + lx + SUBX[k]] = (uint8_t)acc[0];
                }
                acc[1] += (uint32_t)ss_get(&dcv1);
                dc1[mx] = (uint8_t)acc[1];
                acc[2] += (uint32_t)ss_get(&dcv2);
                dc2[mx] = (uint8_t)acc[2];
            }
            else
            {
                acc[0] = acc[1] = acc[2] = 0x7F;
                if (proc.count == 0 && mcbp.base)
                {
                    proc.value ^= 1;
                    proc.count = (uint32_t)ht_get_uovf(tm, &mcbp);
                }
                --proc.count;
                pr = proc.value & 1;
            }
            const uint8_t tag = (uint8_t)(mt << 5 | pr << 4);
            if (pr)
            {
                for (int k = 0; k < 4; ++k) ty0[SUBY[k] * st0 + lx + SUBX[k]] = tag;
                ty1[mx] = ty2[mx] = tag;
                continue;
            }
            for (int k = 0; k < 4; ++k)
            {
                uint8_t *c = &ty0[SUBY[k] * st0 + lx + SUBX[k]];
                if (run_y)
                {
                    *c = tag;
                    --run_y;
                    continue;
                }
                int32_t n = (int16_t)ss_get(&bn0);
                if (n)
                {
                    if (n & ~0xF)
                    {
                        s->err |= SYM_ERR_MCB_TYPE;   /* would corrupt the macroblock bits, h4m:1701 */
                        n &= 0xF;
                    }
                    *c = (uint8_t)(tag | n);
                    count_record(s, *c, 0, my * 2 + SUBY[k], SYM_BAND_SHIFT_LUMA);
                }
                else
                {
                    *c = tag;
                    run_y = (uint32_t)ss_get(&bnr0);
                }
            }
            if (run_c)
            {
                ty1[mx] = ty2[mx] = tag;
                --run_c;
            }
            else
            {
                int32_t n = (int16_t)ss_get(&bn1);
                if (n)
                {
                    ty1[mx] = (uint8_t)(tag | (n & 0xF));
                    ty2[mx] = (uint8_t)(tag | ((n >> 4) & 0xF));
                    count_record(s, ty1[mx], 0, my, SYM_BAND_SHIFT_CHROMA);
                    count_record(s, ty2[mx], 0, my, SYM_BAND_SHIFT_CHROMA);
                }
                else
                {
                    ty1[mx] = ty2[mx] = tag;
                    run_c = (uint32_t)ss_get(&bnr1);
                }
            }
        }
    }
    s->mcbp = mcbp; s->mcbt = mcbt;
    s->q_bn[0] = bn0; s->q_bn[1] = bn1; s->q_bnr[0] = bnr0; s->q_bnr[1] = bnr1;
    s->q_dcv[0] = dcv0; s->q_dcv[1] = dcv1; s->q_dcv[2] = dcv2;
}

/* ------------------------------------------------------------------ block types, data-parallel form
 *
 * Ipic_BasisNumDec (h4m:1073-1130) and decode_PB_cc (h4m:1670-1740) are the same automaton over a
 * SEQUENCE of blocks: take a type symbol; a zero symbol is followed by a run length and leaves
 * that many further blocks without a symbol.  Symbol j therefore lands on sequence element
 *     pos(j) = j + (sum of the first z(j) run lengths),   z(j) = zero symbols before j,
 * which is two prefix sums instead of a serial walk.  The GPU build runs this form with the
 * symbols dealt to the lanes; one lane (host, h4e_seq_set_split_schedule) gives the same maps,
 * and the tests compare it with the serial loops above.
 */
#define RUN_CAP (1u << 23)   /* more than any sequence length (2048 x 2048 blocks); run sums saturate here */

H4E_INL uint32_t sat_run(uint32_t a, uint32_t b) { return a + b > RUN_CAP ? RUN_CAP : a + b; }

/* exclusive saturating prefix sum of v over the lanes */
H4E_INL uint32_t lane_scan_sat(uint32_t v, uint32_t *sum)
{
#if defined(H4E_DEVICE)
    uint32_t inc = v;
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (H4E_LANE >= d) inc = sat_run(inc, o);
    }
    *sum = __shfl_sync(0xFFFFFFFFu, inc, 31);
    const uint32_t ex = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    return H4E_LANE ? ex : 0u;
#else
    *sum = v;
    return 0;
#endif
}

/* replaces the run lengths by their exclusive prefix sums, in place; returns the total */
H4E_FN uint32_t runs_prefix(SymStream *q)
{
    uint32_t carry = 0;
    for (uint32_t b = 0; b < q->n; b += H4E_LANES)
    {
        const uint32_t i = b + (uint32_t)H4E_LANE;
        uint32_t v = 0, sum;
        if (i < q->n)
        {
            v = (uint32_t)q->v[i];
            if (v > RUN_CAP) v = RUN_CAP;
        }
        const uint32_t ex = sat_run(carry, lane_scan_sat(v, &sum));
        if (i < q->n) q->v[i] = (int32_t)ex;
        carry = sat_run(carry, sum);
    }
    return carry;
}

/* sum of the first z run lengths (after runs_prefix) */
H4E_INL uint32_t runs_before(const SymStream *q, uint32_t total, uint32_t z)
{
    if (z < q->n) return (uint32_t)q->v[z];
    if (!q->is_const || z == q->n) return total;          /* an exhausted section reads as 0 (ss_get) */
    uint32_t c = (uint32_t)q->cval;
    if (c > RUN_CAP) c = RUN_CAP;
    const uint64_t r = (uint64_t)total + (uint64_t)(z - q->n) * c;
    return r > RUN_CAP ? RUN_CAP : (uint32_t)r;
}

H4E_INL int32_t ss_peek(const SymStream *q, uint32_t i) { return i < q->n ? q->v[i] : q->is_const ? q->cval : 0; }

enum { SEQ_I_LUMA, SEQ_I_CHROMA, SEQ_PB_LUMA, SEQ_PB_CHROMA, SEQ_I_DC /* + plane */ };

/* the type symbol `sym` belongs to element `pos` of the block sequence */
H4E_INL void place_type(H4Seq *s, int seq, uint32_t pos, int32_t sym)
{
    switch (seq)
    {
    case SEQ_I_LUMA:
    {   /* raster over the luma blocks, h4m:1082-1101 */
        const int by = (int)(pos / (uint32_t)s->bw[0]), bx = (int)(pos % (uint32_t)s->bw[0]);
        s->type[0][cell_at(s, 0, bx, by)] = (uint8_t)sym;
        if ((int16_t)sym != 0) count_record(s, (uint8_t)sym, 1, by, SYM_BAND_SHIFT_LUMA);
        break;
    }
    case SEQ_I_CHROMA:
    {   /* one symbol carries the U and the V nibble, h4m:1103-1129 */
        const int by = (int)(pos / (uint32_t)s->bw[1]), bx = (int)(pos % (uint32_t)s->bw[1]);
        s->type[1][cell_at(s, 1, bx, by)] = sym & 0xF;
        s->type[2][cell_at(s, 2, bx, by)] = (sym >> 4) & 0xF;
        count_record(s, sym & 0xF, 1, by, SYM_BAND_SHIFT_CHROMA);
        count_record(s, (sym >> 4) & 0xF, 1, by, SYM_BAND_SHIFT_CHROMA);
        break;
    }
    case SEQ_PB_LUMA:
    {   /* four luma blocks per listed macroblock, h4m:1689-1712 */
        int32_t n = (int16_t)sym;
        if (!n) break;
        const uint32_t mcb = s->mcb_list[pos >> 2];
        const int k = (int)(pos & 3), my = (int)(mcb / (uint32_t)s->mbw), mx = (int)(mcb % (uint32_t)s->mbw);
        if (n & ~0xF)
        {
            H4E_ERR(s, SYM_ERR_MCB_TYPE);   /* would corrupt the macroblock bits, h4m:1701 */
            n &= 0xF;
        }
        const uint8_t t = (uint8_t)(s->mcb_tag[mcb] | n);
        s->type[0][cell_at(s, 0, mx * 2 + SUBX[k], my * 2 + SUBY[k])] = t;
        count_record(s, t, 0, my * 2 + SUBY[k], SYM_BAND_SHIFT_LUMA);
        break;
    }
    case SEQ_I_DC: case SEQ_I_DC + 1: case SEQ_I_DC + 2:
    {   /* DC delta of a block, raster over plane p (IpicDcvDec, h4m:1132-1164); resolved by ipic_dcs_split */
        const int p = seq - SEQ_I_DC;
        const int by = (int)(pos / (uint32_t)s->bw[p]), bx = (int)(pos % (uint32_t)s->bw[p]);
        s->dc[p][cell_at(s, p, bx, by)] = (uint8_t)sym;
        break;
    }
    default:
    {   /* one symbol per listed macroblock: U and V nibble, h4m:1714-1739 */
        const int32_t n = (int16_t)sym;
        if (!n) break;
        const uint32_t mcb = s->mcb_list[pos];
        const int my = (int)(mcb / (uint32_t)s->mbw), mx = (int)(mcb % (uint32_t)s->mbw);
        const uint8_t tag = s->mcb_tag[mcb];
        const uint8_t tu = (uint8_t)(tag | (n & 0xF)), tv = (uint8_t)(tag | ((n >> 4) & 0xF));
        s->type[1][cell_at(s, 1, mx, my)] = tu;
        s->type[2][cell_at(s, 2, mx, my)] = tv;
        count_record(s, tu, 0, my, SYM_BAND_SHIFT_CHROMA);
        count_record(s, tv, 0, my, SYM_BAND_SHIFT_CHROMA);
        break;
    }
    }
}

/* Runs the automaton over a sequence of seq_len blocks whose cells already hold the "no symbol"
   value.  Collective; bn / bnr are at position 0 (fresh from the flat decode). */
H4E_FN void types_scatter(H4Seq *s, int seq, SymStream *bn, SymStream *bnr, uint32_t seq_len)
{
    const uint32_t total = runs_prefix(bnr);
    H4E_SYNC();
    uint32_t zbase = 0, used = 0, zused = 0;
    for (uint32_t b = 0;; b += H4E_LANES)
    {
        const uint32_t j = b + (uint32_t)H4E_LANE;
        const int32_t sym = ss_peek(bn, j);
        const int isz = seq >= SEQ_I_DC ? sym == 0 : (int16_t)sym == 0;   /* h4m:1146 tests the full value */
        uint32_t zs;
        const uint32_t z = zbase + lane_scan((uint32_t)isz, &zs);
        const uint32_t pos = j + runs_before(bnr, total, z);
        const int act = pos < seq_len;
        if (act) place_type(s, seq, pos, sym);
        const uint32_t n_act = lane_count(act);
        used += n_act;
        zused += lane_count(act && isz);
        zbase += zs;
        if (n_act < H4E_LANES) break;                      /* pos(j) increases with j */
    }
    H4E_SYNC();
    if (H4E_LANE == 0)
    {   /* what the serial walk would have consumed */
        if (used > bn->n && !bn->is_const) bn->over = 1;
        if (zused > bnr->n && !bnr->is_const) bnr->over = 1;
        bn->pos = used < bn->n ? used : bn->n;
        bnr->pos = zused < bnr->n ? zused : bnr->n;
    }
}

/* Ipic_BasisNumDec in the form above */
H4E_FN void ipic_types_split(H4Seq *s)
{
    for (int p = 0; p < 3; ++p)
        for (int by = H4E_LANE; by < s->bh[p]; by += H4E_LANES) memset(s->type[p] + cell_at(s, p, 0, by), 0, (size_t)s->bw[p]);
    H4E_SYNC();
    types_scatter(s, SEQ_I_LUMA, &s->q_bn[0], &s->q_bnr[0], (uint32_t)(s->bw[0] * s->bh[0]));
    types_scatter(s, SEQ_I_CHROMA, &s->q_bn[1], &s->q_bnr[1], (uint32_t)(s->bw[1] * s->bh[1]));
    H4E_SYNC();
}

/* value the lane below produced in the previous wavefront step */
H4E_INL uint32_t lane_above(uint32_t v)
{
#if defined(H4E_DEVICE)
    return __shfl_up_sync(0xFFFFFFFFu, v, 1);
#else
    return v;
#endif
}

/* IpicDcvDec + getDeltaDC (h4m:1043-1058, 1132-1164) in two steps: the deltas land on their blocks
   through the same automaton as the types (a zero delta is followed by a run of blocks without
   one); then the predictor recurrence -- each block starts from the average of its left
   neighbour and the block above its right neighbour -- runs as a wavefront, one row per lane,
   each row two blocks behind the row above. */
H4E_FN void ipic_dcs_split(H4Seq *s)
{
    for (int p = 0; p < 3; ++p)
        for (int by = H4E_LANE; by < s->bh[p]; by += H4E_LANES) memset(s->dc[p] + cell_at(s, p, 0, by), 0, (size_t)s->bw[p]);
    H4E_SYNC();
    for (int p = 0; p < 3; ++p) types_scatter(s, SEQ_I_DC + p, &s->q_dcv[p], &s->q_rle[p], (uint32_t)(s->bw[p] * s->bh[p]));
    H4E_SYNC();
    for (int p = 0; p < 3; ++p)
    {
        const int bw = s->bw[p], bh = s->bh[p], steps = bw + 2 * (H4E_LANES - 1);
        for (int r0 = 0; r0 < bh; r0 += H4E_LANES)
        {
            const int row = r0 + H4E_LANE;
            uint8_t *cur = s->dc[p] + cell_at(s, p, 0, row < bh ? row : 0);
            const uint8_t *up = cur - s->stride[p];
            const uint32_t edge = cur[bw];                  /* border cell right of the row */
            uint32_t v = 0, out = 0;
            for (int t = 0; t < steps; ++t)
            {
                H4E_SYNC();
                const int x = t - 2 * H4E_LANE;
                uint32_t above = lane_above(out);           /* up[x + 1], stored by the lane above one step ago */
                if (row < bh && x >= 0 && x < bw)
                {
                    if (H4E_LANE == 0) above = up[x + 1];   /* the row above belongs to the previous group */
                    if (x == 0) v = up[0];
                    v = (uint8_t)(v + cur[x]);
                    cur[x] = (uint8_t)v;
                    out = v;
                    v = (uint8_t)((v + above + 1) >> 1);
                }
                else if (x >= bw)
                    out = edge;
            }
        }
    }
    H4E_SYNC();
}

/* inclusive running maximum of v over the lanes */
H4E_INL uint32_t lane_scan_max(uint32_t v, uint32_t *last)
{
#if defined(H4E_DEVICE)
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (H4E_LANE >= d && o > v) v = o;
    }
    *last = __shfl_sync(0xFFFFFFFFu, v, 31);
    return v;
#else
    *last = v;
    return v;
#endif
}

/* One run-length coded macroblock attribute -- getMCBtype (h4m:1596-1611, with initMCBtype
   h4m:1551-1559) or getMCBproc (h4m:1613-1628, initMCBproc h4m:1561-1569) -- for `need`
   macroblocks, as run starts + values.  Serial, but only one step per RUN.  Reads exactly what
   the per-macroblock loop would: a run is fetched when a macroblock finds the count at zero, the
   initial run may be empty, and a later zero count wraps around to "the rest". */
H4E_FN uint32_t mcb_runs(H4Seq *s, BR *b, int is_type, uint32_t need, uint32_t *start, uint32_t *value)
{
    if (!b->base)
    {   /* no section: value 0 for everybody (the reference would use an uninitialised type here) */
        if (is_type) s->err |= SYM_ERR_TRUNCATED;
        start[0] = 0;
        value[0] = 0;
        return need ? 1 : 0;
    }
    const HTab *tm = &s->tree[T_MCB];
    uint32_t val = br_bits(b, is_type ? 2 : 1), cnt = (uint32_t)ht_get_uovf(tm, b), n = 0, total = 0;
    for (int first = 1; total < need; first = 0)
    {
        if (cnt || !first)
        {
            start[n] = total;
            value[n] = val;
            ++n;
            if (cnt == 0 || cnt >= need - total) break;
            total += cnt;
        }
        val = is_type ? next_type[br_bit(b)][val & 3] : val ^ 1;
        cnt = (uint32_t)ht_get_uovf(tm, b);
    }
    return n;
}

/* value of the run that covers element j (binary search over the run starts) */
H4E_INL uint32_t run_value_at(const uint32_t *start, const uint32_t *value, uint32_t n, uint32_t j)
{
    uint32_t lo = 0, hi = n;          /* start[lo] <= j < start[hi] */
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (start[mid] <= j) lo = mid;
        else hi = mid;
    }
    return value[lo];
}

/* inclusive prefix sums (mod 2^32) of the first `count` values, in place */
H4E_FN void vals_prefix(SymStream *q, uint32_t count)
{
    if (count > q->n) count = q->n;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < count; b += H4E_LANES)
    {
        const uint32_t i = b + (uint32_t)H4E_LANE;
        const uint32_t v = i < count ? (uint32_t)q->v[i] : 0;
        uint32_t sum;
        const uint32_t inc = carry + lane_scan(v, &sum) + v;
        if (i < count) q->v[i] = (int32_t)inc;
        carry += sum;
    }
}

/* sum of values 0..j of a stream whose first `count` values went through vals_prefix; what lies
   past the end of the section reads as ss_get() would give it */
H4E_INL uint32_t prefix_at(const SymStream *q, uint32_t j)
{
    if (j < q->n) return (uint32_t)q->v[j];
    const uint32_t base = q->n ? (uint32_t)q->v[q->n - 1] : 0;
    return q->is_const ? base + (j - q->n + 1) * (uint32_t)q->cval : base;
}

/* spread_PB_descMap (h4m:1742-1776), data-parallel:
   (1) the type runs, expanded over the macroblocks; (2) the proc runs, expanded over the
   macroblocks that are not intra; (3) the DC chains of the intra macroblocks (decode_PB_dc,
   h4m:1649-1668: every maximal run of intra macroblocks accumulates from 0x7F) as differences
   of prefix sums; (4) the tags spread over the maps; (5) the block types through types_scatter
   (decode_PB_cc).  Only the run decode of (1) and (2) is serial. */
H4E_FN void pb_pass1_split(H4Seq *s)
{
    const uint32_t n_all = (uint32_t)(s->mbw * s->mbh);
    uint32_t *run_start = (uint32_t *)s->mv_raw, *run_value = run_start + n_all;   /* mv_raw is free until pass 2 */
    /* (1) */
    if (H4E_LANE == 0) s->n_list = mcb_runs(s, &s->mcbt, 1, n_all, run_start, run_value);
    H4E_SYNC();
    uint32_t n_inter = 0;
    {
        const uint32_t n_runs = s->n_list;
        for (uint32_t m0 = 0; m0 < n_all; m0 += H4E_LANES)
        {
            const uint32_t mcb = m0 + (uint32_t)H4E_LANE;
            uint32_t mt = 0;
            if (mcb < n_all)
            {
                mt = run_value_at(run_start, run_value, n_runs, mcb);
                if (mt == 3 || (mt == 2 && s->pic_type == SYM_PIC_P))
                {   /* type 3 indexes outside mcbtypetrans in the reference; type 2 in a P picture
                       would predict from the picture being written (h4m:2060) */
                    H4E_ERR(s, SYM_ERR_MCB_TYPE);
                    mt = 1;
                }
                s->mcb_tag[mcb] = (uint8_t)(mt << 5);
            }
            uint32_t sum;
            const uint32_t at = n_inter + lane_scan(mt != 0, &sum);
            if (mt) s->mcb_list[at] = mcb;                 /* the macroblocks that own a proc flag, in order */
            n_inter += sum;
        }
    }
    H4E_SYNC();
    /* (2) */
    if (H4E_LANE == 0) s->n_list = mcb_runs(s, &s->mcbp, 0, n_inter, run_start, run_value);
    H4E_SYNC();
    {
        const uint32_t n_runs = s->n_list;
        for (uint32_t j = (uint32_t)H4E_LANE; j < n_inter; j += H4E_LANES)
            if (run_value_at(run_start, run_value, n_runs, j) & 1) s->mcb_tag[s->mcb_list[j]] |= 0x10;
    }
    H4E_SYNC();
    /* (3) + the list of macroblocks that carry block types */
    const uint32_t n_intra = n_all - n_inter;
    vals_prefix(&s->q_dcv[0], 4 * n_intra);
    vals_prefix(&s->q_dcv[1], n_intra);
    vals_prefix(&s->q_dcv[2], n_intra);
    H4E_SYNC();
    {
        const int st0 = s->stride[0];
        uint32_t seen_intra = 0, seg_carry = 0, n_list = 0;
        for (uint32_t m0 = 0; m0 < n_all; m0 += H4E_LANES)
        {
            const uint32_t mcb = m0 + (uint32_t)H4E_LANE;
            const uint32_t tag = mcb < n_all ? s->mcb_tag[mcb] : 0x20;
            const int intra = (tag >> 5) == 0, typed = mcb < n_all && !(tag & 0x10);
            uint32_t sum, last;
            const uint32_t idx = seen_intra + lane_scan(intra, &sum);
            seen_intra += sum;
            /* 1 + index of the first macroblock of this intra run, carried along the run */
            const int opens = intra && (mcb == 0 || (s->mcb_tag[mcb - 1] >> 5) != 0);
            uint32_t seg = lane_scan_max(opens ? idx + 1 : 0, &last);
            if (seg < seg_carry) seg = seg_carry;
            seg_carry = last > seg_carry ? last : seg_carry;
            if (intra)
            {
                const uint32_t first = seg - 1;
                const int my = (int)(mcb / (uint32_t)s->mbw), mx = (int)(mcb % (uint32_t)s->mbw);
                const uint32_t b0 = first ? prefix_at(&s->q_dcv[0], 4 * first - 1) : 0;
                uint8_t *dc0 = s->dc[0] + cell_at(s, 0, mx * 2, my * 2);
                for (int k = 0; k < 4; ++k)
                    dc0[SUBY[k] * st0 + SUBX[k]] = (uint8_t)(0x7F + prefix_at(&s->q_dcv[0], 4 * idx + (uint32_t)k) - b0);
                for (int p = 1; p < 3; ++p)
                {
                    const uint32_t bp = first ? prefix_at(&s->q_dcv[p], first - 1) : 0;
                    s->dc[p][cell_at(s, p, mx, my)] = (uint8_t)(0x7F + prefix_at(&s->q_dcv[p], idx) - bp);
                }
            }
            const uint32_t at = n_list + lane_scan((uint32_t)typed, &sum);
            n_list += sum;
            if (typed) s->mcb_list[at] = mcb;              /* (2) is done with the old contents */
        }
        H4E_SYNC();
        if (H4E_LANE == 0)
        {
            s->n_list = n_list;
            for (int p = 0; p < 3; ++p)
            {
                SymStream *q = &s->q_dcv[p];
                const uint32_t used = p ? n_intra : 4 * n_intra;
                if (used > q->n && !q->is_const) q->over = 1;
                q->pos = used < q->n ? used : q->n;
            }
        }
    }
    H4E_SYNC();
    const int st0 = s->stride[0], n_mcb = s->mbw * s->mbh;
    for (int mcb = H4E_LANE; mcb < n_mcb; mcb += H4E_LANES)
    {
        const int my = mcb / s->mbw, mx = mcb % s->mbw;
        const uint8_t tag = s->mcb_tag[mcb];
        uint8_t *ty0 = s->type[0] + cell_at(s, 0, mx * 2, my * 2);
        ty0[0] = ty0[1] = ty0[st0] = ty0[st0 + 1] = tag;
        s->type[1][cell_at(s, 1, mx, my)] = s->type[2][cell_at(s, 2, mx, my)] = tag;
    }
    H4E_SYNC();
    types_scatter(s, SEQ_PB_LUMA, &s->q_bn[0], &s->q_bnr[0], 4 * s->n_list);
    types_scatter(s, SEQ_PB_CHROMA, &s->q_bn[1], &s->q_bnr[1], s->n_list);
    H4E_SYNC();
}

/* getMVector, h4m:1846-1860 */
H4E_INL void read_mv(H4Seq *s, BR *b, int32_t *mv, int rbits)
{
    if (rbits > 16) rbits = 16;
    int32_t lim = 1 << (rbits + 5);
    int32_t v = ht_get(&s->tree[T_MV], b) * (1 << rbits);
    v += (int32_t)br_bits(b, rbits);
    /* Modulo 2^32 like the reference's int32 on its targets: with damaged residual-bit counts a step can exceed
       the wrap range and the predictor then grows without bound (the vector is range-checked afterwards). */
    uint32_t acc = (uint32_t)*mv + (uint32_t)v;
    if ((int32_t)acc >= lim) acc -= (uint32_t)lim << 1;
    else if ((int32_t)acc < -lim) acc += (uint32_t)lim << 1;
    *mv = (int32_t)acc;
}

/*
 * Checks that every address the reconstruction of this macroblock will read lies inside
 * the reference surface.  The reference addresses frames linearly with no clamping
 * (h4m:1344, 1866, 1897), so "inside" means inside the contiguous Y|U|V buffer, not inside
 * the plane; recon.cu addresses the same way, which keeps even row-wrapping vectors exact.
 */
H4E_FN int mcb_refs_in_surface(const H4Seq *s, int32_t rx, int32_t ry, int needs_window)
{
    /* fast accept: the whole 9x9 luma patch (and the 70x38 window) strictly inside the luma plane
       implies every chroma access is inside its plane too (positions and sizes halve) */
    {
        const int32_t ix = rx >> 1, iy = ry >> 1;
        if (!needs_window ? (ix >= 0 && iy >= 0 && ix + 9 <= s->width && iy + 9 <= s->height)
                          : (ix >= 32 && iy >= 16 && ix + 38 <= s->width && iy + 22 <= s->height))
            return 1;
    }
    const int64_t total = (int64_t)s->width * s->height * 3 / 2;
    int64_t plane_base = 0;
    int hx = rx & 1, hy = ry & 1;
    for (int p = 0; p < 3; ++p)
    {
        int sh = p ? 1 : 0;
        int64_t w = s->width >> sh;
        int32_t px = rx >> sh, py = ry >> sh;
        if (s->version15) { hx = px & 1; hy = py & 1; }
        int64_t first = plane_base + (int64_t)(py >> 1) * w + (px >> 1);
        int span = p ? 4 : 8;
        int64_t last = first + (int64_t)(span - 1 + hy) * w + span - 1 + hx;
        if (first < 0 || last >= total) return 0;
        plane_base += w * (s->height >> sh);
    }
    if (needs_window)
    {
        /* window origin and size, h4m:1864-1868: 70 x 38 at (-32, -16) in landscape, 38 x 70 at (-16, -32) in portrait */
        const int portrait = s->width < s->height;
        const int NW = portrait ? SYM_NEST_H : SYM_NEST_W, NH = portrait ? SYM_NEST_W : SYM_NEST_H;
        int64_t org = (int64_t)(rx / 2) + (int64_t)(ry / 2 - (portrait ? 32 : 16)) * s->width - (portrait ? 16 : 32);
        if (org < 0 || org + (int64_t)(NH - 1) * s->width + NW - 1 >= total) return 0;
    }
    return 1;
}

/* BpicPlaneDec pass 2, h4m:1922-1967, motion vectors only (the records of the same macroblocks
   are scheduled by schedule_rows()).  The horizontal and the vertical component live in
   separate sections and accumulate separately, so two lanes run the two serial chains
   (getMVector) side by side into mv_raw; then all lanes turn the vectors into checked absolute
   positions.  Needs mcb_tag, i.e. pb_pass1_split. */
H4E_FN void pb_mvs(H4Seq *s, int16_t *mv_out)
{
    const int n_mcb = s->mbw * s->mbh;
#if defined(H4E_DEVICE)
    const int c_lo = H4E_LANE, c_hi = H4E_LANE < 2 ? H4E_LANE + 1 : 0;
#else
    const int c_lo = 0, c_hi = 2;
#endif
    for (int c = c_lo; c < c_hi; ++c)
    {
        BR b = c ? s->mvv : s->mvh;
        int32_t mv = 0;
        int cur_ref = -1;
        for (int mcb = 0; mcb < n_mcb; ++mcb)
        {
            const int mt = (s->mcb_tag[mcb] >> 5) & 3;
            if (mt)
            {
                if (mt - 1 != cur_ref)
                {   /* h4m:1943-1949 */
                    cur_ref = mt - 1;
                    mv = 0;
                }
                read_mv(s, &b, &mv, s->rb[cur_ref][c]);
            }
            s->mv_raw[2 * mcb + c] = mv;
        }
        if (c) s->mvv = b;
        else s->mvh = b;
    }
    H4E_SYNC();
    const int st0 = s->stride[0];
    uint32_t n_inter = 0;
    for (int m0 = 0; m0 < n_mcb; m0 += H4E_LANES)
    {
        const int mcb = m0 + H4E_LANE;
        int inter = 0;
        if (mcb < n_mcb)
        {
            const int my = mcb / s->mbw, mx = mcb % s->mbw, lx = mx * 2;
            const uint8_t tag = s->mcb_tag[mcb];
            int16_t *mvp = mv_out + 2 * (size_t)mcb;
            inter = ((tag >> 5) & 3) != 0;
            if (!inter) mvp[0] = mvp[1] = 0;
            else
            {
                const int32_t rx = mx * 16 + s->mv_raw[2 * mcb], ry = my * 16 + s->mv_raw[2 * mcb + 1];   /* h4m:1954-1955 */
                int needs_window = 0;
                if (!(tag & 0x10))
                {   /* a PrediAot block with bases reads the 70x38 window around the vector (h4m:1334-1348) */
                    const uint8_t *ty0 = s->type[0] + cell_at(s, 0, lx, my * 2);
                    const uint32_t t6[6] = {ty0[0], ty0[st0], ty0[st0 + 1], ty0[1], s->type[1][cell_at(s, 1, mx, my)], s->type[2][cell_at(s, 2, mx, my)]};
                    for (int k = 0; k < 6; ++k)
                    {
                        const uint32_t nib = t6[k] & 0xF;
                        if (nib > 1 && nib != 6) needs_window = 1;
                    }
                }
                if (rx < -32000 || rx > 32000 || ry < -32000 || ry > 32000 || !mcb_refs_in_surface(s, rx, ry, needs_window))
                {
                    H4E_ERR(s, SYM_ERR_MV_RANGE);
                    mvp[0] = mvp[1] = -32768;   /* poison: recon.cu paints the macroblock grey */
                }
                else
                {
                    mvp[0] = (int16_t)rx;
                    mvp[1] = (int16_t)ry;
                }
            }
        }
        n_inter += lane_count(inter);
    }
    if (H4E_LANE == 0) s->n_inter_mcb = n_inter;
}

/* BpicPlaneDec pass 2, h4m:1922-1967, symbol part only: vectors and records in ONE walk.  This is
   what a host thread runs (7% faster there than pb_mvs + schedule_rows, which the GPU uses and
   which h4e_seq_set_split_schedule selects on the host so that tests can compare the two). */
#if !defined(H4E_DEVICE)
H4E_FN void pb_pass2(H4Seq *s, int16_t *mv_out, Work *work, Cursors *cur)
{
    int32_t mvx = 0, mvy = 0;
    int cur_ref = -1;
    const int st0 = s->stride[0];
    s->n_inter_mcb = 0;
    BR mvh = s->mvh, mvv = s->mvv;
    for (int my = 0; my < s->mbh; ++my)
    {
        const uint8_t *ty0 = s->type[0] + cell_at(s, 0, 0, my * 2);
        const uint8_t *ty1 = s->type[1] + cell_at(s, 1, 0, my), *ty2 = s->type[2] + cell_at(s, 2, 0, my);
        for (int mx = 0; mx < s->mbw; ++mx)
        {
            const int lx = mx * 2;
            const uint8_t tag = ty0[lx];
            int16_t *mvp = mv_out + 2 * ((size_t)my * s->mbw + mx);
            const int mt = (tag >> 5) & 3;
            if (mt == 0)
            {   /* MCBlockDecDCNest, h4m:1789-1827 */
                mvp[0] = mvp[1] = 0;
                for (int k = 0; k < 6; ++k)
                {
                    const int p = k < 4 ? 0 : k - 3;
                    const int bx = k < 4 ? lx + SUBX[k] : mx, by = k < 4 ? my * 2 + SUBY[k] : my;
                    const uint32_t t = k < 4 ? ty0[SUBY[k] * st0 + lx + SUBX[k]] : k == 4 ? ty1[mx] : ty2[mx];
                    const uint32_t nib = t & 0xF;
                    if (nib != 0 && nib != 8) schedule_record(s, work, cur, t, 0, p, bx, by);
                }
                continue;
            }
            const int ref = mt - 1;
            s->n_inter_mcb++;
            if (ref != cur_ref)
            {   /* h4m:1943-1949 */
                cur_ref = ref;
                mvx = mvy = 0;
            }
            read_mv(s, &mvh, &mvx, s->rb[ref][0]);
            read_mv(s, &mvv, &mvy, s->rb[ref][1]);
            const int32_t rx = mx * 16 + mvx, ry = my * 16 + mvy;   /* h4m:1954-1955 */
            int needs_window = 0;
            if (!(tag & 0x10))
            {   /* MCBlockDecMCNest, h4m:1871-1909 */
                for (int k = 0; k < 6; ++k)
                {
                    const int p = k < 4 ? 0 : k - 3;
                    const int bx = k < 4 ? lx + SUBX[k] : mx, by = k < 4 ? my * 2 + SUBY[k] : my;
                    const uint32_t t = k < 4 ? ty0[SUBY[k] * st0 + lx + SUBX[k]] : k == 4 ? ty1[mx] : ty2[mx];
                    const uint32_t nib = t & 0xF;
                    if (!nib) continue;
                    schedule_record(s, work, cur, t, 0, p, bx, by);
                    if (nib > 1 && nib != 6) needs_window = 1;
                }
            }
            if (rx < -32000 || rx > 32000 || ry < -32000 || ry > 32000 || !mcb_refs_in_surface(s, rx, ry, needs_window))
            {
                s->err |= SYM_ERR_MV_RANGE;
                mvp[0] = mvp[1] = -32768;   /* poison: recon.cu paints the macroblock grey */
            }
            else
            {
                mvp[0] = (int16_t)rx;
                mvp[1] = (int16_t)ry;
            }
        }
    }
    s->mvh = mvh;
    s->mvv = mvv;
}
#endif

/* ------------------------------------------------------------------ entry points */

/* bulk copy into the blob: memcpy on the host, lane-strided on the GPU */
H4E_FN void blob_copy(uint8_t *dst, const void *src_, size_t n)
{
#if defined(H4E_DEVICE)
    const uint8_t *src = (const uint8_t *)src_;
    size_t words = 0;
    if (!(((uintptr_t)dst | (uintptr_t)src) & 3))
    {
        words = n >> 2;
        for (size_t i = (size_t)H4E_LANE; i < words; i += H4E_LANES) ((uint32_t *)dst)[i] = ((const uint32_t *)src)[i];
    }
    for (size_t i = words * 4 + (size_t)H4E_LANE; i < n; i += H4E_LANES) dst[i] = src[i];
#else
    memcpy(dst, src_, n);
#endif
}

H4E_TABLE uint8_t zero_pic_header[8 + 17 * 4] = {0};


/* serial: header, sections, trees */
H4E_FN void begin_setup(H4Seq *s, int pic_type, const uint8_t *pic, size_t pic_len)
{
    s->pic_type = pic_type;
    s->err = 0;
    s->n_inter_mcb = 0;
    s->blob_bytes = 0;
    s->setup_ok = 0;
    s->nest_x = s->nest_y = 0;
    const int is_i = pic_type == SYM_PIC_I;
    const int nsec = is_i ? 16 : 17;
    if (pic_type != SYM_PIC_I && pic_type != SYM_PIC_P && pic_type != SYM_PIC_B)
    {
        s->err |= SYM_ERR_GEOMETRY;
        s->errors_total |= s->err;
        return;
    }
    if (pic_len < (size_t)(8 + nsec * 4))
    {
        s->err |= SYM_ERR_TRUNCATED;
        pic = zero_pic_header;
        pic_len = sizeof zero_pic_header;
    }
    PROF_T0();
    const uint8_t *tab = pic + 8, *data = tab + nsec * 4;
    const size_t dlen = pic_len - 8 - (size_t)nsec * 4;
    if (is_i)
    {   /* h4m:1973-1977: the I picture keeps dc_shift local; state->dc_shift is P/B only */
        s->dc_shift = pic[0];
        s->unk_shift = pic[1];
        s->nest_x = pic[4] << 8 | pic[5];
        s->nest_y = pic[6] << 8 | pic[7];
    }
    else
    {   /* h4m:2021-2026 */
        s->dc_shift = pic[0];
        s->unk_shift = pic[1];
        s->rb[0][0] = pic[2]; s->rb[0][1] = pic[3];
        s->rb[1][0] = pic[4]; s->rb[1][1] = pic[5];
    }
    if (s->dc_shift > 7 || s->unk_shift > 24)
    {   /* int16 leaf << dc_shift overflows beyond 7 (h4m:616); shifts >= 32 are undefined */
        s->err |= SYM_ERR_GEOMETRY;
        if (s->dc_shift > 7) s->dc_shift = 7;
        if (s->unk_shift > 24) s->unk_shift = 24;
    }
    open_bits(s, &s->bn[0], data, dlen, rd_be32(tab + 0));
    open_bits(s, &s->bnr[0], data, dlen, rd_be32(tab + 4));
    open_bits(s, &s->bn[1], data, dlen, rd_be32(tab + 8));
    open_bits(s, &s->bnr[1], data, dlen, rd_be32(tab + 12));
    for (int p = 0; p < 3; ++p)
    {
        open_bits(s, &s->dcv[p], data, dlen, rd_be32(tab + 16 + 12 * p));
        open_bits(s, &s->sc[p], data, dlen, rd_be32(tab + 20 + 12 * p));
        open_bytes(s, &s->fix[p], data, dlen, rd_be32(tab + 24 + 12 * p));
    }
    if (is_i)
        for (int p = 0; p < 3; ++p) open_bits(s, &s->rle[p], data, dlen, rd_be32(tab + 52 + 4 * p));
    else
    {
        open_bits(s, &s->mvh, data, dlen, rd_be32(tab + 52));
        open_bits(s, &s->mvv, data, dlen, rd_be32(tab + 56));
        open_bits(s, &s->mcbt, data, dlen, rd_be32(tab + 60));
        open_bits(s, &s->mcbp, data, dlen, rd_be32(tab + 64));
    }
    /* tree order as h4m:1996-1999 / 2045-2050 */
    ht_read(&s->tree[T_BNUM], &s->bn[0], (uint32_t)(s->bn[0].end - s->bn[0].base), 0, 0);
    ht_read(&s->tree[T_RUN], &s->bnr[0], (uint32_t)(s->bnr[0].end - s->bnr[0].base), 0, 0);
    ht_read(&s->tree[T_DC], &s->dcv[0], (uint32_t)(s->dcv[0].end - s->dcv[0].base), 1, s->dc_shift);
    ht_read(&s->tree[T_SCALE], &s->sc[0], (uint32_t)(s->sc[0].end - s->sc[0].base), 0, 2);
    if (!is_i)
    {
        ht_read(&s->tree[T_MV], &s->mvh, (uint32_t)(s->mvh.end - s->mvh.base), 1, 0);
        ht_read(&s->tree[T_MCB], &s->mcbt, (uint32_t)(s->mcbt.end - s->mcbt.base), 0, 0);
    }
    for (int t = 0; t < 6; ++t)
        if (s->tree[t].bad) s->err |= SYM_ERR_BAD_TREE;
    s->dc_hi = 0x7F * (1 << s->dc_shift);   /* h4m:2001-2002, 2052-2053 */
    s->dc_lo = -0x80 * (1 << s->dc_shift);
    s->setup_ok = 1;
    PROF_ADD(0);
}

/* flat decode of every symbol-only section: pairs in lock step on a host thread, one section per
   lane on the GPU */
H4E_FN void begin_flat(H4Seq *s, int is_i)
{
    const HTab *tb = &s->tree[T_BNUM], *tr = &s->tree[T_RUN], *td = &s->tree[T_DC], *ts = &s->tree[T_SCALE];
    const int32_t lo = s->dc_lo, hi = s->dc_hi;
#if defined(H4E_DEVICE)
    SymStream *q = 0;
    const HTab *t = 0;
    BR *b = 0;
    int sovf = 0;
    switch (H4E_LANE)
    {   /* the large sections first; every lane then runs the SAME loop on its own section */
    case 0: q = &s->q_dcv[0]; t = td; b = &s->dcv[0]; sovf = 1; break;
    case 1: q = &s->q_bn[0]; t = tb; b = &s->bn[0]; break;
    case 2: q = &s->q_sc[0]; t = ts; b = &s->sc[0]; break;
    case 3: q = &s->q_dcv[1]; t = td; b = &s->dcv[1]; sovf = 1; break;
    case 4: q = &s->q_dcv[2]; t = td; b = &s->dcv[2]; sovf = 1; break;
    case 5: q = &s->q_sc[1]; t = ts; b = &s->sc[1]; break;
    case 6: q = &s->q_sc[2]; t = ts; b = &s->sc[2]; break;
    case 7: q = &s->q_bn[1]; t = tb; b = &s->bn[1]; break;
    case 8: q = &s->q_bnr[0]; t = tr; b = &s->bnr[0]; break;
    case 9: q = &s->q_bnr[1]; t = tr; b = &s->bnr[1]; break;
    case 10: if (is_i) { q = &s->q_rle[0]; t = tr; b = &s->rle[0]; } break;
    case 11: if (is_i) { q = &s->q_rle[1]; t = tr; b = &s->rle[1]; } break;
    case 12: if (is_i) { q = &s->q_rle[2]; t = tr; b = &s->rle[2]; } break;
    default: break;
    }
    ss_decode_lane(q, t, b, sovf, lo, hi);
#else
    /* paired by typical size so that the lock-step part covers most of both sections */
    ss_decode2(&s->q_dcv[0], td, &s->dcv[0], 1, &s->q_bn[0], tb, &s->bn[0], 0, lo, hi);
    ss_decode2(&s->q_sc[0], ts, &s->sc[0], 0, &s->q_bnr[0], tr, &s->bnr[0], 0, lo, hi);
    ss_decode2(&s->q_dcv[1], td, &s->dcv[1], 1, &s->q_dcv[2], td, &s->dcv[2], 1, lo, hi);
    ss_decode2(&s->q_sc[1], ts, &s->sc[1], 0, &s->q_sc[2], ts, &s->sc[2], 0, lo, hi);
    ss_decode2(&s->q_bn[1], tb, &s->bn[1], 0, &s->q_bnr[1], tr, &s->bnr[1], 0, lo, hi);
    if (is_i)
    {
        ss_decode2(&s->q_rle[0], tr, &s->rle[0], 0, &s->q_rle[1], tr, &s->rle[1], 0, lo, hi);
        ss_decode(&s->q_rle[2], tr, &s->rle[2]);
    }
#endif
}

/* maps (I) or pass 1 (P/B), record groups, blob plan */
H4E_FN void begin_maps(H4Seq *s, int is_i)
{
    PROF_T0();
    reset_record_counts(s);
    H4E_SYNC();
#if defined(H4E_DEVICE)
    const int split = 1;
#else
    const int split = s->split_schedule;
#endif
    if (is_i)
    {
        if (split)
        {
            ipic_types_split(s);
            ipic_dcs_split(s);
        }
        else
        {
            ipic_types(s);
            ipic_dcs(s);
        }
        if (H4E_LANE == 0) make_nest(s, s->nest_x, s->nest_y);
    }
    else if (split)
        pb_pass1_split(s);
    else
        pb_pass1(s);
    H4E_SYNC();
    PROF_ADD(1);
    plan_records(s, is_i);
    H4E_SYNC();
    if (H4E_LANE == 0) plan_blob(s);
    H4E_SYNC();
    PROF_ADD(2);
}

H4E_API size_t h4e_parse_begin(H4Seq *s, int pic_type, const uint8_t *pic, size_t pic_len)
{
    if (H4E_LANE == 0) begin_setup(s, pic_type, pic, pic_len);
    H4E_SYNC();
    if (!s->setup_ok) return 0;
    const int is_i = pic_type == SYM_PIC_I;
    {
        PROF_T0();
        begin_flat(s, is_i);
        H4E_SYNC();
        PROF_ADD(5);
    }
    begin_maps(s, is_i);
    return s->blob_bytes;
}

/* resolves the motion vectors (serial) and schedules every record (rows dealt to the lanes) */
H4E_FN void finish_schedule(H4Seq *s, uint8_t *blob)
{
    const SymHeader *h = &s->hdr;
    const int is_i = s->pic_type == SYM_PIC_I;
    PROF_T0();
    if (H4E_LANE == 0)
    {
        s->rec_base = (uint32_t *)(blob + h->off_rec);
        s->cur_work = work_scratch(s, s->n_records);
        if (!s->cur_work && s->n_records) s->err |= SYM_ERR_OVERFLOW;
    }
    H4E_SYNC();
    if (!s->cur_work && s->n_records) return;
#if defined(H4E_DEVICE)
    if (!is_i) pb_mvs(s, (int16_t *)(blob + h->off_mv));
    H4E_SYNC();
#endif
    PROF_ADD(7);
    Cursors cur;
    for (int p = 0; p < 3; ++p)
    {
        cur.fix[p] = s->fix[p].pos;
        cur.sc[p] = s->q_sc[p].pos;
        cur.dcv[p] = s->q_dcv[p].pos;   /* P/B: the pass-1 DC deltas come first in dc_values[p] */
    }
#if !defined(H4E_DEVICE)
    if (!is_i && !s->split_schedule)
    {
        pb_pass2(s, (int16_t *)(blob + h->off_mv), s->cur_work, &cur);
        PROF_ADD(3);
        return;
    }
    if (!is_i) pb_mvs(s, (int16_t *)(blob + h->off_mv));
#endif
    schedule_rows(s, is_i, s->cur_work, &cur);
    H4E_SYNC();
    PROF_ADD(3);
}

H4E_API uint32_t h4e_parse_finish(H4Seq *s, uint8_t *blob)
{
    const SymHeader *h = &s->hdr;
    const int is_i = s->pic_type == SYM_PIC_I;
    if (s->blob_bytes == 0) return s->err;
    finish_schedule(s, blob);
    if (!s->cur_work && s->n_records)
    {
        if (H4E_LANE == 0) s->errors_total |= s->err;
        return s->err;
    }
    PROF_T0();
    fill_records(s, s->cur_work, s->n_records);
    H4E_SYNC();
    PROF_ADD(6);
    blob_copy(blob + h->off_chunks, s->chunks, (size_t)s->n_chunks * 8);
    blob_copy(blob + h->off_bands, s->band_first, (size_t)SYM_REC_CLASSES * (s->nbands + 1) * 4);
    for (int p = 0; p < 3; ++p)
    {
        blob_copy(blob + h->off_type[p], s->type[p], s->map_cells[p]);
        blob_copy(blob + h->off_dc[p], s->dc[p], s->map_cells[p]);
    }
    if (h->has_nest) blob_copy(blob + h->off_nest, s->nest, SYM_NEST_BYTES);
    H4E_SYNC();
    if (H4E_LANE == 0)
    {
        /* every consumer must have stayed inside its section */
        int over = 0;
        for (int i = 0; i < 2; ++i) over |= s->q_bn[i].over | s->q_bnr[i].over;
        for (int p = 0; p < 3; ++p) over |= s->q_dcv[p].over | s->q_sc[p].over | (is_i ? s->q_rle[p].over : 0);
        if (!is_i) over |= br_overrun(&s->mvh) | br_overrun(&s->mvv) | br_overrun(&s->mcbt) | br_overrun(&s->mcbp);
        if (over) s->err |= SYM_ERR_TRUNCATED;
        SymHeader out = *h;
        out.errors = s->err;
        memcpy(blob, &out, sizeof out);
        s->errors_total |= s->err;
    }
    H4E_SYNC();
    PROF_ADD(4);
    return s->err;
}
