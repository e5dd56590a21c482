/*
 * recon_core.h -- per-4x4-block pixel reconstruction, the arithmetic heart of recon.cu.
 *
 * Everything here is integer, two's-complement, 32-bit, and reproduces the reference
 * decoder bit for bit ("h4m:N" = /root/reference/h4m_audio_decode.c line N):
 *   weighted DC fill      h4m:293-383     flat / raw blocks    h4m:281, 543-549
 *   AOT basis + sum       h4m:679-817     intra AOT block      h4m:1358-1377
 *   half-sample MC        h4m:1242-1294   predicted AOT block  h4m:1379-1420
 *   neighbour-DC rule     h4m:1437-1441, 1811-1814
 *   per-plane MC address and phase (1.3 vs 1.5)   h4m:1327-1355, 1862-1910
 *
 * The functions are __host__ __device__ so that tests/emul/ can run the very same code
 * serially on the CPU before a GPU is available (test infrastructure; the product only
 * ever calls them from the CUDA kernels).  A block's result is returned as four 32-bit
 * words, one per row, leftmost pixel in the low byte.
 *
 * Data-parallel formulation (what makes the kernel cheap per block):
 *   - a basis row (4 samples, step 1 or 2) is ONE 32-bit load from a table that holds, for
 *     every nest row y and start x, the 8 nibbles x..x+7 (rc_build_nest_table), or two
 *     64-bit loads from the reference frame for the inter "window" nest;
 *   - the four half-sample filters are one branch-free formula (rc_predict):
 *     (p00 + p01 + p10 + p11 + 2) >> 2 with p01/p10/p11 aliased to p00 when the phase bit
 *     is 0 -- (4a+2)>>2 = a, (2a+2b+2)>>2 = (a+b+1)>>1 -- evaluated on packed bytes.
 */
#ifndef HVQM4_RECON_CORE_H
#define HVQM4_RECON_CORE_H

#include <stdint.h>
#include "symbuf.h"

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ __forceinline__
#define RC_HDM __host__ __device__ __forceinline__   /* member functions */
#else
#define RC_HD static inline
#define RC_HDM inline
#endif

#if defined(__CUDA_ARCH__)
#if defined(RC_PLAIN_LOADS)   /* sweep.cu: symbol data and reference rows are staged in shared memory */
#define RC_LD8(p) (*(const uint8_t *)(p))
#define RC_LD32(p) (*(const uint32_t *)(p))
#define RC_LD64(p) (*(const unsigned long long *)(p))
#else
#define RC_LD8(p) __ldg((const uint8_t *)(p))
#define RC_LD32(p) __ldg((const uint32_t *)(p))
#define RC_LD64(p) __ldg((const unsigned long long *)(p))
#endif
#define RC_PRMT(a, b, s) __byte_perm((a), (b), (s))
#define RC_ADDMIN_U16X2(a, b, c) __viaddmin_u16x2((a), (b), (c))   /* VIADDMNMX.U16x2: per half min((a + b) mod 2^16, c) */
#define RC_DP4A(a, b, c) __dp4a((uint32_t)(a), (uint32_t)(b), (uint32_t)(c))   /* IDP.4A.U8.U8: c + sum of byte products */
#else
#define RC_LD8(p) (*(const uint8_t *)(p))
#define RC_LD32(p) (*(const uint32_t *)(p))
#define RC_LD64(p) (*(const unsigned long long *)(p))
static inline uint32_t rc_prmt_host(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = (uint64_t)b << 32 | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
#define RC_PRMT(a, b, s) rc_prmt_host((a), (b), (s))
static inline uint32_t rc_addmin_u16x2_host(uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t lo = (a + b) & 0xFFFFu, hi = ((a >> 16) + (b >> 16)) & 0xFFFFu;
    const uint32_t clo = c & 0xFFFFu, chi = c >> 16;
    return (lo < clo ? lo : clo) | (hi < chi ? hi : chi) << 16;
}
#define RC_ADDMIN_U16X2(a, b, c) rc_addmin_u16x2_host((a), (b), (c))
static inline uint32_t rc_dp4a_host(uint32_t a, uint32_t b, uint32_t c)
{
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
    return c;
}
#define RC_DP4A(a, b, c) rc_dp4a_host((a), (b), (c))
#endif

#define RC_NEST_PITCH 68                              /* entries per nest row: start columns 0..67 */
#define RC_NEST_PITCH_PORTRAIT 36                     /* portrait nest (38 wide, 70 rows): start columns 0..35; 70 * 36 <= 38 * 68 words */
#define RC_NEST_TABLE_WORDS (SYM_NEST_H * RC_NEST_PITCH)

/* The three lookup tables live at fixed offsets at the start of dynamic shared memory on the
   GPU (no pointer registers); on the CPU they are reached through the view. */
#define RC_SMEM_NEST_OFF 0
#define RC_SMEM_MCDIV_OFF (RC_NEST_TABLE_WORDS * 4)
#define RC_SMEM_DIV_OFF (RC_SMEM_MCDIV_OFF + 256 * 4)
#define RC_SMEM_VIEW_OFF (RC_SMEM_DIV_OFF + 16 * 4)
#define RC_SMEM_TABLE_BYTES (RC_SMEM_VIEW_OFF + 256)   /* ReconView of the CTA's picture lives in shared memory too */
#if defined(__CUDACC__)
extern __shared__ __align__(128) uint8_t rc_smem[];   /* 128: destinations of tensor copies (row.cu) */
#endif
#if defined(__CUDA_ARCH__)
#define RC_NEST_TAB(v) (reinterpret_cast<const uint32_t *>(rc_smem + RC_SMEM_NEST_OFF))
#define RC_MCDIV(v, i) (reinterpret_cast<const int32_t *>(rc_smem + RC_SMEM_MCDIV_OFF)[i])
#define RC_DIV16(v, i) (reinterpret_cast<const int32_t *>(rc_smem + RC_SMEM_DIV_OFF)[i])   /* divTable[i] / 16 */
#else
#define RC_NEST_TAB(v) ((v).nest_tab)
#define RC_MCDIV(v, i) ((v).mcdiv_tab[i])
#define RC_DIV16(v, i) ((v).div_tab[i] >> 4)
#endif

/* Read-only view of one picture job. */
struct ReconView
{
    const uint8_t *blob;       /* symbol buffer (symbuf.h) */
    const uint32_t *nest_tab;  /* rc_build_nest_table() output: shared memory on the GPU */
    const int32_t *div_tab;    /* 16 entries,  h4m:262,270 */
    const int32_t *mcdiv_tab;  /* 256 entries used of h4m:263,272 */
    const uint8_t *ref[2];     /* past, future frame surfaces (Y|U|V contiguous) */
    int width, height;
    int is_ipic, version15;
    int unk_shift;
    uint32_t off_type[3], off_dc[3], off_mv;
    int mcb_w;
    /* used by the kernels only */
    uint8_t *present;
    const uint32_t *rec, *chunks, *bands;
    int nseg, mcb_h, has_nest;
    int portrait;              /* width < height: nest 38 x 70, basis descriptor axes swapped (h4m:700-711, 743-754) */
    uint32_t off_nest, n_chunks, n_chunks_nest, n_bands;
};

RC_HD void rc_make_view(ReconView &v, const uint8_t *blob, const SymHeader &h, const uint32_t *nest_tab,
                        const int32_t *div_tab, const int32_t *mcdiv_tab, const uint8_t *past, const uint8_t *future)
{
    v.blob = blob; v.nest_tab = nest_tab; v.div_tab = div_tab; v.mcdiv_tab = mcdiv_tab;
    v.ref[0] = past; v.ref[1] = future;
    v.width = h.width; v.height = h.height;
    v.is_ipic = h.pic_type == SYM_PIC_I; v.version15 = h.version15;
    v.unk_shift = h.unk_shift;
    for (int p = 0; p < 3; ++p) { v.off_type[p] = h.off_type[p]; v.off_dc[p] = h.off_dc[p]; }
    v.off_mv = h.off_mv;
    v.mcb_w = h.mcb_w;
    v.present = nullptr;
    v.rec = reinterpret_cast<const uint32_t *>(blob + h.off_rec);
    v.chunks = reinterpret_cast<const uint32_t *>(blob + h.off_chunks);
    v.bands = reinterpret_cast<const uint32_t *>(blob + h.off_bands);
    v.n_bands = h.n_bands;
    v.nseg = h.nseg; v.mcb_h = h.mcb_h; v.has_nest = h.has_nest;
    v.portrait = h.portrait;
    v.off_nest = h.off_nest; v.n_chunks = h.n_chunks; v.n_chunks_nest = h.n_chunks_nest;
}

/* Nest lookup table.  rc_nest_table_entry: the nibbles x..x+7 of packed nest row y (zero past
   column 69) as one word; the table the kernels read is derived from it: entry (y, x), x = 0..67,
   = samples x, x+1, x+2, x+3, one per byte and ALREADY MULTIPLIED BY 16 (the sample sits in the
   high nibble, exactly like a reference pixel masked with 0xF0), so that intra and inter bases
   share the arithmetic below.  A step-1 basis row is one load; a step-2 row (samples x, x+2, x+4,
   x+6) is bytes 0 and 2 of entries x and x + 4.  (Until round 2 there was a second, pre-permuted
   step-2 table: 19 KB instead of 10 KB of shared memory per CTA, which the sweep kernel needs for
   its reference windows.) */
RC_HD uint32_t rc_nest_spread_step1(uint32_t nibbles8)
{
    uint32_t x = nibbles8 & 0xFFFFu;                   /* nibbles 0,1,2,3 -> one per byte */
    x = (x | x << 8) & 0x00FF00FFu;
    return ((x | x << 4) & 0x0F0F0F0Fu) << 4;
}

RC_HD uint32_t rc_nest_table_entry(const uint8_t *packed, int y, int x, int row_bytes = SYM_NEST_ROW_BYTES)
{
    const uint8_t *row = packed + y * row_bytes;
    uint64_t bits = 0;
    const int b0 = x >> 1;
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (b0 + i < row_bytes) bits |= (uint64_t)row[b0 + i] << (8 * i);
    return (uint32_t)(bits >> ((x & 1) * 4));
}

/* runtime-indexed picks without a local-memory array */
RC_HD uint32_t rc_pick3(const uint32_t a[3], int i) { return i == 0 ? a[0] : i == 1 ? a[1] : a[2]; }

RC_HD uint32_t rc_clamp255(int32_t x) { return x < 0 ? 0u : x > 255 ? 255u : (uint32_t)x; }

/* ---- weighted DC fill (h4m:293-383) ------------------------------------------------
 * out(r,c) = sat_mean8(8V + rowterm[r] + colterm[c]) with
 *   rowterm = {2T-B-V, V-B, V-T, 2B-T-V},  colterm = {2L-R-V, V-R, V-L, 2R-L-V}
 * which is the sixteen expressions of the reference regrouped.  sat_mean8 divides
 * (sum + 4) by 8 as an UNSIGNED number and then clamps: sums <= -5 -> 255, -4..-1 -> 0. */
RC_HD uint32_t rc_sat_mean8(int32_t sum)
{
    const uint32_t q = ((uint32_t)sum + 4u) >> 3;
    return q > 255u ? 255u : q;
}

/* Two columns per 32-bit word in 16-bit halves (the sixteen sums lie in -765..2805):
 *   s' = sum + 4 + 1024 > 0 in every half, so packed words add without carries between halves
 *        (row term + 516 replicated by one IMAD, column terms + 512 packed pairwise);
 *   q' = s' >> 3 = floor((sum + 4) / 8) + 128;
 *   out = min_u16((q' - 128) mod 2^16, 255): a negative quotient wraps to >= 0xFFA0 and clamps to
 *        255 exactly like the reference's unsigned division (h4m:293-296).
 * 9 instructions per row of four pixels. */
RC_HD void rc_weighted(uint32_t rows[4], int V, int T, int B, int L, int R)
{
    const uint32_t c0 = (uint32_t)(2 * L - R - V + 512), c1 = (uint32_t)(V - R + 512);
    const uint32_t c2 = (uint32_t)(V - L + 512), c3 = (uint32_t)(2 * R - L - V + 512);
    const uint32_t c01 = c1 * 0x10000u + c0, c23 = c3 * 0x10000u + c2;
    /* r_k = 8V + rowterm[k] + 4 + 512 */
    const uint32_t r[4] = {(uint32_t)(7 * V + 2 * T - B + 516), (uint32_t)(9 * V - B + 516),
                           (uint32_t)(9 * V - T + 516), (uint32_t)(7 * V + 2 * B - T + 516)};
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const uint32_t q01 = ((r[k] * 0x10001u + c01) >> 3) & 0x1FFF1FFFu;
        const uint32_t q23 = ((r[k] * 0x10001u + c23) >> 3) & 0x1FFF1FFFu;
        rows[k] = RC_PRMT(RC_ADDMIN_U16X2(q01, 0xFF80FF80u, 0x00FF00FFu), RC_ADDMIN_U16X2(q23, 0xFF80FF80u, 0x00FF00FFu), 0x6420);
    }
}

/* ---- AOT bases (h4m:679-817) --------------------------------------------------------
 * side word: bits 15:0 descriptor ([5:0] x, [10:6] y, [11] x step 2, [12] y step 2,
 * [14:13] scale offset, [15] negate), bits 23:16 scale symbol (>> 2). */

/* Basis rows are handled as four bytes holding 16 * sample (sample in the high nibble).  The
 * reference's factor (scale_sum + offset) * +-divTable[range] (h4m:712-731) is a multiple of 16
 * because every divTable entry is (h4m:270), so  factor * sample == (factor / 16) * (16 * sample)
 * as integers, hence also mod 2^32: the kernels multiply by RC_DIV16 = divTable / 16 and never
 * shift the samples down.
 *
 * Pipe balance (measured on B200, tools/ubench/pipes.cu): LOP3/SHF/PRMT issue at half rate on the
 * ALU pipe, IMAD and IDP.4A at half rate on the FMA pipe, and the two pipes run side by side.  The
 * byte extracts are therefore split: IDP.4A with a one-hot weight (FMA pipe) and PRMT (ALU pipe). */

/* Where reference pixels are read from.  The block functions see a patch as ROWS of aligned 32-bit words:
 * rows.ld(r, k) = aligned word k of row r (r, k compile-time after unrolling).  RcLinearRows / RcLinearWindow
 * address a frame surface in global memory (map, record and band kernels, and the CPU emulation);
 * sweep_core.h adds the same two for the shared-memory ring of the sweep kernel. */
struct RcLinearRows
{
    const uint8_t *base;   /* first byte of row 0, aligned down to 4 */
    int pitch;             /* bytes between the patch's rows */
    RC_HDM uint32_t ld(int r, int k) const { return RC_LD32(base + r * pitch + 4 * k); }
};
/* the 70x38 luma window of a macroblock's reference position (h4m:1864-1868) */
struct RcLinearWindow
{
    const uint8_t *origin;
    int width;
    /* rows oy, oy + ys, ... starting at column ox; a = misalignment of the first byte */
    RC_HDM RcLinearRows rows(int ox, int oy, int ys, uint32_t &a) const
    {
        const uint8_t *p = origin + oy * width + ox;
        a = (uint32_t)((uintptr_t)p & 3);
        return RcLinearRows{p - a, ys * width};
    }
};

/* bytes a .. a+7 of the aligned words w0, w1, w2 (a = 0..3), then samples at step 1 or 2, masked to the high nibbles */
RC_HD uint32_t rc_row_window(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t sel_align, uint32_t sel_step)
{
    const uint32_t lo = RC_PRMT(w0, w1, sel_align), hi = RC_PRMT(w1, w2, sel_align);
    return RC_PRMT(lo, hi, sel_step) & 0xF0F0F0F0u;
}

/* accumulates one basis given its four packed rows (bytes = 16 * sample); kFmaExtracts of the
   four byte extracts per row go through IDP.4A, the rest through PRMT */
template <int kFmaExtracts>
RC_HD void rc_accumulate(const ReconView &v, uint32_t word, const uint32_t R[4], int32_t &scale_sum, int32_t acc[16])
{
    uint32_t b[16];
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x)
            b[y * 4 + x] = x < kFmaExtracts ? RC_DP4A(R[y], 1u << (8 * x), 0u) : RC_PRMT(R[y], 0u, 0x4440 + x);
    uint32_t lo = b[0], hi = b[0];
#pragma unroll
    for (int i = 1; i < 16; ++i)
    {
        lo = b[i] < lo ? b[i] : lo;
        hi = b[i] > hi ? b[i] : hi;
    }
    scale_sum += (int32_t)((word >> 16) & 0xFF) << 2;            /* cumulative within the block, h4m:726,781 */
    int32_t inv = RC_DIV16(v, (hi - lo) >> 4);
    if (word & 0x8000) inv = -inv;
    const uint32_t factor = (uint32_t)(scale_sum + (int32_t)((word >> 13) & 3)) * (uint32_t)inv;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (int32_t)((uint32_t)acc[i] + factor * b[i]);   /* mod 2^32 */
}

/* kInter = false: intra (nest table, the window is not looked at); true: inter (reference luma window).  The window
   policy is passed by value: a pointer to a local policy object put it on the stack (56 bytes of frame, 180 bytes of
   spills in the band kernel at 80 registers). */
struct RcNoWindow { };
template <bool kInter, class Win>
RC_HD void rc_add_basis(const ReconView &v, uint32_t word, const Win window, int32_t &scale_sum, int32_t acc[16])
{
    /* landscape: [5:0] column, [10:6] row, [11] column step 2, [12] row step 2; portrait: the other way round
       (h4m:700-711, 743-754) */
    const int f6 = word & 0x3F, f5 = (word >> 6) & 0x1F;
    const uint32_t s11 = (word >> 11) & 1, s12 = (word >> 12) & 1;
    const bool portrait = v.portrait != 0;
    const int ox = portrait ? f5 : f6, oy = portrait ? f6 : f5;
    const uint32_t xs2 = portrait ? s12 : s11;
    const int ys = 1 + (int)(portrait ? s11 : s12);
    uint32_t R[4];
    if constexpr (kInter)
    {
        /* sample = (pixel >> 4) & 0xF (h4m:756-761): rows are fetched as up to three aligned words */
        uint32_t a;
        const auto rows = window.rows(ox, oy, ys, a);
        const uint32_t sel_align = 0x3210u + 0x1111u * a, sel_step = xs2 ? 0x6420u : 0x3210u;
        const bool need1 = xs2 || a >= 1, need2 = xs2 && a >= 2;
#pragma unroll
        for (int y = 0; y < 4; ++y)
        {
            /* the samples end at byte a + 3 (step 1) or a + 6 (step 2): only the words they reach are fetched */
            const uint32_t w0 = rows.ld(y, 0), w1 = need1 ? rows.ld(y, 1) : 0u, w2 = need2 ? rows.ld(y, 2) : 0u;
            R[y] = rc_row_window(w0, w1, w2, sel_align, sel_step);
        }
        rc_accumulate<4>(v, word, R, scale_sum, acc);
    }
    else
    {
        const int pitch = portrait ? RC_NEST_PITCH_PORTRAIT : RC_NEST_PITCH;
        const uint32_t *tab = RC_NEST_TAB(v) + oy * pitch + ox;
#pragma unroll
        for (int y = 0; y < 4; ++y)
        {
            const uint32_t e0 = tab[y * ys * pitch];
            R[y] = xs2 ? RC_PRMT(e0, tab[y * ys * pitch + 4], 0x6420) : e0;
        }
        rc_accumulate<2>(v, word, R, scale_sum, acc);
    }
}

/* kGen: the side words are read with plain (generic) loads -- the band kernel stages a band's records in shared memory
   when they fit and leaves them in global memory when they do not, one code path for both */
template <bool kGen>
RC_HD uint32_t rc_ld_side(const uint32_t *p)
{
    if (kGen) return *p;
    return RC_LD32(p);
}

template <bool kInter, bool kGen = false, class Win>
RC_HD int32_t rc_aot_sum(const ReconView &v, const uint32_t *side, int n, const Win window, int32_t acc[16])
{
    int32_t scale_sum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    /* the first four basis words are fetched together (one memory latency instead of one per basis) */
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = k < n ? rc_ld_side<kGen>(side + k) : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < n) rc_add_basis<kInter>(v, w[k], window, scale_sum, acc);
    for (int k = 4; k < n; ++k) rc_add_basis<kInter>(v, rc_ld_side<kGen>(side + k), window, scale_sum, acc);
    uint32_t total = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) total += (uint32_t)acc[i];
    return (int32_t)total >> 4;                                   /* arithmetic, h4m:793,815 */
}

/* IntraAotBlock, h4m:1358-1377 */
template <bool kGen = false>
RC_HD void rc_intra_aot(const ReconView &v, uint32_t rows[4], const uint32_t *side, int n, int V)
{
    int32_t acc[16];
    const int32_t mean = rc_aot_sum<false, kGen>(v, side, n, RcNoWindow{}, acc);
    /* modulo 2^32 like the reference's int32 on its targets (damaged scale symbols overflow it) */
    const uint32_t delta = ((uint32_t)V << v.unk_shift) - (uint32_t)mean;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        uint32_t out = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) out |= rc_clamp255((int32_t)((uint32_t)acc[r * 4 + c] + delta) >> v.unk_shift) << (8 * c);
        rows[r] = out;
    }
}

/* ---- half-sample prediction (h4m:1242-1294) ----------------------------------------------
 * Rows are read as two aligned 32-bit words (covers the 5 bytes a row can need).  Two exact
 * formulations, chosen per WARP so that lanes do not diverge:
 *   no lane needs both half steps:  out = (a + o + 1) >> 1 with o = the tap one step right, or
 *                                   one step down, or a itself ((a + a + 1) >> 1 = a);
 *   otherwise:                      out = (p00 + p01 + p10 + p11 + 2) >> 2 with taps aliased when
 *                                   a phase bit is 0 ((4a+2)>>2 = a, (2a+2b+2)>>2 = (a+b+1)>>1). */
RC_HD uint32_t rc_avg4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) & 0xFEFEFEFEu) >> 1); }

/* rows W[2r], W[2r+1] = the two aligned words that cover the 5 bytes of reference row r; a = misalignment of
   the patch's first byte */
/* kSkipSecond: an aligned row without a horizontal half step ends in its first word, the second is then not
   requested (dense content, band kernel: +0.4 %; the ALU-bound map kernel on sparse content loses 0.6 % to the
   predicate, so it keeps the unconditional pair) */
template <bool kSkipSecond, class Rows>
RC_HD void rc_predict_load_rows(uint32_t W[10], const Rows &rows, uint32_t a, int hx, int hy)
{
    const bool second = !kSkipSecond || a != 0 || hx;
#pragma unroll
    for (int r = 0; r < 5; ++r)
    {
        if (r < 4 || hy)
        {
            W[2 * r] = rows.ld(r, 0);
            W[2 * r + 1] = second ? rows.ld(r, 1) : 0u;
        }
        else
            W[2 * r] = W[2 * r + 1] = 0;
    }
}

template <bool kSkipSecond>
RC_HD void rc_predict_load(uint32_t W[10], const uint8_t *src, int stride, int hx, int hy)
{
    const uint32_t a = (uint32_t)((uintptr_t)src & 3);
    rc_predict_load_rows<kSkipSecond>(W, RcLinearRows{src - a, stride}, a, hx, hy);
}

/* a = src & 3.  Alignment is a byte permute (A = bytes a..a+3, B = bytes a+1..a+4 of a row).
 *   no lane needs both half steps:  out = (A + O + 1) >> 1 byte-wise with O = B, the next row's A,
 *                                   or A itself;
 *   otherwise:  every output pixel is two IDP.4A (FMA pipe) -- 2 + this row's taps, + the next
 *               row's taps -- with per-lane weight words that encode the phase: hx picks taps
 *               {x, x+1} with weight 1 or tap {x} with weight 2, hy moves half of the weight to
 *               the next row or doubles this row's.  (4a+2)>>2 = a, (2a+2b+2)>>2 = (a+b+1)>>1, so
 *               this is exactly the four _MotionComp_xy variants (h4m:1242-1294), without a
 *               divergent branch and with the ALU pipe left to the alignment and the packing. */
RC_HD void rc_predict_filter(uint32_t rows[4], const uint32_t W[10], uint32_t a, int hx, int hy, bool any_diag)
{
    const uint32_t sel = 0x3210u + 0x1111u * a;
    uint32_t A[5], B[5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
    {
        A[r] = RC_PRMT(W[2 * r], W[2 * r + 1], sel);
        B[r] = RC_PRMT(W[2 * r], W[2 * r + 1], sel + 0x1111u);
    }
    if (!any_diag)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r) rows[r] = rc_avg4(A[r], hx ? B[r] : hy ? A[r + 1] : A[r]);
        return;
    }
    /* weights of pixel 0 on A (pixels 1, 2: shifted left by 8, 16) and of pixel 3 on B */
    const uint32_t wx = hx ? 0x0101u : 0x0002u, wx3 = hx ? 0x01010000u : 0x00020000u;
    const uint32_t m0 = hy ? 1u : 2u, m1 = hy ? 1u : 0u;     /* this row, next row */
    const uint32_t u0 = wx * m0, u3 = wx3 * m0, v0 = wx * m1, v3 = wx3 * m1;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const uint32_t p0 = RC_DP4A(A[r + 1], v0, RC_DP4A(A[r], u0, 2u));
        const uint32_t p1 = RC_DP4A(A[r + 1], v0 << 8, RC_DP4A(A[r], u0 << 8, 2u));
        const uint32_t p2 = RC_DP4A(A[r + 1], v0 << 16, RC_DP4A(A[r], u0 << 16, 2u));
        const uint32_t p3 = RC_DP4A(B[r + 1], v3, RC_DP4A(B[r], u3, 2u));
        /* two 16-bit lanes per word (sums <= 1022), >> 2, then bytes 0 and 2 of each */
        rows[r] = RC_PRMT((p1 * 0x10000u + p0) >> 2, (p3 * 0x10000u + p2) >> 2, 0x6420);
    }
}

RC_HD void rc_predict(uint32_t rows[4], const uint8_t *src, int stride, int hx, int hy)
{
#if defined(__CUDA_ARCH__)
    const bool any_diag = __any_sync(__activemask(), hx & hy);
#else
    const bool any_diag = hx & hy;
#endif
    uint32_t W[10];
    rc_predict_load<false>(W, src, stride, hx, hy);
    rc_predict_filter(rows, W, (uint32_t)((uintptr_t)src & 3), hx, hy, any_diag);
}

/* ---- predicted AOT block (h4m:1379-1420); rows[] holds the MC prediction on entry ------
 * The prediction stays packed in rows[]: its sum is four byte-wise dot products with 1, its
 * range a min/max over transient byte extracts; samples are re-extracted in the output loop,
 * so only the 16 accumulators stay live across the basis loop. */
RC_HD uint32_t rc_sum4(uint32_t packed, uint32_t acc)
{
#if defined(__CUDA_ARCH__)
    return __dp4a(packed, 0x01010101u, acc);
#else
    return acc + (packed & 0xFF) + ((packed >> 8) & 0xFF) + ((packed >> 16) & 0xFF) + (packed >> 24);
#endif
}

template <bool kGen = false, class Win>
RC_HD void rc_predicted_aot(const ReconView &v, uint32_t rows[4], const uint32_t *side, int nibble, const Win window)
{
    int32_t acc[16];
    const uint32_t aot_mean = (uint32_t)rc_aot_sum<true, kGen>(v, side, nibble - 1, window, acc);
    const uint32_t pair = rc_ld_side<kGen>(side + nibble - 1);
    const int32_t mean = (int32_t)(rc_sum4(rows[3], rc_sum4(rows[2], rc_sum4(rows[1], rc_sum4(rows[0], 8u)))) >> 4);
    int32_t lo = 255, hi = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        const int32_t m = (int32_t)((rows[i >> 2] >> ((i & 3) * 8)) & 0xFF);
        lo = m < lo ? m : lo;
        hi = m > hi ? m : hi;
    }
    const int32_t s1 = (int32_t)(int16_t)(pair & 0xFFFF), s2 = (int32_t)(int16_t)(pair >> 16);
    const uint32_t factor = (uint32_t)s2 * (uint32_t)RC_MCDIV(v, hi - lo);   /* range of (m - mean) = range of m */
    /* acc + addend + (m - mean) * factor  =  acc + (addend - mean * factor) + m * factor */
    const uint32_t addend = ((uint32_t)s1 << v.unk_shift) - aot_mean - (uint32_t)mean * factor;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        uint32_t out = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
        {
            const uint32_t m = (rows[r] >> (8 * c)) & 0xFF;
            const int32_t res = (int32_t)((uint32_t)acc[r * 4 + c] + addend + m * factor);
            out |= rc_clamp255((res >> v.unk_shift) + (int32_t)m) << (8 * c);
        }
        rows[r] = out;
    }
}

/* ---- block classes -----------------------------------------------------------------------
 * RC_DIRECT  flat DC or raw: a handful of instructions, done while classifying
 * RC_WEIGHTED weighted DC fill (needs the four neighbour cells)
 * RC_MC      motion compensation only (proc-1 macroblock, or nibble 0 of a proc-0 one)
 * RC_AOT_INTRA / RC_AOT_INTER  blocks with a basis loop */
enum { RC_DIRECT = 0, RC_WEIGHTED = 1, RC_MC = 2, RC_AOT_INTRA = 3, RC_AOT_INTER = 4, RC_NCLASS = 5 };

RC_HD int rc_classify(uint32_t t, int is_ipic)
{
    const uint32_t nib = is_ipic ? t : (t & 0xF);
    if (!is_ipic && (t & 0x60))
    {
        if (t & 0x10) return RC_MC;
        return nib == 0 ? RC_MC : nib == 6 ? RC_DIRECT : RC_AOT_INTER;
    }
    return nib == 0 ? RC_WEIGHTED : (nib == 8 || nib == 6) ? RC_DIRECT : RC_AOT_INTRA;
}

/* ---- motion of one 4x4 block of an inter macroblock -----------------------------------------
 * Resolved in two steps so that the band kernel can do the first one while it classifies blocks
 * (vector table rows read coalesced) and queue the result:
 *   rc_motion_pack   (vector word, block coordinates) -> one word: [26:0] byte offset of the
 *                    integer-sample position inside the reference surface, [27] hx, [28] hy,
 *                    [29] reference (0 past, 1 future), [30] poisoned (SYM_ERR_MV_RANGE: never
 *                    dereferenced, painted grey)
 *   rc_mc_packed     that word -> the prediction                                              */
#define RC_MP_HX (1u << 27)
#define RC_MP_HY (1u << 28)
#define RC_MP_FUTURE (1u << 29)
#define RC_MP_POISON (1u << 30)
#define RC_MP_OFFSET(w) ((w) & 0x07FFFFFFu)

RC_HD uint32_t rc_mv_word(const ReconView &v, int plane, int bx, int by)
{
    const int sh = plane ? 1 : 0;
    const int mx = bx >> (1 - sh), my = by >> (1 - sh);
    return RC_LD32(v.blob + v.off_mv + 4 * (my * v.mcb_w + mx));
}

RC_HD uint32_t rc_motion_pack(const ReconView &v, int plane, int bx, int by, uint32_t t, uint32_t mvw)
{
    const int sh = plane ? 1 : 0;
    const int pw = v.width >> sh;
    const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
    if (rx == -32768) return RC_MP_POISON;
    const int px = rx >> sh, py = ry >> sh;
    /* 1.3: luma phase for every plane (h4m:1329-1330,1869-1870); 1.5: per-plane phase (h4m:1337-1343,1890-1896) */
    const uint32_t hx = (uint32_t)(v.version15 ? px : rx) & 1u, hy = (uint32_t)(v.version15 ? py : ry) & 1u;
    const int plane_off = plane == 0 ? 0 : plane == 1 ? v.width * v.height : v.width * v.height + (v.width >> 1) * (v.height >> 1);
    /* linear addressing, no clamping (h4m:1344,1897); sub-block offset = pb_offset (h4m:866-869) */
    const int subx = plane == 0 ? (bx & 1) * 4 : 0, suby = plane == 0 ? (by & 1) * 4 : 0;
    const uint32_t off = (uint32_t)(plane_off + ((py >> 1) + suby) * pw + (px >> 1) + subx);
    return (off & 0x07FFFFFFu) | hx << 27 | hy << 28 | (((t >> 5) & 3) == 2 ? RC_MP_FUTURE : 0u);
}

/* origin of the 70x38 luma window of the macroblock's reference position (h4m:1864-1868); NULL if poisoned */
RC_HD const uint8_t *rc_motion_window(const ReconView &v, uint32_t t, uint32_t mvw)
{
    const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
    if (rx == -32768) return nullptr;
    const uint8_t *ref = ((t >> 5) & 3) == 2 ? v.ref[1] : v.ref[0];
    /* 70 x 38 at (-32, -16) in landscape, 38 x 70 at (-16, -32) in portrait pictures */
    return v.portrait ? ref + rx / 2 + (ry / 2 - 32) * v.width - 16 : ref + rx / 2 + (ry / 2 - 16) * v.width - 32;
}

RC_HD void rc_mc_packed(const ReconView &v, int plane, uint32_t mp, uint32_t rows[4])
{
    if (mp & RC_MP_POISON) { rows[0] = rows[1] = rows[2] = rows[3] = 0x80808080u; return; }
    const uint8_t *ref = (mp & RC_MP_FUTURE) ? v.ref[1] : v.ref[0];
    rc_predict(rows, ref + RC_MP_OFFSET(mp), v.width >> (plane ? 1 : 0), (int)((mp >> 27) & 1), (int)((mp >> 28) & 1));
}

RC_HD void rc_mc_packed2(const ReconView &v, int plane0, uint32_t mp0, uint32_t rows0[4], int plane1, uint32_t mp1, uint32_t rows1[4])
{
    const int hx0 = (int)((mp0 >> 27) & 1), hy0 = (int)((mp0 >> 28) & 1), hx1 = (int)((mp1 >> 27) & 1), hy1 = (int)((mp1 >> 28) & 1);
#if defined(__CUDA_ARCH__)
    const bool any_diag = __any_sync(__activemask(), (hx0 & hy0) | (hx1 & hy1));
#else
    const bool any_diag = (hx0 & hy0) | (hx1 & hy1);
#endif
    const uint8_t *src0 = ((mp0 & RC_MP_FUTURE) ? v.ref[1] : v.ref[0]) + RC_MP_OFFSET(mp0);
    const uint8_t *src1 = ((mp1 & RC_MP_FUTURE) ? v.ref[1] : v.ref[0]) + RC_MP_OFFSET(mp1);
    uint32_t W0[10], W1[10];
    const bool ok0 = !(mp0 & RC_MP_POISON), ok1 = !(mp1 & RC_MP_POISON);
    if (ok0) rc_predict_load<true>(W0, src0, v.width >> (plane0 ? 1 : 0), hx0, hy0);
    if (ok1) rc_predict_load<true>(W1, src1, v.width >> (plane1 ? 1 : 0), hx1, hy1);
    if (ok0) rc_predict_filter(rows0, W0, (uint32_t)((uintptr_t)src0 & 3), hx0, hy0, any_diag);
    else rows0[0] = rows0[1] = rows0[2] = rows0[3] = 0x80808080u;
    if (ok1) rc_predict_filter(rows1, W1, (uint32_t)((uintptr_t)src1 & 3), hx1, hy1, any_diag);
    else rows1[0] = rows1[1] = rows1[2] = rows1[3] = 0x80808080u;
}

/* tmap / dmap point at the block's own cell of the bordered type and DC maps (row pitch bstride) */
RC_HD void rc_weighted_at(const uint8_t *tmap, const uint8_t *dmap, int bstride, int is_ipic, uint32_t rows[4])
{
    const int V = RC_LD8(dmap);
    /* neighbour DC only if (type & 0x77) == 0, else own DC; borders carry type 0xFF.
       In I pictures the left neighbour is tracked as "type 0 or 8" (h4m:1441-1454). */
    const uint32_t tT = RC_LD8(tmap - bstride), tB = RC_LD8(tmap + bstride), tL = RC_LD8(tmap - 1), tR = RC_LD8(tmap + 1);
    const int T = (tT & 0x77) ? V : RC_LD8(dmap - bstride);
    const int B = (tB & 0x77) ? V : RC_LD8(dmap + bstride);
    const int R = (tR & 0x77) ? V : RC_LD8(dmap + 1);
    const bool left_ok = is_ipic ? (tL == 0 || tL == 8) : !(tL & 0x77);
    const int L = left_ok ? RC_LD8(dmap - 1) : V;
    rc_weighted(rows, V, T, B, L, R);
}

RC_HD void rc_weighted_block(const ReconView &v, int plane, int bx, int by, uint32_t rows[4])
{
    const int bstride = ((v.width >> (plane ? 1 : 0)) >> 2) + 2;
    const uint8_t *tmap = v.blob + rc_pick3(v.off_type, plane) + (by + 1) * bstride + bx + 1;
    const uint8_t *dmap = v.blob + rc_pick3(v.off_dc, plane) + (by + 1) * bstride + bx + 1;
    const int V = RC_LD8(dmap);
    /* neighbour DC only if (type & 0x77) == 0, else own DC; borders carry type 0xFF.
       In I pictures the left neighbour is tracked as "type 0 or 8" (h4m:1441-1454). */
    const uint32_t tT = RC_LD8(tmap - bstride), tB = RC_LD8(tmap + bstride), tL = RC_LD8(tmap - 1), tR = RC_LD8(tmap + 1);
    const int T = (tT & 0x77) ? V : RC_LD8(dmap - bstride);
    const int B = (tB & 0x77) ? V : RC_LD8(dmap + bstride);
    const int R = (tR & 0x77) ? V : RC_LD8(dmap + 1);
    const bool left_ok = v.is_ipic ? (tL == 0 || tL == 8) : !(tL & 0x77);
    const int L = left_ok ? RC_LD8(dmap - 1) : V;
    rc_weighted(rows, V, T, B, L, R);
}

RC_HD void rc_mc_block(const ReconView &v, int plane, int bx, int by, uint32_t t, uint32_t rows[4])
{
    rc_mc_packed(v, plane, rc_motion_pack(v, plane, bx, by, t, rc_mv_word(v, plane, bx, by)), rows);
}

/* ---- MAP kernel: what can be reconstructed from the maps (and reference frames) alone -------
 * Returns false when the block is left to the record kernel (raw, intra AOT).  Predicted-AOT
 * blocks get their motion-compensated prediction here; the record kernel reads it back from
 * the picture and adds the AOT residual (h4m:1387-1417 use the same prediction `mdst`). */
/* mp = the block's resolved motion (rc_motion_pack), computed by the caller ahead of time; it is only
   looked at for inter blocks */
RC_HD bool rc_map_block_mp(const ReconView &v, int plane, int bx, int by, uint32_t t, uint32_t mp, uint32_t rows[4])
{
    switch (rc_classify(t, v.is_ipic))
    {
    case RC_WEIGHTED:
        rc_weighted_block(v, plane, bx, by, rows);
        return true;
    case RC_MC:
    case RC_AOT_INTER:
        rc_mc_packed(v, plane, mp, rows);
        return true;
    case RC_DIRECT:
        if ((v.is_ipic ? t : (t & 0xF)) == 8)
        {
            const int bstride = ((v.width >> (plane ? 1 : 0)) >> 2) + 2;
            const uint32_t V = RC_LD8(v.blob + rc_pick3(v.off_dc, plane) + (by + 1) * bstride + bx + 1);
            rows[0] = rows[1] = rows[2] = rows[3] = V * 0x01010101u;
            return true;
        }
        return false;
    default:
        return false;
    }
}

RC_HD bool rc_map_block(const ReconView &v, int plane, int bx, int by, uint32_t t, uint32_t rows[4])
{
    const bool inter = !v.is_ipic && (t & 0x60);
    return rc_map_block_mp(v, plane, bx, by, t, inter ? rc_motion_pack(v, plane, bx, by, t, rc_mv_word(v, plane, bx, by)) : 0u, rows);
}

/* ---- RECORD kernel: one record (header word + payload) --------------------------------------
 * rows[] must hold the block's current pixels for SYM_REC_INTER (the prediction written by the
 * map kernel); it is ignored for the other classes. */
RC_HD void rc_record_coords(uint32_t hdr, uint32_t &t, int &plane, int &bx, int &by)
{
    t = hdr & 0xFF;
    plane = (int)((hdr >> 8) & 3);
    bx = (int)((hdr >> 10) & 0x7FF);
    by = (int)(hdr >> 21);
}

/* what a record needs from outside its own words: the macroblock's vector word (predicted AOT) or the block's
   DC (intra AOT); 0 for raw records.  Separate from rc_record_block_pre so that a caller can fetch it early. */
RC_HD uint32_t rc_record_extra(const ReconView &v, int cls, uint32_t hdr)
{
    uint32_t t;
    int plane, bx, by;
    rc_record_coords(hdr, t, plane, bx, by);
    if (cls == SYM_REC_INTRA)
    {
        const int bstride = ((v.width >> (plane ? 1 : 0)) >> 2) + 2;
        return RC_LD8(v.blob + rc_pick3(v.off_dc, plane) + (by + 1) * bstride + bx + 1);
    }
    return cls == SYM_REC_INTER ? rc_mv_word(v, plane, bx, by) : 0u;
}

/* hdr = rec[0], extra = rc_record_extra(v, cls, hdr) */
template <bool kGen = false>
RC_HD void rc_record_block_pre(const ReconView &v, int cls, uint32_t len, const uint32_t *rec, uint32_t hdr, uint32_t extra, uint32_t rows[4])
{
    if (cls == SYM_REC_RAW)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r) rows[r] = rc_ld_side<kGen>(rec + 1 + r);
    }
    else if (cls == SYM_REC_INTRA)
        rc_intra_aot<kGen>(v, rows, rec + 1, (int)len - 1, (int)extra);
    else
    {
        const uint8_t *window = rc_motion_window(v, hdr & 0xFF, extra);
        if (!window) return;                          /* the map work painted it grey */
        rc_predicted_aot<kGen>(v, rows, rec + 1, (int)len - 1, RcLinearWindow{window, v.width});
    }
}

RC_HD void rc_record_block(const ReconView &v, int cls, uint32_t len, const uint32_t *rec, uint32_t rows[4])
{
    uint32_t t;
    int plane, bx, by;
    rc_record_coords(RC_LD32(rec), t, plane, bx, by);
    if (cls == SYM_REC_RAW)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r) rows[r] = RC_LD32(rec + 1 + r);
    }
    else if (cls == SYM_REC_INTRA)
    {
        const int bstride = ((v.width >> (plane ? 1 : 0)) >> 2) + 2;
        const int V = RC_LD8(v.blob + rc_pick3(v.off_dc, plane) + (by + 1) * bstride + bx + 1);
        rc_intra_aot(v, rows, rec + 1, (int)len - 1, V);
    }
    else
    {
        const uint8_t *window = rc_motion_window(v, t, rc_mv_word(v, plane, bx, by));
        if (!window) return;                          /* the map work painted it grey */
        rc_predicted_aot(v, rows, rec + 1, (int)len - 1, RcLinearWindow{window, v.width});
    }
}

#endif
