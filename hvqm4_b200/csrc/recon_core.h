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
 * words, one per row, leftmost pixel in the low byte, ready for a 4-byte store.
 */
#ifndef HVQM4_RECON_CORE_H
#define HVQM4_RECON_CORE_H

#include <stdint.h>
#include "symbuf.h"

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ __forceinline__
#else
#define RC_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define RC_LD8(p) __ldg((const uint8_t *)(p))
#define RC_LD32(p) __ldg((const uint32_t *)(p))
#else
#define RC_LD8(p) (*(const uint8_t *)(p))
#define RC_LD32(p) (*(const uint32_t *)(p))
#endif

/* Read-only view of one picture job. */
struct ReconView
{
    const uint8_t *blob;       /* symbol buffer (symbuf.h) */
    const uint8_t *nest;       /* packed nest: shared memory on the GPU, blob memory on the CPU */
    const int32_t *div_tab;    /* 16 entries,  h4m:262,270 */
    const int32_t *mcdiv_tab;  /* 512 entries, h4m:263,272 */
    const uint8_t *ref[2];     /* past, future frame surfaces (Y|U|V contiguous) */
    int width, height;
    int is_ipic, version15;
    int unk_shift;
    uint32_t off_type[3], off_dc[3], off_mv;
    int mcb_w;
};

RC_HD void rc_make_view(ReconView &v, const uint8_t *blob, const SymHeader &h, const uint8_t *nest,
                        const int32_t *div_tab, const int32_t *mcdiv_tab, const uint8_t *past, const uint8_t *future)
{
    v.blob = blob; v.nest = nest; v.div_tab = div_tab; v.mcdiv_tab = mcdiv_tab;
    v.ref[0] = past; v.ref[1] = future;
    v.width = h.width; v.height = h.height;
    v.is_ipic = h.pic_type == SYM_PIC_I; v.version15 = h.version15;
    v.unk_shift = h.unk_shift;
    for (int p = 0; p < 3; ++p) { v.off_type[p] = h.off_type[p]; v.off_dc[p] = h.off_dc[p]; }
    v.off_mv = h.off_mv;
    v.mcb_w = h.mcb_w;
}

RC_HD uint32_t rc_clamp255(int32_t x) { return x < 0 ? 0u : x > 255 ? 255u : (uint32_t)x; }

/* byte-wise (a + b + 1) >> 1 on four packed bytes: (a|b) - (((a^b) & 0xFE..) >> 1) */
RC_HD uint32_t rc_avg4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) & 0xFEFEFEFEu) >> 1); }

/* byte-wise (a + b + c + d + 2) >> 2 on four packed bytes, via two 16-bit lanes */
RC_HD uint32_t rc_avg4x4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    const uint32_t M = 0x00FF00FFu;
    uint32_t lo = (a & M) + (b & M) + (c & M) + (d & M) + 0x00020002u;
    uint32_t hi = ((a >> 8) & M) + ((b >> 8) & M) + ((c >> 8) & M) + ((d >> 8) & M) + 0x00020002u;
    return ((lo >> 2) & M) | (((hi >> 2) & M) << 8);
}

/* ---- weighted DC fill (h4m:293-383) ------------------------------------------------
 * out(r,c) = sat_mean8(8V + rowterm[r] + colterm[c]) with
 *   rowterm = {2T-B-V, V-B, V-T, 2B-T-V},  colterm = {2L-R-V, V-R, V-L, 2R-L-V}
 * which is the sixteen expressions of the reference regrouped.  sat_mean8 divides
 * (sum + 4) by 8 as an UNSIGNED number and then clamps: sums <= -5 -> 255, -4..-1 -> 0. */
RC_HD uint32_t rc_sat_mean8(int32_t sum)
{
    uint32_t q = ((uint32_t)sum + 4u) >> 3;
    return q > 255u ? 255u : q;
}

RC_HD void rc_weighted(uint32_t rows[4], int V, int T, int B, int L, int R)
{
    const int rt[4] = {2 * T - B - V, V - B, V - T, 2 * B - T - V};
    const int ct[4] = {2 * L - R - V, V - R, V - L, 2 * R - L - V};
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int base = 8 * V + rt[r];
        rows[r] = rc_sat_mean8(base + ct[0]) | rc_sat_mean8(base + ct[1]) << 8 |
                  rc_sat_mean8(base + ct[2]) << 16 | rc_sat_mean8(base + ct[3]) << 24;
    }
}

/* ---- AOT bases (h4m:679-817) --------------------------------------------------------
 * word: bits 15:0 descriptor ([5:0] x, [10:6] y, [11] x step 2, [12] y step 2,
 * [14:13] scale offset, [15] negate), bits 23:16 scale symbol (>> 2). */

/* sample of the packed I-picture nest */
RC_HD int rc_nest_at(const uint8_t *nest, int x, int y)
{
    return (nest[y * SYM_NEST_ROW_BYTES + (x >> 1)] >> ((x & 1) * 4)) & 0xF;
}

template <bool kWindow>
RC_HD void rc_add_basis(const ReconView &v, uint32_t word, const uint8_t *src, int src_stride,
                        int32_t &scale_sum, int32_t acc[16])
{
    const int ox = word & 0x3F, oy = (word >> 6) & 0x1F;
    const int xs = 1 + ((word >> 11) & 1), ys = 1 + ((word >> 12) & 1);
    int b[16];
    int lo = 15, hi = 0;
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x)
        {
            int s;
            if (kWindow) s = (RC_LD8(src + (oy + y * ys) * src_stride + ox + x * xs) >> 4) & 0xF;   /* h4m:756-761 */
            else s = rc_nest_at(src, ox + x * xs, oy + y * ys);
            b[y * 4 + x] = s;
            lo = s < lo ? s : lo;
            hi = s > hi ? s : hi;
        }
    scale_sum += (int32_t)((word >> 16) & 0xFF) << 2;            /* cumulative within the block, h4m:726,781 */
    int32_t inv = v.div_tab[hi - lo];
    if (word & 0x8000) inv = -inv;
    const uint32_t factor = (uint32_t)(scale_sum + (int32_t)((word >> 13) & 3)) * (uint32_t)inv;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (int32_t)((uint32_t)acc[i] + factor * (uint32_t)b[i]);   /* mod 2^32 */
}

template <bool kWindow>
RC_HD int32_t rc_aot_sum(const ReconView &v, const uint32_t *side, int n, const uint8_t *src, int src_stride, int32_t acc[16])
{
    int32_t scale_sum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    for (int k = 0; k < n; ++k) rc_add_basis<kWindow>(v, RC_LD32(side + k), src, src_stride, scale_sum, acc);
    uint32_t total = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) total += (uint32_t)acc[i];
    return (int32_t)total >> 4;                                   /* arithmetic, h4m:793,815 */
}

/* ---- half-sample prediction (h4m:1242-1294) -------------------------------------------
 * Fetches the 4x4 prediction whose top-left integer sample is `src` with phase (hx,hy).
 * Rows are read as two aligned 32-bit words (covers the 5 bytes a row can need). */
RC_HD void rc_predict(uint32_t rows[4], const uint8_t *src, int stride, int hx, int hy)
{
    const uint32_t a = (uint32_t)((uintptr_t)src & 3);
    const uint8_t *base = src - a;
    uint32_t A[5], Bx[5];
    const int nrow = 4 + hy;
#pragma unroll
    for (int r = 0; r < 5; ++r)
    {
        if (r < nrow)
        {
            const uint32_t w0 = RC_LD32(base + r * stride), w1 = RC_LD32(base + r * stride + 4);
            const uint64_t w = ((uint64_t)w1 << 32 | w0) >> (8 * a);
            A[r] = (uint32_t)w;
            Bx[r] = (uint32_t)(w >> 8);
        }
        else
            A[r] = Bx[r] = 0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        if (!hx && !hy) rows[r] = A[r];
        else if (hx && !hy) rows[r] = rc_avg4(A[r], Bx[r]);
        else if (!hx && hy) rows[r] = rc_avg4(A[r], A[r + 1]);
        else rows[r] = rc_avg4x4(A[r], Bx[r], A[r + 1], Bx[r + 1]);
    }
}

/* ---- predicted AOT block (h4m:1379-1420) ---------------------------------------------- */
RC_HD void rc_predicted_aot(const ReconView &v, uint32_t rows[4], const uint32_t *side, int nibble,
                            const uint8_t *window, int window_stride)
{
    int32_t acc[16];
    const uint32_t aot_mean = (uint32_t)rc_aot_sum<true>(v, side, nibble - 1, window, window_stride, acc);
    const uint32_t pair = RC_LD32(side + nibble - 1);
    int m[16];
    int32_t mean = 8;
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        m[i] = (rows[i >> 2] >> ((i & 3) * 8)) & 0xFF;
        mean += m[i];
    }
    mean >>= 4;                                                   /* non-negative: same as /16 */
    int32_t lo = m[0] - mean, hi = lo;
#pragma unroll
    for (int i = 1; i < 16; ++i)
    {
        const int32_t d = m[i] - mean;
        lo = d < lo ? d : lo;
        hi = d > hi ? d : hi;
    }
    const int32_t s1 = (int32_t)(int16_t)(pair & 0xFFFF), s2 = (int32_t)(int16_t)(pair >> 16);
    const uint32_t addend = ((uint32_t)s1 << v.unk_shift) - aot_mean;
    const uint32_t factor = (uint32_t)s2 * (uint32_t)v.mcdiv_tab[hi - lo];
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        uint32_t out = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
        {
            const int i = r * 4 + c;
            const int32_t res = (int32_t)((uint32_t)acc[i] + addend + (uint32_t)(m[i] - mean) * factor);
            out |= rc_clamp255((res >> v.unk_shift) + m[i]) << (8 * c);
        }
        rows[r] = out;
    }
}

/* ---- one block, all cases --------------------------------------------------------------
 * plane/bx/by: block position; t: its type byte; side: its words in the side array. */
RC_HD void rc_block(const ReconView &v, int plane, int bx, int by, uint32_t t, const uint32_t *side, uint32_t rows[4])
{
    const int sh = plane ? 1 : 0;
    const int pw = v.width >> sh;
    const int bstride = (pw >> 2) + 2;
    const bool inter = !v.is_ipic && (t & 0x60);
    const uint32_t nib = v.is_ipic ? t : (t & 0xF);

    if (!inter)
    {
        const uint8_t *tmap = v.blob + v.off_type[plane] + (by + 1) * bstride + bx + 1;
        const uint8_t *dmap = v.blob + v.off_dc[plane] + (by + 1) * bstride + bx + 1;
        const int V = RC_LD8(dmap);
        if (nib == 0)
        {
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
        else if (nib == 8)
        {
            rows[0] = rows[1] = rows[2] = rows[3] = (uint32_t)V * 0x01010101u;
        }
        else if (nib == 6)
        {
#pragma unroll
            for (int r = 0; r < 4; ++r) rows[r] = RC_LD32(side + r);
        }
        else
        {   /* IntraAotBlock, h4m:1358-1377 */
            int32_t acc[16];
            const int32_t mean = rc_aot_sum<false>(v, side, (int)nib, v.nest, 0, acc);
            const int32_t delta = (int32_t)((uint32_t)V << v.unk_shift) - mean;
#pragma unroll
            for (int r = 0; r < 4; ++r)
            {
                uint32_t out = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) out |= rc_clamp255((acc[r * 4 + c] + delta) >> v.unk_shift) << (8 * c);
                rows[r] = out;
            }
        }
        return;
    }

    /* inter macroblock */
    if (nib == 6 && !(t & 0x10))
    {
#pragma unroll
        for (int r = 0; r < 4; ++r) rows[r] = RC_LD32(side + r);
        return;
    }
    const int mx = bx >> (1 - sh), my = by >> (1 - sh);
    const uint32_t mvw = RC_LD32(v.blob + v.off_mv + 4 * (my * v.mcb_w + mx));
    const int rx = (int16_t)(mvw & 0xFFFF), ry = (int16_t)(mvw >> 16);
    if (rx == -32768)
    {   /* poisoned by the host stage (SYM_ERR_MV_RANGE): never dereference */
        rows[0] = rows[1] = rows[2] = rows[3] = 0x80808080u;
        return;
    }
    const uint8_t *ref = v.ref[((t >> 5) & 3) - 1];
    const int px = rx >> sh, py = ry >> sh;
    int hx = rx & 1, hy = ry & 1;                    /* 1.3: luma phase for every plane (h4m:1329-1330,1869-1870) */
    if (v.version15) { hx = px & 1; hy = py & 1; }   /* 1.5: per-plane phase (h4m:1337-1343,1890-1896) */
    const int plane_off = plane == 0 ? 0 : plane == 1 ? v.width * v.height : v.width * v.height + (v.width >> 1) * (v.height >> 1);
    /* linear addressing, no clamping (h4m:1344,1897); sub-block offset = pb_offset (h4m:866-869) */
    const int subx = plane == 0 ? (bx & 1) * 4 : 0, suby = plane == 0 ? (by & 1) * 4 : 0;
    const uint8_t *src = ref + plane_off + ((py >> 1) + suby) * pw + (px >> 1) + subx;
    rc_predict(rows, src, pw, hx, hy);
    if ((t & 0x10) || nib == 0) return;
    /* 70x38 window of the reference luma, origin (rx/2 - 32, ry/2 - 16) (h4m:1864-1868) */
    const uint8_t *window = ref + rx / 2 + (ry / 2 - 16) * v.width - 32;
    rc_predicted_aot(v, rows, side, (int)nib, window, v.width);
}

#endif
