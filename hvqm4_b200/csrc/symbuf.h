/*
 * symbuf.h -- layout of the per-picture SYMBOL BUFFER, the only thing that crosses
 * from the host serial stage (entropy.c) to the CUDA reconstruction kernels
 * (recon.cu).  One picture = one contiguous, self-contained blob, so a batch of
 * pictures is uploaded with a single cudaMemcpyAsync.
 *
 * The blob carries exactly what pixel reconstruction consumes in the reference
 * (/root/reference/h4m_audio_decode.c, "h4m:N" below) and nothing that needs a bit
 * reader:
 *
 *   type/dc maps   the reference's BlockData{value,type} maps (h4m:432-436) split into
 *                  two byte planes per picture plane, each with the 1-cell border
 *                  {dc 0x7F, type 0xFF} of h4m:951-955.  type byte layout as h4m:1296-1304:
 *                  bits 6:5 macroblock type (0 intra, 1 past, 2 future), bit 4 proc,
 *                  bits 3:0 block nibble (I pictures: the whole byte is the basis count).
 *                  The MAP kernel reconstructs every block that needs nothing else from
 *                  these two maps: weighted-DC, flat-DC and motion-compensated blocks.
 *   mv table       per macroblock, absolute half-sample luma position of the prediction
 *                  (ref_x, ref_y of h4m:1954-1955) as int16 x 2; predictor chain,
 *                  wrap-around and reference switches (h4m:1846-1860,1943-1949) are
 *                  already resolved by the host.
 *   nest           the 70x38 (portrait pictures: 38x70) 4-bit nest of the last I picture (h4m:1166-1239),
 *                  packed two samples per byte (even x in the low nibble); present in every I picture
 *                  and in P/B pictures that contain intra AOT blocks.
 *   records        one variable-length record per block that carries side data, i.e. raw
 *                  blocks (h4m:543-549) and blocks with an AOT basis loop (h4m:1358-1420):
 *                    word 0     [7:0] type byte, [9:8] plane, [20:10] block x, [31:21] block y
 *                    raw        4 words = the 16 fixvl bytes, row by row
 *                    intra AOT  n basis words
 *                    inter AOT  (nibble-1) basis words + 1 pair word
 *                  basis word: bits 15:0 descriptor exactly as read from fixvl (h4m:691),
 *                              bits 23:16 scale symbol = decodeHuff(bufTree0) >> 2 (h4m:726)
 *                  pair word:  low 16 = S1 >> dc_shift, high 16 = S2 >> dc_shift (int16; h4m:1405-1406)
 *                  The host already knows every block's type before it emits side data, so it
 *                  stores the records GROUPED: by class (raw, intra AOT, inter AOT), then by
 *                  band of SYM_BAND_MCB_ROWS macroblock rows (locality), then by record
 *                  length.  Records of a group have the same length, so record i of a group is
 *                  at first + i * length -- no prefix sums on the GPU -- and every lane of a
 *                  warp of the RECORD kernel runs the same number of basis iterations.
 *   chunk table    the record kernel's work list: one entry per <= 32 records of one group:
 *                    word 0  offset of the first record (words from off_rec)
 *                    word 1  [7:0] count, [15:8] record length - 1, [23:16] class (SYM_REC_*)
 *                  Chunks [0, n_chunks_nest) are raw/intra (need the nest), the rest inter.
 *   band table     first chunk of every (class, band) pair, so that one CTA can reconstruct a
 *                  whole band of the picture -- its map blocks first, then exactly the records
 *                  that lie in it -- while the band's output sectors are still in cache.
 */
#ifndef HVQM4_SYMBUF_H
#define HVQM4_SYMBUF_H

#include <stdint.h>

#define SYM_MAGIC 0x42533448u /* "H4SB" */
#define SYM_SEG_MCBS 16       /* macroblocks per segment of the map kernel (= 32 luma blocks = one warp) */
#define SYM_BAND_MCB_ROWS 8   /* record groups are formed per band of this many macroblock rows -- the band kernel's unit: few, full
                                 chunks -- or, chosen per stream at creation (h4e_set_band_rows), of ONE row, so that a kernel can take
                                 any consecutive macroblock rows as its unit (sweep and row kernels); SymHeader.n_bands says which */
#define SYM_CHUNK 32          /* records per chunk (= one warp of the record kernel) */
#define SYM_LEN_BUCKETS 18    /* record lengths 1..17 words get their own group; longer ones one chunk each */
#define SYM_NEST_W 70
#define SYM_NEST_H 38
#define SYM_NEST_ROW_BYTES 35 /* packed nibbles */
#define SYM_NEST_BYTES (SYM_NEST_ROW_BYTES * SYM_NEST_H)

enum { SYM_PIC_I = 0x10, SYM_PIC_P = 0x20, SYM_PIC_B = 0x30 };
enum { SYM_REC_RAW = 0, SYM_REC_INTRA = 1, SYM_REC_INTER = 2, SYM_REC_CLASSES = 3 };

/* error bits raised by the host stage (the SDK entry points return void) */
enum
{
    SYM_ERR_TRUNCATED   = 1 << 0,  /* a section or tree ran past its declared size */
    SYM_ERR_BAD_TREE    = 1 << 1,  /* tree with more than 256 internal nodes */
    SYM_ERR_MCB_TYPE    = 1 << 2,  /* macroblock type 3, or type 2 inside a P picture */
    SYM_ERR_MV_RANGE    = 1 << 3,  /* prediction or nest window leaves the frame surface */
    SYM_ERR_PAIR_RANGE  = 1 << 4,  /* S1/S2 of a predicted-AOT block do not fit int16 */
    SYM_ERR_GEOMETRY    = 1 << 5,  /* unsupported size / sampling */
    SYM_ERR_OVERFLOW    = 1 << 6,  /* symbol buffer capacity exceeded */
};

typedef struct SymHeader
{
    uint32_t magic;
    uint32_t total_bytes;      /* whole blob, multiple of 16 */
    uint16_t width, height;    /* luma samples */
    uint8_t  pic_type;         /* SYM_PIC_* */
    uint8_t  version15;        /* 1: per-plane half-sample phase (h4m:1337-1343) */
    uint8_t  dc_shift;         /* P/B only (h4m:2021) */
    uint8_t  unk_shift;        /* h4m:1974, 2022 */
    uint8_t  has_nest;
    uint8_t  portrait;         /* width < height: the nest is 38 x 70 (19 packed bytes per row) and the axes of the basis
                                  descriptors swap (h4m:700-711, 743-754, 965-975) */
    uint8_t  pad0[2];
    uint32_t errors;           /* SYM_ERR_* */
    uint16_t mcb_w, mcb_h;     /* macroblocks */
    uint16_t nseg;             /* segments per macroblock row = ceil(mcb_w / 16) */
    uint16_t pad1;
    uint32_t off_type[3];      /* byte offsets from the blob start; bordered (bw+2)x(bh+2) */
    uint32_t off_dc[3];
    uint32_t off_mv;           /* int16[2] per macroblock; 0 for I pictures */
    uint32_t off_chunks;       /* uint32[2] per chunk */
    uint32_t off_nest;         /* SYM_NEST_BYTES; 0 if !has_nest */
    uint32_t off_rec;          /* records, uint32 words */
    uint32_t n_rec_words;
    uint32_t n_chunks;
    uint32_t n_chunks_nest;    /* leading chunks that are raw / intra AOT */
    uint32_t n_records;
    uint32_t off_bands;        /* uint32 [SYM_REC_CLASSES][n_bands + 1]: first chunk of (class, band); the
                                  chunks of a (class, band) pair are contiguous */
    uint32_t n_bands;
    uint32_t pad2[8];
} SymHeader;                   /* 128 bytes */

#ifdef __cplusplus
static_assert(sizeof(SymHeader) == 128, "SymHeader must be 128 bytes");
#else
_Static_assert(sizeof(SymHeader) == 128, "SymHeader must be 128 bytes");
#endif

#if defined(__CUDACC__)
#define SYM_HD __host__ __device__ __forceinline__
#else
#define SYM_HD static inline
#endif

/*
 * Record class and length (in words, header included) of a block, from its type byte alone;
 * length 0 = the block has no record (the map kernel reconstructs it completely).
 *   I picture:           type is the full byte: 0/8 -> none, 6 -> raw, n -> n bases
 *   P/B intra MCB:       nibble as above
 *   P/B inter, proc 1:   none (whole-macroblock motion compensation, h4m:1673-1683)
 *   P/B inter, proc 0:   nibble 0 -> none, 6 -> raw, k -> (k-1) bases + 1 pair
 */
SYM_HD uint32_t sym_record_len(uint32_t type, int is_ipic, int *cls)
{
    const uint32_t nib = is_ipic ? type : (type & 0xF);
    if (!is_ipic && (type & 0x60))
    {
        if ((type & 0x10) || nib == 0) return 0;
        if (nib == 6) { *cls = SYM_REC_RAW; return 5; }
        *cls = SYM_REC_INTER;
        return 1 + nib;
    }
    if (nib == 0 || nib == 8) return 0;
    if (nib == 6) { *cls = SYM_REC_RAW; return 5; }
    *cls = SYM_REC_INTRA;
    return 1 + nib;
}

SYM_HD uint32_t sym_record_header(uint32_t type, int plane, int bx, int by)
{
    return (type & 0xFF) | (uint32_t)plane << 8 | (uint32_t)bx << 10 | (uint32_t)by << 21;
}

#endif
