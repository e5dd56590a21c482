/*
 * entropy.h -- host serial stage of the HVQM4 picture decoder (internal interface).
 *
 * Runs everything in the reference that needs a bit reader -- section setup, Huffman
 * trees, block-type / DC / macroblock maps, motion vectors and the symbol halves of
 * the AOT, predicted-AOT and raw blocks (/root/reference/h4m_audio_decode.c:552-677,
 * 1043-1164, 1551-1776, 1846-1860 and the read16/decodeHuff/decodeSOvfSym calls at
 * 691, 726, 738, 767, 1405-1406, 543-549) -- and emits the symbol buffer of symbuf.h.
 * It never touches a pixel; reconstruction is CUDA only (recon.cu).
 */
#ifndef HVQM4_ENTROPY_H
#define HVQM4_ENTROPY_H

#include <stddef.h>
#include <stdint.h>
#include "symbuf.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct H4Seq H4Seq;

/* Per-stream persistent state: bordered type/DC maps, last I-picture nest, the six
   Huffman leaf tables (they persist across pictures in the reference, h4m:474). */
H4Seq *h4e_seq_create(int width, int height, int h_samp, int v_samp, int version15);
void h4e_seq_destroy(H4Seq *s);
/* macroblock rows per record band (symbuf.h) of the streams created from now on: 8 (default) or 1; process-wide */
void h4e_set_band_rows(int rows);
void h4e_seq_set_version(H4Seq *s, int version15);
uint32_t h4e_seq_errors(const H4Seq *s);   /* OR of SYM_ERR_* since creation */

/*
 * Two-step parse so the caller can allocate the blob exactly (pinned arenas):
 *   h4e_parse_begin   header, sections, trees, type/DC maps (I) or pass 1 (P/B),
 *                     work-order offsets; returns the blob size in bytes (0 on a
 *                     fatal geometry error).
 *   h4e_parse_finish  pass 2: motion vectors and per-block side words, written
 *                     straight into `blob` (blob_bytes from parse_begin).
 * `pic` points at the picture header, i.e. 4 bytes into the frame record, exactly
 * what the SDK entry points receive (h4m:2100); pic_len is the number of readable
 * bytes from there.  Returns the SYM_ERR_* bits of this picture.
 */
size_t h4e_parse_begin(H4Seq *s, int pic_type, const uint8_t *pic, size_t pic_len);
uint32_t h4e_parse_finish(H4Seq *s, uint8_t *blob);

/* inter-coded macroblocks of the last parsed picture (for bandwidth accounting) */
/* test hook: use the split pass 2 of the GPU build (serial vector chain + row-wise scheduling) on the host */
void h4e_seq_set_split_schedule(H4Seq *s, int on);
uint32_t h4e_last_inter_mcbs(const H4Seq *s);
/* record-kernel chunks of the picture planned by the last h4e_parse_begin */
uint32_t h4e_last_chunks(const H4Seq *s);

/* geometry helpers */
size_t h4e_frame_bytes(const H4Seq *s);    /* planar Y|U|V bytes = W*H*3/2 */
void h4e_seq_dims(const H4Seq *s, int out[6]); /* width,height,mcb_w,mcb_h,nseg,version15 */

#ifdef __cplusplus
}
#endif
#endif
