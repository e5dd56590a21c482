/* recon.h -- launch interface of the CUDA reconstruction kernels (internal). */
#ifndef HVQM4_RECON_H
#define HVQM4_RECON_H
#include <stdint.h>
#include <cuda_runtime.h>

/* chunks (of <= 32 records) one CTA of the record kernel works through */
#ifndef HVQM4_REC_CHUNKS_PER_CTA
#define HVQM4_REC_CHUNKS_PER_CTA 32
#endif

/* One picture to reconstruct.  All pointers are device pointers; surfaces are planar
   Y|U|V, contiguous, stride = plane width (the reference's frame layout, h4m:2343-2349). */
typedef struct ReconJob
{
    const uint8_t *blob;     /* symbol buffer, 16-byte aligned (symbuf.h) */
    uint8_t *present;
    const uint8_t *past;
    const uint8_t *future;
    uint32_t rec_cta_begin;  /* first CTA of the record kernel that belongs to this picture (exclusive prefix) */
    uint32_t n_chunks;       /* chunk-table entries of the picture */
    uint32_t pad[2];
} ReconJob;

#ifdef __cplusplus
extern "C" {
#endif
/* CTAs of the record kernel a picture with n_chunks chunks needs */
static inline uint32_t hvqm4_rec_ctas(uint32_t n_chunks)
{
    return (n_chunks + HVQM4_REC_CHUNKS_PER_CTA - 1) / HVQM4_REC_CHUNKS_PER_CTA;
}
/* Reconstructs n_jobs pictures of identical geometry: per sub-batch of pictures one launch of
   the map kernel, then one of the record kernel.  h_rec_prefix (host memory, n_jobs + 1 entries)
   is the exclusive prefix of hvqm4_rec_ctas(n_chunks) over the pictures, the same values the
   jobs carry in rec_cta_begin.  Returns a cudaError_t.  *launches is incremented by the number
   of kernels launched. */
/* 0 = choose by batch size (default), > 0 = always the fused band kernel, < 0 = always the
   map + record kernel pair */
void hvqm4_recon_set_mode(int band_mode);
long long hvqm4_recon_band_launches(void);
int hvqm4_recon_band_rows(void);

/* rgb.cu: planar Y|U|V 4:2:0 surfaces -> interleaved RGB (the reference's dumpRGB, h4m:895-926) */
int hvqm4_rgb_launch(const uint8_t *const *d_frames, int n, uint8_t *d_dst, size_t dst_stride, int width, int height, cudaStream_t stream);
/* one launch of the fused band kernel; jobs whose blob is NULL are skipped */
/* slab: base of the registered surface slab (hvqm4_row_register_slab) the pictures' surfaces lie in, or NULL: the row
   kernel reads reference patches through tensor maps of that slab and is not used without one */
int hvqm4_recon_launch_band(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const void *slab, int band_rows, cudaStream_t stream);
/* band_rows: macroblock rows per record band the pictures' streams were created with (hvqm4_recon_band_rows at that time) */
int hvqm4_recon_launch(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, const uint32_t *h_rec_prefix, const void *slab, int band_rows,
                       cudaStream_t stream, int *launches);
int hvqm4_row_register_slab(const void *base, size_t stride, int count, int width, int height);
void hvqm4_row_unregister_slab(const void *base);
#ifdef __cplusplus
}
#endif
#endif
