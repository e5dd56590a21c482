/* recon.h -- launch interface of the CUDA reconstruction kernels (internal). */
#ifndef HVQM4_RECON_H
#define HVQM4_RECON_H
#include <stdint.h>
#include <cuda_runtime.h>

/* One picture to reconstruct.  All pointers are device pointers; surfaces are planar
   Y|U|V, contiguous, stride = plane width (the reference's frame layout, h4m:2343-2349). */
typedef struct ReconJob
{
    const uint8_t *blob;     /* symbol buffer, 16-byte aligned (symbuf.h) */
    uint8_t *present;
    const uint8_t *past;
    const uint8_t *future;
} ReconJob;

#ifdef __cplusplus
extern "C" {
#endif
/* Reconstructs n_jobs pictures of identical geometry in one launch.  Returns a cudaError_t. */
int hvqm4_recon_launch(const ReconJob *d_jobs, int n_jobs, int mcb_w, int mcb_h, cudaStream_t stream);
#ifdef __cplusplus
}
#endif
#endif
