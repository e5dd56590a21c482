/* entropy_dev.h -- launch interface of the GPU-resident entropy stage (internal). */
#ifndef HVQM4_ENTROPY_DEV_H
#define HVQM4_ENTROPY_DEV_H
#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

/* one picture to parse on the GPU: raw bytes (device pointer, starting at the picture header =
   4 bytes into the frame record, like the SDK calls), its length, frame type, owning stream */
typedef struct H4DevPicture
{
    const uint8_t *data;
    uint32_t bytes;
    int32_t pic_type;
    int32_t stream;
    int32_t pad;
} H4DevPicture;

/* one picture to fetch from page-locked, mapped host memory into the step's device arena */
typedef struct H4Gather
{
    const uint8_t *src;     /* device-visible address of the picture bytes */
    uint32_t dst_off;       /* byte offset in the arena; (dst_off & 15) == (src & 15) */
    uint32_t bytes;
} H4Gather;

struct ReconJob;

#ifdef __cplusplus
extern "C" {
#endif
/* Parser slots per stream = parse kernels of consecutive steps that may be in flight together. */
#ifndef H4_PARSE_SLOTS
#define H4_PARSE_SLOTS 2
#endif

/* bytes of device memory ONE parser slot needs (0 on unsupported geometry); a stream owns
   H4_PARSE_SLOTS slots, used by consecutive steps in turn, so the arena is H4_PARSE_SLOTS * n_streams slots */
/* macroblock rows per record band of the device streams created from now on (entropy.h: h4e_set_band_rows) */
void hvqm4_dev_entropy_set_band_rows(int rows);
size_t hvqm4_dev_entropy_slot_bytes(int width, int height, uint32_t sym_cap, uint32_t work_cap);
int hvqm4_dev_entropy_init(uint8_t *arena, size_t slot_bytes, int n_streams, int width, int height, int version15,
                           uint32_t sym_cap, uint32_t work_cap, cudaStream_t stream);
/* parses n_pics pictures (one warp each) in the slots number `parity` (0 .. H4_PARSE_SLOTS - 1), bump-allocates their symbol
   buffers in blob_arena and fills jobs[i].blob / jobs[i].n_chunks (the surface pointers of jobs[]
   are filled by the host) */
int hvqm4_dev_entropy_parse(uint8_t *arena, size_t slot_bytes, const H4DevPicture *d_pics, int n_pics, int parity, uint8_t *blob_arena,
                            unsigned long long *d_blob_used, unsigned long long blob_cap, struct ReconJob *d_jobs,
                            uint32_t *d_errors, cudaStream_t stream);
/* copies n pictures out of mapped host memory (one CTA each, 16-byte words) and clears 16 bytes behind each */
int hvqm4_dev_gather(const H4Gather *d_descs, int n, uint8_t *d_base, cudaStream_t stream);
void hvqm4_dev_entropy_profile(unsigned long long out[8]);
#ifdef __cplusplus
}
#endif
#endif
