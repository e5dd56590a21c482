/*
 * entropy_dev.cu -- the serial bitstream stage, run on the GPU.
 *
 * The host stage (entropy.c) is the end-to-end bottleneck: one B200 reconstructs a dense picture
 * in ~1 us, one host core parses it in ~0.6 ms.  A picture's parse is serial, but a batch has a
 * thousand independent pictures, so the same C code is compiled here as __device__ functions
 * (entropy.c is #include'd with H4E_DEVICE; see the macro block at its top) and run with ONE
 * PICTURE PER WARP against per-stream state that lives in device memory.  The entry points are
 * warp-collective: lane 0 runs what is inherently serial (section table, trees, run decode), the
 * rest uses all lanes -- one symbol section per lane in SIMT lock step, block types and DC
 * values through prefix sums, record scheduling by rows, lane-strided fill and copies (DESIGN.md
 * section 4c has the table).  The symbol buffer is written straight into a device arena that the
 * reconstruction kernel reads next; nothing but the raw picture bytes crosses PCIe on the way in.
 * A warp is latency bound (~8 ms per dense picture) but a step keeps a thousand of them in
 * flight.
 *
 * Because it is the same source, parity with the host stage (and through it with the reference)
 * is structural; tests/test_gpu_parity.py still checks the decoded frames in this mode.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#define H4E_DEVICE 1
#include "entropy.c"

#include "entropy_dev.h"
#include "recon.h"

namespace {

__device__ __forceinline__ H4Seq *slot_seq(uint8_t *arena, size_t slot_bytes, int stream)
{
    return reinterpret_cast<H4Seq *>(arena + (size_t)stream * slot_bytes);
}

__global__ void dev_measure_kernel(int width, int height, uint32_t sym_cap, uint32_t work_cap, unsigned long long *out)
{
    if (threadIdx.x || blockIdx.x) return;
    H4Seq tmp;
    memset(&tmp, 0, sizeof tmp);
    if (!seq_geometry_ok(width, height, 2, 2))
    {
        *out = 0;
        return;
    }
    seq_set_dims(&tmp, width, height, 1);
    *out = (unsigned long long)(align16(sizeof(H4Seq)) + seq_carve(&tmp, nullptr, sym_cap, work_cap));
}

__global__ void dev_init_kernel(uint8_t *arena, size_t slot_bytes, int n_streams, int width, int height, int version15,
                                uint32_t sym_cap, uint32_t work_cap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_streams) return;
    H4Seq *s = slot_seq(arena, slot_bytes, i);
    memset(s, 0, sizeof(H4Seq));
    seq_set_dims(s, width, height, version15);
    seq_carve(s, reinterpret_cast<uint8_t *>(s) + align16(sizeof(H4Seq)), sym_cap, work_cap);
    seq_init_maps(s);
}

/* one warp per picture: the entry points of entropy.c are warp-collective in the device build.
   Every stream owns H4_PARSE_SLOTS parser slots (scratch + state) used by consecutive steps in
   turn (`parity`), so that the parse kernels of consecutive steps can be in flight together; the
   only state a picture inherits is the nest of the last I picture, which an I picture therefore
   also writes into the sibling slots (the runtime keeps I-picture steps exclusive, api.cpp). */
__global__ void __launch_bounds__(32)
dev_parse_kernel(uint8_t *arena, size_t slot_bytes, const H4DevPicture *pics, int n_pics, int parity, uint8_t *blob_arena,
                 unsigned long long *blob_used, unsigned long long blob_cap, ReconJob *jobs, uint32_t *errors)
{
    const int i = blockIdx.x;
    if (i >= n_pics) return;
    const int lane = threadIdx.x;
    const H4DevPicture pic = pics[i];
    if (pic.pic_type == 0) return;     /* a picture of the host's share (api.cpp): its job is already complete */
    H4Seq *s = slot_seq(arena, slot_bytes, pic.stream * H4_PARSE_SLOTS + parity);
    const size_t bytes = h4e_parse_begin(s, pic.pic_type, pic.data, pic.bytes);
    uint32_t err = 0;
    uint8_t *blob = nullptr;
    if (bytes)
    {
        const unsigned long long need = (bytes + 127) & ~127ull;
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(blob_used, need);
        at = __shfl_sync(0xFFFFFFFFu, at, 0);
        if (at + need <= blob_cap)
        {
            blob = blob_arena + at;
            err = h4e_parse_finish(s, blob);
            if (pic.pic_type == SYM_PIC_I)
            {
                __syncwarp();
                for (int j = 0; j < H4_PARSE_SLOTS; ++j)
                {
                    if (j == parity) continue;
                    H4Seq *sibling = slot_seq(arena, slot_bytes, pic.stream * H4_PARSE_SLOTS + j);
                    for (int k = lane; k < SYM_NEST_BYTES; k += 32) sibling->nest[k] = s->nest[k];
                }
            }
        }
        else
            err = s->err | SYM_ERR_OVERFLOW;
    }
    else
        err = s->err | SYM_ERR_GEOMETRY;
    if (lane == 0)
    {
        jobs[i].blob = blob;
        jobs[i].n_chunks = blob ? s->n_chunks : 0;
        if (err) atomicOr(errors, err);
    }
}

/* The bitstreams of a step fetched by the GPU itself from page-locked, mapped host memory
   (HVQM4HostRegister): no host thread touches the picture bytes, only a 16-byte descriptor per
   picture is uploaded.  Source and destination share their alignment modulo 16, so the body is
   16-byte words; every thread keeps several PCIe reads in flight. */
__global__ void __launch_bounds__(128)
dev_gather_kernel(const H4Gather *__restrict__ descs, uint8_t *__restrict__ base)
{
    const H4Gather d = descs[blockIdx.x];
    if (d.bytes == 0) return;          /* nothing to fetch (a picture parsed by the host) */
    const uint8_t *src = d.src;
    uint8_t *dst = base + d.dst_off;
    const uint32_t head = min((16u - (uint32_t)((uintptr_t)src & 15u)) & 15u, d.bytes);
    const uint32_t n16 = (d.bytes - head) >> 4, tail0 = head + (n16 << 4);
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    const uint4 *s16 = reinterpret_cast<const uint4 *>(src + head);
    uint4 *d16 = reinterpret_cast<uint4 *>(dst + head);
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < n16; i += 128) d16[i] = s16[i];
    /* tail bytes, then 16 bytes of zeros (the bit reader's slack, h4m:2080-2082) */
    if (threadIdx.x < d.bytes - tail0 + 16)
    {
        const uint32_t o = tail0 + threadIdx.x;
        dst[o] = o < d.bytes ? src[o] : (uint8_t)0;
    }
}

}  // namespace

extern "C" int hvqm4_dev_gather(const H4Gather *d_descs, int n, uint8_t *d_base, cudaStream_t stream)
{
    if (n <= 0) return 0;
    dev_gather_kernel<<<n, 128, 0, stream>>>(d_descs, d_base);
    return (int)cudaGetLastError();
}

/* per-phase cycle totals of the GPU parser since the last call (diagnostics): hdr+trees, pass 1,
   plan, pass 2 scheduling, map copies, flat decode, record fill */
extern "C" void hvqm4_dev_entropy_profile(unsigned long long out[8])
{
    cudaMemcpyFromSymbol(out, h4e_dev_prof, sizeof(unsigned long long) * 8);
    unsigned long long zero[8] = {0};
    cudaMemcpyToSymbol(h4e_dev_prof, zero, sizeof zero);
}

extern "C" size_t hvqm4_dev_entropy_slot_bytes(int width, int height, uint32_t sym_cap, uint32_t work_cap)
{
    unsigned long long *d = nullptr, h = 0;
    if (cudaMalloc((void **)&d, sizeof h) != cudaSuccess) return 0;
    dev_measure_kernel<<<1, 1>>>(width, height, sym_cap, work_cap, d);
    const cudaError_t e = cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return 0;
    return (size_t)((h + 255) & ~255ull);
}

extern "C" void hvqm4_dev_entropy_set_band_rows(int rows)
{
    const int shift = rows <= 1 ? 0 : rows <= 4 ? 2 : 3;
    cudaMemcpyToSymbol(g_h4e_band_shift, &shift, sizeof shift);
}

extern "C" int hvqm4_dev_entropy_init(uint8_t *arena, size_t slot_bytes, int n_streams, int width, int height, int version15,
                                      uint32_t sym_cap, uint32_t work_cap, cudaStream_t stream)
{
    const int n_slots = H4_PARSE_SLOTS * n_streams;   /* see dev_parse_kernel */
    dev_init_kernel<<<(n_slots + 63) / 64, 64, 0, stream>>>(arena, slot_bytes, n_slots, width, height, version15, sym_cap, work_cap);
    return (int)cudaGetLastError();
}

extern "C" int hvqm4_dev_entropy_parse(uint8_t *arena, size_t slot_bytes, const H4DevPicture *d_pics, int n_pics, int parity, uint8_t *blob_arena,
                                       unsigned long long *d_blob_used, unsigned long long blob_cap, ReconJob *d_jobs,
                                       uint32_t *d_errors, cudaStream_t stream)
{
    if (n_pics <= 0) return 0;
    dev_parse_kernel<<<n_pics, 32, 0, stream>>>(arena, slot_bytes, d_pics, n_pics, parity, blob_arena, d_blob_used, blob_cap, d_jobs, d_errors);
    return (int)cudaGetLastError();
}
