/*
 * audio.cu -- the IMA-ADPCM audio track of .h4m files, batched over independent streams.
 *
 * Reference: decode_audio, /root/reference/h4m_audio_decode.c:185-258 (upstream keeps the call
 * disabled, h4m:2486-2507, but the function is complete), state reset per GOP block h4m:2446-2452.
 * One audio frame record payload = BE32 sample count, then -- only in the first audio frame of a
 * GOP block -- a 2-byte seed per channel (channels in DESCENDING order: high byte of the
 * predictor, then bit 7 = bit 7 of its low byte, bits 6:0 = step index, which must be <= 88),
 * then 4-bit codes, high nibble first, again channels descending inside every sample; each frame
 * starts on a fresh byte.  The recurrence is serial per stream (all channels of a stream share the
 * byte stream), so a stream is ONE thread; a batch of streams is data parallel.  The work is tiny
 * next to the video path (a few KB per frame): the kernel exists so that a file decoder built on
 * this library never has to leave the GPU runtime, not because it is a hot spot.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

#include "../../include/hvqm4.h"

namespace {

__constant__ int16_t c_ima_steps[89] = {      /* the standard IMA table (h4m:150-168) */
    7, 8, 9, 10, 11, 12, 13, 14, 16, 17, 19, 21, 23, 25, 28, 31, 34, 37, 41, 45, 50, 55, 60, 66, 73, 80, 88, 97, 107, 118, 130, 143,
    157, 173, 190, 209, 230, 253, 279, 307, 337, 371, 408, 449, 494, 544, 598, 658, 724, 796, 876, 963, 1060, 1166, 1282, 1411,
    1552, 1707, 1878, 2066, 2272, 2499, 2749, 3024, 3327, 3660, 4026, 4428, 4871, 5358, 5894, 6484, 7132, 7845, 8630, 9493, 10442,
    11487, 12635, 13899, 15289, 16818, 18500, 20350, 22385, 24623, 27086, 29794, 32767};

struct AudioJob
{
    uint32_t in_off, in_bytes;      /* payload inside the staging buffer (behind the sample count) */
    uint32_t out_off;               /* int16 units */
    uint32_t samples;
    int32_t first;
    uint32_t error;
};

__global__ void adpcm_kernel(const uint8_t *__restrict__ in, int16_t *__restrict__ out, AudioJob *jobs, HVQM4AudioState *states, int n, int channels)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    AudioJob job = jobs[i];
    HVQM4AudioState st = states[i];
    const uint8_t *p = in + job.in_off, *end = p + job.in_bytes;
    int16_t *o = out + job.out_off;
    uint32_t s = 0, err = 0;
    if (job.first && job.samples)
    {   /* h4m:190-212 */
        for (int c = channels - 1; c >= 0; --c)
        {
            if (end - p < 2) { err = HVQM4_ERR_TRUNCATED; break; }
            const uint32_t hi = p[0], b = p[1];
            p += 2;
            st.hist[c] = (int16_t)(hi << 8 | (b & 0x80));
            st.idx[c] = (int8_t)(b & 0x7F);
            if (st.idx[c] > 88) { err = HVQM4_ERR_ARGUMENT; st.idx[c] = 88; }   /* the reference exits here */
        }
        for (int c = 0; c < channels; ++c) o[c] = st.hist[c];
        s = 1;
    }
    uint32_t b = 0;
    int bitsleft = 0;
    for (; s < job.samples && !err; ++s)
    {   /* h4m:216-246 */
        for (int c = channels - 1; c >= 0; --c)
        {
            if (bitsleft == 0)
            {
                if (p == end) { err = HVQM4_ERR_TRUNCATED; break; }
                b = *p++;
                bitsleft = 8;
            }
            const int32_t step = c_ima_steps[st.idx[c]];
            int32_t delta = step >> 3;
            if (b & 0x10) delta += step >> 2;
            if (b & 0x20) delta += step >> 1;
            if (b & 0x40) delta += step;
            int32_t h = (b & 0x80) ? st.hist[c] - delta : st.hist[c] + delta;
            h = h > 32767 ? 32767 : h < -32768 ? -32768 : h;
            st.hist[c] = (int16_t)h;
            const int nib = (int)((b & 0xF0) >> 4) & 7;
            int idx = st.idx[c] + (nib < 4 ? -1 : 2 * (nib - 3));       /* IMA_IndexTable, h4m:170-176 */
            st.idx[c] = (int8_t)(idx > 88 ? 88 : idx < 0 ? 0 : idx);
            b = (b << 4) & 0xFF;
            bitsleft -= 4;
        }
        if (err) break;
        for (int c = 0; c < channels; ++c) o[s * channels + c] = st.hist[c];
    }
    states[i] = st;
    jobs[i].error = err;
    jobs[i].samples = s;     /* samples actually produced */
}

struct AudioScratch
{
    std::mutex lock;
    uint8_t *h = nullptr, *d = nullptr;
    size_t cap = 0;
    cudaStream_t stream = nullptr;
    int device = -1;
};
AudioScratch g_audio;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

extern "C" __attribute__((visibility("default")))
int HVQM4DecodeAudioBatch(int n, int channels, HVQM4AudioState *states, const int32_t *first, const uint8_t *const *frames,
                          const uint32_t *frame_bytes, int16_t *const *pcm, const uint32_t *pcm_capacity, uint32_t *samples_out)
{
    if (n <= 0 || channels < 1 || channels > HVQM4_AUDIO_MAX_CHANNELS || !states || !first || !frames || !frame_bytes || !pcm || !pcm_capacity)
        return HVQM4_ERR_ARGUMENT;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return HVQM4_ERR_NO_DEVICE;
    /* staging layout: [jobs | states | payloads | pcm] */
    const size_t jobs_bytes = align_up(sizeof(AudioJob) * (size_t)n, 256), st_bytes = align_up(sizeof(HVQM4AudioState) * (size_t)n, 256);
    size_t in_total = 0, out_total = 0;
    for (int i = 0; i < n; ++i)
    {
        if (!frames[i] || frame_bytes[i] < 4 || !pcm[i]) return HVQM4_ERR_ARGUMENT;
        const uint32_t samples = (uint32_t)frames[i][0] << 24 | (uint32_t)frames[i][1] << 16 | (uint32_t)frames[i][2] << 8 | frames[i][3];
        if (samples > pcm_capacity[i]) return HVQM4_ERR_OVERFLOW;
        in_total += align_up(frame_bytes[i] - 4, 16);
        out_total += align_up((size_t)samples * channels * 2, 16);
    }
    const size_t in_off = jobs_bytes + st_bytes, out_off = in_off + align_up(in_total, 256), total = out_off + out_total;
    if (total > 0xFFFFFFFFull) return HVQM4_ERR_OVERFLOW;
    std::lock_guard<std::mutex> guard(g_audio.lock);
    AudioScratch &a = g_audio;
    if (a.device != device || a.cap < total)
    {
        if (a.h) cudaFreeHost(a.h);
        if (a.d) cudaFree(a.d);
        if (a.stream) cudaStreamDestroy(a.stream);
        a.h = a.d = nullptr;
        a.stream = nullptr;
        a.cap = 0;
        const size_t cap = align_up(total + total / 2, 1 << 16);
        if (cudaHostAlloc((void **)&a.h, cap, cudaHostAllocDefault) != cudaSuccess) return HVQM4_ERR_NO_DEVICE;
        if (cudaMalloc((void **)&a.d, cap) != cudaSuccess || cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking) != cudaSuccess)
        {
            cudaFreeHost(a.h);
            if (a.d) cudaFree(a.d);
            a.h = a.d = nullptr;
            return HVQM4_ERR_NOMEM;
        }
        a.cap = cap;
        a.device = device;
    }
    AudioJob *jobs = reinterpret_cast<AudioJob *>(a.h);
    HVQM4AudioState *st = reinterpret_cast<HVQM4AudioState *>(a.h + jobs_bytes);
    size_t ip = 0, op = 0;
    for (int i = 0; i < n; ++i)
    {
        const uint32_t samples = (uint32_t)frames[i][0] << 24 | (uint32_t)frames[i][1] << 16 | (uint32_t)frames[i][2] << 8 | frames[i][3];
        jobs[i].in_off = (uint32_t)ip;
        jobs[i].in_bytes = frame_bytes[i] - 4;
        jobs[i].out_off = (uint32_t)(op / 2);
        jobs[i].samples = samples;
        jobs[i].first = first[i];
        jobs[i].error = 0;
        st[i] = states[i];
        if (first[i]) memset(&st[i], 0, sizeof st[i]);                    /* calloc per GOP block, h4m:2452 */
        memcpy(a.h + in_off + ip, frames[i] + 4, frame_bytes[i] - 4);
        ip += align_up(frame_bytes[i] - 4, 16);
        op += align_up((size_t)samples * channels * 2, 16);
    }
    bool ok = cudaMemcpyAsync(a.d, a.h, out_off, cudaMemcpyHostToDevice, a.stream) == cudaSuccess;
    if (ok)
    {
        adpcm_kernel<<<(n + 63) / 64, 64, 0, a.stream>>>(a.d + in_off, reinterpret_cast<int16_t *>(a.d + out_off),
                                                         reinterpret_cast<AudioJob *>(a.d), reinterpret_cast<HVQM4AudioState *>(a.d + jobs_bytes), n, channels);
        ok = cudaGetLastError() == cudaSuccess;
    }
    ok = ok && cudaMemcpyAsync(a.h, a.d, in_off, cudaMemcpyDeviceToHost, a.stream) == cudaSuccess;
    ok = ok && (out_total == 0 || cudaMemcpyAsync(a.h + out_off, a.d + out_off, out_total, cudaMemcpyDeviceToHost, a.stream) == cudaSuccess);
    ok = ok && cudaStreamSynchronize(a.stream) == cudaSuccess;
    if (!ok) return HVQM4_ERR_CUDA;
    uint32_t err = 0;
    for (int i = 0; i < n; ++i)
    {
        err |= jobs[i].error;
        states[i] = st[i];
        if (samples_out) samples_out[i] = jobs[i].samples;
        memcpy(pcm[i], a.h + out_off + (size_t)jobs[i].out_off * 2, (size_t)jobs[i].samples * channels * 2);
    }
    return (int)err;
}
