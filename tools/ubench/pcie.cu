// D2H read-back probe: 1024 frames of 460800 bytes as (a) one pitched 2-D copy out of a 3-surface
// slab, (b) one contiguous copy, (c) a gather kernel into a contiguous staging buffer + contiguous copy,
// (d) a kernel storing straight into mapped pinned host memory; each alone and next to a 109 MB H2D.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie pcie.cu && ./pcie
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void gather(const uint4 *__restrict__ src, size_t src_pitch16, uint4 *__restrict__ dst, size_t row16)
{
    const uint4 *s = src + blockIdx.y * src_pitch16;
    uint4 *d = dst + blockIdx.y * row16;
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < row16; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}

int main()
{
    const size_t n = 1024, frame = 460800, pitch = 3 * (frame + 320);
    uint8_t *d_slab, *d_stage, *d_up, *h, *h_up;
    cudaMalloc(&d_slab, n * pitch);
    cudaMalloc(&d_stage, n * frame);
    cudaMalloc(&d_up, 109 << 20);
    cudaHostAlloc(&h, n * frame, cudaHostAllocMapped);
    cudaHostAlloc(&h_up, 109 << 20, cudaHostAllocDefault);
    uint8_t *h_dev;
    cudaHostGetDevicePointer(&h_dev, h, 0);
    cudaStream_t s0, s1;
    cudaStreamCreate(&s0);
    cudaStreamCreate(&s1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int with_up = 0; with_up < 2; ++with_up)
        for (int mode = 0; mode < 4; ++mode)
        {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep)
            {
                cudaDeviceSynchronize();
                cudaEventRecord(e0, s0);
                if (with_up) cudaMemcpyAsync(d_up, h_up, 109 << 20, cudaMemcpyHostToDevice, s1);
                if (mode == 0) cudaMemcpy2DAsync(h, frame, d_slab, pitch, frame, n, cudaMemcpyDeviceToHost, s0);
                if (mode == 1) cudaMemcpyAsync(h, d_stage, n * frame, cudaMemcpyDeviceToHost, s0);
                if (mode == 2)
                {
                    gather<<<dim3(8, n), 256, 0, s0>>>((const uint4 *)d_slab, pitch / 16, (uint4 *)d_stage, frame / 16);
                    cudaMemcpyAsync(h, d_stage, n * frame, cudaMemcpyDeviceToHost, s0);
                }
                if (mode == 3) gather<<<dim3(8, n), 256, 0, s0>>>((const uint4 *)d_slab, pitch / 16, (uint4 *)h_dev, frame / 16);
                cudaEventRecord(e1, s0);
                cudaEventSynchronize(e1);
                cudaStreamSynchronize(s1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const char *names[] = {"2-D pitched copy", "contiguous copy", "gather kernel + contiguous copy", "kernel stores to mapped host memory"};
            printf("%-38s %s: %.2f ms  %.1f GB/s\n", names[mode], with_up ? "with 109 MB H2D" : "alone          ", best, n * frame / best / 1e6);
        }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
