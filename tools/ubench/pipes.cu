// Instruction-throughput probe for the integer ops the reconstruction kernels are made of (sm_100a).
// Prints warp-instructions per cycle per SM sub-partition for each op (8 independent chains per thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CHAINS 8
#define ITERS 512

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    if (OP == 0) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 1) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 2) asm volatile("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(c));
    else if (OP == 3) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 4) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 5) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 6) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 7) asm volatile("max.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 8) { float f; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); r = __float_as_uint(f); }
    else if (OP == 9) asm volatile("bfe.u32 %0, %1, 12, 4;" : "=r"(r) : "r"(a));
    else if (OP == 10) { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(a)); r = __float_as_uint(f); }
    else if (OP == 11) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 12) asm volatile("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 13) asm volatile("{.reg .b32 t; add.u16x2 t, %1, %2; min.u16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 14) asm volatile("shl.b32 %0, %1, 3;" : "=r"(r) : "r"(a));
    else if (OP == 15) asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(a));
    else if (OP == 16) asm volatile("{.reg .pred p; setp.ne.u32 p, %3, 0; selp.b32 %0, %1, %2, p;}" : "=r"(r) : "r"(a), "r"(b), "r"(c & 1));
    else r = a;
    return r;
}

template <int OP, int OP2 = OP>
__global__ void __launch_bounds__(256) probe(uint32_t *out, const uint32_t *in, long long *cycles)
{
    uint32_t x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = in[threadIdx.x + i];
    const uint32_t b = in[40], c = in[41];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it)
    {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) x[i] = (i & 1) ? op<OP2>(x[i], b, c) : op<OP>(x[i], b, c);
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP, int OP2 = OP>
void run(const char *name, uint32_t *out, uint32_t *in, long long *cyc)
{
    const int blocks = 148 * 2;    // 2 CTAs of 8 warps per SM = 4 warps per sub-partition
    probe<OP, OP2><<<blocks, 256>>>(out, in, cyc);
    probe<OP, OP2><<<blocks, 256>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h[148 * 2];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    // per sub-partition: 4 warps x ITERS x 4 x CHAINS instructions in `avg` cycles
    printf("%-28s %.3f warp-inst/cycle/SMSP\n", name, 4.0 * ITERS * 4 * CHAINS / avg);
}

int main()
{
    uint32_t *out, *in;
    long long *cyc;
    cudaMalloc(&out, 148 * 2 * 256 * 4);
    cudaMalloc(&in, 4096);
    cudaMalloc(&cyc, 148 * 2 * 8);
    uint32_t h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0x01234567u * (i + 1) | 1;
    h[41] = 0x5410;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    run<0>("prmt", out, in, cyc);
    run<1>("lop3", out, in, cyc);
    run<2>("shr", out, in, cyc);
    run<3>("mad.lo (IMAD)", out, in, cyc);
    run<4>("mul.hi (IMAD.HI)", out, in, cyc);
    run<5>("dp4a (IDP.4A)", out, in, cyc);
    run<6>("add", out, in, cyc);
    run<7>("max.u32 (VIMNMX)", out, in, cyc);
    run<8>("fma.f32 (FFMA)", out, in, cyc);
    run<9>("bfe.u32", out, in, cyc);
    run<10>("cvt.f32.u32 (I2F)", out, in, cyc);
    run<11>("shf.r.wrap (funnel)", out, in, cyc);
    run<12>("add.u16x2 (VIADD.16x2)", out, in, cyc);
    run<13>("add+min.u16x2 (VIADDMNMX)", out, in, cyc);
    run<14>("shl imm", out, in, cyc);
    run<15>("popc", out, in, cyc);
    run<16>("selp", out, in, cyc);
    /* pairs, interleaved 1:1: which ops share a pipe? */
    run<1, 3>("lop3 + IMAD", out, in, cyc);
    run<3, 5>("IMAD + IDP.4A", out, in, cyc);
    run<1, 5>("lop3 + IDP.4A", out, in, cyc);
    run<1, 6>("lop3 + add", out, in, cyc);
    run<3, 6>("IMAD + add", out, in, cyc);
    run<0, 2>("prmt + shr", out, in, cyc);
    run<6, 7>("add + max", out, in, cyc);
    run<3, 8>("IMAD + FFMA", out, in, cyc);
    run<1, 8>("lop3 + FFMA", out, in, cyc);
    run<1, 10>("lop3 + I2F", out, in, cyc);
    run<3, 10>("IMAD + I2F", out, in, cyc);
    run<0, 12>("prmt + VIADD.16x2", out, in, cyc);
    run<3, 12>("IMAD + VIADD.16x2", out, in, cyc);
    return 0;
}
