// Microbenchmark (not product code): issue cost of packed FFMA2 vs scalar FFMA on sm_100a, alone and mixed with
// ALU-pipe integer work.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void k(float *out, int iters, float s) {
    float a[16];
    unsigned long long p[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = pk(a[2 * i], a[2 * i + 1]); u[i] = threadIdx.x + i; }
    const unsigned long long ps = pk(s, s);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {  // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.5f);
        }
        if (MODE == 1 || MODE == 3) {  // 8 FFMA2 (= 16 fma)
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], ps, ps);
        }
        if (MODE == 2 || MODE == 3 || MODE == 4) {  // 8 ALU ops (LOP3)
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = (u[i] ^ (u[(i + 1) & 7] >> 3)) & 0x7fffffffu;
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]) + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
void run(const char *name, float *out) {
    const int iters = 20000, blocks = 148 * 8, threads = 256;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(out, 100, 1.0001f);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double warp_iters = (double)blocks * threads / 32 * iters;
    printf("%-28s %8.3f ms  %7.2f cycles/iter/SMSP-warp-slot (at 1.9 GHz, 148 SMs x 4 SMSP)\n", name, ms,
           ms * 1e-3 * 1.9e9 * 148 * 4 / warp_iters);
}
int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("16 FFMA", out);
    run<1>("8 FFMA2", out);
    run<4>("8 LOP3", out);
    run<2>("16 FFMA + 8 LOP3", out);
    run<3>("8 FFMA2 + 8 LOP3", out);
    return 0;
}
