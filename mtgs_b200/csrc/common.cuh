// Shared device/host helpers for libb200splat (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200splat.h"

#define B2S_ALPHA_MAX 0.999f
#define B2S_ALPHA_MIN (1.0f / 255.0f)
#define B2S_T_EPS 1e-4f
#define B2S_LOG2E 1.4426950408889634f
#define B2S_LN2 0.6931471805599453f

// Diagnostic launch counter (api.cu): the library's only mutable global; atomic, never read by a kernel launch path.
void b2s_count_launch(int n);

static inline int b2s_check_launch() {
    cudaError_t e = cudaGetLastError();
    b2s_count_launch(1);
    return e == cudaSuccess ? B2S_OK : -(int)e - 1000;
}

#define B2S_LAUNCH_CHECK()                 \
    do {                                   \
        int _rc = b2s_check_launch();      \
        if (_rc != B2S_OK) return _rc;     \
    } while (0)

static inline int b2s_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// 16-byte vector reduction to global memory (sm_90+): one L2 atomic op for four floats.
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}
// streaming (read-once) loads that do not pollute L1
__device__ __forceinline__ float4 ldg_stream4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// Peer-memory description of the multi-GPU gradient exchange (exchange.cu), passed to kernels by value.
#define B2S_MAX_WORLD 8
struct B2sExchange {
    int world, rank;
    int shard;              // Gaussians (rows) owned per rank; multiple of 256
    int d_col;              // colour floats per row that take part in the exchange (d_in, or 0: colours stay local)
    long long slot_floats;  // floats per (owner, source) staging slot = (11 + d_col) * shard
    unsigned epoch;         // step counter written into the flags
    long long timeout_cycles;  // spin-loop budget of the flag waits (SM clock cycles)
    float *stage[B2S_MAX_WORLD];     // stage[r]: rank r's staging buffer [world][slot_floats] (peer-mapped)
    float *arena[B2S_MAX_WORLD];     // arena[r]: rank r's reduced-gradient arena (peer-mapped)
    unsigned *flags[B2S_MAX_WORLD];  // flags[r]: rank r's flag words [2 phases][B2S_MAX_WORLD] (peer-mapped)
};

// tile rows per row group / tile columns per column group of the tile-list hierarchy (tilelists.cu)
int b2s_tl_shifts(int tile_w, int tile_h, int *rg_shift, int *cg_shift);
