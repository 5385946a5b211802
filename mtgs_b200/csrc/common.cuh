// Shared device/host helpers for libb200splat (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200splat.h"

#define B2S_ALPHA_MAX 0.999f
#define B2S_ALPHA_MIN (1.0f / 255.0f)
#define B2S_OVERFLOW_RECORDS 9  // value of the overflow word when the blend ran out of record blocks (1..4: list levels)
#define B2S_T_EPS 1e-4f
#define B2S_LOG2E 1.4426950408889634f
#define B2S_LN2 0.6931471805599453f
#define B2S_N_TOTALS 5  // list sizes written by the projection: see b2s_project_fwd in include/b200splat.h

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

// Can this Gaussian reach alpha >= 1/255 at any pixel centre of the rectangle [rx0,rx1]x[ry0,ry1]?
// Conservative (never drops a contributing pair): exact box-constrained minimum of the convex quadratic,
// compared with log2(255 * opacity) plus a slack that dominates fp32 evaluation error.
__device__ __forceinline__ bool tile_keep(float mx, float my, float A, float B, float C, float opac, float rx0,
                                          float ry0, float rx1, float ry1) {
    if (!(opac >= 0.0039f)) return false;  // opac * e^-sigma <= opac < 1/255 (also drops NaN)
    float ex = fminf(fmaxf(mx, rx0), rx1) - mx;  // nearest pixel-centre coordinate minus mean (0 if inside)
    float ey = fminf(fmaxf(my, ry0), ry1) - my;
    if (ex == 0.f && ey == 0.f) return true;
    float tau = __log2f(opac * 255.0f);
    float smin = 3.0e38f, mag = 0.f;
    if (ex != 0.f) {  // facing vertical edge: dx fixed, minimise over dy
        float dy = fminf(fmaxf(-B * ex / (2.f * C), ry0 - my), ry1 - my);
        float t0 = A * ex * ex, t1 = B * ex * dy, t2 = C * dy * dy;
        float s = t0 + t1 + t2;
        if (s < smin) { smin = s; mag = fabsf(t0) + fabsf(t1) + fabsf(t2); }
    }
    if (ey != 0.f) {  // facing horizontal edge
        float dx = fminf(fmaxf(-B * ey / (2.f * A), rx0 - mx), rx1 - mx);
        float t0 = A * dx * dx, t1 = B * dx * ey, t2 = C * ey * ey;
        float s = t0 + t1 + t2;
        if (s < smin) { smin = s; mag = fabsf(t0) + fabsf(t1) + fabsf(t2); }
    }
    return !(smin > tau + 0.02f + 2e-5f * mag);  // NaN-safe: keep on NaN
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) and mbarriers ----
__device__ __forceinline__ unsigned b2s_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void b2s_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b2s_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void b2s_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b2s_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void b2s_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "B2S_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra B2S_DONE;\n"
        "bra B2S_WAIT;\n"
        "B2S_DONE:\n"
        "}\n" ::"r"(b2s_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void b2s_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     b2s_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(b2s_smem_u32(bar))
                 : "memory");
}
// shared -> global through the bulk async-group of the issuing thread
__device__ __forceinline__ void b2s_bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(b2s_smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void b2s_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's bulk stores have finished READING shared memory (the buffer may be overwritten)
__device__ __forceinline__ void b2s_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA)
__device__ __forceinline__ void b2s_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Peer-memory description of the multi-GPU gradient exchange (exchange.cu), passed to kernels by value.
#define B2S_MAX_WORLD 8
struct B2sExchange {
    int world, rank;
    int shard;              // Gaussians (rows) owned per rank; multiple of 256
    int d_col;              // colour floats per row that take part in the exchange (d_in, or 0: colours stay local)
    long long slot_floats;  // floats per (owner, source) staging slot = (11 + d_col) * shard
    int cta_rot;            // first 256-row block the fused projection backward works on (see k_project_bwd)
    unsigned epoch;         // step counter written into the flags (host-counted mode)
    unsigned *epoch_dev;    // device-counted mode (or null): this rank's own step counter, incremented by the first
                            // signal kernel of a step -- nothing of the launch depends on a host-side value, so the
                            // whole step can be replayed from a CUDA graph
    long long timeout_cycles;  // spin-loop budget of the flag waits (SM clock cycles)
    float *stage[B2S_MAX_WORLD];     // stage[r]: rank r's staging buffer [world][slot_floats] (peer-mapped)
    float *arena[B2S_MAX_WORLD];     // arena[r]: rank r's reduced-gradient arena (peer-mapped)
    unsigned *flags[B2S_MAX_WORLD];  // flags[r]: rank r's flag words [2 phases][B2S_MAX_WORLD] (peer-mapped)
};

// tile rows per row group / tile columns per column group of the tile-list hierarchy (tilelists.cu)
int b2s_tl_shifts(int tile_w, int tile_h, int *rg_shift, int *cg_shift);
