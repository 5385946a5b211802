// Real spherical-harmonics colour evaluation, forward and backward (sm_100a).
//
// Replaces upstream gsplat v1.4.0 compute_sh_{fwd,bwd}_kernel (kernels K1/K2) as called from
// mtgs/scene_model/gaussian_model/vanilla_gaussian_splatting.py:317 (and multi_color_gaussian_splatting.py:96,
// rigid_node.py:248, deformable_node.py:125).  MTGS adds the +0.5 and clamp outside this call.
//
// HBM-bound streaming: 12*K bytes of coefficients per Gaussian dominate.  One thread per Gaussian evaluates the
// basis once and contracts all three channels; a CTA's coefficient block (contiguous in memory) is staged
// through shared memory with coalesced 16-byte loads so every 32-byte sector is fetched once, and the rows
// are padded to an odd number of 16-byte units so the per-thread float4 reads are bank-conflict free.
#include "common.cuh"

constexpr int SH_THREADS = 128;
constexpr int SH_MAXK = 25;

#include "sh_basis.cuh"

// Stage the CTA's coefficient rows: global [g][K*3] floats -> smem rows of `stride` floats (stride*4 bytes
// = odd multiple of 16 B).  K*3 floats per row; rows are 16-byte aligned when K*3 % 4 == 0, else scalar path.
__device__ __forceinline__ void stage_rows(const float *__restrict__ src, float *s_rows, int rows, int row_floats,
                                           int stride, bool vec_ok) {
    if (vec_ok) {
        const int q_per_row = row_floats / 4;
        const int total = rows * q_per_row;
        const float4 *src4 = reinterpret_cast<const float4 *>(src);
        for (int i = threadIdx.x; i < total; i += SH_THREADS) {
            const int r = i / q_per_row, qd = i - r * q_per_row;
            *reinterpret_cast<float4 *>(s_rows + r * stride + 4 * qd) = ldg_stream4(src4 + i);
        }
    } else {
        const int total = rows * row_floats;
        for (int i = threadIdx.x; i < total; i += SH_THREADS) {
            const int r = i / row_floats, e = i - r * row_floats;
            s_rows[r * stride + e] = src[i];
        }
    }
}

// ---- TMA (bulk asynchronous copy) helpers: cp.async.bulk moves a contiguous, 16-byte aligned run between global and
// shared memory without occupying the threads; completion of loads is signalled on an mbarrier (transaction bytes),
// completion of stores through the bulk async-group.
__device__ __forceinline__ unsigned sh_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sh_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sh_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sh_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sh_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sh_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SH_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SH_DONE;\n"
        "bra SH_WAIT;\n"
        "SH_DONE:\n"
        "}\n" ::"r"(sh_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void sh_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sh_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(sh_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sh_bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(sh_smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}

static inline int sh_row_stride(int K) {
    int f = K * 3;
    int units = (f + 3) / 4;      // 16-byte units
    if ((units & 1) == 0) ++units;  // odd -> conflict-free float4 row reads across 8 consecutive rows
    return units * 4;
}

template <int DEG>
__global__ void __launch_bounds__(SH_THREADS)
k_sh_fwd(const float *__restrict__ dirs, const float *__restrict__ coeffs, const uint8_t *__restrict__ masks, int N,
         int K, int stride, float *__restrict__ colors) {
    extern __shared__ __align__(16) float s_rows[];
    __shared__ __align__(8) unsigned long long s_bar;
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const int g0 = blockIdx.x * SH_THREADS;
    const int rows = min(SH_THREADS, N - g0);
    const int rf = K * 3;
    // rows of 16-byte multiples (K = 16: 192 B) are fetched by the TMA engine, one bulk copy per row into the padded
    // shared-memory row, while the threads load their directions and evaluate the basis; other K: cooperative loads
    const bool tma = (rf & 3) == 0 && ((size_t)coeffs & 15) == 0;
    if (tma) {
        if (threadIdx.x == 0) sh_mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) sh_mbar_expect_tx(&s_bar, (unsigned)(rows * rf * 4));
        if ((int)threadIdx.x < rows)
            sh_bulk_g2s(s_rows + threadIdx.x * stride, coeffs + (size_t)(g0 + threadIdx.x) * rf, (unsigned)(rf * 4), &s_bar);
    } else {
        stage_rows(coeffs + (size_t)g0 * rf, s_rows, rows, rf, stride, (rf & 3) == 0);
        __syncthreads();
    }
    const int g = g0 + threadIdx.x;
    const bool live = g < N && !(masks && !masks[g]);
    float B[SH_MAXK];
    if (live) {
        float x = dirs[3 * g], y = dirs[3 * g + 1], z = dirs[3 * g + 2];
        if (DEG >= 1) {
            const float inorm = rsqrtf(x * x + y * y + z * z);
            x *= inorm; y *= inorm; z *= inorm;
        }
        sh_basis<DEG>(x, y, z, B);
    }
    if (tma) sh_mbar_wait(&s_bar, 0);  // every thread observes the completed transaction before reading the rows
    if (!live) return;
    const float *row = s_rows + threadIdx.x * stride;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        r0 = fmaf(B[k], row[3 * k], r0);
        r1 = fmaf(B[k], row[3 * k + 1], r1);
        r2 = fmaf(B[k], row[3 * k + 2], r2);
    }
    colors[3 * g] = r0;
    colors[3 * g + 1] = r1;
    colors[3 * g + 2] = r2;
}

template <int DEG>
__global__ void __launch_bounds__(SH_THREADS)
k_sh_bwd(const float *__restrict__ dirs, const float *__restrict__ coeffs, const uint8_t *__restrict__ masks, int N,
         int K, int stride, const float *__restrict__ v_colors, float *__restrict__ v_coeffs,
         float *__restrict__ v_dirs) {
    extern __shared__ __align__(16) float s_rows[];
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const int g0 = blockIdx.x * SH_THREADS;
    const int rows = min(SH_THREADS, N - g0);
    const int rf = K * 3;
    const bool need_dirs = v_dirs != nullptr && DEG >= 1;
    if (need_dirs) {
        stage_rows(coeffs + (size_t)g0 * rf, s_rows, rows, rf, stride, (rf & 3) == 0);
        __syncthreads();
    }
    const int g = g0 + threadIdx.x;
    const bool live = g < N && !(masks && !masks[g]);
    float B[SH_MAXK];
#pragma unroll
    for (int k = 0; k < SH_MAXK; ++k) B[k] = 0.f;
    float vc0 = 0.f, vc1 = 0.f, vc2 = 0.f;
    if (live) {
        float x = dirs[3 * g], y = dirs[3 * g + 1], z = dirs[3 * g + 2];
        float inorm = 1.f;
        if (DEG >= 1) {
            inorm = rsqrtf(x * x + y * y + z * z);
            x *= inorm; y *= inorm; z *= inorm;
        }
        sh_basis<DEG>(x, y, z, B);
        vc0 = v_colors[3 * g]; vc1 = v_colors[3 * g + 1]; vc2 = v_colors[3 * g + 2];
        if (need_dirs) {
            float3 dB[SH_MAXK];
            sh_basis_grad<DEG>(x, y, z, dB);
            const float *row = s_rows + threadIdx.x * stride;
            float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
            for (int k = 1; k < NB; ++k) {
                const float w = row[3 * k] * vc0 + row[3 * k + 1] * vc1 + row[3 * k + 2] * vc2;
                gx = fmaf(dB[k].x, w, gx);
                gy = fmaf(dB[k].y, w, gy);
                gz = fmaf(dB[k].z, w, gz);
            }
            const float dotp = gx * x + gy * y + gz * z;  // through n = d / |d|
            v_dirs[3 * g] = (gx - dotp * x) * inorm;
            v_dirs[3 * g + 1] = (gy - dotp * y) * inorm;
            v_dirs[3 * g + 2] = (gz - dotp * z) * inorm;
        }
    } else if (g < N && v_dirs != nullptr) {
        v_dirs[3 * g] = v_dirs[3 * g + 1] = v_dirs[3 * g + 2] = 0.f;
    }
    if (live && v_dirs != nullptr && DEG < 1) v_dirs[3 * g] = v_dirs[3 * g + 1] = v_dirs[3 * g + 2] = 0.f;
    // v_coeffs rows: write through shared memory so the global stores are coalesced 16-byte streams
    __syncthreads();
    if (g < N) {
        float *row = s_rows + threadIdx.x * stride;
#pragma unroll
        for (int k = 0; k < SH_MAXK; ++k) {
            if (k < K) {
                const float b = (k < NB) ? B[k] : 0.f;
                row[3 * k] = b * vc0;
                row[3 * k + 1] = b * vc1;
                row[3 * k + 2] = b * vc2;
            }
        }
    }
    float *dst = v_coeffs + (size_t)g0 * rf;
    if ((rf & 3) == 0 && ((size_t)v_coeffs & 15) == 0) {
        // each thread hands its finished 16-byte-multiple row to the TMA engine (shared -> global bulk store)
        if (g < N) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my generic-proxy writes -> async proxy
            sh_bulk_s2g(dst + (size_t)threadIdx.x * rf, s_rows + threadIdx.x * stride, (unsigned)(rf * 4));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the read
        }
        return;
    }
    __syncthreads();
    if ((rf & 3) == 0) {
        const int q_per_row = rf / 4;
        const int total = rows * q_per_row;
        for (int i = threadIdx.x; i < total; i += SH_THREADS) {
            const int r = i / q_per_row, qd = i - r * q_per_row;
            reinterpret_cast<float4 *>(dst)[i] = *reinterpret_cast<const float4 *>(s_rows + r * stride + 4 * qd);
        }
    } else {
        const int total = rows * rf;
        for (int i = threadIdx.x; i < total; i += SH_THREADS) {
            const int r = i / rf, e = i - r * rf;
            dst[i] = s_rows[r * stride + e];
        }
    }
}

static int sh_check(int degree, int N, int K) {
    if (N < 0 || K <= 0) return B2S_ERR_ARG;
    if (degree < 0 || degree > 4 || K > SH_MAXK) return B2S_ERR_UNSUPPORTED;
    if ((degree + 1) * (degree + 1) > K) return B2S_ERR_ARG;
    return B2S_OK;
}

extern "C" int b2s_sh_fwd(int degree, const float *dirs, const float *coeffs, const uint8_t *masks, int N, int K,
                          float *colors, b2s_stream_t stream) {
    int rc = sh_check(degree, N, K);
    if (rc) return rc;
    if (N == 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int stride = sh_row_stride(K);
    const size_t smem = (size_t)SH_THREADS * stride * 4;
    const int grid = b2s_div_up(N, SH_THREADS);
#define L(D)                                                                                              \
    case D:                                                                                               \
        cudaFuncSetAttribute(k_sh_fwd<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        k_sh_fwd<D><<<grid, SH_THREADS, smem, st>>>(dirs, coeffs, masks, N, K, stride, colors);           \
        break;
    switch (degree) { L(0) L(1) L(2) L(3) L(4) }
#undef L
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_sh_bwd(int degree, const float *dirs, const float *coeffs, const uint8_t *masks, int N, int K,
                          const float *v_colors, float *v_coeffs, float *v_dirs, b2s_stream_t stream) {
    int rc = sh_check(degree, N, K);
    if (rc) return rc;
    if (N == 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int stride = sh_row_stride(K);
    const size_t smem = (size_t)SH_THREADS * stride * 4;
    const int grid = b2s_div_up(N, SH_THREADS);
#define L(D)                                                                                              \
    case D:                                                                                               \
        cudaFuncSetAttribute(k_sh_bwd<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        k_sh_bwd<D><<<grid, SH_THREADS, smem, st>>>(dirs, coeffs, masks, N, K, stride, v_colors, v_coeffs, v_dirs); \
        break;
    switch (degree) { L(0) L(1) L(2) L(3) L(4) }
#undef L
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
