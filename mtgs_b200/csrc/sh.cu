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

template <int DEG>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float *B) {
    B[0] = 0.2820947917738781f;
    if (DEG < 1) return;
    B[1] = -0.48860251190292f * y;
    B[2] = 0.48860251190292f * z;
    B[3] = -0.48860251190292f * x;
    if (DEG < 2) return;
    const float z2 = z * z;
    const float fTmp0B = -1.092548430592079f * z;
    const float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    B[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    B[7] = fTmp0B * x;
    B[5] = fTmp0B * y;
    B[8] = 0.5462742152960395f * fC1;
    B[4] = 0.5462742152960395f * fS1;
    if (DEG < 3) return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float fTmp1B = 1.445305721320277f * z;
    const float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    B[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
    B[13] = fTmp0C * x;
    B[11] = fTmp0C * y;
    B[14] = fTmp1B * fC1;
    B[10] = fTmp1B * fS1;
    B[15] = -0.5900435899266435f * fC2;
    B[9] = -0.5900435899266435f * fS2;
    if (DEG < 4) return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float fTmp2B = -1.770130769779931f * z;
    const float fC3 = x * fC2 - y * fS2, fS3 = x * fS2 + y * fC2;
    B[20] = 1.984313483298443f * z * B[12] + -1.006230589874905f * B[6];
    B[21] = fTmp0D * x;
    B[19] = fTmp0D * y;
    B[22] = fTmp1C * fC1;
    B[18] = fTmp1C * fS1;
    B[23] = fTmp2B * fC2;
    B[17] = fTmp2B * fS2;
    B[24] = 0.6258357354491763f * fC3;
    B[16] = 0.6258357354491763f * fS3;
}

// d(basis_k)/d(x,y,z) for the unit direction; dB[k] = (dx, dy, dz)
template <int DEG>
__device__ __forceinline__ void sh_basis_grad(float x, float y, float z, float3 *dB) {
    dB[0] = make_float3(0.f, 0.f, 0.f);
    if (DEG < 1) return;
    dB[1] = make_float3(0.f, -0.48860251190292f, 0.f);
    dB[2] = make_float3(0.f, 0.f, 0.48860251190292f);
    dB[3] = make_float3(-0.48860251190292f, 0.f, 0.f);
    if (DEG < 2) return;
    const float z2 = z * z;
    const float c0B = -1.092548430592079f, c1 = 0.5462742152960395f;
    const float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    // fC1: (2x, -2y, 0)   fS1: (2y, 2x, 0)
    dB[4] = make_float3(c1 * 2.f * y, c1 * 2.f * x, 0.f);
    dB[5] = make_float3(0.f, c0B * z, c0B * y);
    dB[6] = make_float3(0.f, 0.f, 2.f * 0.9461746957575601f * z);
    dB[7] = make_float3(c0B * z, 0.f, c0B * x);
    dB[8] = make_float3(c1 * 2.f * x, -c1 * 2.f * y, 0.f);
    if (DEG < 3) return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float dTmp0C = -2.f * 2.285228997322329f * z;
    const float c1B = 1.445305721320277f, c3 = -0.5900435899266435f;
    const float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    // fC2 = x^3 - 3 x y^2 : (3(x^2-y^2), -6xy, 0) = (3 fC1, -3 fS1, 0)
    // fS2 = 3 x^2 y - y^3 : (6xy, 3(x^2-y^2), 0) = (3 fS1, 3 fC1, 0)
    dB[9] = make_float3(c3 * 3.f * fS1, c3 * 3.f * fC1, 0.f);
    dB[10] = make_float3(c1B * z * 2.f * y, c1B * z * 2.f * x, c1B * fS1);
    dB[11] = make_float3(0.f, fTmp0C, dTmp0C * y);
    dB[12] = make_float3(0.f, 0.f, 3.f * 1.865881662950577f * z2 - 1.119528997770346f);
    dB[13] = make_float3(fTmp0C, 0.f, dTmp0C * x);
    dB[14] = make_float3(c1B * z * 2.f * x, -c1B * z * 2.f * y, c1B * fC1);
    dB[15] = make_float3(c3 * 3.f * fC1, -c3 * 3.f * fS1, 0.f);
    if (DEG < 4) return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float dTmp0D = -3.f * 4.683325804901025f * z2 + 2.007139630671868f;
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float dTmp1C = 2.f * 3.31161143515146f * z;
    const float c2B = -1.770130769779931f, c4 = 0.6258357354491763f;
    // fC3 = x fC2 - y fS2 : d/dx = fC2 + x*3fC1 - y*3fS1 = 4 fC2 ; d/dy = -3x fS1 - fS2 - 3y fC1 = -4 fS2
    // fS3 = x fS2 + y fC2 : d/dx = fS2 + 3x fS1 + 3y fC1 = 4 fS2 ; d/dy = 3x fC1 + fC2 - 3y fS1 = 4 fC2
    const float B12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    const float dB12 = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    const float dB6 = 2.f * 0.9461746957575601f * z;
    dB[16] = make_float3(c4 * 4.f * fS2, c4 * 4.f * fC2, 0.f);
    dB[17] = make_float3(c2B * z * 3.f * fS1, c2B * z * 3.f * fC1, c2B * fS2);
    dB[18] = make_float3(fTmp1C * 2.f * y, fTmp1C * 2.f * x, dTmp1C * fS1);
    dB[19] = make_float3(0.f, fTmp0D, dTmp0D * y);
    dB[20] = make_float3(0.f, 0.f, 1.984313483298443f * (B12 + z * dB12) - 1.006230589874905f * dB6);
    dB[21] = make_float3(fTmp0D, 0.f, dTmp0D * x);
    dB[22] = make_float3(fTmp1C * 2.f * x, -fTmp1C * 2.f * y, dTmp1C * fC1);
    dB[23] = make_float3(c2B * z * 3.f * fC1, -c2B * z * 3.f * fS1, c2B * fC2);
    dB[24] = make_float3(c4 * 4.f * fC2, -c4 * 4.f * fS2, 0.f);
}

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
