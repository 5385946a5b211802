// Gaussian arena: node gather + activations + SH colour in ONE kernel each way (sm_100a) -- SURVEY.md 8f row f1.
//
// Reference: every node model turns its raw parameters into rasterizer inputs with its own chain of small torch ops --
//   scales.exp(), quats / ||quats||, sigmoid(opacities)            gaussian_model/vanilla_gaussian_splatting.py:299-307
//   SH colours: cat(features_dc, features_rest), viewdirs = means.detach() - camera centre, normalise,
//   spherical_harmonics(n, viewdirs, colors), clamp(rgb + 0.5, 0, 1)                              :309-322
//   rigid nodes: means @ R_node^T + t_node, quat_mult(q_node, quats / ||quats||)        rigid_node.py:206-215, 243-252
// -- and the scene graph then concatenates the per-node results attribute by attribute, every step
// (mtgs/scene_model/mtgs_scene_graph.py:408-461, the torch.cat at :454-455).
// Here all nodes live in ONE structure-of-arrays arena (a node = a row range + a pose); one kernel reads the raw rows
// and writes the rasterizer inputs of the whole scene (no cat, no per-node launches, the SH coefficient block is read
// exactly once), and one kernel maps the rasterizer-input gradients back to the raw parameters (activation VJPs, the
// node rotation, the SH transpose).  HBM-bound streaming: 204 of the ~270 bytes per Gaussian are the SH coefficients,
// staged through shared memory with coalesced 16-byte loads / stores (rows padded to an odd number of 16-byte units:
// conflict-free per-thread row access).  No tensor cores (a per-row 16x3 contraction with a per-row basis).
#include "common.cuh"
#include "sh_basis.cuh"

constexpr int AR_THREADS = 128;
constexpr int AR_MAXK = 25;

// node pose record: R (9, row-major) | t (3) | q (4, wxyz)
constexpr int AR_POSE_FLOATS = 16;

__device__ __forceinline__ int ar_row_stride(int K) {
    int units = (K * 3 + 3) / 4;   // 16-byte units per row
    if ((units & 1) == 0) ++units;  // odd -> conflict-free float4 row reads
    return units * 4;
}

template <int DEG>
__device__ __forceinline__ void ar_basis(float x, float y, float z, float *B) {
    sh_basis<DEG>(x, y, z, B);
}

// unit view direction of a world-space point (0 when the point sits on the camera centre, like x / ||x|| -> nan guard
// is NOT applied by the reference; a Gaussian exactly at the camera centre is culled by the near plane anyway)
__device__ __forceinline__ void ar_viewdir(const float *mw, const float *campos, float *d) {
    d[0] = mw[0] - campos[0]; d[1] = mw[1] - campos[1]; d[2] = mw[2] - campos[2];
    const float inv = rsqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] *= inv; d[1] *= inv; d[2] *= inv;
}

template <int DEG>
__global__ void __launch_bounds__(AR_THREADS)
k_arena_fwd(const float *__restrict__ means, const float *__restrict__ scales_raw, const float *__restrict__ quats_raw,
            const float *__restrict__ opac_raw, const float *__restrict__ sh /* [N, K, 3] */,
            const int32_t *__restrict__ node_of /* [N] or null: everything in node 0 */,
            const float *__restrict__ poses /* [n_nodes, 16] */, const float *__restrict__ campos /* [3] */, int N, int K,
            float *__restrict__ means_w, float4 *__restrict__ quats_w, float *__restrict__ scales,
            float *__restrict__ opac, float *__restrict__ colors /* [N, 3] */, uint8_t *__restrict__ clamp_mask) {
    extern __shared__ __align__(16) float s_rows[];
    const int g0 = blockIdx.x * AR_THREADS;
    const int rows = min(AR_THREADS, N - g0);
    const int rf = K * 3, stride = ar_row_stride(K);
    // stage the CTA's coefficient block (contiguous in memory), coalesced
    {
        const float *src = sh + (size_t)g0 * rf;
        if ((rf & 3) == 0 && (((uintptr_t)sh) & 15) == 0) {
            const int qpr = rf / 4, total = rows * qpr;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            for (int i = threadIdx.x; i < total; i += AR_THREADS) {
                const int r = i / qpr, qd = i - r * qpr;
                *reinterpret_cast<float4 *>(s_rows + r * stride + 4 * qd) = ldg_stream4(src4 + i);
            }
        } else {
            const int total = rows * rf;
            for (int i = threadIdx.x; i < total; i += AR_THREADS) s_rows[(i / rf) * stride + (i % rf)] = src[i];
        }
    }
    __syncthreads();
    const int g = g0 + threadIdx.x;
    if (g >= N) return;
    const float *P = poses + (size_t)(node_of != nullptr ? node_of[g] : 0) * AR_POSE_FLOATS;
    // means: local @ R^T + t
    const float m0 = means[3 * g], m1 = means[3 * g + 1], m2 = means[3 * g + 2];
    float mw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) mw[i] = P[3 * i] * m0 + P[3 * i + 1] * m1 + P[3 * i + 2] * m2 + P[9 + i];
    means_w[3 * g] = mw[0]; means_w[3 * g + 1] = mw[1]; means_w[3 * g + 2] = mw[2];
    // quats: q_node (x) q / ||q||
    const float4 qr = reinterpret_cast<const float4 *>(quats_raw)[g];
    const float inv = 1.0f / sqrtf(qr.x * qr.x + qr.y * qr.y + qr.z * qr.z + qr.w * qr.w);
    const float w2 = qr.x * inv, x2 = qr.y * inv, y2 = qr.z * inv, z2 = qr.w * inv;
    const float w1 = P[12], x1 = P[13], y1 = P[14], z1 = P[15];
    quats_w[g] = make_float4(w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                             w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2);
    // scales, opacity
#pragma unroll
    for (int i = 0; i < 3; ++i) scales[3 * g + i] = expf(scales_raw[3 * g + i]);
    opac[g] = 1.0f / (1.0f + expf(-opac_raw[g]));
    // colour
    const float *row = s_rows + threadIdx.x * stride;
    float rgb[3];
    if (DEG == 0) {  // reference: sigmoid(features_dc) when sh_degree == 0
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = 1.0f / (1.0f + expf(-row[c]));
        clamp_mask[g] = 7;
    } else {
        float d[3], B[AR_MAXK];
        const float cp[3] = {campos[0], campos[1], campos[2]};
        ar_viewdir(mw, cp, d);
        ar_basis<DEG>(d[0], d[1], d[2], B);
        constexpr int NB = (DEG + 1) * (DEG + 1);
        unsigned mask = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < NB; ++k) a += B[k] * row[3 * k + c];
            a += 0.5f;
            if (a > 0.f && a < 1.f) mask |= 1u << c;  // clamp(., 0, 1) passes gradient strictly inside
            rgb[c] = fminf(fmaxf(a, 0.f), 1.f);
        }
        clamp_mask[g] = (uint8_t)mask;
    }
    colors[3 * g] = rgb[0]; colors[3 * g + 1] = rgb[1]; colors[3 * g + 2] = rgb[2];
}

template <int DEG>
__global__ void __launch_bounds__(AR_THREADS)
k_arena_bwd(const float *__restrict__ means, const float *__restrict__ scales_raw, const float *__restrict__ quats_raw,
            const float *__restrict__ opac_raw, const float *__restrict__ sh, const int32_t *__restrict__ node_of,
            const float *__restrict__ poses, const float *__restrict__ campos, int N, int K,
            const uint8_t *__restrict__ clamp_mask, const float *__restrict__ v_means_w,
            const float4 *__restrict__ v_quats_w, const float *__restrict__ v_scales, const float *__restrict__ v_opac,
            const float *__restrict__ v_colors, float *__restrict__ g_means, float4 *__restrict__ g_quats,
            float *__restrict__ g_scales, float *__restrict__ g_opac, float *__restrict__ g_sh) {
    extern __shared__ __align__(16) float s_rows[];
    const int g0 = blockIdx.x * AR_THREADS;
    const int rows = min(AR_THREADS, N - g0);
    const int rf = K * 3, stride = ar_row_stride(K);
    const int g = g0 + threadIdx.x;
    float *row = s_rows + threadIdx.x * stride;
    if (g < N) {
        const float *P = poses + (size_t)(node_of != nullptr ? node_of[g] : 0) * AR_POSE_FLOATS;
        // means: v_local = R^T v_world   (view directions are detached in the reference: no colour term)
        const float a0 = v_means_w[3 * g], a1 = v_means_w[3 * g + 1], a2 = v_means_w[3 * g + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j) g_means[3 * g + j] = P[j] * a0 + P[3 + j] * a1 + P[6 + j] * a2;
        // quats: q_w = L(q_node) q_n, q_n = q / ||q||
        const float4 qr = reinterpret_cast<const float4 *>(quats_raw)[g];
        const float inv = 1.0f / sqrtf(qr.x * qr.x + qr.y * qr.y + qr.z * qr.z + qr.w * qr.w);
        const float qn[4] = {qr.x * inv, qr.y * inv, qr.z * inv, qr.w * inv};
        const float w1 = P[12], x1 = P[13], y1 = P[14], z1 = P[15];
        const float4 vw = v_quats_w[g];
        float vn[4];  // L^T v
        vn[0] = w1 * vw.x + x1 * vw.y + y1 * vw.z + z1 * vw.w;
        vn[1] = -x1 * vw.x + w1 * vw.y + z1 * vw.z - y1 * vw.w;
        vn[2] = -y1 * vw.x - z1 * vw.y + w1 * vw.z + x1 * vw.w;
        vn[3] = -z1 * vw.x + y1 * vw.y - x1 * vw.z + w1 * vw.w;
        const float dotp = vn[0] * qn[0] + vn[1] * qn[1] + vn[2] * qn[2] + vn[3] * qn[3];
        g_quats[g] = make_float4((vn[0] - dotp * qn[0]) * inv, (vn[1] - dotp * qn[1]) * inv, (vn[2] - dotp * qn[2]) * inv,
                                 (vn[3] - dotp * qn[3]) * inv);
#pragma unroll
        for (int i = 0; i < 3; ++i) g_scales[3 * g + i] = v_scales[3 * g + i] * expf(scales_raw[3 * g + i]);
        const float sg = 1.0f / (1.0f + expf(-opac_raw[g]));
        g_opac[g] = v_opac[g] * sg * (1.0f - sg);
        // SH coefficients: v_coeff[k][c] = B_k v_rgb[c] inside the clamp, 0 elsewhere and for unused bases
        const unsigned mask = clamp_mask[g];
        float vc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) vc[c] = ((mask >> c) & 1u) ? v_colors[3 * g + c] : 0.f;
        if (DEG == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float s0 = 1.0f / (1.0f + expf(-sh[(size_t)g * rf + c]));
                row[c] = v_colors[3 * g + c] * s0 * (1.0f - s0);
            }
            for (int e = 3; e < rf; ++e) row[e] = 0.f;
        } else {
            const float m0 = means[3 * g], m1 = means[3 * g + 1], m2 = means[3 * g + 2];
            float mw[3], d[3], B[AR_MAXK];
#pragma unroll
            for (int i = 0; i < 3; ++i) mw[i] = P[3 * i] * m0 + P[3 * i + 1] * m1 + P[3 * i + 2] * m2 + P[9 + i];
            const float cp[3] = {campos[0], campos[1], campos[2]};
            ar_viewdir(mw, cp, d);
            ar_basis<DEG>(d[0], d[1], d[2], B);
            constexpr int NB = (DEG + 1) * (DEG + 1);
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                row[3 * k] = B[k] * vc[0];
                row[3 * k + 1] = B[k] * vc[1];
                row[3 * k + 2] = B[k] * vc[2];
            }
            for (int e = 3 * NB; e < rf; ++e) row[e] = 0.f;
        }
    }
    __syncthreads();
    // coalesced store of the CTA's gradient block
    float *dst = g_sh + (size_t)g0 * rf;
    if ((rf & 3) == 0 && (((uintptr_t)g_sh) & 15) == 0) {
        const int qpr = rf / 4, total = rows * qpr;
        for (int i = threadIdx.x; i < total; i += AR_THREADS) {
            const int r = i / qpr, qd = i - r * qpr;
            reinterpret_cast<float4 *>(dst)[i] = *reinterpret_cast<const float4 *>(s_rows + r * stride + 4 * qd);
        }
    } else {
        const int total = rows * rf;
        for (int i = threadIdx.x; i < total; i += AR_THREADS) dst[i] = s_rows[(i / rf) * stride + (i % rf)];
    }
}

static inline int ar_stride_host(int K) {
    int units = (K * 3 + 3) / 4;
    if ((units & 1) == 0) ++units;
    return units * 4;
}

#define AR_DISPATCH(KERNEL, ...)                                                          \
    do {                                                                                  \
        switch (degree) {                                                                 \
            case 0: KERNEL<0><<<grid, AR_THREADS, smem, st>>>(__VA_ARGS__); break;         \
            case 1: KERNEL<1><<<grid, AR_THREADS, smem, st>>>(__VA_ARGS__); break;         \
            case 2: KERNEL<2><<<grid, AR_THREADS, smem, st>>>(__VA_ARGS__); break;         \
            case 3: KERNEL<3><<<grid, AR_THREADS, smem, st>>>(__VA_ARGS__); break;         \
            default: KERNEL<4><<<grid, AR_THREADS, smem, st>>>(__VA_ARGS__); break;        \
        }                                                                                 \
    } while (0)

extern "C" int b2s_arena_fwd(const float *means, const float *scales_raw, const float *quats_raw, const float *opac_raw,
                             const float *sh, const int32_t *node_of, const float *poses, const float *campos, int N,
                             int K, int degree, float *means_w, float *quats_w, float *scales, float *opac, float *colors,
                             uint8_t *clamp_mask, b2s_stream_t stream) {
    if (N < 0 || K < 1 || K > AR_MAXK || !poses || !campos) return B2S_ERR_ARG;
    if (degree < 0 || degree > 4 || (degree + 1) * (degree + 1) > K) return B2S_ERR_UNSUPPORTED;
    if (N == 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = b2s_div_up(N, AR_THREADS);
    const size_t smem = (size_t)AR_THREADS * ar_stride_host(K) * sizeof(float);
    AR_DISPATCH(k_arena_fwd, means, scales_raw, quats_raw, opac_raw, sh, node_of, poses, campos, N, K, means_w,
                (float4 *)quats_w, scales, opac, colors, clamp_mask);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_arena_bwd(const float *means, const float *scales_raw, const float *quats_raw, const float *opac_raw,
                             const float *sh, const int32_t *node_of, const float *poses, const float *campos, int N,
                             int K, int degree, const uint8_t *clamp_mask, const float *v_means_w, const float *v_quats_w,
                             const float *v_scales, const float *v_opac, const float *v_colors, float *g_means,
                             float *g_quats, float *g_scales, float *g_opac, float *g_sh, b2s_stream_t stream) {
    if (N < 0 || K < 1 || K > AR_MAXK || !poses || !campos) return B2S_ERR_ARG;
    if (degree < 0 || degree > 4 || (degree + 1) * (degree + 1) > K) return B2S_ERR_UNSUPPORTED;
    if (N == 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = b2s_div_up(N, AR_THREADS);
    const size_t smem = (size_t)AR_THREADS * ar_stride_host(K) * sizeof(float);
    AR_DISPATCH(k_arena_bwd, means, scales_raw, quats_raw, opac_raw, sh, node_of, poses, campos, N, K, clamp_mask,
                v_means_w, (const float4 *)v_quats_w, v_scales, v_opac, v_colors, g_means, (float4 *)g_quats, g_scales,
                g_opac, g_sh);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
