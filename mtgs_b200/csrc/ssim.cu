// Masked SSIM, forward and backward, as two fused kernels (sm_100a).
//
// SURVEY.md 8f row f3 (the first "next" row after the rasterizer): MTGS evaluates the SSIM term of its RGB loss at
// every training step through mtgs/utils/ssim.py (MaskedSSIM; call site mtgs/scene_model/mtgs_scene_graph.py:822-840):
// ten depthwise conv launches (two 1-D passes for each of X, Y, XX, YY, XY; ssim.py:28-53, 88-96), ~15 elementwise
// launches for the SSIM map (ssim.py:90-99), a masked_select and a mean (ssim.py:101-104), and the same again through
// autograd in the backward.  Here:
//   k_ssim_fwd   one pass over the image pair: a CTA stages a (TY + R - 1) x (TX + R - 1) patch of X and Y in shared
//                memory, runs the separable R-tap Gaussian over the five products, forms the SSIM value per output
//                pixel of the "valid" region (no padding, as the reference), accumulates the (masked) sum and count per
//                image plane, and keeps the three partial derivatives needed by the backward (dS/dmu, dS/dE[yy],
//                dS/dE[xy]; already multiplied by the mask) instead of the five filtered maps.
//   k_ssim_bwd   the adjoint of the separable filter applied to those three maps, fused with the chain rule:
//                grad_y[q] = scale * (F^T a + 2 y_q F^T b + x_q F^T c)[q].
// Both kernels are HBM streams (2 reads + 3 writes of an image plane forward, 5 reads + 1 write backward) with
// ~200 FMAs per pixel from shared memory; no tensor cores (an 11-tap depthwise filter has no contraction dimension
// worth a tcgen05 tile).
#include "common.cuh"

constexpr int SS_TX = 32, SS_TY = 8;  // output tile of a CTA (one thread per pixel)
constexpr int SS_RMAX = 15;           // largest window
constexpr int SS_PX = SS_TX + SS_RMAX - 1, SS_PY = SS_TY + SS_RMAX - 1;

__global__ void __launch_bounds__(SS_TX * SS_TY)
k_ssim_fwd(const float *__restrict__ X, const float *__restrict__ Y, const uint8_t *__restrict__ mask,
           long long mask_n_stride, long long mask_c_stride, int C, int H, int W, const float *__restrict__ win, int R,
           float C1, float C2, float *__restrict__ map_a, float *__restrict__ map_b, float *__restrict__ map_c,
           float *__restrict__ map_ax, double *__restrict__ acc) {
    __shared__ float s_x[SS_PY][SS_PX + 1], s_y[SS_PY][SS_PX + 1];
    __shared__ float s_h[5][SS_PY][SS_TX + 1];
    __shared__ float s_win[SS_RMAX];
    __shared__ float s_sum[SS_TX * SS_TY / 32], s_cnt[SS_TX * SS_TY / 32];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * SS_TX + tx;
    const int plane = blockIdx.z, n = plane / C, c = plane - n * C;
    const int Ho = H - R + 1, Wo = W - R + 1;
    const int ox0 = blockIdx.x * SS_TX, oy0 = blockIdx.y * SS_TY;
    const float *xp = X + (size_t)plane * H * W, *yp = Y + (size_t)plane * H * W;
    if (tid < R) s_win[tid] = win[tid];
    const int py = SS_TY + R - 1, px = SS_TX + R - 1;
    for (int i = tid; i < py * px; i += SS_TX * SS_TY) {
        const int r = i / px, q = i - r * px;
        const int gy = oy0 + r, gx = ox0 + q;
        const bool in = gy < H && gx < W;
        s_x[r][q] = in ? xp[(size_t)gy * W + gx] : 0.f;
        s_y[r][q] = in ? yp[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    // horizontal pass over every staged row
    for (int i = tid; i < py * SS_TX; i += SS_TX * SS_TY) {
        const int r = i / SS_TX, q = i - r * SS_TX;
        float hx = 0.f, hy = 0.f, hxx = 0.f, hyy = 0.f, hxy = 0.f;
        for (int k = 0; k < R; ++k) {
            const float w = s_win[k], a = s_x[r][q + k], b = s_y[r][q + k];
            hx = fmaf(w, a, hx);
            hy = fmaf(w, b, hy);
            hxx = fmaf(w, a * a, hxx);
            hyy = fmaf(w, b * b, hyy);
            hxy = fmaf(w, a * b, hxy);
        }
        s_h[0][r][q] = hx; s_h[1][r][q] = hy; s_h[2][r][q] = hxx; s_h[3][r][q] = hyy; s_h[4][r][q] = hxy;
    }
    __syncthreads();
    // vertical pass + SSIM of this thread's output pixel
    const int oy = oy0 + ty, ox = ox0 + tx;
    float contrib = 0.f, cnt = 0.f;
    if (oy < Ho && ox < Wo) {
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
        for (int k = 0; k < R; ++k) {
            const float w = s_win[k];
            mu1 = fmaf(w, s_h[0][ty + k][tx], mu1);
            mu2 = fmaf(w, s_h[1][ty + k][tx], mu2);
            e11 = fmaf(w, s_h[2][ty + k][tx], e11);
            e22 = fmaf(w, s_h[3][ty + k][tx], e22);
            e12 = fmaf(w, s_h[4][ty + k][tx], e12);
        }
        const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
        const float A1 = 2.f * mu1 * mu2 + C1, A2 = 2.f * s12 + C2;
        const float B1 = mu1 * mu1 + mu2 * mu2 + C1, B2 = s11 + s22 + C2;
        const float rB1 = 1.f / B1, rB2 = 1.f / B2;
        const float S = A1 * rB1 * A2 * rB2;
        const int half = R >> 1;
        float m = 1.f;
        if (mask) m = mask[(size_t)n * mask_n_stride + (size_t)c * mask_c_stride + (size_t)(oy + half) * W + (ox + half)] ? 1.f : 0.f;
        // partial derivatives of S w.r.t. the filtered statistics of Y (mu2, E[yy], E[xy]) and mu1
        const float dS_dmu2 = (2.f * mu1 * (A2 - A1)) * rB1 * rB2 - S * 2.f * mu2 * (rB1 - rB2);
        const float dS_dmu1 = (2.f * mu2 * (A2 - A1)) * rB1 * rB2 - S * 2.f * mu1 * (rB1 - rB2);
        const size_t o = ((size_t)plane * Ho + oy) * Wo + ox;
        map_a[o] = m * dS_dmu2;
        map_b[o] = m * (-S * rB2);
        map_c[o] = m * (2.f * A1 * rB1 * rB2);
        if (map_ax) map_ax[o] = m * dS_dmu1;
        contrib = m * S;
        cnt = m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((tid & 31) == 0) {
        s_sum[tid >> 5] = contrib;
        s_cnt[tid >> 5] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < SS_TX * SS_TY / 32; ++w) {
            a += (double)s_sum[w];
            b += (double)s_cnt[w];
        }
        if (b != 0.0 || a != 0.0) {
            atomicAdd(acc + 2 * plane, a);
            atomicAdd(acc + 2 * plane + 1, b);
        }
    }
}

// grad[q] = plane_scale * (F^T a + 2 self_q F^T b + other_q F^T c)[q], F^T = adjoint of the valid separable filter
__global__ void __launch_bounds__(SS_TX * SS_TY)
k_ssim_bwd(const float *__restrict__ self, const float *__restrict__ other, const float *__restrict__ map_a,
           const float *__restrict__ map_b, const float *__restrict__ map_c, const float *__restrict__ plane_scale, int H,
           int W, const float *__restrict__ win, int R, float *__restrict__ grad) {
    __shared__ float s_m[3][SS_PY][SS_PX + 1];
    __shared__ float s_h[3][SS_PY][SS_TX + 1];
    __shared__ float s_win[SS_RMAX];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * SS_TX + tx;
    const int plane = blockIdx.z;
    const int Ho = H - R + 1, Wo = W - R + 1;
    const int ix0 = blockIdx.x * SS_TX, iy0 = blockIdx.y * SS_TY;  // tile of INPUT pixels
    if (tid < R) s_win[tid] = win[tid];
    const int py = SS_TY + R - 1, px = SS_TX + R - 1;
    const float *ma = map_a + (size_t)plane * Ho * Wo, *mb = map_b + (size_t)plane * Ho * Wo,
                *mc = map_c + (size_t)plane * Ho * Wo;
    // output pixels (oy, ox) with oy in [iy - R + 1, iy], ox in [ix - R + 1, ix] cover input pixel (iy, ix)
    for (int i = tid; i < py * px; i += SS_TX * SS_TY) {
        const int r = i / px, q = i - r * px;
        const int oy = iy0 - (R - 1) + r, ox = ix0 - (R - 1) + q;
        const bool in = oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
        const size_t o = (size_t)oy * Wo + ox;
        s_m[0][r][q] = in ? ma[o] : 0.f;
        s_m[1][r][q] = in ? mb[o] : 0.f;
        s_m[2][r][q] = in ? mc[o] : 0.f;
    }
    __syncthreads();
    // horizontal adjoint: input column ix0 + q gathers staged columns q .. q + R - 1 with reversed taps
    for (int i = tid; i < py * SS_TX; i += SS_TX * SS_TY) {
        const int r = i / SS_TX, q = i - r * SS_TX;
        float ha = 0.f, hb = 0.f, hc = 0.f;
        for (int k = 0; k < R; ++k) {
            const float w = s_win[R - 1 - k];
            ha = fmaf(w, s_m[0][r][q + k], ha);
            hb = fmaf(w, s_m[1][r][q + k], hb);
            hc = fmaf(w, s_m[2][r][q + k], hc);
        }
        s_h[0][r][q] = ha; s_h[1][r][q] = hb; s_h[2][r][q] = hc;
    }
    __syncthreads();
    const int iy = iy0 + ty, ix = ix0 + tx;
    if (iy < H && ix < W) {
        float ga = 0.f, gb = 0.f, gc = 0.f;
        for (int k = 0; k < R; ++k) {
            const float w = s_win[R - 1 - k];
            ga = fmaf(w, s_h[0][ty + k][tx], ga);
            gb = fmaf(w, s_h[1][ty + k][tx], gb);
            gc = fmaf(w, s_h[2][ty + k][tx], gc);
        }
        const size_t o = ((size_t)plane * H + iy) * W + ix;
        grad[o] = plane_scale[plane] * (ga + 2.f * self[o] * gb + other[o] * gc);
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path: window size known at compile time (R = 11 is what MTGS uses).  The generic kernels above are
// instruction-issue bound (IPC 3.3, ~1100 instructions per pixel: runtime tap loops, five scalar FMAs and two scalar
// shared loads per tap).  Here the taps are unrolled with the weights in registers, the five products are formed once
// per staged pixel, and the accumulations run on packed fp32x2 FMAs: (x, y) and (xx, yy) travel as float2 halves, so a
// tap is one 16-byte + one 4-byte shared load and two FFMA2 + one FFMA.  CTA tile 32 x 16 (two rows per thread).
// ------------------------------------------------------------------------------------------------
constexpr int SF_TX = 32, SF_TY = 16, SF_NT = 256;

__device__ __forceinline__ float2 ss_fma2(float w, float2 v, float2 acc) {
    return __ffma2_rn(make_float2(w, w), v, acc);
}

template <int R>
__global__ void __launch_bounds__(SF_NT)
k_ssim_fwd_fast(const float *__restrict__ X, const float *__restrict__ Y, const uint8_t *__restrict__ mask,
                long long mask_n_stride, long long mask_c_stride, int C, int H, int W, const float *__restrict__ win,
                float C1, float C2, float *__restrict__ map_a, float *__restrict__ map_b, float *__restrict__ map_c,
                float *__restrict__ map_ax, double *__restrict__ acc) {
    constexpr int PY = SF_TY + R - 1, PX = SF_TX + R - 1;
    __shared__ float4 s_p4[PY][PX + 1];  // x, y, xx, yy
    __shared__ float s_p1[PY][PX + 1];   // xy
    __shared__ float4 s_h4[PY][SF_TX + 1];
    __shared__ float s_h1[PY][SF_TX + 1];
    __shared__ float s_sum[SF_NT / 32], s_cnt[SF_NT / 32];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int plane = blockIdx.z, n = plane / C, c = plane - n * C;
    const int Ho = H - R + 1, Wo = W - R + 1;
    const int ox0 = blockIdx.x * SF_TX, oy0 = blockIdx.y * SF_TY;
    const float *xp = X + (size_t)plane * H * W, *yp = Y + (size_t)plane * H * W;
    float w[R];
#pragma unroll
    for (int k = 0; k < R; ++k) w[k] = __ldg(win + k);
    for (int i = tid; i < PY * PX; i += SF_NT) {
        const int r = i / PX, q = i - r * PX;
        const int gy = oy0 + r, gx = ox0 + q;
        const bool in = gy < H && gx < W;
        const float a = in ? xp[(size_t)gy * W + gx] : 0.f, b = in ? yp[(size_t)gy * W + gx] : 0.f;
        s_p4[r][q] = make_float4(a, b, a * a, b * b);
        s_p1[r][q] = a * b;
    }
    __syncthreads();
    for (int i = tid; i < PY * SF_TX; i += SF_NT) {
        const int r = i >> 5, q = i & 31;
        float2 h01 = make_float2(0.f, 0.f), h23 = make_float2(0.f, 0.f);
        float h4 = 0.f;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const float4 p = s_p4[r][q + k];
            h01 = ss_fma2(w[k], make_float2(p.x, p.y), h01);
            h23 = ss_fma2(w[k], make_float2(p.z, p.w), h23);
            h4 = fmaf(w[k], s_p1[r][q + k], h4);
        }
        s_h4[r][q] = make_float4(h01.x, h01.y, h23.x, h23.y);
        s_h1[r][q] = h4;
    }
    __syncthreads();
    const int ox = ox0 + tx;
    float contrib = 0.f, cnt = 0.f;
#pragma unroll
    for (int j = 0; j < SF_TY / 8; ++j) {
        const int ly = ty + 8 * j, oy = oy0 + ly;
        if (!(oy < Ho && ox < Wo)) continue;
        float2 m01 = make_float2(0.f, 0.f), m23 = make_float2(0.f, 0.f);
        float e12 = 0.f;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const float4 p = s_h4[ly + k][tx];
            m01 = ss_fma2(w[k], make_float2(p.x, p.y), m01);
            m23 = ss_fma2(w[k], make_float2(p.z, p.w), m23);
            e12 = fmaf(w[k], s_h1[ly + k][tx], e12);
        }
        const float mu1 = m01.x, mu2 = m01.y, e11 = m23.x, e22 = m23.y;
        const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
        const float A1 = 2.f * mu1 * mu2 + C1, A2 = 2.f * s12 + C2;
        const float B1 = mu1 * mu1 + mu2 * mu2 + C1, B2 = s11 + s22 + C2;
        const float rB1 = 1.f / B1, rB2 = 1.f / B2;
        const float S = A1 * rB1 * A2 * rB2;
        float m = 1.f;
        if (mask)
            m = mask[(size_t)n * mask_n_stride + (size_t)c * mask_c_stride + (size_t)(oy + R / 2) * W + (ox + R / 2)] ? 1.f : 0.f;
        const float t = 2.f * (A2 - A1) * rB1 * rB2, u = 2.f * S * (rB1 - rB2);
        const size_t o = ((size_t)plane * Ho + oy) * Wo + ox;
        map_a[o] = m * (mu1 * t - mu2 * u);
        map_b[o] = m * (-S * rB2);
        map_c[o] = m * (2.f * A1 * rB1 * rB2);
        if (map_ax) map_ax[o] = m * (mu2 * t - mu1 * u);
        contrib += m * S;
        cnt += m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (tx == 0) {
        s_sum[ty] = contrib;
        s_cnt[ty] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int v = 0; v < SF_NT / 32; ++v) {
            a += (double)s_sum[v];
            b += (double)s_cnt[v];
        }
        if (b != 0.0 || a != 0.0) {
            atomicAdd(acc + 2 * plane, a);
            atomicAdd(acc + 2 * plane + 1, b);
        }
    }
}

template <int R>
__global__ void __launch_bounds__(SF_NT)
k_ssim_bwd_fast(const float *__restrict__ self, const float *__restrict__ other, const float *__restrict__ map_a,
                const float *__restrict__ map_b, const float *__restrict__ map_c, const float *__restrict__ plane_scale,
                int H, int W, const float *__restrict__ win, float *__restrict__ grad) {
    constexpr int PY = SF_TY + R - 1, PX = SF_TX + R - 1;
    __shared__ float2 s_m2[PY][PX + 1];  // a, b
    __shared__ float s_m1[PY][PX + 1];   // c
    __shared__ float2 s_h2[PY][SF_TX + 1];
    __shared__ float s_h1[PY][SF_TX + 1];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int plane = blockIdx.z;
    const int Ho = H - R + 1, Wo = W - R + 1;
    const int ix0 = blockIdx.x * SF_TX, iy0 = blockIdx.y * SF_TY;
    float w[R];  // reversed taps (adjoint)
#pragma unroll
    for (int k = 0; k < R; ++k) w[k] = __ldg(win + (R - 1 - k));
    const float *ma = map_a + (size_t)plane * Ho * Wo, *mb = map_b + (size_t)plane * Ho * Wo,
                *mc = map_c + (size_t)plane * Ho * Wo;
    for (int i = tid; i < PY * PX; i += SF_NT) {
        const int r = i / PX, q = i - r * PX;
        const int oy = iy0 - (R - 1) + r, ox = ix0 - (R - 1) + q;
        const bool in = oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
        const size_t o = (size_t)oy * Wo + ox;
        s_m2[r][q] = in ? make_float2(ma[o], mb[o]) : make_float2(0.f, 0.f);
        s_m1[r][q] = in ? mc[o] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < PY * SF_TX; i += SF_NT) {
        const int r = i >> 5, q = i & 31;
        float2 h = make_float2(0.f, 0.f);
        float hc = 0.f;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            h = ss_fma2(w[k], s_m2[r][q + k], h);
            hc = fmaf(w[k], s_m1[r][q + k], hc);
        }
        s_h2[r][q] = h;
        s_h1[r][q] = hc;
    }
    __syncthreads();
    const int ix = ix0 + tx;
    const float scale = plane_scale[plane];
#pragma unroll
    for (int j = 0; j < SF_TY / 8; ++j) {
        const int ly = ty + 8 * j, iy = iy0 + ly;
        if (!(iy < H && ix < W)) continue;
        float2 g = make_float2(0.f, 0.f);
        float gc = 0.f;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            g = ss_fma2(w[k], s_h2[ly + k][tx], g);
            gc = fmaf(w[k], s_h1[ly + k][tx], gc);
        }
        const size_t o = ((size_t)plane * H + iy) * W + ix;
        grad[o] = scale * (g.x + 2.f * self[o] * g.y + other[o] * gc);
    }
}

extern "C" int b2s_ssim_fwd(const float *X, const float *Y, const uint8_t *mask, long long mask_n_stride,
                            long long mask_c_stride, int N, int C, int H, int W, const float *win, int win_size, float C1,
                            float C2, float *map_a, float *map_b, float *map_c, float *map_ax, double *acc,
                            b2s_stream_t stream) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || !X || !Y || !win || !map_a || !map_b || !map_c || !acc) return B2S_ERR_ARG;
    if (win_size < 1 || win_size > SS_RMAX || !(win_size & 1)) return B2S_ERR_UNSUPPORTED;
    if (H < win_size || W < win_size) return B2S_ERR_UNSUPPORTED;  // the reference skips the filter along such a dimension
    if ((long long)N * C > 65535) return B2S_ERR_UNSUPPORTED;
    const int Ho = H - win_size + 1, Wo = W - win_size + 1;
    if (win_size == 11) {
        dim3 gridf(b2s_div_up(Wo, SF_TX), b2s_div_up(Ho, SF_TY), N * C);
        k_ssim_fwd_fast<11><<<gridf, SF_NT, 0, (cudaStream_t)stream>>>(X, Y, mask, mask_n_stride, mask_c_stride, C, H, W, win,
                                                                       C1, C2, map_a, map_b, map_c, map_ax, acc);
        B2S_LAUNCH_CHECK();
        return B2S_OK;
    }
    dim3 grid(b2s_div_up(Wo, SS_TX), b2s_div_up(Ho, SS_TY), N * C), block(SS_TX, SS_TY);
    k_ssim_fwd<<<grid, block, 0, (cudaStream_t)stream>>>(X, Y, mask, mask_n_stride, mask_c_stride, C, H, W, win, win_size, C1,
                                                         C2, map_a, map_b, map_c, map_ax, acc);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_ssim_bwd(const float *self, const float *other, const float *map_a, const float *map_b,
                            const float *map_c, const float *plane_scale, int N, int C, int H, int W, const float *win,
                            int win_size, float *grad, b2s_stream_t stream) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || !self || !other || !win || !map_a || !map_b || !map_c || !plane_scale || !grad)
        return B2S_ERR_ARG;
    if (win_size < 1 || win_size > SS_RMAX || !(win_size & 1)) return B2S_ERR_UNSUPPORTED;
    if (H < win_size || W < win_size) return B2S_ERR_UNSUPPORTED;
    if ((long long)N * C > 65535) return B2S_ERR_UNSUPPORTED;
    if (win_size == 11) {
        dim3 gridf(b2s_div_up(W, SF_TX), b2s_div_up(H, SF_TY), N * C);
        k_ssim_bwd_fast<11><<<gridf, SF_NT, 0, (cudaStream_t)stream>>>(self, other, map_a, map_b, map_c, plane_scale, H, W,
                                                                       win, grad);
        B2S_LAUNCH_CHECK();
        return B2S_OK;
    }
    dim3 grid(b2s_div_up(W, SS_TX), b2s_div_up(H, SS_TY), N * C), block(SS_TX, SS_TY);
    k_ssim_bwd<<<grid, block, 0, (cudaStream_t)stream>>>(self, other, map_a, map_b, map_c, plane_scale, H, W, win, win_size,
                                                         grad);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
