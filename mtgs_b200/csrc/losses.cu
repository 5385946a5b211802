// Image-space losses of the training step as fused kernels (sm_100a) -- SURVEY.md 8f row f3.
//
// Reference (all plain torch op sequences, several of them with a host sync through boolean indexing):
//   masked L1 on RGB / normals          mtgs/scene_model/mtgs_scene_graph.py:825-828, 929
//       torch.abs(gt - pred)[mask].mean()
//   LiDAR depth loss, InverseL1 / L1    mtgs_scene_graph.py:875-884
//       torch.abs(1 / (gt + 1e-5) - 1 / (pred + 1e-5))[mask].mean()
//   total-variation loss                mtgs/utils/geometric_loss.py:287-303 (TVLoss)
//   patch NCC between depth maps        mtgs/utils/geometric_loss.py:322-348 (calculate_depth_ncc_loss)
//   normals from a depth image          mtgs/utils/geometric_loss.py:350-388 (pcd_to_normal, normal_from_depth_image)
//                                       + mtgs/utils/camera_utils.py:74-148 (pixel-centre back-projection)
// Each loss is ONE streaming pass forward (sum and count accumulate in doubles, so the mean is formed on the device
// and nothing synchronises) and ONE streaming pass backward that rebuilds the sign / weight from the inputs instead of
// storing an intermediate.  HBM-bound elementwise / stencil work: coalesced loads, no tensor cores.
#include "common.cuh"

constexpr int LS_THREADS = 256;

__device__ __forceinline__ double ls_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block sum of two doubles -> atomicAdd into acc[0], acc[1]
__device__ __forceinline__ void ls_block_accumulate(double a, double b, double *acc) {
    __shared__ double s_a[LS_THREADS / 32], s_b[LS_THREADS / 32];
    a = ls_warp_sum(a);
    b = ls_warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_a[warp] = a;
        s_b[warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int w = 0; w < LS_THREADS / 32; ++w) {
            ta += s_a[w];
            tb += s_b[w];
        }
        if (ta != 0.0) atomicAdd(acc, ta);
        if (tb != 0.0) atomicAdd(acc + 1, tb);
    }
}

// ---- masked L1 -------------------------------------------------------------------------------------------------
// pred, gt: [P, C]; mask: uint8 [P] (non-zero = selected) or null; MODE 0: |gt - pred|, MODE 1: |1/(gt+eps) - 1/(pred+eps)|
template <int MODE>
__device__ __forceinline__ float ls_l1_term(float p, float g, float eps) {
    if (MODE == 0) return fabsf(g - p);
    return fabsf(1.0f / (g + eps) - 1.0f / (p + eps));
}

template <int MODE>
__global__ void __launch_bounds__(LS_THREADS)
k_masked_l1_fwd(const float *__restrict__ pred, const float *__restrict__ gt, const uint8_t *__restrict__ mask,
                long long P, int C, float eps, double *__restrict__ acc /* [2]: sum, count */) {
    double sum = 0.0, cnt = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        if (mask != nullptr && mask[p] == 0) continue;
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += ls_l1_term<MODE>(pred[p * C + c], gt[p * C + c], eps);
        sum += (double)s;
        cnt += (double)C;
    }
    ls_block_accumulate(sum, cnt, acc);
}

// grad_pred[p, c] = (*grad_out / count) * d term / d pred   (0 where the mask is off)
template <int MODE>
__global__ void __launch_bounds__(LS_THREADS)
k_masked_l1_bwd(const float *__restrict__ pred, const float *__restrict__ gt, const uint8_t *__restrict__ mask,
                long long P, int C, float eps, const double *__restrict__ acc, const float *__restrict__ grad_out,
                float *__restrict__ grad_pred) {
    const float scale = (float)((double)(*grad_out) / acc[1]);
    const long long n = P * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C;
        float gr = 0.f;
        if (mask == nullptr || mask[p] != 0) {
            const float pv = pred[i], gv = gt[i];
            if (MODE == 0) {
                const float d = pv - gv;
                gr = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
            } else {
                const float ip = 1.0f / (pv + eps);
                const float u = 1.0f / (gv + eps) - ip;
                gr = (u > 0.f ? scale : (u < 0.f ? -scale : 0.f)) * ip * ip;
            }
        }
        grad_pred[i] = gr;
    }
}

// ---- total variation (geometric_loss.py:287-303) ---------------------------------------------------------------
// pred [B, H, W, C]: mean |pred[:, :, :-1] - pred[:, :, 1:]| + mean |pred[:, :-1] - pred[:, 1:]|
__global__ void __launch_bounds__(LS_THREADS)
k_tv_fwd(const float *__restrict__ pred, int B, int H, int W, int C, double *__restrict__ acc /* [2]: sum along W, along H */) {
    const long long n = (long long)B * H * W * C;
    double sw = 0.0, sh = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)((i / C) % W), y = (int)((i / ((long long)C * W)) % H);
        const float v = pred[i];
        if (x + 1 < W) sw += (double)fabsf(v - pred[i + C]);
        if (y + 1 < H) sh += (double)fabsf(v - pred[i + (long long)C * W]);
    }
    ls_block_accumulate(sw, sh, acc);
}

__device__ __forceinline__ float ls_sign(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(LS_THREADS)
k_tv_bwd(const float *__restrict__ pred, int B, int H, int W, int C, const float *__restrict__ grad_out,
         float *__restrict__ grad_pred) {
    const long long n = (long long)B * H * W * C;
    const float go = *grad_out;
    const float gw = W > 1 ? go / (float)((double)B * H * (W - 1) * C) : 0.f;
    const float gh = H > 1 ? go / (float)((double)B * (H - 1) * W * C) : 0.f;
    const long long sx = C, sy = (long long)C * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)((i / C) % W), y = (int)((i / sy) % H);
        const float v = pred[i];
        float g = 0.f;
        if (x + 1 < W) g += gw * ls_sign(v - pred[i + sx]);
        if (x > 0) g -= gw * ls_sign(pred[i - sx] - v);
        if (y + 1 < H) g += gh * ls_sign(v - pred[i + sy]);
        if (y > 0) g -= gh * ls_sign(pred[i - sy] - v);
        grad_pred[i] = g;
    }
}

// ---- patch NCC between two depth maps (geometric_loss.py:322-348) ----------------------------------------------
// Patches of patch x patch pixels at stride `stride`, zero padding patch / 2 (F.unfold); a patch counts only when
// every one of its mask entries is set (padding counts as unset).  One warp per patch.
// stats[patch] = (pred mean, 1 / pred std, gt mean, 1 / gt std, ncc, valid)
__global__ void __launch_bounds__(LS_THREADS)
k_ncc_fwd(const float *__restrict__ pred, const float *__restrict__ gt, const uint8_t *__restrict__ mask, int H, int W,
          int patch, int stride, int npy, int npx, float *__restrict__ stats, double *__restrict__ acc /* sum ncc, count */) {
    const int lane = threadIdx.x & 31;
    const int pid = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    double my_ncc = 0.0, my_cnt = 0.0;
    if (pid < npy * npx) {
        const int py = pid / npx, pxi = pid - py * npx;
        const int pad = patch / 2;
        const int y0 = py * stride - pad, x0 = pxi * stride - pad;
        const int n = patch * patch;
        bool ok = y0 >= 0 && x0 >= 0 && y0 + patch <= H && x0 + patch <= W;  // padding is never valid
        if (ok) {
            for (int e = lane; e < n; e += 32) {
                const int yy = y0 + e / patch, xx = x0 + e % patch;
                ok = ok && mask[(size_t)yy * W + xx] != 0;
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        float sp = 0.f, sg = 0.f;
        if (ok) {
            for (int e = lane; e < n; e += 32) {
                const size_t o = (size_t)(y0 + e / patch) * W + (x0 + e % patch);
                sp += pred[o];
                sg += gt[o];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sp += __shfl_xor_sync(0xffffffffu, sp, o);
            sg += __shfl_xor_sync(0xffffffffu, sg, o);
        }
        const float inv_n = 1.0f / (float)n;
        const float mp = sp * inv_n, mg = sg * inv_n;
        float vpp = 0.f, vgg = 0.f, vpg = 0.f;
        if (ok) {
            for (int e = lane; e < n; e += 32) {
                const size_t o = (size_t)(y0 + e / patch) * W + (x0 + e % patch);
                const float a = pred[o] - mp, b = gt[o] - mg;
                vpp += a * a;
                vgg += b * b;
                vpg += a * b;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vpp += __shfl_xor_sync(0xffffffffu, vpp, o);
            vgg += __shfl_xor_sync(0xffffffffu, vgg, o);
            vpg += __shfl_xor_sync(0xffffffffu, vpg, o);
        }
        const float isp = 1.0f / sqrtf(vpp * inv_n + 1e-8f), isg = 1.0f / sqrtf(vgg * inv_n + 1e-8f);
        const float ncc = vpg * inv_n * isp * isg;
        if (lane == 0) {
            float *st = stats + (size_t)pid * 6;
            st[0] = mp; st[1] = isp; st[2] = mg; st[3] = isg; st[4] = ok ? ncc : 0.f; st[5] = ok ? 1.f : 0.f;
            if (ok) {
                my_ncc = (double)ncc;
                my_cnt = 1.0;
            }
        }
    }
    ls_block_accumulate(my_ncc, my_cnt, acc);
}

// loss = 1 - mean ncc.  d ncc / d pred_j = (g^_j - p^_j ncc) / (n sigma_p) for the pixels of a valid patch; a pixel
// gathers the contributions of every patch that covers it (no atomics).
__global__ void __launch_bounds__(LS_THREADS)
k_ncc_bwd(const float *__restrict__ pred, const float *__restrict__ gt, int H, int W, int patch, int stride, int npy,
          int npx, const float *__restrict__ stats, const double *__restrict__ acc, const float *__restrict__ grad_out,
          float *__restrict__ grad_pred) {
    const long long n_pix = (long long)H * W;
    const float scale = (float)(-(double)(*grad_out) / acc[1]) / (float)(patch * patch);
    const int pad = patch / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        // patches with y0 <= y < y0 + patch, y0 = py * stride - pad
        const int py_hi = min(npy - 1, (y + pad) / stride);
        const int py_lo = max(0, (y + pad - patch + stride) / stride);  // ceil((y + pad - patch + 1) / stride)
        const int px_hi = min(npx - 1, (x + pad) / stride);
        const int px_lo = max(0, (x + pad - patch + stride) / stride);
        const float pv = pred[i], gv = gt[i];
        float g = 0.f;
        for (int py = py_lo; py <= py_hi; ++py)
            for (int pxi = px_lo; pxi <= px_hi; ++pxi) {
                const float *st = stats + ((size_t)py * npx + pxi) * 6;
                if (st[5] != 0.f) {
                    const float ph = (pv - st[0]) * st[1], gh = (gv - st[2]) * st[3];
                    g += (gh - ph * st[4]) * st[1];
                }
            }
        grad_pred[i] = g * scale;
    }
}

// ---- normals from a depth image (geometric_loss.py:350-388, camera_utils.py:74-148) ------------------------------
// means3d = ((u + 0.5 - cx) d / fx, (v + 0.5 - cy) d / fy, d) @ A + t   (A = inv(c2w[:3, :3]), t = c2w[:3, 3]: the
// reference's own convention); normal = normalize(cross(right - left, top - bottom)), zero on the 1-pixel border.
__device__ __forceinline__ void ls_backproject(const float *__restrict__ depth, int x, int y, int W, float fx, float fy,
                                               float cx, float cy, const float *A, const float *t, float *out) {
    const float d = depth[(size_t)y * W + x];
    const float px = ((float)x + 0.5f - cx) * d / fx, py = ((float)y + 0.5f - cy) * d / fy, pz = d;
#pragma unroll
    for (int j = 0; j < 3; ++j) out[j] = px * A[0 * 3 + j] + py * A[1 * 3 + j] + pz * A[2 * 3 + j] + t[j];
}

__global__ void __launch_bounds__(LS_THREADS)
k_normal_from_depth(const float *__restrict__ depth, int H, int W, float fx, float fy, float cx, float cy,
                    const float *__restrict__ A_t /* [12]: A row-major, then t */, float *__restrict__ normals) {
    const long long n = (long long)H * W;
    float A[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = A_t[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = A_t[9 + k];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (x > 0 && y > 0 && x < W - 1 && y < H - 1) {
            float r[3], l[3], tp[3], b[3];
            ls_backproject(depth, x + 1, y, W, fx, fy, cx, cy, A, t, r);
            ls_backproject(depth, x - 1, y, W, fx, fy, cx, cy, A, t, l);
            ls_backproject(depth, x, y - 1, W, fx, fy, cx, cy, A, t, tp);
            ls_backproject(depth, x, y + 1, W, fx, fy, cx, cy, A, t, b);
            const float ax = r[0] - l[0], ay = r[1] - l[1], az = r[2] - l[2];      // left_to_right
            const float bx = tp[0] - b[0], by = tp[1] - b[1], bz = tp[2] - b[2];   // bottom_to_top
            nx = ay * bz - az * by;
            ny = az * bx - ax * bz;
            nz = ax * by - ay * bx;
            const float inv = 1.0f / fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);  // F.normalize eps
            nx *= inv; ny *= inv; nz *= inv;
        }
        normals[i * 3] = nx;
        normals[i * 3 + 1] = ny;
        normals[i * 3 + 2] = nz;
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static inline int ls_grid(long long n) {
    long long g = (n + LS_THREADS - 1) / LS_THREADS;
    const long long cap = 148LL * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int b2s_masked_l1_fwd(const float *pred, const float *gt, const uint8_t *mask, long long P, int C, int mode,
                                 float eps, double *acc, b2s_stream_t stream) {
    if (P < 0 || C < 1 || acc == nullptr) return B2S_ERR_ARG;
    if (mode != 0 && mode != 1) return B2S_ERR_UNSUPPORTED;
    if (P == 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) k_masked_l1_fwd<0><<<ls_grid(P), LS_THREADS, 0, st>>>(pred, gt, mask, P, C, eps, acc);
    else k_masked_l1_fwd<1><<<ls_grid(P), LS_THREADS, 0, st>>>(pred, gt, mask, P, C, eps, acc);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_masked_l1_bwd(const float *pred, const float *gt, const uint8_t *mask, long long P, int C, int mode,
                                 float eps, const double *acc, const float *grad_out, float *grad_pred,
                                 b2s_stream_t stream) {
    if (P < 0 || C < 1 || acc == nullptr || grad_out == nullptr) return B2S_ERR_ARG;
    if (mode != 0 && mode != 1) return B2S_ERR_UNSUPPORTED;
    if (P == 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0)
        k_masked_l1_bwd<0><<<ls_grid(P * C), LS_THREADS, 0, st>>>(pred, gt, mask, P, C, eps, acc, grad_out, grad_pred);
    else
        k_masked_l1_bwd<1><<<ls_grid(P * C), LS_THREADS, 0, st>>>(pred, gt, mask, P, C, eps, acc, grad_out, grad_pred);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_tv_fwd(const float *pred, int B, int H, int W, int C, double *acc, b2s_stream_t stream) {
    if (B < 0 || H < 1 || W < 1 || C < 1 || acc == nullptr) return B2S_ERR_ARG;
    if (B == 0) return B2S_OK;
    k_tv_fwd<<<ls_grid((long long)B * H * W * C), LS_THREADS, 0, (cudaStream_t)stream>>>(pred, B, H, W, C, acc);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_tv_bwd(const float *pred, int B, int H, int W, int C, const float *grad_out, float *grad_pred,
                          b2s_stream_t stream) {
    if (B < 0 || H < 1 || W < 1 || C < 1 || grad_out == nullptr) return B2S_ERR_ARG;
    if (B == 0) return B2S_OK;
    k_tv_bwd<<<ls_grid((long long)B * H * W * C), LS_THREADS, 0, (cudaStream_t)stream>>>(pred, B, H, W, C, grad_out,
                                                                                         grad_pred);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_ncc_patch_grid(int H, int W, int patch, int stride, int *npy, int *npx) {
    if (H < 1 || W < 1 || patch < 1 || stride < 1 || !npy || !npx) return B2S_ERR_ARG;
    const int pad = patch / 2;
    *npy = (H + 2 * pad - patch) / stride + 1;  // F.unfold output size
    *npx = (W + 2 * pad - patch) / stride + 1;
    return (*npy > 0 && *npx > 0) ? B2S_OK : B2S_ERR_ARG;
}

extern "C" int b2s_ncc_fwd(const float *pred, const float *gt, const uint8_t *mask, int H, int W, int patch, int stride,
                           float *stats, double *acc, b2s_stream_t stream) {
    int npy, npx;
    if (b2s_ncc_patch_grid(H, W, patch, stride, &npy, &npx) != B2S_OK || !mask || !stats || !acc) return B2S_ERR_ARG;
    const long long threads = (long long)npy * npx * 32;
    k_ncc_fwd<<<(int)((threads + LS_THREADS - 1) / LS_THREADS), LS_THREADS, 0, (cudaStream_t)stream>>>(
        pred, gt, mask, H, W, patch, stride, npy, npx, stats, acc);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_ncc_bwd(const float *pred, const float *gt, int H, int W, int patch, int stride, const float *stats,
                           const double *acc, const float *grad_out, float *grad_pred, b2s_stream_t stream) {
    int npy, npx;
    if (b2s_ncc_patch_grid(H, W, patch, stride, &npy, &npx) != B2S_OK || !stats || !acc || !grad_out) return B2S_ERR_ARG;
    k_ncc_bwd<<<ls_grid((long long)H * W), LS_THREADS, 0, (cudaStream_t)stream>>>(pred, gt, H, W, patch, stride, npy, npx,
                                                                                 stats, acc, grad_out, grad_pred);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_normal_from_depth(const float *depth, int H, int W, float fx, float fy, float cx, float cy,
                                     const float *A_t, float *normals, b2s_stream_t stream) {
    if (H < 1 || W < 1 || !A_t) return B2S_ERR_ARG;
    k_normal_from_depth<<<ls_grid((long long)H * W), LS_THREADS, 0, (cudaStream_t)stream>>>(depth, H, W, fx, fy, cx, cy,
                                                                                           A_t, normals);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
