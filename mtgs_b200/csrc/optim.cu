// Fused multi-tensor Adam and densification primitives (sm_100a) -- SURVEY.md 8f row f2.
//
// Reference: one torch.optim.Adam per (node, attribute) parameter group, stepped one after the other
// (mtgs/scene_model/custom_trainer.py:115-136 builds them; nerfstudio's Optimizers.optimizer_step_all runs them), so a
// scene graph with hundreds of vehicle nodes spends its optimizer phase in kernel launches; and the densification
// bookkeeping of mtgs/scene_model/gaussian_model/vanilla_gaussian_splatting.py:448-474 (after_train statistics),
// :392-446 (optimizer-state surgery) and :476-699 (cull / split / duplicate) as boolean-indexing torch ops with
// host synchronisations.  Here:
//   k_adam_multi        ONE launch steps every tensor of every group: a chunk table maps 4096-element chunks to
//                       tensors (per-tensor lr / betas / eps / bias corrections in a small device table)
//   k_densify_stats     xys_grad_norm / vis_counts / max_2Dsize update from info["means2d"].absgrad and radii
//                       (mtgs_scene_graph.py:1171-1178 + vanilla_gaussian_splatting.py:455-474) in one pass
//   k_mask_*            order-preserving stream compaction of rows by a keep mask (block counts -> scan -> scatter);
//                       used for parameters AND their Adam moments (remove_from_optim)
// HBM-bound streaming kernels: 16-byte vector loads where the tensor allows it, no tensor cores.
#include "common.cuh"

constexpr int OP_THREADS = 256;
constexpr int OP_CHUNK = 4096;

__global__ void __launch_bounds__(OP_THREADS)
k_adam_multi(const B2sAdamTensor *__restrict__ tensors, const int32_t *__restrict__ chunk_tensor,
             const long long *__restrict__ chunk_start) {
    const B2sAdamTensor t = tensors[chunk_tensor[blockIdx.x]];
    const long long s0 = chunk_start[blockIdx.x];
    const long long s1 = s0 + OP_CHUNK < t.n ? s0 + OP_CHUNK : t.n;
    const float b2 = t.beta2, ob1 = t.one_minus_beta1, ob2 = t.one_minus_beta2;
    const float step_size = t.step_size, bc2_sqrt = t.bias_correction2_sqrt;
    auto upd = [&](float &p, float g, float &m, float &v) {
        if (t.weight_decay != 0.f) g += t.weight_decay * p;
        m = m + (g - m) * ob1;            // torch: exp_avg.lerp_(grad, 1 - beta1)
        v = v * b2 + ob2 * g * g;          // torch: exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        const float denom = sqrtf(v) / bc2_sqrt + t.eps;  // torch: (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
        p = p - step_size * (m / denom);   // torch: param.addcdiv_(exp_avg, denom, value=-step_size)
    };
    const bool vec = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0;
    if (vec) {
        for (long long i = s0 + 4 * threadIdx.x; i < s1; i += 4 * OP_THREADS) {
            if (i + 3 < s1) {
                float4 p = *reinterpret_cast<float4 *>(t.p + i), m = *reinterpret_cast<float4 *>(t.m + i),
                       v = *reinterpret_cast<float4 *>(t.v + i);
                const float4 g = *reinterpret_cast<const float4 *>(t.g + i);
                upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
                *reinterpret_cast<float4 *>(t.p + i) = p;
                *reinterpret_cast<float4 *>(t.m + i) = m;
                *reinterpret_cast<float4 *>(t.v + i) = v;
            } else {
                for (long long k = i; k < s1; ++k) upd(t.p[k], t.g[k], t.m[k], t.v[k]);
            }
        }
    } else {
        for (long long i = s0 + threadIdx.x; i < s1; i += OP_THREADS) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
    }
}

// after_train statistics: for visible Gaussians (radii > 0)
//   xys_grad_norm += || absgrad * (W, H) / 2 ||,  vis_counts += 1,  max_2Dsize = max(max_2Dsize, radii)
__global__ void __launch_bounds__(OP_THREADS)
k_densify_stats(const float *__restrict__ grad2d, int grad_stride, const int32_t *__restrict__ radii, int N, float half_w,
                float half_h, float *__restrict__ xys_grad_norm, float *__restrict__ vis_counts,
                float *__restrict__ max_2dsize) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const int r = radii[g];
    if (r <= 0) return;
    const float gx = grad2d[(size_t)g * grad_stride] * half_w, gy = grad2d[(size_t)g * grad_stride + 1] * half_h;
    xys_grad_norm[g] += sqrtf(gx * gx + gy * gy);
    vis_counts[g] += 1.0f;
    max_2dsize[g] = fmaxf(max_2dsize[g], (float)r);
}

// ---- order-preserving compaction by a keep mask ----
__global__ void __launch_bounds__(OP_THREADS)
k_mask_block_count(const uint8_t *__restrict__ keep, long long N, int32_t *__restrict__ block_cnt) {
    const long long i = (long long)blockIdx.x * OP_THREADS + threadIdx.x;
    const int c = __syncthreads_count(i < N && keep[i] != 0);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = c;
}

// single CTA: exclusive scan of the block counts in place, total -> *total
__global__ void __launch_bounds__(1024)
k_mask_scan_blocks(int32_t *__restrict__ block_cnt, int nblk, int32_t *__restrict__ total) {
    __shared__ int s_w[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int carry = 0;
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblk ? block_cnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            s_w[lane] = wi - w;
            if (lane == 31) s_w[32] = wi;
        }
        __syncthreads();
        if (i < nblk) block_cnt[i] = carry + s_w[warp] + incl - v;
        carry += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// dst[rank(i)] = src[i] for kept rows (rows of `row_floats` floats); rank from the block offsets + an in-block scan
__global__ void __launch_bounds__(OP_THREADS)
k_mask_gather_rows(const uint8_t *__restrict__ keep, long long N, const int32_t *__restrict__ block_off,
                   const float *__restrict__ src, float *__restrict__ dst, int row_floats) {
    __shared__ int s_w[OP_THREADS / 32];
    __shared__ int s_rank[OP_THREADS];
    const long long i = (long long)blockIdx.x * OP_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool k = i < N && keep[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int off = block_off[blockIdx.x];
    for (int w = 0; w < warp; ++w) off += s_w[w];
    s_rank[threadIdx.x] = k ? off + __popc(bal & lanemask_lt()) : -1;
    __syncthreads();
    // the CTA copies its kept rows cooperatively: consecutive threads move consecutive floats of a row
    const long long row0 = (long long)blockIdx.x * OP_THREADS;
    const int total = OP_THREADS * row_floats;
    for (int e = threadIdx.x; e < total; e += OP_THREADS) {
        const int r = e / row_floats, c = e - r * row_floats;
        const int dr = s_rank[r];
        if (dr >= 0) dst[(size_t)dr * row_floats + c] = src[(size_t)(row0 + r) * row_floats + c];
    }
}

extern "C" int b2s_adam_chunk(void) { return OP_CHUNK; }

extern "C" int b2s_adam_multi(const B2sAdamTensor *tensors_dev, const int32_t *chunk_tensor_dev,
                              const long long *chunk_start_dev, int n_chunks, b2s_stream_t stream) {
    if (n_chunks < 0 || (n_chunks > 0 && (!tensors_dev || !chunk_tensor_dev || !chunk_start_dev))) return B2S_ERR_ARG;
    if (n_chunks == 0) return B2S_OK;
    k_adam_multi<<<n_chunks, OP_THREADS, 0, (cudaStream_t)stream>>>(tensors_dev, chunk_tensor_dev, chunk_start_dev);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_densify_stats(const float *grad2d, int grad_stride, const int32_t *radii, int N, int W, int H,
                                 float *xys_grad_norm, float *vis_counts, float *max_2dsize, b2s_stream_t stream) {
    if (N < 0 || grad_stride < 2) return B2S_ERR_ARG;
    if (N == 0) return B2S_OK;
    k_densify_stats<<<b2s_div_up(N, OP_THREADS), OP_THREADS, 0, (cudaStream_t)stream>>>(
        grad2d, grad_stride, radii, N, 0.5f * (float)W, 0.5f * (float)H, xys_grad_norm, vis_counts, max_2dsize);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" size_t b2s_mask_workspace_bytes(long long N) {
    return (size_t)((N + OP_THREADS - 1) / OP_THREADS + 1) * sizeof(int32_t) + 256;
}

extern "C" int b2s_mask_scan(const uint8_t *keep, long long N, void *workspace, size_t workspace_bytes, int32_t *total,
                             b2s_stream_t stream) {
    if (N < 0 || !total) return B2S_ERR_ARG;
    if (workspace_bytes < b2s_mask_workspace_bytes(N)) return B2S_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        cudaMemsetAsync(total, 0, sizeof(int32_t), st);
        return B2S_OK;
    }
    const int nblk = (int)((N + OP_THREADS - 1) / OP_THREADS);
    k_mask_block_count<<<nblk, OP_THREADS, 0, st>>>(keep, N, (int32_t *)workspace);
    B2S_LAUNCH_CHECK();
    k_mask_scan_blocks<<<1, 1024, 0, st>>>((int32_t *)workspace, nblk, total);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_mask_gather_rows(const uint8_t *keep, long long N, const void *workspace, const float *src,
                                    float *dst, int row_floats, b2s_stream_t stream) {
    if (N < 0 || row_floats < 1) return B2S_ERR_ARG;
    if (N == 0) return B2S_OK;
    const int nblk = (int)((N + OP_THREADS - 1) / OP_THREADS);
    k_mask_gather_rows<<<nblk, OP_THREADS, 0, (cudaStream_t)stream>>>(keep, N, (const int32_t *)workspace, src, dst,
                                                                     row_floats);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
