// Multi-GPU exchange of the shared-node gradients, fused with the projection backward (sm_100a, NVLink peer memory).
//
// Scope: SURVEY.md 8e.  MTGS renders one camera per step (mtgs/scene_model/mtgs_scene_graph.py:548); with one
// traversal camera per GPU over replicated shared nodes, the only exchange step of the path is the sum of the
// per-rank gradients of the rasterizer inputs (means 3 + quats 4 + scales 3 + opacity 1 + colours d_in floats per
// shared Gaussian).  The reference has no such data path (its DDP hook is unused); a library all-reduce
// after the backward is the baseline this file replaces:
//
//   K1  k_project_bwd<.., EXCH = true>  (project.cu)   every rank computes its partial gradient rows and stores them
//       straight into the staging slot [owner][source] of the rank that owns the rows (16-byte coalesced peer stores)
//       -- the reduce-scatter traffic overlaps the projection backward's arithmetic; no local partial-gradient
//       buffer is written, zero-filled or re-read.  A one-warp kernel then raises flag[phase 0][rank] at every peer
//       (the kernel boundary has completed the peer stores, so no CTA pays a system fence).
//   K2  k_grad_reduce_bcast             every rank waits for all phase-0 flags, sums the `world` partial slots of ITS
//       rows from local HBM, scales (1 / world for the mean) and stores the result into EVERY rank's gradient arena
//       (peer stores: the all-gather half).
//   K3  k_exchange_signal + k_exchange_wait   raise flag[phase 1][rank] at every peer, then wait for all phase-1
//       flags in stream order: afterwards this rank's arena holds
//       the reduced gradient of every shared row and the stream continues (activations' VJPs, optimizer).
//
// Rows at or beyond n_shared (rank-local nodes, e.g. the vehicles of this rank's traversal; reference
// rigid_node.py:87, 259-261) are written to the local arena by the plain kernel and never leave the GPU.
// Flags carry a monotonically increasing epoch, so nothing is reset between steps.  The spin loops block like a library
// collective would (ranks must stay in lock-step); they give up only after the caller's timeout (default 120 s) so that a
// dead peer cannot hang the GPU for ever.  A timeout is LOUD: the status word (which may live in pinned host memory, so
// the host reads it without synchronising) becomes non-zero and the reduced rows of this rank's arena are filled with
// NaN, so an optimizer can never silently step on stale or partial gradients.
#include <cstring>

#include "common.cuh"

int b2s_launch_project_bwd_exchange(const float *means, const float *quats, const float *scales, const float *opacities,
                                    const float *viewmat, const float *K, int n_rows, int W, int H, float eps2d,
                                    int calc_comp, int d_in, int with_depth, int cdim, const int32_t *radii,
                                    const float *geo, const float *comps, const float *v_means2d, int v_means2d_stride,
                                    const float *v_geo, const float *v_colpack, float *v_viewmat, const B2sExchange &ex,
                                    cudaStream_t st);

// spin until *flag == epoch; false on timeout
__device__ __forceinline__ bool ex_wait_flag(const volatile unsigned *flag, unsigned epoch, long long timeout_cycles) {
    const long long t0 = clock64();
    while (*flag != epoch) {
        if (clock64() - t0 > timeout_cycles) return false;
        __nanosleep(200);
    }
    return true;
}

// NaN-fill of the shared rows of this rank's arena (timeout path only)
__device__ void ex_poison_arena(const B2sExchange &ex, int n_shared, long long rows_cap) {
    const float nan = __int_as_float(0x7fc00000);
    float *a = ex.arena[ex.rank];
    const int widths[5] = {3, 4, 3, 1, ex.d_col};
    long long blk = 0;
    for (int p = 0; p < 5; ++p) {
        const long long n = (long long)widths[p] * n_shared;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            a[blk * rows_cap + i] = nan;
        blk += widths[p];
    }
}

__global__ void __launch_bounds__(256)
k_grad_reduce_bcast(const B2sExchange ex, int n_shared, long long rows_cap, float scale,
                    unsigned *__restrict__ status) {
    __shared__ unsigned s_ok;
    if (threadIdx.x == 0) s_ok = 1u;
    __syncthreads();
    if (threadIdx.x < ex.world) {
        const unsigned epoch = ex.epoch_dev != nullptr ? *(volatile unsigned *)ex.epoch_dev : ex.epoch;
        if (!ex_wait_flag(ex.flags[ex.rank] + threadIdx.x, epoch, ex.timeout_cycles)) {
            s_ok = 0u;
            *(volatile unsigned *)status = 1u;
            __threadfence_system();
        }
    }
    __syncthreads();
    __threadfence();
    if (!s_ok) {
        ex_poison_arena(ex, n_shared, rows_cap);
    } else {
        const long long n4 = ex.slot_floats >> 2;
        const float *mine = ex.stage[ex.rank];
        const long long b1 = 3LL * ex.shard, b2 = 7LL * ex.shard, b3 = 10LL * ex.shard, b4 = 11LL * ex.shard;
        for (long long e4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; e4 < n4;
             e4 += (long long)gridDim.x * blockDim.x) {
            const long long e = e4 << 2;
            // block p of the slot (means | quats | scales | opacity | colours) and the float index inside it
            int a_p = 0, wp = 3;
            long long b0 = 0;
            if (e >= b4) { a_p = 11; wp = ex.d_col; b0 = b4; }
            else if (e >= b3) { a_p = 10; wp = 1; b0 = b3; }
            else if (e >= b2) { a_p = 7; wp = 3; b0 = b2; }
            else if (e >= b1) { a_p = 3; wp = 4; b0 = b1; }
            // all the sources' loads in flight before the first add (same summation order as a plain loop)
            float4 v[B2S_MAX_WORLD];
#pragma unroll
            for (int src = 0; src < B2S_MAX_WORLD; ++src)
                v[src] = src < ex.world ? __ldcs(reinterpret_cast<const float4 *>(mine + (size_t)src * ex.slot_floats + e))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int src = 0; src < B2S_MAX_WORLD; ++src)
                if (src < ex.world) { s.x += v[src].x; s.y += v[src].y; s.z += v[src].z; s.w += v[src].w; }
            s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
            const long long gi = (long long)wp * ex.rank * ex.shard + (e - b0);  // float index inside arena block p
            const long long limit = (long long)wp * n_shared;
            const long long dst = (long long)a_p * rows_cap + gi;
            if (gi + 3 < limit) {
                for (int r = 0; r < ex.world; ++r) *reinterpret_cast<float4 *>(ex.arena[r] + dst) = s;
            } else if (gi < limit) {
                const float vals[4] = {s.x, s.y, s.z, s.w};
                for (int r = 0; r < ex.world; ++r)
                    for (int k = 0; k < 4 && gi + k < limit; ++k) ex.arena[r][dst + k] = vals[k];
            }
        }
    }
}

// One warp: raise this rank's flag of `phase` at every rank.  Launched right after the kernel whose peer stores it
// publishes: the kernel boundary has completed those stores system-wide, so no CTA of the producer ever fences.
__global__ void __launch_bounds__(32)
k_exchange_signal(const B2sExchange ex, int phase) {
    __threadfence_system();
    unsigned epoch = ex.epoch;
    if (ex.epoch_dev != nullptr) {  // device-counted steps: the phase-0 signal opens a new step
        if (phase == 0 && threadIdx.x == 0) *(volatile unsigned *)ex.epoch_dev = *(volatile unsigned *)ex.epoch_dev + 1u;
        __syncwarp();
        epoch = *(volatile unsigned *)ex.epoch_dev;
    }
    if (threadIdx.x < ex.world) {
        volatile unsigned *f = ex.flags[threadIdx.x] + phase * B2S_MAX_WORLD + ex.rank;
        *f = epoch;
    }
}

// One warp: wait until every rank has published "my reduced rows are in every arena" (phase 1).
__global__ void __launch_bounds__(256)
k_exchange_wait(const B2sExchange ex, int n_shared, long long rows_cap, unsigned *__restrict__ status) {
    __shared__ unsigned s_ok;
    if (threadIdx.x == 0) s_ok = 1u;
    __syncthreads();
    if (threadIdx.x < ex.world) {
        const unsigned epoch = ex.epoch_dev != nullptr ? *(volatile unsigned *)ex.epoch_dev : ex.epoch;
        if (!ex_wait_flag(ex.flags[ex.rank] + B2S_MAX_WORLD + threadIdx.x, epoch, ex.timeout_cycles)) {
            s_ok = 0u;
            *(volatile unsigned *)status = 2u;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (!s_ok) ex_poison_arena(ex, n_shared, rows_cap);  // a peer's rows never arrived: nothing in the arena is trustworthy
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int b2s_exchange_shard_rows(int n_shared, int world) {
    if (n_shared < 0 || world < 1 || world > B2S_MAX_WORLD) return B2S_ERR_ARG;
    const long long per = ((long long)n_shared + world - 1) / world;
    const long long s = (per + 255) / 256 * 256;
    return (int)(s > 0 ? s : 256);
}

// Peer-shared buffers must be plain cudaMalloc allocations to be exportable through CUDA IPC, so these two calls are
// the one place where the library allocates device memory (set-up time only, never on the hot path).
extern "C" int b2s_peer_alloc(size_t bytes, void **dev_ptr) {
    if (!dev_ptr || bytes == 0) return B2S_ERR_ARG;
    cudaError_t e = cudaMalloc(dev_ptr, bytes);
    if (e != cudaSuccess) return -(int)e - 1000;
    e = cudaMemset(*dev_ptr, 0, bytes);
    return e == cudaSuccess ? B2S_OK : -(int)e - 1000;
}
extern "C" int b2s_peer_free(void *dev_ptr) {
    cudaError_t e = cudaFree(dev_ptr);
    return e == cudaSuccess ? B2S_OK : -(int)e - 1000;
}
extern "C" int b2s_ipc_export(void *dev_ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr);
    if (e != cudaSuccess) return -(int)e - 1000;
    memcpy(handle_out, &h, 64);
    return B2S_OK;
}
extern "C" int b2s_ipc_import(const unsigned char handle[64], void **peer_ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? B2S_OK : -(int)e - 1000;
}
extern "C" int b2s_ipc_close(void *peer_ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
    return e == cudaSuccess ? B2S_OK : -(int)e - 1000;
}

extern "C" int b2s_project_bwd_exchange(
    const float *means, const float *quats, const float *scales, const float *opacities, const float *viewmat,
    const float *K, int N, int W, int H, float eps2d, int calc_comp, int d_in, int with_depth, int cdim,
    const int32_t *radii, const float *geo, const float *comps, const float *v_means2d, int v_means2d_stride,
    const float *v_geo, const float *v_colpack, float *v_viewmat, int n_shared, int exchange_colors, int world, int rank,
    long long rows_cap, float scale, unsigned epoch, int phases, float timeout_s,
    const unsigned long long *stage_ptrs_host, const unsigned long long *arena_ptrs_host,
    const unsigned long long *flag_ptrs_host, unsigned *status, b2s_stream_t stream) {
    if (N < 0 || n_shared < 0 || n_shared > N || world < 1 || world > B2S_MAX_WORLD || rank < 0 || rank >= world)
        return B2S_ERR_ARG;
    if (rows_cap < N || (rows_cap & 3) || d_in < 0 || d_in > 8) return B2S_ERR_ARG;
    if (cdim != 4 && cdim != 8) return B2S_ERR_UNSUPPORTED;
    if (calc_comp && comps == nullptr) return B2S_ERR_ARG;
    if (v_means2d_stride < 2 || (v_means2d_stride & 1)) return B2S_ERR_ARG;
    if (!stage_ptrs_host || !arena_ptrs_host || !flag_ptrs_host || !status) return B2S_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    B2sExchange ex = {};
    ex.world = world;
    ex.rank = rank;
    ex.shard = b2s_exchange_shard_rows(n_shared, world);
    ex.d_col = exchange_colors ? d_in : 0;
    ex.slot_floats = (long long)(11 + ex.d_col) * ex.shard;
    ex.cta_rot = ((rank + 1) % world) * (ex.shard / 256);
    ex.epoch = epoch;
    // epoch == 0: device-counted steps; the counter is word 32 of this rank's own flag buffer (peers never touch it)
    ex.epoch_dev = epoch == 0 ? (unsigned *)(uintptr_t)flag_ptrs_host[rank] + 32 : nullptr;
    ex.timeout_cycles = (long long)((timeout_s > 0.f ? (double)timeout_s : 120.0) * 1.9e9);
    for (int r = 0; r < world; ++r) {
        ex.stage[r] = (float *)(uintptr_t)stage_ptrs_host[r];
        ex.arena[r] = (float *)(uintptr_t)arena_ptrs_host[r];
        ex.flags[r] = (unsigned *)(uintptr_t)flag_ptrs_host[r];
        if (!ex.stage[r] || !ex.arena[r] || !ex.flags[r]) return B2S_ERR_ARG;
    }
    float *arena = ex.arena[rank];
    int rc;
    if (n_shared > 0 && (phases & 1)) {
        rc = b2s_launch_project_bwd_exchange(means, quats, scales, opacities, viewmat, K, n_shared, W, H, eps2d, calc_comp,
                                             d_in, with_depth, cdim, radii, geo, comps, v_means2d, v_means2d_stride,
                                             v_geo, v_colpack, v_viewmat, ex, st);
        if (rc != B2S_OK) return rc;
    }
    if (N > n_shared && (phases & 1)) {  // rank-local rows: plain kernel, written behind the shared rows of the local arena
        const size_t o = (size_t)n_shared;
        rc = b2s_project_bwd(means + 3 * o, quats + 4 * o, scales + 3 * o, opacities + o, viewmat, K, N - n_shared, W, H,
                             eps2d, calc_comp, d_in, with_depth, cdim, radii + o, geo + 4 * o,
                             comps ? comps + o : nullptr, v_means2d + (size_t)v_means2d_stride * o, v_means2d_stride,
                             v_geo + 4 * o, v_colpack + (size_t)cdim * o, arena + 3 * o, arena + 3 * rows_cap + 4 * o,
                             arena + 7 * rows_cap + 3 * o, arena + 10 * rows_cap + o, nullptr, v_viewmat, stream);
        if (rc != B2S_OK) return rc;
    }
    if (n_shared > 0 && (phases & 1)) {
        k_exchange_signal<<<1, 32, 0, st>>>(ex, 0);
        B2S_LAUNCH_CHECK();
    }
    if (n_shared > 0 && (phases & 2)) {
        int device = 0, sms = 148;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        k_grad_reduce_bcast<<<sms * 4, 256, 0, st>>>(ex, n_shared, rows_cap, scale, status);
        B2S_LAUNCH_CHECK();
    }
    if (n_shared > 0 && (phases & 4)) {
        k_exchange_signal<<<1, 32, 0, st>>>(ex, 1);
        B2S_LAUNCH_CHECK();
    }
    if (n_shared > 0 && (phases & 8)) {
        k_exchange_wait<<<1, 256, 0, st>>>(ex, n_shared, rows_cap, status);
        B2S_LAUNCH_CHECK();
    }
    return B2S_OK;
}
