// Device-wide scan utilities shared by the binning kernels (sm_100a).
//
// Replaces upstream gsplat v1.4.0  isect_tiles (pass 2) + cub::DeviceRadixSort::SortPairs on 64-bit
// (camera | tile | depth-bits) keys + isect_offset_encode  (kernels K5-K7), as reached from
// mtgs/scene_model/mtgs_scene_graph.py:641-662.
//
// B200-first formulation.  The upstream order is "stable sort of all M intersections by (tile, depth
// bits)", ties in emission order (= ascending Gaussian id).  A stable sort by a composite key equals a
// stable sort by the minor key followed by a stable sort by the major key, and all intersections of one
// Gaussian share its depth, so:
//   1. stable LSD radix sort (3 x 11 bits) of the visible Gaussians by their depth bits (depthsort.cu)
//   2. ordered bucket fill of the M (tile, gaussian) intersections by tile    (tilelists.cu; M items written once)
// gives bit-identical flatten_ids / isect_offsets without ever sorting M 64-bit keys.
// isect_ids (int64) are not needed by the blend; b2s_bin_isect_ids rebuilds them on request.
//
// All kernels are HBM/L2-bound integer work: coalesced 4-byte streams, warp-match ranking, no tensor cores.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// device-wide exclusive scan of int32 (reduce-then-scan, 3 launches)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// block-wide exclusive scan of one int per thread; returns exclusive prefix, total via *total (all threads)
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int *total, int *s_warp /* THREADS/32 + 1 ints */) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < THREADS / 32 ? s_warp[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        if (lane < THREADS / 32) s_warp[lane] = wi - w;
        if (lane == THREADS / 32 - 1) s_warp[THREADS / 32] = wi;
    }
    __syncthreads();
    int res = s_warp[warp] + incl - v;
    *total = s_warp[THREADS / 32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ int scan_load(const int32_t *in, const int32_t *gather, int i, int n) {
    if (i >= n) return 0;
    return gather ? in[gather[i]] : in[i];
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const int32_t *in, const int32_t *__restrict__ gather, int n, int32_t *__restrict__ block_sums) {
    __shared__ int s_warp[SCAN_THREADS / 32 + 1];
    int base = blockIdx.x * SCAN_TILE;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) sum += scan_load(in, gather, base + k * SCAN_THREADS + threadIdx.x, n);
    int total;
    block_excl_scan<SCAN_THREADS>(sum, &total, s_warp);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of block_sums[nb]; grand total (int64) to *total_out if non-null
__global__ void __launch_bounds__(1024) k_scan_top(int32_t *__restrict__ block_sums, int nb, int64_t *__restrict__ total_out) {
    __shared__ int s_warp[1024 / 32 + 1];
    long long carry = 0;
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? block_sums[i] : 0;
        int total;
        int ex = block_excl_scan<1024>(v, &total, s_warp);
        if (i < nb) block_sums[i] = (int)(carry + ex);
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_down(const int32_t *in, const int32_t *__restrict__ gather, int n,
            const int32_t *__restrict__ block_sums, int32_t *out) {  // in may alias out
    __shared__ int s_warp[SCAN_THREADS / 32 + 1];
    // thread-contiguous items so that the per-thread serial scan is in memory order
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = scan_load(in, gather, base + k, n);
        sum += v[k];
    }
    int total;
    int ex = block_excl_scan<SCAN_THREADS>(sum, &total, s_warp) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}

size_t b2s_scan_ws_ints(int n) { return (size_t)b2s_div_up(n > 0 ? n : 1, SCAN_TILE) + 1; }

// exclusive scan of in[gather[i]] (or in[i]) into out; ws needs b2s_scan_ws_ints(n) ints
int b2s_device_excl_scan(const int32_t *in, const int32_t *gather, int n, int32_t *out, int64_t *total_out,
                            int32_t *ws, cudaStream_t st) {
    int nb = b2s_div_up(n, SCAN_TILE);
    k_scan_reduce<<<nb, SCAN_THREADS, 0, st>>>(in, gather, n, ws);
    B2S_LAUNCH_CHECK();
    k_scan_top<<<1, 1024, 0, st>>>(ws, nb, total_out);
    B2S_LAUNCH_CHECK();
    k_scan_down<<<nb, SCAN_THREADS, 0, st>>>(in, gather, n, ws, out);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

