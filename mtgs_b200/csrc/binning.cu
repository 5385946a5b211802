// Tile binning and depth ordering (sm_100a).
//
// Replaces upstream gsplat v1.4.0  isect_tiles (pass 2) + cub::DeviceRadixSort::SortPairs on 64-bit
// (camera | tile | depth-bits) keys + isect_offset_encode  (kernels K5-K7), as reached from
// mtgs/scene_model/mtgs_scene_graph.py:641-662.
//
// B200-first formulation.  The upstream order is "stable sort of all M intersections by (tile, depth
// bits)", ties in emission order (= ascending Gaussian id).  A stable sort by a composite key equals a
// stable sort by the minor key followed by a stable sort by the major key, and all intersections of one
// Gaussian share its depth, so:
//   1. stable LSD radix sort (3 x 11 bits) of the N Gaussians by their depth bits (this file; N items, 8 B each)
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

// ------------------------------------------------------------------------------------------------
// stable LSD radix pass on (uint32 key, uint32 value) pairs, 11-bit digit (3 passes cover 32 bits)
// ------------------------------------------------------------------------------------------------
constexpr int RDX_THREADS = 256;
constexpr int RDX_WARPS = RDX_THREADS / 32;
constexpr int RDX_ROUNDS = 16;                              // items per thread
constexpr int RDX_WARP_ITEMS = 32 * RDX_ROUNDS;             // 512 consecutive items per warp
constexpr int RDX_TILE = RDX_THREADS * RDX_ROUNDS;          // 4096 items per block
constexpr int RDX_BITS = 11;
constexpr int RDX_BINS = 1 << RDX_BITS;                     // 2048
constexpr unsigned RDX_MASK = RDX_BINS - 1;
constexpr size_t RDX_SCATTER_SMEM = (size_t)RDX_WARPS * RDX_BINS * sizeof(int);  // 64 KB of per-warp counters

__global__ void __launch_bounds__(RDX_THREADS)
k_radix_hist(const uint32_t *__restrict__ keys, long long n, int shift, int32_t *__restrict__ table, int nblocks) {
    __shared__ int s_hist[RDX_BINS];
    for (int d = threadIdx.x; d < RDX_BINS; d += RDX_THREADS) s_hist[d] = 0;
    __syncthreads();
    long long base = (long long)blockIdx.x * RDX_TILE;
#pragma unroll 4
    for (int k = 0; k < RDX_ROUNDS; ++k) {
        long long i = base + (long long)k * RDX_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[(keys[i] >> shift) & RDX_MASK], 1);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RDX_BINS; d += RDX_THREADS) table[(size_t)d * nblocks + blockIdx.x] = s_hist[d];
}

// table_scanned[d * nblocks + b] = number of items with digit < d, plus items with digit d in blocks < b.
// Ranking: in every 32-wide round the lanes with equal digits are grouped with match.any; the group leader
// reserves popc(group) slots of the warp's private counter with ONE shared-memory atomic.  Atomics of one warp
// to one address are applied in issue order, so earlier rounds get smaller ranks (stable) and the 16 rounds
// pipeline instead of forming a load->store dependency chain.
__global__ void __launch_bounds__(RDX_THREADS)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, long long n, int shift,
                const int32_t *__restrict__ table_scanned, int nblocks) {
    extern __shared__ int s_cnt[];  // [RDX_WARPS][RDX_BINS]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RDX_WARPS * RDX_BINS; i += RDX_THREADS) s_cnt[i] = 0;
    __syncthreads();
    int *my_cnt = s_cnt + warp * RDX_BINS;

    const long long wbase = (long long)blockIdx.x * RDX_TILE + (long long)warp * RDX_WARP_ITEMS;
    uint32_t key[RDX_ROUNDS];
    int wrank[RDX_ROUNDS];
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < RDX_ROUNDS; ++r) {
        long long i = wbase + r * 32 + lane;
        key[r] = i < n ? keys_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RDX_ROUNDS; ++r) {
        long long i = wbase + r * 32 + lane;
        bool valid = i < n;
        int d = valid ? (int)((key[r] >> shift) & RDX_MASK) : RDX_BINS;  // invalid lanes share a dummy digit
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        int old = 0;
        if (lane == leader && valid) old = atomicAdd(&my_cnt[d], __popc(peers));
        old = __shfl_sync(0xffffffffu, old, leader);
        wrank[r] = old + __popc(peers & lt);
    }
    __syncthreads();
    // turn per-warp counts into global bases (stable: warp 0's items first)
    for (int d = threadIdx.x; d < RDX_BINS; d += RDX_THREADS) {
        int running = table_scanned[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RDX_WARPS; ++w) {
            int c = s_cnt[w * RDX_BINS + d];
            s_cnt[w * RDX_BINS + d] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RDX_ROUNDS; ++r) {
        long long i = wbase + r * 32 + lane;
        if (i < n) {
            int d = (int)((key[r] >> shift) & RDX_MASK);
            int pos = my_cnt[d] + wrank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = vals_in[i];
        }
    }
}

static inline int radix_nblocks(long long n) { return b2s_div_up(n > 0 ? n : 1, RDX_TILE); }
// ints needed: table (RDX_BINS * nblocks) + scan workspace for that table
static inline size_t radix_ws_ints(long long n) {
    size_t nb = (size_t)radix_nblocks(n);
    return nb * RDX_BINS + b2s_scan_ws_ints((int)(nb * RDX_BINS));
}

static int radix_pass(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout, long long n,
                      int shift, int32_t *ws, cudaStream_t st) {
    int nb = radix_nblocks(n);
    int32_t *table = ws;
    int32_t *scan_ws = ws + (size_t)nb * RDX_BINS;
    k_radix_hist<<<nb, RDX_THREADS, 0, st>>>(kin, n, shift, table, nb);
    B2S_LAUNCH_CHECK();
    int rc = b2s_device_excl_scan(table, nullptr, nb * RDX_BINS, table, nullptr, scan_ws, st);
    if (rc != B2S_OK) return rc;
    cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RDX_SCATTER_SMEM);
    k_radix_scatter<<<nb, RDX_THREADS, RDX_SCATTER_SMEM, st>>>(kin, vin, kout, vout, n, shift, table, nb);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

// ------------------------------------------------------------------------------------------------
// step 1: depth order of the Gaussians + exclusive scan of their tile counts in that order
// ------------------------------------------------------------------------------------------------
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t b2s_bin_depth_workspace_bytes(int N) {
    size_t n = (size_t)(N > 0 ? N : 1);
    // key ping, key pong, val pong  + radix table/scan ints + scan ints
    return 3 * align256(n * 4) + align256(radix_ws_ints(N) * 4) + align256(b2s_scan_ws_ints(N) * 4) + 1024;
}

extern "C" int b2s_bin_sort_depth(const uint32_t *sort_keys, const uint32_t *sort_vals,
                                  const int32_t *tiles_per_gauss, int N, int32_t *order, int32_t *cum,
                                  int64_t *total, void *workspace, size_t workspace_bytes, b2s_stream_t stream) {
    if (N < 0) return B2S_ERR_ARG;
    if (workspace_bytes < b2s_bin_depth_workspace_bytes(N)) return B2S_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        cudaMemsetAsync(total, 0, sizeof(int64_t), st);
        return B2S_OK;
    }
    char *w = (char *)workspace;
    size_t n4 = align256((size_t)N * 4);
    uint32_t *kA = (uint32_t *)w; w += n4;
    uint32_t *kB = (uint32_t *)w; w += n4;
    uint32_t *vB = (uint32_t *)w; w += n4;
    int32_t *rws = (int32_t *)w; w += align256(radix_ws_ints(N) * 4);
    int32_t *sws = (int32_t *)w;
    uint32_t *vA = (uint32_t *)order;  // final pass lands here
    int rc;  // 3 passes of 11 bits; the last one lands in (kA, order)
    rc = radix_pass(sort_keys, sort_vals, kA, vA, N, 0, rws, st);             if (rc) return rc;
    rc = radix_pass(kA, vA, kB, vB, N, RDX_BITS, rws, st);                    if (rc) return rc;
    rc = radix_pass(kB, vB, kA, vA, N, 2 * RDX_BITS, rws, st);                if (rc) return rc;
    return b2s_device_excl_scan(tiles_per_gauss, order, N, cum, total, sws, st);
}

