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
//   1. stable LSD radix sort of the N Gaussians by their 32 depth bits        (N items, 8 B each)
//   2. emit (tile, gaussian) pairs walking Gaussians in that order            (M items, written once)
//   3. stable LSD radix sort of the M pairs by tile id only (<= 16 bits -> 2 passes instead of 6)
// gives bit-identical flatten_ids / isect_offsets while moving ~1/3 of the bytes of the 64-bit sort.
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

static inline size_t scan_ws_ints(int n) { return (size_t)b2s_div_up(n > 0 ? n : 1, SCAN_TILE) + 1; }

// exclusive scan of in[gather[i]] (or in[i]) into out; ws needs scan_ws_ints(n) ints
static int device_excl_scan(const int32_t *in, const int32_t *gather, int n, int32_t *out, int64_t *total_out,
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
// stable LSD radix pass on (uint32 key, uint32 value) pairs, 8-bit digit
// ------------------------------------------------------------------------------------------------
constexpr int RDX_THREADS = 256;
constexpr int RDX_WARPS = RDX_THREADS / 32;
constexpr int RDX_ROUNDS = 16;                              // items per thread
constexpr int RDX_WARP_ITEMS = 32 * RDX_ROUNDS;             // 512 consecutive items per warp
constexpr int RDX_TILE = RDX_THREADS * RDX_ROUNDS;          // 4096 items per block
constexpr int RDX_BINS = 256;

__global__ void __launch_bounds__(RDX_THREADS)
k_radix_hist(const uint32_t *__restrict__ keys, long long n, int shift, int32_t *__restrict__ table, int nblocks) {
    __shared__ int s_hist[RDX_BINS];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    long long base = (long long)blockIdx.x * RDX_TILE;
#pragma unroll 4
    for (int k = 0; k < RDX_ROUNDS; ++k) {
        long long i = base + (long long)k * RDX_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[(keys[i] >> shift) & 0xFFu], 1);
    }
    __syncthreads();
    table[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_hist[threadIdx.x];
}

// table_scanned[d * nblocks + b] = number of items with digit < d, plus items with digit d in blocks < b.
__global__ void __launch_bounds__(RDX_THREADS)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, long long n, int shift,
                const int32_t *__restrict__ table_scanned, int nblocks) {
    __shared__ int s_cnt[RDX_WARPS][RDX_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RDX_WARPS * RDX_BINS; i += RDX_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();

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
        int d = valid ? (int)((key[r] >> shift) & 0xFFu) : RDX_BINS;  // invalid lanes share a dummy digit
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        int old = 0;
        if (lane == leader && valid) {
            old = s_cnt[warp][d];
            s_cnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        wrank[r] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: turn per-warp counts into global bases (stable: warp 0's items first)
        int d = threadIdx.x;
        int running = table_scanned[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RDX_WARPS; ++w) {
            int c = s_cnt[w][d];
            s_cnt[w][d] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RDX_ROUNDS; ++r) {
        long long i = wbase + r * 32 + lane;
        if (i < n) {
            int d = (int)((key[r] >> shift) & 0xFFu);
            int pos = s_cnt[warp][d] + wrank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = vals_in[i];
        }
    }
}

static inline int radix_nblocks(long long n) { return b2s_div_up(n > 0 ? n : 1, RDX_TILE); }
// ints needed: table (256 * nblocks) + scan workspace for that table
static inline size_t radix_ws_ints(long long n) {
    size_t nb = (size_t)radix_nblocks(n);
    return nb * RDX_BINS + scan_ws_ints((int)(nb * RDX_BINS));
}

static int radix_pass(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout, long long n,
                      int shift, int32_t *ws, cudaStream_t st) {
    int nb = radix_nblocks(n);
    int32_t *table = ws;
    int32_t *scan_ws = ws + (size_t)nb * RDX_BINS;
    k_radix_hist<<<nb, RDX_THREADS, 0, st>>>(kin, n, shift, table, nb);
    B2S_LAUNCH_CHECK();
    int rc = device_excl_scan(table, nullptr, nb * RDX_BINS, table, nullptr, scan_ws, st);
    if (rc != B2S_OK) return rc;
    k_radix_scatter<<<nb, RDX_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, table, nb);
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
    return 3 * align256(n * 4) + align256(radix_ws_ints(N) * 4) + align256(scan_ws_ints(N) * 4) + 1024;
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
    // 4 passes: (keys, vals) -> (kA, vA) -> (kB, vB) -> (kA, vA) -> (kB, vB)?  arrange so pass 4 writes `order`
    int rc;
    rc = radix_pass(sort_keys, sort_vals, kB, vB, N, 0, rws, st);  if (rc) return rc;
    rc = radix_pass(kB, vB, kA, vA, N, 8, rws, st);                if (rc) return rc;
    rc = radix_pass(kA, vA, kB, vB, N, 16, rws, st);               if (rc) return rc;
    rc = radix_pass(kB, vB, kA, vA, N, 24, rws, st);               if (rc) return rc;
    return device_excl_scan(tiles_per_gauss, order, N, cum, total, sws, st);
}

// ------------------------------------------------------------------------------------------------
// step 2: emit (tile, gaussian) in depth order, stable sort by tile, offsets
// ------------------------------------------------------------------------------------------------
// one warp per 32 consecutive depth-ordered Gaussians; lanes stride over the tiles of each rectangle so that
// the writes of one Gaussian are contiguous (upstream isect_tiles pass 2 emission order: row-major).
__global__ void __launch_bounds__(256)
k_emit(const float2 *__restrict__ means2d, const int32_t *__restrict__ radii, const int32_t *__restrict__ order,
       const int32_t *__restrict__ cum, int N, int tile_w, int tile_h, uint32_t *__restrict__ tile_out,
       uint32_t *__restrict__ gid_out) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int i = gwarp * 32 + lane;
    int g = -1, x0 = 0, y0 = 0, wdt = 0, cnt = 0, start = 0;
    if (i < N) {
        g = order[i];
        int r = radii[g];
        if (r > 0) {
            float2 m = means2d[g];
            float tr = __fdiv_rn((float)r, 16.0f);
            float txc = __fdiv_rn(m.x, 16.0f), tyc = __fdiv_rn(m.y, 16.0f);
            float fx0 = floorf(__fsub_rn(txc, tr)), fy0 = floorf(__fsub_rn(tyc, tr));
            float fx1 = ceilf(__fadd_rn(txc, tr)), fy1 = ceilf(__fadd_rn(tyc, tr));
            x0 = fx0 <= 0.f ? 0 : (fx0 >= (float)tile_w ? tile_w : (int)fx0);
            y0 = fy0 <= 0.f ? 0 : (fy0 >= (float)tile_h ? tile_h : (int)fy0);
            int x1 = fx1 <= 0.f ? 0 : (fx1 >= (float)tile_w ? tile_w : (int)fx1);
            int y1 = fy1 <= 0.f ? 0 : (fy1 >= (float)tile_h ? tile_h : (int)fy1);
            wdt = x1 - x0;
            cnt = (y1 - y0) * wdt;
            start = cum[i];
        }
    }
    unsigned any = __ballot_sync(0xffffffffu, cnt > 0);
    while (any) {
        int src = __ffs(any) - 1;
        any &= any - 1;
        int c = __shfl_sync(0xffffffffu, cnt, src);
        int sx0 = __shfl_sync(0xffffffffu, x0, src);
        int sy0 = __shfl_sync(0xffffffffu, y0, src);
        int sw = __shfl_sync(0xffffffffu, wdt, src);
        int sst = __shfl_sync(0xffffffffu, start, src);
        int sg = __shfl_sync(0xffffffffu, g, src);
        for (int k = lane; k < c; k += 32) {
            int ty = sy0 + k / sw, tx = sx0 + k % sw;
            tile_out[sst + k] = (uint32_t)(ty * tile_w + tx);
            gid_out[sst + k] = (uint32_t)sg;
        }
    }
}

// upstream isect_offset_encode on the sorted tile keys
__global__ void __launch_bounds__(256)
k_offsets(const uint32_t *__restrict__ tile_keys, long long M, int n_tiles, int32_t *__restrict__ offsets) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    int cur = (int)tile_keys[idx];
    if (idx == 0) {
        for (int t = 0; t <= cur; ++t) offsets[t] = 0;
    } else {
        int prev = (int)tile_keys[idx - 1];
        for (int t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)idx;
    }
    if (idx == M - 1) {
        for (int t = cur + 1; t < n_tiles; ++t) offsets[t] = (int32_t)M;
    }
}

static inline int tile_passes(int n_tiles) {
    int bits = 0;
    while ((1LL << bits) < (long long)n_tiles) ++bits;  // ids in [0, n_tiles)
    return bits <= 8 ? 1 : (bits <= 16 ? 2 : (bits <= 24 ? 3 : 4));
}

extern "C" size_t b2s_bin_tiles_workspace_bytes(int N, long long M) {
    (void)N;
    size_t m = (size_t)(M > 0 ? M : 1);
    // tile ping, gid ping, gid pong (+ tile pong is the caller's tile_keys) + radix ints
    return 3 * align256(m * 4) + align256(radix_ws_ints(M) * 4) + 1024;
}

extern "C" int b2s_bin_tiles(const float *means2d, const int32_t *radii, const int32_t *order,
                             const int32_t *cum, int N, long long M, int tile_size, int tile_w, int tile_h,
                             int32_t *flatten_ids, uint32_t *tile_keys, int32_t *isect_offsets, void *workspace,
                             size_t workspace_bytes, b2s_stream_t stream) {
    if (N < 0 || M < 0 || M >= (1LL << 31)) return B2S_ERR_ARG;
    if (tile_size != 16) return B2S_ERR_UNSUPPORTED;
    if (workspace_bytes < b2s_bin_tiles_workspace_bytes(N, M)) return B2S_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int n_tiles = tile_w * tile_h;
    if (M == 0 || N == 0) {
        cudaMemsetAsync(isect_offsets, 0, sizeof(int32_t) * (size_t)n_tiles, st);
        return B2S_OK;
    }
    char *w = (char *)workspace;
    size_t m4 = align256((size_t)M * 4);
    uint32_t *tA = (uint32_t *)w; w += m4;
    uint32_t *gA = (uint32_t *)w; w += m4;
    uint32_t *gB = (uint32_t *)w; w += m4;
    int32_t *rws = (int32_t *)w;
    uint32_t *tOut = tile_keys;
    uint32_t *gOut = (uint32_t *)flatten_ids;
    int passes = tile_passes(n_tiles);
    // choose the emit target so that the last pass writes (tile_keys, flatten_ids)
    //   1 pass : emit -> (tA,gA) ; pass0 -> out
    //   2 pass : emit -> (tA,gA) ; pass0 -> (tOut?,..)   we need ping-pong: A -> B -> A ...
    // buffers: A = (tA, gA), B = (tOut, gB) for intermediate; final must be (tOut, gOut).
    // odd #passes : emit->A, A->OUT                      (1) ; emit->A, A->B', B'->A, A->OUT (3, B'=(tOut,gB))
    // even #passes: emit->B', B'->A, A->OUT              (2) ; ...
    uint32_t *tk[2] = {tA, tOut};
    uint32_t *gk[2] = {gA, gB};
    int cur = (passes % 2 == 1) ? 0 : 1;
    k_emit<<<b2s_div_up(N, 256), 256, 0, st>>>((const float2 *)means2d, radii, order, cum, N, tile_w, tile_h,
                                                   tk[cur], gk[cur]);
    B2S_LAUNCH_CHECK();
    for (int p = 0; p < passes; ++p) {
        int nxt = cur ^ 1;
        bool last = (p == passes - 1);
        uint32_t *to = last ? tOut : tk[nxt];
        uint32_t *go = last ? gOut : gk[nxt];
        int rc = radix_pass(tk[cur], gk[cur], to, go, M, 8 * p, rws, st);
        if (rc) return rc;
        cur = nxt;
    }
    k_offsets<<<b2s_div_up(M, 256), 256, 0, st>>>(tile_keys, M, n_tiles, isect_offsets);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

__global__ void __launch_bounds__(256)
k_isect_ids(const uint32_t *__restrict__ tile_keys, const int32_t *__restrict__ flatten_ids,
            const float *__restrict__ depths, long long M, int64_t *__restrict__ isect_ids) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    // upstream: (cam << (32 + tile_bits)) | (tile << 32) | (int64)(int32 view of depth); cam = 0
    int32_t dbits = __float_as_int(depths[flatten_ids[i]]);
    isect_ids[i] = ((int64_t)tile_keys[i] << 32) | (int64_t)dbits;
}

extern "C" int b2s_bin_isect_ids(const uint32_t *tile_keys, const int32_t *flatten_ids, const float *depths,
                                 long long M, int64_t *isect_ids, b2s_stream_t stream) {
    if (M < 0) return B2S_ERR_ARG;
    if (M == 0) return B2S_OK;
    k_isect_ids<<<b2s_div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(tile_keys, flatten_ids, depths, M, isect_ids);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
