// Depth order of the visible Gaussians in ONE cooperative kernel (sm_100a).
//
// First half of the two-level replacement of upstream gsplat v1.4.0's isect_tiles pass 2 + 64-bit radix sort of
// all tile intersections + isect_offset_encode (reached from mtgs_scene_graph.py:641-662).
// The upstream order is "stable sort of all M intersections by (tile, depth bits)", ties in emission order
// (= ascending Gaussian id).  A stable sort by a composite key equals a stable sort by the minor key followed by a
// stable sort by the major key, and all intersections of one Gaussian share its depth, so a stable sort of the
// GAUSSIANS by depth bits followed by an order-preserving bucketing of their intersections by tile (tilelists.cu)
// gives bit-identical flatten_ids / isect_offsets without ever sorting M 64-bit keys.  This file produces
//     order[0 .. n_vis)  Gaussian ids in stable depth order (ties: ascending id), culled Gaussians dropped
//     n_vis              (the list sizes of the tile-list hierarchy are summed by the projection kernel, so that their
//                        device->host copy overlaps this sort)
//
// A multi-launch LSD radix sort of 2 M keys spent most of its time in launch/drain gaps between ~15 small
// kernels.  Here a single persistent grid (one launch, cudaLaunchCooperativeKernel, all CTAs co-resident) runs
// 3 passes of an 11-bit stable radix sort separated by grid-wide barriers:
//     histogram of the CTA's slice -> grid.sync -> per-digit exclusive scan over CTAs (one warp per digit row)
//     -> grid.sync -> digit bases (redundantly per CTA) + stable ranking (match.any groups, per-warp counters)
//     + scatter -> grid.sync
// Pass 0 reads the projection's keys directly and drops the culled ones (key 0xFFFFFFFF), so the later passes and
// everything downstream only touch n_vis items; the value payload of pass 0 is the index itself.  The M / S totals
// ride on the same grid.  Integer work on L2-resident data; no tensor cores.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

constexpr int DS_THREADS = 256;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_ROUNDS = 16;                     // items per thread per sub-tile
constexpr int DS_TILE = DS_THREADS * DS_ROUNDS;   // 4096 items per sub-tile
constexpr int DS_BITS = 11;
constexpr int DS_BINS = 1 << DS_BITS;
constexpr unsigned DS_MASK = DS_BINS - 1;
constexpr int DS_PASSES = 3;                      // 33 bits >= 32
constexpr unsigned DS_CULLED = 0xFFFFFFFFu;
// shared memory: per-warp counters [8][2048] (also the histogram) + running digit bases [2048] + scan scratch
constexpr size_t DS_SMEM = ((size_t)DS_WARPS * DS_BINS + DS_BINS + 64) * sizeof(int);

__device__ __forceinline__ int ds_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// block-wide exclusive scan of one int per thread (256 threads); s_w needs DS_WARPS + 1 ints
__device__ __forceinline__ int ds_block_excl_scan(int v, int *total, int *s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = ds_warp_incl_scan(v, lane);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < DS_WARPS ? s_w[lane] : 0;
        int wi = ds_warp_incl_scan(w, lane);
        if (lane < DS_WARPS) s_w[lane] = wi - w;
        if (lane == DS_WARPS - 1) s_w[DS_WARPS] = wi;
    }
    __syncthreads();
    const int res = s_w[warp] + incl - v;
    *total = s_w[DS_WARPS];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(DS_THREADS)
k_depth_sort_coop(const uint32_t *__restrict__ keys_in, int N,
                  uint32_t *kA, uint32_t *vA, uint32_t *kB, uint32_t *vB /* == order */, int32_t *table /* [BINS][G] */,
                  int32_t *digit_tot /* [BINS] */, int32_t *nvis_out) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ int ds_smem[];
    int *s_cnt = ds_smem;                           // [DS_WARPS][DS_BINS]
    int *s_run = ds_smem + DS_WARPS * DS_BINS;      // [DS_BINS]
    int *s_w = s_run + DS_BINS;                     // scan scratch
    const int G = gridDim.x, b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = lanemask_lt();

    int n = N;  // items entering the current pass
    for (int pass = 0; pass < DS_PASSES; ++pass) {
        const int shift = DS_BITS * pass;
        const uint32_t *ksrc = pass == 0 ? keys_in : (pass == 1 ? kB : kA);
        const uint32_t *vsrc = pass == 0 ? nullptr : (pass == 1 ? vB : vA);
        uint32_t *kdst = pass == 1 ? kA : kB;
        uint32_t *vdst = pass == 1 ? vA : vB;
        const int per = (n + G - 1) / G;
        const int begin = min(n, b * per), end = min(n, begin + per);

        // ---- phase 1: histogram of this CTA's slice
        for (int d = tid; d < DS_BINS; d += DS_THREADS) s_cnt[d] = 0;
        __syncthreads();
        for (int i = begin + tid; i < end; i += DS_THREADS) {
            const uint32_t k = ksrc[i];
            if (pass > 0 || k != DS_CULLED) atomicAdd(&s_cnt[(k >> shift) & DS_MASK], 1);
        }
        __syncthreads();
        for (int d = tid; d < DS_BINS; d += DS_THREADS) table[(size_t)d * G + b] = s_cnt[d];
        grid.sync();

        // ---- phase 2: per-digit exclusive scan over the CTAs (one warp per digit row), digit totals
        for (int row = b * DS_WARPS + warp; row < DS_BINS; row += G * DS_WARPS) {
            int32_t *r = table + (size_t)row * G;
            int carry = 0;
            for (int x = 0; x < G; x += 32) {
                const int v = (x + lane < G) ? r[x + lane] : 0;
                const int incl = ds_warp_incl_scan(v, lane);
                if (x + lane < G) r[x + lane] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) digit_tot[row] = carry;
        }
        grid.sync();

        // ---- phase 3: digit bases (every CTA, redundantly) -> running cursor of this CTA per digit
        {
            int loc[DS_BINS / DS_THREADS];
            int sum = 0;
#pragma unroll
            for (int k = 0; k < DS_BINS / DS_THREADS; ++k) {
                loc[k] = digit_tot[tid * (DS_BINS / DS_THREADS) + k];
                sum += loc[k];
            }
            int tot;
            int ex = ds_block_excl_scan(sum, &tot, s_w);
#pragma unroll
            for (int k = 0; k < DS_BINS / DS_THREADS; ++k) {
                const int d = tid * (DS_BINS / DS_THREADS) + k;
                s_run[d] = ex + table[(size_t)d * G + b];
                ex += loc[k];
            }
            if (pass == 0) {
                if (b == 0 && tid == 0) *nvis_out = tot;
                n = tot;  // later passes (and their slices) only see the visible Gaussians
            }
        }
        __syncthreads();
        // stable ranking + scatter, sub-tile by sub-tile in slice order
        for (int sub = begin; sub < end; sub += DS_TILE) {
            for (int i = tid; i < DS_WARPS * DS_BINS; i += DS_THREADS) s_cnt[i] = 0;
            __syncthreads();
            int *my_cnt = s_cnt + warp * DS_BINS;
            const int wbase = sub + warp * (32 * DS_ROUNDS);
            uint32_t key[DS_ROUNDS];
            int wrank[DS_ROUNDS];
#pragma unroll
            for (int r = 0; r < DS_ROUNDS; ++r) {
                const int i = wbase + r * 32 + lane;
                key[r] = i < end ? ksrc[i] : DS_CULLED;
            }
#pragma unroll
            for (int r = 0; r < DS_ROUNDS; ++r) {
                const int i = wbase + r * 32 + lane;
                const bool valid = i < end && (pass > 0 || key[r] != DS_CULLED);
                const int d = valid ? (int)((key[r] >> shift) & DS_MASK) : DS_BINS;  // invalid lanes: dummy digit
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                const int leader = __ffs(peers) - 1;
                int old = 0;
                // one atomic per (round, digit group); the shuffle below makes the next round wait for it,
                // so earlier rounds get smaller ranks (stable)
                if (lane == leader && valid) old = atomicAdd(&my_cnt[d], __popc(peers));
                old = __shfl_sync(0xffffffffu, old, leader);
                wrank[r] = valid ? old + __popc(peers & lt) : -1;
            }
            __syncthreads();
            for (int d = tid; d < DS_BINS; d += DS_THREADS) {
                int running = s_run[d];
#pragma unroll
                for (int w = 0; w < DS_WARPS; ++w) {
                    const int c = s_cnt[w * DS_BINS + d];
                    s_cnt[w * DS_BINS + d] = running;
                    running += c;
                }
                s_run[d] = running;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < DS_ROUNDS; ++r) {
                if (wrank[r] >= 0) {
                    const int i = wbase + r * 32 + lane;
                    const int pos = my_cnt[(key[r] >> shift) & DS_MASK] + wrank[r];
                    kdst[pos] = key[r];
                    vdst[pos] = pass == 0 ? (uint32_t)i : vsrc[i];
                }
            }
            __syncthreads();
        }
        grid.sync();
    }

}

static inline size_t ds_align256(size_t x) { return (x + 255) & ~(size_t)255; }

// co-resident grid size for this device (cached per device; immutable after first use)
static int ds_max_grid(int device) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device] > 0) return cached[device];
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaFuncSetAttribute(k_depth_sort_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DS_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_depth_sort_coop, DS_THREADS, DS_SMEM);
    int g = sms * (per_sm > 0 ? per_sm : 1);
    if (g < 1) g = 1;
    if (device >= 0 && device < 64) cached[device] = g;
    return g;
}

extern "C" size_t b2s_bin_depth_workspace_bytes(int N) {
    size_t n = (size_t)(N > 0 ? N : 1);
    // kA, vA, kB + table [BINS][G<=2048] + digit totals
    return 3 * ds_align256(n * 4) + ds_align256((size_t)DS_BINS * 2048 * 4) + ds_align256(DS_BINS * 4) + 1024;
}

extern "C" int b2s_bin_sort_depth(const uint32_t *sort_keys, int N, int32_t *order, int32_t *n_vis,
                                  void *workspace, size_t workspace_bytes, b2s_stream_t stream) {
    if (N < 0) return B2S_ERR_ARG;
    if (workspace_bytes < b2s_bin_depth_workspace_bytes(N)) return B2S_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        cudaMemsetAsync(n_vis, 0, sizeof(int32_t), st);
        return B2S_OK;
    }
    int device = 0;
    cudaGetDevice(&device);
    int G = ds_max_grid(device);
    const int by_work = b2s_div_up(N, DS_TILE);
    if (G > by_work) G = by_work;
    if (G > 2048) G = 2048;
    char *w = (char *)workspace;
    const size_t n4 = ds_align256((size_t)N * 4);
    uint32_t *kA = (uint32_t *)w; w += n4;
    uint32_t *vA = (uint32_t *)w; w += n4;
    uint32_t *kB = (uint32_t *)w; w += n4;
    int32_t *table = (int32_t *)w; w += ds_align256((size_t)DS_BINS * 2048 * 4);
    int32_t *digit_tot = (int32_t *)w;
    uint32_t *vB = (uint32_t *)order;
    void *args[] = {(void *)&sort_keys, (void *)&N,     (void *)&kA,        (void *)&vA,   (void *)&kB,
                    (void *)&vB,        (void *)&table, (void *)&digit_tot, (void *)&n_vis};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_depth_sort_coop, dim3(G), dim3(DS_THREADS), args,
                                                DS_SMEM, st);
    if (e != cudaSuccess) {
        b2s_count_launch(1);
        return -(int)e - 1000;
    }
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
