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
// One persistent grid (cudaLaunchCooperativeKernel, all CTAs co-resident, two per SM) runs 4 passes of an 8-bit
// stable LSD radix sort; the whole working set (keys + ids of the visible Gaussians, ~10 MB) lives in the 126 MB L2.
// Per pass and CTA (a CTA owns a contiguous slice of the input):
//     (a) digit histogram of the slice (warp-aggregated shared-memory atomics) -> table[digit][cta]   -> grid barrier
//     (a2) exclusive scan of every digit row over the CTAs (one warp per row, rows spread over the grid) -> grid barrier
//     (b) every CTA reads the global start of ITS run of every digit (256 loads + one block scan over the digit
//         totals), then sub-tile by sub-tile (4096 items): stable ranks from match.any groups and
//         per-warp digit counters, a pass through shared memory that makes each digit's items contiguous, and a
//         scatter whose stores are coalesced runs (average run = 16 items at 8 bits per digit)        -> grid barrier
// The round-1 version used 11-bit digits (3 passes): 2048 counters per 4096-item sub-tile made the cross-warp prefix
// cost more shared-memory operations than there are items, and runs of 2-4 items made the scatter a sector-granular
// 4-byte write (profiles/r01g_kernels_ncu.txt: IPC 0.40, barrier + scoreboard stalls, 167 us).
// Pass 0 reads the projection's keys directly and drops the culled ones (key 0xFFFFFFFF), so the later passes and
// everything downstream only touch n_vis items; the value payload of pass 0 is the index itself.
// Integer work on L2-resident data; no tensor cores.
#include "common.cuh"

constexpr int DS_THREADS = 512;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_ROUNDS = 8;                      // items per thread per sub-tile
constexpr int DS_TILE = DS_THREADS * DS_ROUNDS;   // 4096 items per sub-tile
constexpr int DS_BITS = 8;
constexpr int DS_BINS = 1 << DS_BITS;
constexpr unsigned DS_MASK = DS_BINS - 1;
constexpr int DS_PASSES = 4;
constexpr int DS_MAX_GRID = 512;   // >= 2 x SM count; rows of the digit table are scanned by one warp (16 loads per lane)
constexpr unsigned DS_CULLED = 0xFFFFFFFFu;
// shared memory: reordered keys + values of a sub-tile, per-warp digit counters, per-digit bookkeeping, scan scratch
constexpr size_t DS_SMEM = (size_t)DS_TILE * 8 + (size_t)DS_WARPS * DS_BINS * 4 + (size_t)DS_BINS * 4 * 4 + 256;

__device__ __forceinline__ int ds_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// exclusive scan over the DS_BINS values held by threads 0 .. DS_BINS-1 (one per thread); s_w: DS_BINS / 32 + 1 ints.
// Returns the exclusive prefix for threads < DS_BINS and the grand total through *total.  Contains CTA barriers.
__device__ __forceinline__ int ds_bins_excl_scan(int v, int *total, int *s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = DS_BINS / 32;
    int incl = 0;
    if (warp < NW) {
        incl = ds_warp_incl_scan(v, lane);
        if (lane == 31) s_w[warp] = incl;
    }
    __syncthreads();
    if (warp == 0) {
        int w = lane < NW ? s_w[lane] : 0;
        int wi = ds_warp_incl_scan(w, lane);
        if (lane < NW) s_w[lane] = wi - w;
        if (lane == NW - 1) s_w[NW] = wi;
    }
    __syncthreads();
    const int res = warp < NW ? s_w[warp] + incl - v : 0;
    *total = s_w[NW];
    __syncthreads();
    return res;
}

// Grid-wide barrier of the co-resident grid: one release-add per CTA on a counter that only grows, then an acquire
// spin until the counter reaches `target` (= CTAs x barriers passed so far).  cooperative_groups' grid.sync() was
// measured at ~5-10 us per barrier for 296 CTAs of 512 threads (tools/sort_phases.py: 11 barriers were most of the
// kernel); this one needs one L2 atomic and one polled line per CTA.
__device__ __forceinline__ void ds_grid_barrier(unsigned *counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(DS_THREADS, 2)
k_depth_sort_coop(const uint32_t *__restrict__ keys_in, int N, uint32_t *kA, uint32_t *vA, uint32_t *kB,
                  uint32_t *vB /* == order */, int32_t *table /* [BINS][G] */, int32_t *digit_tot /* [BINS] */,
                  int32_t *nvis_out, unsigned *barrier /* zeroed before the launch */,
                  unsigned long long *phase_ns /* optional diagnostic: [1 + 3 * passes] timestamps */) {
    unsigned n_bar = 0;
    extern __shared__ __align__(16) unsigned char ds_smem_raw[];
    uint32_t *s_k = reinterpret_cast<uint32_t *>(ds_smem_raw);    // [DS_TILE] keys in digit-major order
    uint32_t *s_v = s_k + DS_TILE;                                // [DS_TILE] values
    int *s_cnt = reinterpret_cast<int *>(s_v + DS_TILE);          // [DS_WARPS][DS_BINS]
    int *s_run = s_cnt + DS_WARPS * DS_BINS;                      // [DS_BINS] global cursor of this CTA's run per digit
    int *s_tcnt = s_run + DS_BINS;                                // [DS_BINS] items per digit (all CTAs / this sub-tile)
    int *s_toff = s_tcnt + DS_BINS;                               // [DS_BINS] start of the digit's run inside the sub-tile
    int *s_gdst = s_toff + DS_BINS;                               // [DS_BINS] s_run - s_toff (global = local + this)
    int *s_w = s_gdst + DS_BINS;                                  // scan scratch
    const int G = gridDim.x, b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = lanemask_lt();

    auto stamp = [&](int slot) {
        if (phase_ns != nullptr && b == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            phase_ns[slot] = t;
        }
    };
    stamp(0);
    int n = N;  // items entering the current pass
    for (int pass = 0; pass < DS_PASSES; ++pass) {
        const int shift = DS_BITS * pass;
        // ping-pong: keys_in -> A -> B -> A -> B (= order)
        const uint32_t *ksrc = pass == 0 ? keys_in : ((pass & 1) ? kA : kB);
        const uint32_t *vsrc = pass == 0 ? nullptr : ((pass & 1) ? vA : vB);
        uint32_t *kdst = (pass & 1) ? kB : kA;
        uint32_t *vdst = (pass & 1) ? vB : vA;
        // slices are multiples of 4 items so that 16-byte loads stay aligned
        const int per = (((n + G - 1) / G) + 3) & ~3;
        const int begin = min(n, b * per), end = min(n, begin + per);

        // ---- (a) histogram of this CTA's slice
        for (int d = tid; d < DS_WARPS * DS_BINS; d += DS_THREADS) s_cnt[d] = 0;
        __syncthreads();
        for (int w0 = begin + warp * 128; w0 < end; w0 += DS_THREADS * 4) {  // warp-uniform trip count
            const int i0 = w0 + 4 * lane;
            uint32_t k4[4];
            if (i0 + 3 < end) {
                const uint4 q = *reinterpret_cast<const uint4 *>(ksrc + i0);
                k4[0] = q.x; k4[1] = q.y; k4[2] = q.z; k4[3] = q.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) k4[j] = i0 + j < end ? ksrc[i0 + j] : DS_CULLED;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool valid = i0 + j < end && (pass > 0 || k4[j] != DS_CULLED);
                // plain shared atomics on the warp's private counters: a full same-address conflict costs ~32 cycles
                // per warp instruction, less than grouping the lanes first
                if (valid) atomicAdd(&s_cnt[warp * DS_BINS + ((k4[j] >> shift) & DS_MASK)], 1);
            }
        }
        __syncthreads();
        for (int d = tid; d < DS_BINS; d += DS_THREADS) {
            int c = 0;
#pragma unroll
            for (int w = 0; w < DS_WARPS; ++w) c += s_cnt[w * DS_BINS + d];
            s_tcnt[d] = c;
        }
        __syncthreads();
        for (int d = tid; d < DS_BINS; d += DS_THREADS) table[(size_t)d * G + b] = s_tcnt[d];
        ds_grid_barrier(barrier, ++n_bar * gridDim.x);
        stamp(1 + 3 * pass);

        // ---- (a2) exclusive scan of every digit row over the CTAs, in place (one warp per row, rows spread over the
        // grid: every table entry is touched once -- letting each CTA sum the rows itself costs O(G^2) loads and was
        // measured to be more than half of all instructions of the kernel), digit totals
        // rows are dealt round-robin to the CTAs (row r goes to CTA r % G) so that at full grid size every CTA scans
        // at most one row; the whole row is loaded before the scan starts (independent loads, G <= DS_MAX_GRID = 32 x 32)
        for (int row = b + warp * G; row < DS_BINS; row += G * DS_WARPS) {
            {
                int32_t *r = table + (size_t)row * G;
                int v[DS_MAX_GRID / 32];
#pragma unroll
                for (int j = 0; j < DS_MAX_GRID / 32; ++j) v[j] = (j * 32 + lane < G) ? r[j * 32 + lane] : 0;
                int carry = 0;
#pragma unroll
                for (int j = 0; j < DS_MAX_GRID / 32; ++j) {
                    if (j * 32 < G) {
                        const int incl = ds_warp_incl_scan(v[j], lane);
                        if (j * 32 + lane < G) r[j * 32 + lane] = carry + incl - v[j];
                        carry += __shfl_sync(0xffffffffu, incl, 31);
                    }
                }
                if (lane == 0) digit_tot[row] = carry;
            }
        }
        ds_grid_barrier(barrier, ++n_bar * gridDim.x);
        stamp(2 + 3 * pass);

        // ---- (b1) global start of this CTA's run of every digit = digits before it + this digit in the CTAs before it
        {
            int total;
            const int ex = ds_bins_excl_scan(tid < DS_BINS ? digit_tot[tid] : 0, &total, s_w);
            if (tid < DS_BINS) s_run[tid] = ex + table[(size_t)tid * G + b];
            if (pass == 0) {
                if (b == 0 && tid == 0) *nvis_out = total;
                n = total;  // later passes (and their slices) only see the visible Gaussians
            }
            __syncthreads();
        }

        // ---- (b2) stable ranking + coalesced scatter, sub-tile by sub-tile in slice order
        for (int sub = begin; sub < end; sub += DS_TILE) {
            for (int i = tid; i < DS_WARPS * DS_BINS; i += DS_THREADS) s_cnt[i] = 0;
            __syncthreads();
            int *my_cnt = s_cnt + warp * DS_BINS;
            const int wbase = sub + warp * (32 * DS_ROUNDS);
            uint32_t key[DS_ROUNDS];
            unsigned wrank2[DS_ROUNDS / 2];  // ranks inside the warp's digit group: < 512, two per register (0xffff = invalid)
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
                // lanes holding the same digit: one ballot per digit bit (+1 for validity); match.any takes one
                // iteration per distinct value, i.e. up to 32 for the random low digits
                unsigned peers = __ballot_sync(0xffffffffu, valid);
                if (!valid) peers = ~peers;
#pragma unroll
                for (int bit = 0; bit < DS_BITS; ++bit) {
                    const bool one = (d >> bit) & 1;
                    const unsigned bal = __ballot_sync(0xffffffffu, one);
                    peers &= one ? bal : ~bal;
                }
                const int leader = __ffs(peers) - 1;
                int old = 0;
                // one atomic per (round, digit group); the shuffle below makes the next round wait for it,
                // so earlier rounds get smaller ranks (stable)
                if (lane == leader && valid) old = atomicAdd(&my_cnt[d], __popc(peers));
                old = __shfl_sync(0xffffffffu, old, leader);
                const unsigned rk = valid ? (unsigned)(old + __popc(peers & lt)) : 0xffffu;
                if (r & 1) wrank2[r >> 1] |= rk << 16;
                else wrank2[r >> 1] = rk;
            }
            __syncthreads();
            // per digit: exclusive prefix over the warps (in place), items of the digit in this sub-tile
            if (tid < DS_BINS) {
                int running = 0;
#pragma unroll
                for (int w = 0; w < DS_WARPS; ++w) {
                    const int c = s_cnt[w * DS_BINS + tid];
                    s_cnt[w * DS_BINS + tid] = running;
                    running += c;
                }
                s_tcnt[tid] = running;
            }
            __syncthreads();
            int tile_n;
            const int toff = ds_bins_excl_scan(tid < DS_BINS ? s_tcnt[tid] : 0, &tile_n, s_w);
            if (tid < DS_BINS) {
                s_toff[tid] = toff;
                s_gdst[tid] = s_run[tid] - toff;
                s_run[tid] += s_tcnt[tid];
            }
            __syncthreads();
            // digit-major order inside shared memory
#pragma unroll
            for (int r = 0; r < DS_ROUNDS; ++r) {
                const unsigned rk = (wrank2[r >> 1] >> (16 * (r & 1))) & 0xffffu;
                if (rk != 0xffffu) {
                    const int i = wbase + r * 32 + lane;
                    const int d = (int)((key[r] >> shift) & DS_MASK);
                    const int lp = s_toff[d] + my_cnt[d] + (int)rk;
                    s_k[lp] = key[r];
                    s_v[lp] = pass == 0 ? (uint32_t)i : vsrc[i];
                }
            }
            __syncthreads();
            // coalesced runs out
            for (int i = tid; i < tile_n; i += DS_THREADS) {
                const uint32_t k = s_k[i];
                const int pos = s_gdst[(k >> shift) & DS_MASK] + i;
                kdst[pos] = k;
                vdst[pos] = s_v[i];
            }
            __syncthreads();
        }
        if (pass + 1 < DS_PASSES) ds_grid_barrier(barrier, ++n_bar * gridDim.x);
        stamp(3 + 3 * pass);
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
    if (g > DS_MAX_GRID) g = DS_MAX_GRID;
    if (device >= 0 && device < 64) cached[device] = g;
    return g;
}

extern "C" size_t b2s_bin_depth_workspace_bytes(int N) {
    size_t n = (size_t)(N > 0 ? N : 1);
    // kA, vA, kB (+16 bytes each for vector loads at the slice ends) + table [BINS][G <= DS_MAX_GRID] + digit totals
    return 3 * ds_align256(n * 4 + 16) + ds_align256((size_t)DS_BINS * DS_MAX_GRID * 4) + ds_align256(DS_BINS * 4) + 1024;
}

static int ds_launch(const uint32_t *sort_keys, int N, int32_t *order, int32_t *n_vis, void *workspace,
                     size_t workspace_bytes, unsigned long long *phase_ns, b2s_stream_t stream);

extern "C" int b2s_bin_sort_depth(const uint32_t *sort_keys, int N, int32_t *order, int32_t *n_vis,
                                  void *workspace, size_t workspace_bytes, b2s_stream_t stream) {
    return ds_launch(sort_keys, N, order, n_vis, workspace, workspace_bytes, nullptr, stream);
}

// Diagnostic variant (tools/sort_phases.py): CTA 0 also writes 13 %globaltimer stamps (kernel start, then after the
// histogram barrier, the row-scan barrier and the scatter of each of the 4 passes) to phase_ns (device uint64[13]).
extern "C" int b2s_debug_sort_depth_phases(const uint32_t *sort_keys, int N, int32_t *order, int32_t *n_vis,
                                           void *workspace, size_t workspace_bytes, unsigned long long *phase_ns,
                                           b2s_stream_t stream) {
    return ds_launch(sort_keys, N, order, n_vis, workspace, workspace_bytes, phase_ns, stream);
}

static int ds_launch(const uint32_t *sort_keys, int N, int32_t *order, int32_t *n_vis, void *workspace,
                     size_t workspace_bytes, unsigned long long *phase_ns, b2s_stream_t stream) {
    if (N < 0) return B2S_ERR_ARG;
    if (workspace_bytes < b2s_bin_depth_workspace_bytes(N)) return B2S_ERR_WORKSPACE;
    if (((uintptr_t)sort_keys & 15) || ((uintptr_t)workspace & 15)) return B2S_ERR_ARG;
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
    char *w = (char *)workspace;
    const size_t n4 = ds_align256((size_t)N * 4 + 16);
    uint32_t *kA = (uint32_t *)w; w += n4;
    uint32_t *vA = (uint32_t *)w; w += n4;
    uint32_t *kB = (uint32_t *)w; w += n4;
    int32_t *table = (int32_t *)w; w += ds_align256((size_t)DS_BINS * DS_MAX_GRID * 4);
    int32_t *digit_tot = (int32_t *)w; w += ds_align256(DS_BINS * 4);
    unsigned *barrier = (unsigned *)w;
    cudaMemsetAsync(barrier, 0, sizeof(unsigned), st);
    uint32_t *vB = (uint32_t *)order;
    void *args[] = {(void *)&sort_keys, (void *)&N, (void *)&kA, (void *)&vA, (void *)&kB, (void *)&vB, (void *)&table,
                    (void *)&digit_tot, (void *)&n_vis, (void *)&barrier, (void *)&phase_ns};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_depth_sort_coop, dim3(G), dim3(DS_THREADS), args,
                                                DS_SMEM, st);
    if (e != cudaSuccess) {
        b2s_count_launch(1);
        return -(int)e - 1000;
    }
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
