// Per-tile depth-ordered Gaussian lists without sorting the intersections (sm_100a).
//
// Replaces, together with binning.cu's depth sort, upstream gsplat v1.4.0 isect_tiles pass 2 + the 64-bit
// cub radix sort + isect_offset_encode reached from mtgs_scene_graph.py:641-662.
//
// Input: Gaussians in stable depth order (order[], from b2s_bin_sort_depth), the exclusive scan cum[] of
// their tile counts in that order, and each Gaussian's tile rectangle (tile_rects, from the projection kernel).
// The (Gaussian x tile) incidence is a sparse matrix given row by row (row = rectangle of one Gaussian);
// upstream's sorted list is its column-major transpose with rows kept in order.  We build it directly with
// "tile-owner lanes":
//   * the depth-ordered Gaussian stream is cut into R chunks of equal cost (k_chunk_bounds);
//   * the tile grid is cut into bands of <= 24 tile rows x 128 tile columns; a CTA handles one (chunk, band):
//     warp w owns tile row w of the band, lane l owns the tiles (row, l + 32 k), k = 0..3, and keeps their list
//     cursors in REGISTERS -- no shared-memory counters, no atomics, no ranking;
//   * the CTA streams its chunk through shared memory in batches (one coalesced gather of the 8-byte
//     rectangles per batch, software-pipelined); every warp ballots the batch entries that cover its row and,
//     for each hit, the lanes whose column lies in [x0, x1) append the Gaussian id at their own cursor.
//     Entries are visited in stream order, so every tile list is in depth order (ties: ascending id):
//     bit-identical to the stable global sort.
//   * pass 1 (k_tile_count) gets count[chunk][tile] from a 2-D difference array (4 corner updates per Gaussian,
//     independent of its size); k_tile_prefix turns it into per-chunk bases and per-tile totals; their
//     exclusive scan IS isect_offsets; pass 2 (k_tile_fill) writes flatten_ids.
// Work ~ (rows covered by the rectangles) ~ M / mean width instead of M; M x 4 B written once.
// Integer work, issue-bound at full occupancy; no tensor cores.
#include <cstdlib>

#include "common.cuh"

constexpr int TR_MAX_ROWS = 24;  // tile rows per band = warps per CTA
constexpr int TR_NG = 4;         // 32-tile column groups per lane -> 128 tile columns per band
constexpr int TR_COST_C = 16;    // chunk cost = tiles + TR_COST_C per Gaussian
constexpr int TR_TARGET_CTAS = 148 * 12;  // chunks per band: rows near the horizon carry most hits, so many short chunks

// f(i) = cum[i] + c * i is the cost of the stream before Gaussian i; first i in [0, nv) with f(i) >= target, else nv
__device__ __forceinline__ int lower_bound_cost(const int32_t *__restrict__ cum, int nv, long long c, long long target) {
    int lo = 0, hi = nv;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((long long)cum[mid] + c * mid < target) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// bounds[r] (r = 0..R) = first Gaussian of chunk r in the depth-ordered stream; bounds[R] = n_vis.
__global__ void __launch_bounds__(1024)
k_chunk_bounds(const int32_t *__restrict__ cum, const int32_t *__restrict__ n_vis, long long M, int R,
               int32_t *__restrict__ bounds) {
    const int nv = *n_vis;  // Gaussians that own intersections (depthsort.cu drops the culled ones)
    const long long C = M + (long long)TR_COST_C * nv;
    for (int r = threadIdx.x; r <= R; r += blockDim.x)
        bounds[r] = (r == R) ? nv : lower_bound_cost(cum, nv, TR_COST_C, ((long long)r * C + R - 1) / R);
}

// ------------------------------------------------------------------------------------------------
// pass 1: count[chunk][tile] for one (chunk, band) through a 2-D difference array in shared memory:
// every thread adds ONE rectangle (4 corner updates), independent of how many tiles it covers.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * TR_MAX_ROWS)
k_tile_count(const int2 *__restrict__ rects, const int32_t *__restrict__ order, const int32_t *__restrict__ bounds,
             int tile_w, int tile_h, int rows_per_band, int row_bands, int R, int32_t *__restrict__ table /* [R][T] */) {
    constexpr int CW = 32 * TR_NG + 1;  // difference array width (columns of the band + 1)
    extern __shared__ int s_d[];        // (rows_per_band + 1) x CW
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x % R;
    const int band = blockIdx.x / R;
    const int band_y = band % row_bands, band_x = band / row_bands;
    const int T = tile_w * tile_h;
    const int by0 = band_y * rows_per_band, by1 = min(tile_h, by0 + rows_per_band);
    const int bx0 = band_x * (32 * TR_NG), bx1 = min(tile_w, bx0 + 32 * TR_NG);
    const int rows = by1 - by0, cols = bx1 - bx0;
    for (int t = tid; t < (rows_per_band + 1) * CW; t += blockDim.x) s_d[t] = 0;
    __syncthreads();
    const int i_begin = bounds[chunk], i_end = bounds[chunk + 1];
    for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
        const int2 rc = rects[order[i]];
        const int x0 = max(rc.x & 0xffff, bx0), x1 = min((rc.x >> 16) & 0xffff, bx1);
        const int y0 = max(rc.y & 0xffff, by0), y1 = min((rc.y >> 16) & 0xffff, by1);
        if (x1 > x0 && y1 > y0) {
            atomicAdd(&s_d[(y0 - by0) * CW + (x0 - bx0)], 1);
            atomicAdd(&s_d[(y0 - by0) * CW + (x1 - bx0)], -1);
            atomicAdd(&s_d[(y1 - by0) * CW + (x0 - bx0)], -1);
            atomicAdd(&s_d[(y1 - by0) * CW + (x1 - bx0)], 1);
        }
    }
    __syncthreads();
    // integrate along x: warp w owns row w (warp scan with carry)
    if (warp < rows) {
        int carry = 0;
        for (int x = 0; x < cols; x += 32) {
            int v = (x + lane < cols) ? s_d[warp * CW + x + lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += n;
            }
            v += carry;
            if (x + lane < cols) s_d[warp * CW + x + lane] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // integrate along y: thread x owns a column; table rows are written coalesced
    for (int x = tid; x < cols; x += blockDim.x) {
        int run = 0;
        for (int r = 0; r < rows; ++r) {
            run += s_d[r * CW + x];
            table[(size_t)chunk * T + (size_t)(by0 + r) * tile_w + bx0 + x] = run;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2: ordered fill.  Phase A compacts (in stream order) the chunk's rectangles that touch the band into a
// shared-memory list; phase B lets every warp (= tile row) walk that list on its own, lanes appending at
// register cursors.  The two phases alternate until the chunk is consumed.
// ------------------------------------------------------------------------------------------------
constexpr int TR_LIST_CAP = 2048;  // band-filtered entries buffered per round (32 KB)

__global__ void __launch_bounds__(32 * TR_MAX_ROWS)
k_tile_fill(const int2 *__restrict__ rects, const int32_t *__restrict__ order, const int32_t *__restrict__ bounds,
            int tile_w, int tile_h, int rows_per_band, int row_bands, int R, const int32_t *__restrict__ table,
            const int32_t *__restrict__ offsets, int32_t *__restrict__ flatten_ids) {
    __shared__ int4 s_list[TR_LIST_CAP];       // (x0 | x1 << 16, y0 | y1 << 16, gid, -), band-filtered, in order
    __shared__ int s_wcnt[2][TR_MAX_ROWS];
    const int B = blockDim.x, nwarps = B >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = lanemask_lt();
    const int chunk = blockIdx.x % R;
    const int band = blockIdx.x / R;
    const int band_y = band % row_bands, band_x = band / row_bands;
    const int T = tile_w * tile_h;
    const int by0 = band_y * rows_per_band, by1 = min(tile_h, by0 + rows_per_band);
    const int bx0 = band_x * (32 * TR_NG), bx1 = min(tile_w, bx0 + 32 * TR_NG);
    const int y = by0 + warp;
    const bool row_ok = y < by1;

    int cur[TR_NG];
#pragma unroll
    for (int k = 0; k < TR_NG; ++k) {
        const int x = bx0 + lane + 32 * k;
        cur[k] = 0;
        if (row_ok && x < bx1) cur[k] = offsets[y * tile_w + x] + table[(size_t)chunk * T + y * tile_w + x];
    }
    const int i_begin = bounds[chunk], i_end = bounds[chunk + 1];

    int base = i_begin;
    int2 rc = make_int2(0, 0);
    int gid = 0;
    if (base + tid < i_end) {
        gid = order[base + tid];
        rc = rects[gid];
    }
    int par = 0;
    while (true) {
        // ---- phase A: append whole batches while they fit
        int fill = 0;
        while (base < i_end) {
            const bool keep = (base + tid < i_end) && (rc.x & 0xffff) < bx1 && ((rc.x >> 16) & 0xffff) > bx0 &&
                              (rc.y & 0xffff) < by1 && ((rc.y >> 16) & 0xffff) > by0;
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_wcnt[par][warp] = __popc(bal);
            __syncthreads();
            int off = 0, tot = 0;
            for (int w = 0; w < nwarps; ++w) {
                const int n = s_wcnt[par][w];
                off += (w < warp) ? n : 0;
                tot += n;
            }
            par ^= 1;
            if (fill + tot > TR_LIST_CAP) break;  // uniform: this batch stays in registers for the next round
            if (keep) s_list[fill + off + __popc(bal & lt)] = make_int4(rc.x, rc.y, gid, 0);
            fill += tot;
            base += B;
            rc = make_int2(0, 0);
            gid = 0;
            if (base + tid < i_end) {
                gid = order[base + tid];
                rc = rects[gid];
            }
        }
        __syncthreads();
        // ---- phase B: every warp walks the list for its own tile row, no barriers
        if (row_ok) {
            for (int j0 = 0; j0 < fill; j0 += 32) {
                bool hit = false;
                if (j0 + lane < fill) {
                    const int ey = s_list[j0 + lane].y;
                    hit = y >= (ey & 0xffff) && y < ((ey >> 16) & 0xffff);
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int4 e = s_list[j0 + src];  // broadcast
                    const int rel = bx0 + lane - (e.x & 0xffff);
                    const unsigned wdt = (unsigned)(((e.x >> 16) & 0xffff) - (e.x & 0xffff));
                    // (a warp-uniform "single 32-tile group" fast path was measured SLOWER than these four
                    // predicated appends: 0.45 vs 0.39 ms for the whole stage)
#pragma unroll
                    for (int k = 0; k < TR_NG; ++k) {
                        if ((unsigned)(rel + 32 * k) < wdt) {
                            flatten_ids[cur[k]] = e.z;
                            ++cur[k];
                        }
                    }
                }
            }
        }
        if (base >= i_end) break;
        __syncthreads();  // list fully consumed before phase A overwrites it
    }
}

// For every tile: exclusive prefix of table[.][tile] over the chunks (in place) and the tile's total.
// Block = 8 warps x 32 consecutive tiles; warp w owns the chunks [w*R/8, (w+1)*R/8).
__global__ void __launch_bounds__(256)
k_tile_prefix(int32_t *__restrict__ table, int R, int T, int32_t *__restrict__ tile_total) {
    __shared__ int s_part[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    const int r0 = (int)((long long)w * R / 8), r1 = (int)((long long)(w + 1) * R / 8);
    int sum = 0;
    if (t < T) {
#pragma unroll 4
        for (int r = r0; r < r1; ++r) sum += table[(size_t)r * T + t];
    }
    s_part[w][lane] = sum;
    __syncthreads();
    int run = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int v = s_part[k][lane];
        run += (k < w) ? v : 0;
        tot += v;
    }
    if (t < T) {
        if (w == 0) tile_total[t] = tot;
        for (int r = r0; r < r1; ++r) {
            const size_t a = (size_t)r * T + t;
            const int v = table[a];
            table[a] = run;
            run += v;
        }
    }
}

static inline size_t tl_align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TlPlan {
    int rows_per_band, row_bands, col_bands, R;
};

static TlPlan tl_plan(int N, int tile_w, int tile_h) {
    TlPlan p;
    p.row_bands = (tile_h + 19) / 20;                         // ~20 rows per band ...
    p.rows_per_band = (tile_h + p.row_bands - 1) / p.row_bands;  // ... evenly split (1080p: 4 x 17)
    p.col_bands = (tile_w + 32 * TR_NG - 1) / (32 * TR_NG);
    int bands = p.row_bands * p.col_bands;
    static const int target_ctas = [] {  // tuning knob (read once): B2S_TL_CTAS overrides the CTA budget
        const char *e = getenv("B2S_TL_CTAS");
        int v = e ? atoi(e) : 0;
        return v > 0 ? v : TR_TARGET_CTAS;
    }();
    int r = target_ctas / bands;
    int by_work = N / 512;  // no point in chunks shorter than a batch
    if (r > by_work) r = by_work;
    if (r > 1023) r = 1023;
    if (r < 1) r = 1;
    p.R = r;
    return p;
}

static inline bool tl_supported(int tile_w, int tile_h) {
    return tile_w > 0 && tile_h > 0 && tile_w <= 32767 && tile_h <= 32767;
}

extern "C" size_t b2s_bin_tiles_workspace_bytes(int N, long long M, int tile_w, int tile_h) {
    (void)M;
    if (!tl_supported(tile_w, tile_h)) return 0;
    TlPlan p = tl_plan(N, tile_w, tile_h);
    size_t T = (size_t)tile_w * tile_h;
    return tl_align256((size_t)p.R * T * 4) + tl_align256(T * 4) + tl_align256(b2s_scan_ws_ints((int)T) * 4) +
           tl_align256((size_t)(p.R + 1) * 4) + 1024;
}

extern "C" int b2s_bin_tiles(const int32_t *tile_rects, const int32_t *order, const int32_t *cum,
                             const int32_t *n_vis, int N, long long M, int tile_size, int tile_w, int tile_h, int32_t *flatten_ids, int32_t *isect_offsets,
                             void *workspace, size_t workspace_bytes, b2s_stream_t stream) {
    if (N < 0 || M < 0 || M >= (1LL << 31) || tile_w <= 0 || tile_h <= 0) return B2S_ERR_ARG;
    if (tile_size != 16 || !tl_supported(tile_w, tile_h)) return B2S_ERR_UNSUPPORTED;
    if (workspace_bytes < b2s_bin_tiles_workspace_bytes(N, M, tile_w, tile_h)) return B2S_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = tile_w * tile_h;
    if (M == 0 || N == 0) {
        cudaMemsetAsync(isect_offsets, 0, sizeof(int32_t) * (size_t)T, st);
        return B2S_OK;
    }
    const TlPlan p = tl_plan(N, tile_w, tile_h);
    char *w = (char *)workspace;
    int32_t *table = (int32_t *)w; w += tl_align256((size_t)p.R * T * 4);
    int32_t *tile_total = (int32_t *)w; w += tl_align256((size_t)T * 4);
    int32_t *sws = (int32_t *)w; w += tl_align256(b2s_scan_ws_ints(T) * 4);
    int32_t *bounds = (int32_t *)w;
    const int threads = 32 * p.rows_per_band;
    const int grid = p.R * p.row_bands * p.col_bands;
    const size_t smem_count = (size_t)(p.rows_per_band + 1) * (32 * TR_NG + 1) * sizeof(int);
    k_chunk_bounds<<<1, 1024, 0, st>>>(cum, n_vis, M, p.R, bounds);
    B2S_LAUNCH_CHECK();
    k_tile_count<<<grid, threads, smem_count, st>>>((const int2 *)tile_rects, order, bounds, tile_w, tile_h,
                                                     p.rows_per_band, p.row_bands, p.R, table);
    B2S_LAUNCH_CHECK();
    k_tile_prefix<<<b2s_div_up(T, 32), 256, 0, st>>>(table, p.R, T, tile_total);
    B2S_LAUNCH_CHECK();
    int rc = b2s_device_excl_scan(tile_total, nullptr, T, isect_offsets, nullptr, sws, st);
    if (rc != B2S_OK) return rc;
    k_tile_fill<<<grid, threads, 0, st>>>((const int2 *)tile_rects, order, bounds, tile_w, tile_h, p.rows_per_band,
                                           p.row_bands, p.R, table, isect_offsets, flatten_ids);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

// upstream key layout: (cam << (32 + tile_bits)) | (tile << 32) | (int64)(int32 view of depth); cam = 0.
// The tile of sorted entry i is the last t with offsets[t] <= i.
__global__ void __launch_bounds__(256)
k_isect_ids(const int32_t *__restrict__ offsets, int T, const int32_t *__restrict__ flatten_ids,
            const float *__restrict__ depths, long long M, int64_t *__restrict__ isect_ids) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int lo = 0, hi = T;  // first t with offsets[t] > i
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((long long)offsets[mid] <= i) lo = mid + 1;
        else hi = mid;
    }
    const int tile = lo - 1;
    const int32_t dbits = __float_as_int(depths[flatten_ids[i]]);
    isect_ids[i] = ((int64_t)tile << 32) | (int64_t)dbits;
}

extern "C" int b2s_bin_isect_ids(const int32_t *isect_offsets, int n_tiles, const int32_t *flatten_ids,
                                 const float *depths, long long M, int64_t *isect_ids, b2s_stream_t stream) {
    if (M < 0 || n_tiles <= 0) return B2S_ERR_ARG;
    if (M == 0) return B2S_OK;
    k_isect_ids<<<b2s_div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(isect_offsets, n_tiles, flatten_ids, depths, M,
                                                                       isect_ids);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
