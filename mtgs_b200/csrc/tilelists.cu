// Per-tile depth-ordered Gaussian lists without sorting the intersections (sm_100a).
//
// Replaces, together with depthsort.cu, upstream gsplat v1.4.0 isect_tiles pass 2 + the 64-bit cub radix sort +
// isect_offset_encode reached from mtgs_scene_graph.py:641-662.
//
// Input: the Gaussians in stable depth order (order[], from b2s_bin_sort_depth) and each Gaussian's tile
// rectangle (tile_rects, from the projection kernel).  The (Gaussian x tile) incidence is a sparse matrix given
// row by row (row = rectangle of one Gaussian); upstream's sorted list is its column-major transpose with the
// rows kept in order.  It is built by two order-preserving "interval multisplits", each perfectly load balanced:
//   stage 1 (rows):  the depth-ordered stream is cut into chunks of equal WORK (TL_C1 appended hits; one warp
//                    each); every Gaussian is appended to the list of each tile ROW its rectangle covers, as an
//                    8-byte hit (gaussian id, x0 | x1 << 16).  S = sum of rectangle heights hits in total.
//   stage 2 (tiles): every row list is cut into chunks of equal work (~TL_C2 appended entries; one warp each);
//                    every hit is appended to the list of each TILE of that row in [x0, x1) -> flatten_ids.
// Both stages are count (difference array in shared memory: 2 atomics per item, independent of its extent) ->
// exclusive prefix over the chunks -> fill.  The fill never ranks or sorts: a warp takes 32 items, every lane ORs
// its lane bit into the shared-memory word of each bin its interval covers (transposing the 32 x bins incidence),
// then lane l OWNS bins l, l+32, ... with their list cursors in registers and appends the items of its words in
// bit order.  Items are visited in stream order, so every list is in depth order (ties: ascending id): bit-identical
// to the stable global sort, with M x 4 B written exactly once and no atomics on global memory.
// Chunks are cut by appended entries, not by items or image area, so neither tile rows near the horizon (several
// times the mean load) nor screen-filling Gaussians close to the camera (hundreds of tiles each, all at the front
// of every list) unbalance the warps.  Integer work; no tensor cores.
#include <cstdlib>

#include "common.cuh"

constexpr int TL_C1 = 1024;     // tile-row hits appended per stage-1 chunk (one warp)
constexpr int TL_C2 = 2048;     // tile-list entries appended per stage-2 chunk (one warp)
constexpr int TL_WARPS = 2;     // warps (= chunks) per CTA: small CTAs so that every chunk is resident at once

__device__ __forceinline__ int tl_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// ---- interval counting: one warp, items [begin, end), interval of item i over bins = iv(i) -> (lo, hi).
// s_d: nbins + 1 ints of this warp's shared memory.  Afterwards lane-strided s_d[b] = number of items covering bin b.
template <class IV>
__device__ __forceinline__ void tl_warp_count(int *s_d, int nbins, int begin, int end, int lane, IV iv) {
    for (int b = lane; b <= nbins; b += 32) s_d[b] = 0;
    __syncwarp();
    for (int i = begin + lane; i < end; i += 32) {
        int lo, hi;
        iv(i, lo, hi);
        if (hi > lo) {
            atomicAdd(&s_d[lo], 1);
            atomicAdd(&s_d[hi], -1);
        }
    }
    __syncwarp();
    int carry = 0;
    for (int b0 = 0; b0 < nbins; b0 += 32) {
        const int v = (b0 + lane < nbins) ? s_d[b0 + lane] : 0;
        const int incl = tl_warp_incl_scan(v, lane) + carry;
        if (b0 + lane < nbins) s_d[b0 + lane] = incl;
        carry = __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
}

// ---- ordered fill: one warp, items [begin, end) in stream order.  fetch(i) loads the raw item i and
// decode(item, lo, hi, payload) turns it into its bin interval and payload; cursor0(bin) is the output index of
// the first item this chunk appends to `bin`; out receives the payloads.
// s_words: 32 * NG ints, s_pay: 32 payloads (this warp's shared memory).  NG = bins owned per lane (32 * NG bins
// per pass over the chunk; more bins: more passes).  The next batch's raw items are fetched while the current one
// is expanded (decoded only when needed, so the load latency is hidden), and the expansion is a warp-convergent,
// branch-free loop in which a lane pops one bit from each of its NG words per round (NG independent
// shared-load -> predicated-store chains).
template <class PAY, int NG, class ITEM, class FETCH, class DECODE, class CUR>
__device__ __forceinline__ void tl_warp_fill(unsigned *s_words, PAY *s_pay, int nbins, int begin, int end, int lane,
                                             FETCH fetch, DECODE decode, CUR cursor0, PAY *__restrict__ out) {
    constexpr int BAND = 32 * NG;
    for (int band = 0; band < nbins; band += BAND) {  // one pass unless there are more than 32 * NG bins
        PAY *cur[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            const int b = band + lane + 32 * k;
            cur[k] = out + (b < nbins ? cursor0(b) : 0);
        }
        ITEM nxt = ITEM();
        bool nvalid = begin + lane < end;
        if (nvalid) nxt = fetch(begin + lane);
        for (int i0 = begin; i0 < end; i0 += 32) {
            int lo = 0, hi = 0;
            PAY pay = PAY();
            if (nvalid) decode(nxt, lo, hi, pay);
            lo = max(lo - band, 0);
            hi = min(hi - band, BAND);
            s_pay[lane] = pay;
#pragma unroll
            for (int k = 0; k < NG; ++k) s_words[lane + 32 * k] = 0u;
            nvalid = i0 + 32 + lane < end;
            if (nvalid) nxt = fetch(i0 + 32 + lane);  // prefetch the next batch (raw; decoded next round)
            __syncwarp();
            for (int b = lo; b < hi; ++b) atomicOr(&s_words[b], 1u << lane);
            __syncwarp();
            unsigned w[NG];
            unsigned any = 0u;
#pragma unroll
            for (int k = 0; k < NG; ++k) {
                w[k] = s_words[lane + 32 * k];
                any |= w[k];
            }
            while (__any_sync(0xffffffffu, any != 0u)) {
                PAY v[NG];
#pragma unroll
                for (int k = 0; k < NG; ++k) v[k] = s_pay[(__ffs(w[k]) - 1) & 31];
                any = 0u;
#pragma unroll
                for (int k = 0; k < NG; ++k) {
                    if (w[k]) *cur[k] = v[k];
                    cur[k] += w[k] ? 1 : 0;
                    w[k] &= w[k] - 1;  // 0 stays 0
                    any |= w[k];
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// stage 1: rows.  Chunk c = the Gaussians whose exclusive row-hit prefix cum_rows[i] lies in
// [c * TL_C1, (c + 1) * TL_C1): equal WORK (appends) per warp, whatever the sizes of the Gaussians.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bounds1(const int32_t *__restrict__ cum_rows, const int32_t *__restrict__ n_vis, int nc1,
          int32_t *__restrict__ bounds1 /* [nc1 + 1] */) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nc1) return;
    const int nv = *n_vis;
    int lo = 0, hi = nv;  // first i with cum_rows[i] >= c * TL_C1
    if (c == nc1) lo = nv;
    const long long target = (long long)c * TL_C1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum_rows[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    bounds1[c] = lo;
}

// count1[y * nc1s + c] = Gaussians of chunk c covering tile row y; count1w[..] = tiles of row y they cover
__global__ void __launch_bounds__(32 * TL_WARPS)
k_rows_count(const int2 *__restrict__ rects, const int32_t *__restrict__ order, const int32_t *__restrict__ bounds1,
             int nc1, int tile_w, int tile_h, int nc1s, int32_t *__restrict__ count1, int32_t *__restrict__ count1w) {
    extern __shared__ int s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * TL_WARPS + warp;
    if (c >= nc1) return;
    const int begin = bounds1[c], end = bounds1[c + 1];
    int *s_d = s_dyn + warp * 2 * (tile_h + 1);
    int *s_w = s_d + tile_h + 1;
    for (int b = lane; b <= tile_h; b += 32) s_d[b] = s_w[b] = 0;
    __syncwarp();
    for (int i = begin + lane; i < end; i += 32) {
        const int2 rc = rects[order[i]];
        const int lo = rc.y & 0xffff, hi = min((rc.y >> 16) & 0xffff, tile_h);
        const int wd = max(0, min((rc.x >> 16) & 0xffff, tile_w) - (rc.x & 0xffff));
        if (hi > lo) {
            atomicAdd(&s_d[lo], 1);
            atomicAdd(&s_d[hi], -1);
            atomicAdd(&s_w[lo], wd);
            atomicAdd(&s_w[hi], -wd);
        }
    }
    __syncwarp();
    int carry = 0, carry_w = 0;
    for (int b0 = 0; b0 < tile_h; b0 += 32) {
        const int y = b0 + lane;
        const int v = y < tile_h ? s_d[y] : 0, vw = y < tile_h ? s_w[y] : 0;
        const int incl = tl_warp_incl_scan(v, lane) + carry, incl_w = tl_warp_incl_scan(vw, lane) + carry_w;
        if (y < tile_h) {
            count1[(size_t)y * nc1s + c] = incl;
            count1w[(size_t)y * nc1s + c] = incl_w;
        }
        carry = __shfl_sync(0xffffffffu, incl, 31);
        carry_w = __shfl_sync(0xffffffffu, incl_w, 31);
    }
}

// CTA (y, which): in-place exclusive scan over the chunks of count1[y][.] (which = 0; total -> row_len[y]) or
// count1w[y][.] (which = 1; total -> row_app[y])
__global__ void __launch_bounds__(256)
k_rows_prefix(int32_t *__restrict__ count1, int32_t *__restrict__ count1w, int nc1, int nc1s,
              int32_t *__restrict__ row_len, int32_t *__restrict__ row_app) {
    __shared__ int s_w[9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int32_t *r = (blockIdx.y ? count1w : count1) + (size_t)blockIdx.x * nc1s;
    const int per = (nc1 + 255) / 256;
    const int b = min(nc1, tid * per), e = min(nc1, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += r[i];
    const int incl = tl_warp_incl_scan(sum, lane);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < 8 ? s_w[lane] : 0;
        const int wi = tl_warp_incl_scan(w, lane);
        if (lane < 8) s_w[lane] = wi - w;
        if (lane == 7) s_w[8] = wi;
    }
    __syncthreads();
    int run = s_w[warp] + incl - sum;
    for (int i = b; i < e; ++i) {
        const int v = r[i];
        r[i] = run;
        run += v;
    }
    if (tid == 0) (blockIdx.y ? row_app : row_len)[blockIdx.x] = s_w[8];
}

// one warp: row_off = exclusive scan of row_len, chunk_off = exclusive scan of ceil(row_app / TL_C2); entry
// [tile_h] holds the totals (S and the number of stage-2 chunks).
__global__ void __launch_bounds__(32)
k_rows_offsets(const int32_t *__restrict__ row_len, const int32_t *__restrict__ row_app, int tile_h,
               int32_t *__restrict__ row_off, int32_t *__restrict__ chunk_off) {
    const int lane = threadIdx.x;
    int carry_r = 0, carry_c = 0;
    for (int y0 = 0; y0 < tile_h; y0 += 32) {
        const int y = y0 + lane;
        const int len = y < tile_h ? row_len[y] : 0;
        const int nch = y < tile_h ? (row_app[y] + TL_C2 - 1) / TL_C2 : 0;
        const int ir = tl_warp_incl_scan(len, lane), ic = tl_warp_incl_scan(nch, lane);
        if (y < tile_h) {
            row_off[y] = carry_r + ir - len;
            chunk_off[y] = carry_c + ic - nch;
        }
        carry_r += __shfl_sync(0xffffffffu, ir, 31);
        carry_c += __shfl_sync(0xffffffffu, ic, 31);
    }
    if (lane == 0) {
        row_off[tile_h] = carry_r;
        chunk_off[tile_h] = carry_c;
    }
}

// row_list[row_off[y] ..) = the Gaussians covering tile row y, in depth order, as (id, x0 | x1 << 16)
template <int NG>
__global__ void __launch_bounds__(32 * TL_WARPS)
k_rows_fill(const int2 *__restrict__ rects, const int32_t *__restrict__ order, const int32_t *__restrict__ bounds1,
            int nc1, int tile_h, int nc1s, const int32_t *__restrict__ count1,
            const int32_t *__restrict__ row_off, int2 *__restrict__ row_list) {
    __shared__ unsigned s_words[TL_WARPS][32 * NG];
    __shared__ int2 s_pay[TL_WARPS][32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * TL_WARPS + warp;
    if (c >= nc1) return;
    const int begin = bounds1[c], end = bounds1[c + 1];
    tl_warp_fill<int2, NG, int4>(
        s_words[warp], s_pay[warp], tile_h, begin, end, lane,
        [&](int i) {
            const int g = order[i];
            const int2 rc = rects[g];
            return make_int4(g, rc.x, rc.y, 0);
        },
        [&](const int4 &it, int &lo, int &hi, int2 &pay) {
            lo = it.z & 0xffff;
            hi = min((it.z >> 16) & 0xffff, tile_h);
            pay = make_int2(it.x, it.y);
        },
        [&](int y) { return row_off[y] + count1[(size_t)y * nc1s + c]; }, row_list);
}

// ------------------------------------------------------------------------------------------------
// stage 2: tiles.  The list of row y is cut at stage-1 cell boundaries (cell = hits of one stage-1 chunk in
// row y, a handful of hits) into chunks of ~TL_C2 appends: chunk k of row y = the cells whose exclusive append
// prefix count1w[y][c] lies in [k * TL_C2, (k + 1) * TL_C2).  bounds2[c2] = (first hit, row) of chunk c2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bounds2(const int32_t *__restrict__ count1, const int32_t *__restrict__ count1w, int nc1, int nc1s, int tile_h,
          const int32_t *__restrict__ row_off, const int32_t *__restrict__ chunk_off, int2 *__restrict__ bounds2) {
    const int c2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c2 >= chunk_off[tile_h]) return;
    int lo = 0, hi = tile_h;  // last y with chunk_off[y] <= c2
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= c2) lo = mid;
        else hi = mid;
    }
    const int y = lo;
    const long long target = (long long)(c2 - chunk_off[y]) * TL_C2;
    const int32_t *rw = count1w + (size_t)y * nc1s;
    int a = 0, b = nc1;  // first cell with append prefix >= target
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (rw[mid] < target) a = mid + 1;
        else b = mid;
    }
    const int first = a < nc1 ? row_off[y] + count1[(size_t)y * nc1s + a] : row_off[y + 1];
    bounds2[c2] = make_int2(first, y);
}

__device__ __forceinline__ bool tl_chunk2(int c2, int tile_h, const int32_t *__restrict__ row_off,
                                          const int32_t *__restrict__ chunk_off, const int2 *__restrict__ bounds2,
                                          int &y, int &begin, int &end) {
    if (c2 >= chunk_off[tile_h]) return false;
    const int2 b = bounds2[c2];
    y = b.y;
    begin = b.x;
    end = (c2 + 1 < chunk_off[y + 1]) ? bounds2[c2 + 1].x : row_off[y + 1];
    return true;
}

// table2[c2 * tile_w + x] = number of hits of chunk c2 covering tile x of its row
__global__ void __launch_bounds__(32 * TL_WARPS)
k_tiles_count(const int2 *__restrict__ row_list, const int32_t *__restrict__ row_off,
              const int32_t *__restrict__ chunk_off, const int2 *__restrict__ bounds2, int tile_w, int tile_h,
              int32_t *__restrict__ table2) {
    extern __shared__ int s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c2 = blockIdx.x * TL_WARPS + warp;
    int y, begin, end;
    if (!tl_chunk2(c2, tile_h, row_off, chunk_off, bounds2, y, begin, end)) return;
    int *s_d = s_dyn + warp * (tile_w + 1);
    tl_warp_count(s_d, tile_w, begin, end, lane, [&](int i, int &lo, int &hi) {
        const int rx = row_list[i].y;
        lo = rx & 0xffff;
        hi = min((rx >> 16) & 0xffff, tile_w);
    });
    for (int x = lane; x < tile_w; x += 32) table2[(size_t)c2 * tile_w + x] = s_d[x];
}

// CTA = (tile row y, 32 consecutive tiles of it); 8 warps split the row's chunks.  In-place exclusive prefix of
// table2[.][x] over the chunks of row y, and the tile's total.
__global__ void __launch_bounds__(256)
k_tiles_prefix(int32_t *__restrict__ table2, const int32_t *__restrict__ chunk_off, int tile_w,
               int32_t *__restrict__ tile_total) {
    __shared__ int s_part[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int y = blockIdx.y;
    const int x = blockIdx.x * 32 + lane;
    const int c0 = chunk_off[y], nc = chunk_off[y + 1] - c0;
    const int r0 = c0 + (int)((long long)w * nc / 8), r1 = c0 + (int)((long long)(w + 1) * nc / 8);
    int sum = 0;
    if (x < tile_w) {
#pragma unroll 4
        for (int r = r0; r < r1; ++r) sum += table2[(size_t)r * tile_w + x];
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
    if (x < tile_w) {
        if (w == 0) tile_total[y * tile_w + x] = tot;
        for (int r = r0; r < r1; ++r) {
            const size_t a = (size_t)r * tile_w + x;
            const int v = table2[a];
            table2[a] = run;
            run += v;
        }
    }
}

// single CTA: isect_offsets = exclusive scan of tile_total (T = 8160 at 1080p, 32400 at 4K)
__global__ void __launch_bounds__(1024)
k_tiles_offsets(const int32_t *__restrict__ tile_total, int T, int32_t *__restrict__ offsets) {
    __shared__ int s_w[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < T; base += 1024 * 8) {
        const int i0 = base + tid * 8;
        int v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = (i0 + k < T) ? tile_total[i0 + k] : 0;
            sum += v[k];
        }
        const int incl = tl_warp_incl_scan(sum, lane);
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int wv = s_w[lane];
            const int wi = tl_warp_incl_scan(wv, lane);
            s_w[lane] = wi - wv;
            if (lane == 31) s_w[32] = wi;
        }
        __syncthreads();
        int ex = carry + s_w[warp] + incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (i0 + k < T) offsets[i0 + k] = ex;
            ex += v[k];
        }
        carry += s_w[32];
        __syncthreads();
    }
}

// flatten_ids[isect_offsets[t] ..) = the Gaussians intersecting tile t, in depth order
template <int NG>
__global__ void __launch_bounds__(32 * TL_WARPS)
k_tiles_fill(const int2 *__restrict__ row_list, const int32_t *__restrict__ row_off,
             const int32_t *__restrict__ chunk_off, const int2 *__restrict__ bounds2, int tile_w, int tile_h,
             const int32_t *__restrict__ table2, const int32_t *__restrict__ offsets,
             int32_t *__restrict__ flatten_ids) {
    __shared__ unsigned s_words[TL_WARPS][32 * NG];
    __shared__ int s_pay[TL_WARPS][32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c2 = blockIdx.x * TL_WARPS + warp;
    int y, begin, end;
    if (!tl_chunk2(c2, tile_h, row_off, chunk_off, bounds2, y, begin, end)) return;
    tl_warp_fill<int, NG, int2>(
        s_words[warp], s_pay[warp], tile_w, begin, end, lane, [&](int i) { return row_list[i]; },
        [&](const int2 &h, int &lo, int &hi, int &pay) {
            lo = h.y & 0xffff;
            hi = min((h.y >> 16) & 0xffff, tile_w);
            pay = h.x;
        },
        [&](int x) { return offsets[y * tile_w + x] + table2[(size_t)c2 * tile_w + x]; }, flatten_ids);
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static inline size_t tl_align256(size_t x) { return (x + 255) & ~(size_t)255; }

static inline bool tl_supported(int tile_w, int tile_h) {
    return tile_w > 0 && tile_h > 0 && tile_w <= 32767 && tile_h <= 32767;
}

struct TlLayout {
    int nc1;         // stage-1 chunks
    long long nc2s;  // stage-2 chunks (upper bound from M)
    size_t bounds1, count1, count1w, row_len, row_app, row_off, chunk_off, row_list, bounds2, table2, tile_total, total;
};

static TlLayout tl_layout(long long M, long long S, int tile_w, int tile_h) {
    TlLayout L;
    L.nc1 = (int)(S / TL_C1) + 1;
    L.nc2s = M / TL_C2 + tile_h;
    size_t o = 0;
    L.bounds1 = o; o += tl_align256((size_t)(L.nc1 + 1) * 4);
    L.count1 = o; o += tl_align256((size_t)tile_h * L.nc1 * 4);
    L.count1w = o; o += tl_align256((size_t)tile_h * L.nc1 * 4);
    L.row_len = o; o += tl_align256((size_t)tile_h * 4);
    L.row_app = o; o += tl_align256((size_t)tile_h * 4);
    L.row_off = o; o += tl_align256((size_t)(tile_h + 1) * 4);
    L.chunk_off = o; o += tl_align256((size_t)(tile_h + 1) * 4);
    L.row_list = o; o += tl_align256((size_t)(S > 0 ? S : 1) * 8);
    L.bounds2 = o; o += tl_align256((size_t)(L.nc2s + 1) * 8);
    L.table2 = o; o += tl_align256((size_t)L.nc2s * tile_w * 4);
    L.tile_total = o; o += tl_align256((size_t)tile_w * tile_h * 4);
    L.total = o + 1024;
    return L;
}

extern "C" size_t b2s_bin_tiles_workspace_bytes(int N, long long M, long long S, int tile_w, int tile_h) {
    (void)N;
    if (!tl_supported(tile_w, tile_h) || S < 0 || M < 0) return 0;
    return tl_layout(M, S, tile_w, tile_h).total;
}

extern "C" int b2s_bin_tiles(const int32_t *tile_rects, const int32_t *order, const int32_t *cum_rows,
                             const int32_t *n_vis, int N, long long M, long long S, int tile_size, int tile_w,
                             int tile_h, int32_t *flatten_ids, int32_t *isect_offsets, void *workspace,
                             size_t workspace_bytes, b2s_stream_t stream) {
    if (N < 0 || M < 0 || S < 0 || S >= (1LL << 31) || M >= (1LL << 31) || tile_w <= 0 || tile_h <= 0) return B2S_ERR_ARG;
    if (tile_size != 16 || !tl_supported(tile_w, tile_h)) return B2S_ERR_UNSUPPORTED;
    if (workspace_bytes < b2s_bin_tiles_workspace_bytes(N, M, S, tile_w, tile_h)) return B2S_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = tile_w * tile_h;
    if (M == 0 || N == 0) {
        cudaMemsetAsync(isect_offsets, 0, sizeof(int32_t) * (size_t)T, st);
        return B2S_OK;
    }
    const TlLayout L = tl_layout(M, S, tile_w, tile_h);
    char *w = (char *)workspace;
    int32_t *bounds1 = (int32_t *)(w + L.bounds1);
    int32_t *count1 = (int32_t *)(w + L.count1);
    int32_t *count1w = (int32_t *)(w + L.count1w);
    int32_t *row_len = (int32_t *)(w + L.row_len);
    int32_t *row_app = (int32_t *)(w + L.row_app);
    int32_t *row_off = (int32_t *)(w + L.row_off);
    int32_t *chunk_off = (int32_t *)(w + L.chunk_off);
    int2 *row_list = (int2 *)(w + L.row_list);
    int2 *bounds2 = (int2 *)(w + L.bounds2);
    int32_t *table2 = (int32_t *)(w + L.table2);
    int32_t *tile_total = (int32_t *)(w + L.tile_total);
    const int2 *rects = (const int2 *)tile_rects;

    const int nc1 = L.nc1;
    const int grid1 = b2s_div_up(nc1, TL_WARPS);
    const int grid2 = b2s_div_up(L.nc2s, TL_WARPS);
    const size_t smem1 = (size_t)TL_WARPS * 2 * (tile_h + 1) * sizeof(int);
    const size_t smem2 = (size_t)TL_WARPS * (tile_w + 1) * sizeof(int);
    if (smem1 > 48 * 1024 || smem2 > 48 * 1024) return B2S_ERR_UNSUPPORTED;
    k_bounds1<<<b2s_div_up(nc1 + 1, 256), 256, 0, st>>>(cum_rows, n_vis, nc1, bounds1);
    B2S_LAUNCH_CHECK();
    k_rows_count<<<grid1, 32 * TL_WARPS, smem1, st>>>(rects, order, bounds1, nc1, tile_w, tile_h, nc1, count1, count1w);
    B2S_LAUNCH_CHECK();
    k_rows_prefix<<<dim3(tile_h, 2), 256, 0, st>>>(count1, count1w, nc1, nc1, row_len, row_app);
    B2S_LAUNCH_CHECK();
    k_rows_offsets<<<1, 32, 0, st>>>(row_len, row_app, tile_h, row_off, chunk_off);
    B2S_LAUNCH_CHECK();
    // bins owned per lane: 4 (<= 128 tile rows / columns, i.e. <= 2048 px) or 8; larger grids take several passes
    if (tile_h <= 128)
        k_rows_fill<4><<<grid1, 32 * TL_WARPS, 0, st>>>(rects, order, bounds1, nc1, tile_h, nc1, count1, row_off, row_list);
    else
        k_rows_fill<8><<<grid1, 32 * TL_WARPS, 0, st>>>(rects, order, bounds1, nc1, tile_h, nc1, count1, row_off, row_list);
    B2S_LAUNCH_CHECK();
    k_bounds2<<<b2s_div_up(L.nc2s, 256), 256, 0, st>>>(count1, count1w, nc1, nc1, tile_h, row_off, chunk_off, bounds2);
    B2S_LAUNCH_CHECK();
    k_tiles_count<<<grid2, 32 * TL_WARPS, smem2, st>>>(row_list, row_off, chunk_off, bounds2, tile_w, tile_h, table2);
    B2S_LAUNCH_CHECK();
    k_tiles_prefix<<<dim3(b2s_div_up(tile_w, 32), tile_h), 256, 0, st>>>(table2, chunk_off, tile_w, tile_total);
    B2S_LAUNCH_CHECK();
    k_tiles_offsets<<<1, 1024, 0, st>>>(tile_total, T, isect_offsets);
    B2S_LAUNCH_CHECK();
    if (tile_w <= 128)
        k_tiles_fill<4><<<grid2, 32 * TL_WARPS, 0, st>>>(row_list, row_off, chunk_off, bounds2, tile_w, tile_h, table2,
                                                         isect_offsets, flatten_ids);
    else
        k_tiles_fill<8><<<grid2, 32 * TL_WARPS, 0, st>>>(row_list, row_off, chunk_off, bounds2, tile_w, tile_h, table2,
                                                         isect_offsets, flatten_ids);
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
