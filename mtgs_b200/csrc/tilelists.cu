// Per-tile depth-ordered Gaussian lists without sorting the intersections (sm_100a).
//
// Replaces, together with depthsort.cu, upstream gsplat v1.4.0 isect_tiles pass 2 + the 64-bit cub radix sort +
// isect_offset_encode reached from mtgs_scene_graph.py:641-662.
//
// Input: the Gaussians in stable depth order (order[], from b2s_bin_sort_depth) and each Gaussian's tile
// rectangle (tile_rects, from the projection kernel).  The (Gaussian x tile) incidence is a sparse matrix given
// row by row (row = rectangle of one Gaussian); upstream's sorted list is its column-major transpose with the
// rows kept in order.  It is built by a hierarchy of ORDER-PRESERVING FILTERS, each splitting every list of the
// previous level into <= 16 (32 for very large images) child lists:
//     level 1  depth-ordered Gaussians        -> row groups  (8 or 16 tile rows)
//     level 2  row-group lists                -> tile rows
//     level 3  tile-row lists                 -> column groups (16 tile columns)
//     level 4  (row, column-group) lists      -> tiles  = flatten_ids, isect_offsets
// A level is three launches: count (per 1024/2048-item chunk and child), prefix (over the chunks of a list; the last
// CTA turns the child lengths into child offsets and into the chunk table of the next level) and fill.  Count and
// fill walk a chunk 32 items at a time; for every child b the warp ballots "item covers b" and the covering lanes
// write their payload at cursor_b + (rank among the set lanes): stream order is preserved (depth order, ties by
// ascending id: bit-identical to the stable global sort), every store instruction writes one contiguous run, and
// the work per chunk does not depend on how large the Gaussians are -- a 4-byte scatter of the M entries by
// "tile-owner lanes" (the previous design) spent its time in the SM's store-address path (lg/mio-throttle stalls,
// ~2.5 cycles per entry).  Cost ~ sum over levels of items x children / 32 ballots instead of M scattered stores.
// Integer work; no tensor cores.
#include <cstdlib>

#include "common.cuh"

#ifndef B2S_CG_SHIFT_MIN
#define B2S_CG_SHIFT_MIN 4  // smallest column-group width (log2 tiles): the blend scans whole (row, column-group) lists
#endif
constexpr int TL_WARPS = 8;   // warps per CTA; a CTA owns one chunk, warp w the w-th slice of it
// items per chunk at level K = TL_WARPS slices of 128 (levels 1-2) or 256 (levels 3-4) items: short slices = many
// warps in flight (a warp is a serial chain of dependent loads and the early levels have few items), CTA-sized
// chunks = short per-list chunk tables (the prefix over a list's chunks is a serial dependency)
template <int K> struct TlCh {
    static constexpr int slice = K <= 2 ? 128 : 256;
    static constexpr int v = slice * TL_WARPS;
};
static inline int tl_ch(int k) { return (k <= 2 ? 128 : 256) * TL_WARPS; }

struct TlGeom {
    int tile_w, tile_h;
    int W, H;                // image size in pixels (exact mode: pixel-centre boxes of the tiles)
    int exact;               // level 4 keeps only (Gaussian, tile) pairs that can reach alpha >= 1/255 (tile_keep)
    int offs_total;          // isect_offsets has tile_w * tile_h + 1 entries; the last one receives the list length
    int *overflow;           // capacity mode (or null): set when a level needs more room than the caller provided;
                             // every later kernel of the build then returns at once (no out-of-bounds write)
    int rg_shift, cg_shift;  // tile rows per row group = 1 << rg_shift, tile columns per column group = 1 << cg_shift
    int nrg, ncg;            // number of row groups / column groups (<= 32)
};

// smallest shift >= lo with ceil(n / 2^shift) <= 16 (<= 32 at the largest shift); -1 if the grid is too large
static inline int tl_shift(int n, int lo) {
    for (int s = lo; s <= 5; ++s)
        if (((n + (1 << s) - 1) >> s) <= 16) return s;
    return ((n + 31) >> 5) <= 32 ? 5 : -1;
}
static inline bool tl_geom(int tile_w, int tile_h, TlGeom &g) {
    if (tile_w <= 0 || tile_h <= 0) return false;
    g.tile_w = tile_w;
    g.tile_h = tile_h;
    g.rg_shift = tl_shift(tile_h, 3);
    g.cg_shift = tl_shift(tile_w, B2S_CG_SHIFT_MIN);
    if (g.rg_shift < 0 || g.cg_shift < 0) return false;
    g.nrg = (tile_h + (1 << g.rg_shift) - 1) >> g.rg_shift;
    g.ncg = (tile_w + (1 << g.cg_shift) - 1) >> g.cg_shift;
    return true;
}

__device__ __forceinline__ int tl_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// children per list and geometry of level K
template <int K>
__device__ __forceinline__ int tl_nb(const TlGeom &g) {
    return K == 1 ? g.nrg : K == 2 ? (1 << g.rg_shift) : K == 3 ? g.ncg : (1 << g.cg_shift);
}
// children [lo, hi) of list L covered by an item whose packed range (lo | hi << 16 along the level's axis) is r
template <int K>
__device__ __forceinline__ void tl_bins(int r, int L, const TlGeom &g, int &lo, int &hi) {
    const int a = r & 0xffff, b = (r >> 16) & 0xffff;
    if (K == 1) {
        lo = a >> g.rg_shift;
        hi = b > a ? ((b - 1) >> g.rg_shift) + 1 : lo;
    } else if (K == 2) {
        const int base = L << g.rg_shift;
        lo = max(a - base, 0);
        hi = min(b - base, 1 << g.rg_shift);
    } else if (K == 3) {
        lo = a >> g.cg_shift;
        hi = b > a ? ((b - 1) >> g.cg_shift) + 1 : lo;
    } else {
        const int base = (L % g.ncg) << g.cg_shift;
        lo = max(a - base, 0);
        hi = min(b - base, 1 << g.cg_shift);
    }
}

// chunk c of level K -> list L and item range [begin, end)
template <int K>
__device__ __forceinline__ bool tl_chunk(int c, int nlists, const int32_t *__restrict__ list_off,
                                         const int32_t *__restrict__ chunk_off, const int32_t *__restrict__ n_vis,
                                         int &L, int &begin, int &end) {
    if (K == 1) {  // one list: the depth-ordered stream
        const int nv = *n_vis;
        L = 0;
        begin = c * TlCh<K>::v;
        end = min(nv, begin + TlCh<K>::v);
        return begin < nv;
    }
    if (c >= chunk_off[nlists]) return false;
    int lo = 0, hi = nlists;  // last L with chunk_off[L] <= c
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= c) lo = mid;
        else hi = mid;
    }
    L = lo;
    begin = list_off[L] + (c - chunk_off[L]) * TlCh<K>::v;
    end = min(list_off[L + 1], begin + TlCh<K>::v);
    return true;
}

// item i of level K as (gaussian id, packed range along the level's axis)
template <int K>
__device__ __forceinline__ int2 tl_item(int i, const int2 *__restrict__ in, const int32_t *__restrict__ order,
                                        const int2 *__restrict__ rects) {
    if (K == 1) {
        const int g = order[i];
        return make_int2(g, rects[g].y);
    }
    return in[i];
}

// children-per-slice counting: lane b of the warp ends with the number of items of the slice [sb, se) that cover
// child b.  Counting needs no ordering, so it uses a difference array (+1 at the first covered child, -1 behind the
// last: two shared-memory atomics per item, whatever its extent) and one warp scan, instead of one ballot per child.
template <int K>
__device__ __forceinline__ int tl_slice_count(const TlGeom &g, int NB, int L, int sb, int se, int lane, int *s_d /* 33 */,
                                              const int2 *__restrict__ in, const int32_t *__restrict__ order,
                                              const int2 *__restrict__ rects) {
    s_d[lane] = 0;
    if (lane == 0) s_d[32] = 0;
    __syncwarp();
    int2 nxt = make_int2(0, 0);
    if (sb + lane < se) nxt = tl_item<K>(sb + lane, in, order, rects);
    for (int i0 = sb; i0 < se; i0 += 32) {
        int lo = 0, hi = 0;
        if (i0 + lane < se) tl_bins<K>(nxt.y, L, g, lo, hi);
        if (i0 + 32 + lane < se) nxt = tl_item<K>(i0 + 32 + lane, in, order, rects);  // prefetch
        if (hi > lo) {
            atomicAdd(&s_d[lo], 1);
            atomicAdd(&s_d[hi], -1);  // hi <= NB <= 32
        }
    }
    __syncwarp();
    return tl_warp_incl_scan(s_d[lane], lane);  // lanes >= NB hold garbage-free zeros or are ignored by the callers
}

// Exact mode, level 4: bit b of the result = "the Gaussian of item `it` can reach alpha >= 1/255 at a pixel centre of
// child tile b of list L" (tile_keep, the test the blend kernels used to run while staging; conservative, so the
// image is unchanged).  Only the children [lo, hi) of the item's tight column range are tested.
__device__ __forceinline__ unsigned tl_tile_mask(const TlGeom &g, int L, int2 it, const float2 *__restrict__ means2d,
                                                 const float4 *__restrict__ geo) {
    int lo, hi;
    tl_bins<4>(it.y, L, g, lo, hi);
    if (hi <= lo) return 0u;
    const float2 m = means2d[it.x];
    const float4 ge = geo[it.x];
    const float A = 0.5f * B2S_LOG2E * ge.x, B = B2S_LOG2E * ge.y, C = 0.5f * B2S_LOG2E * ge.z;
    const int ty = L / g.ncg, tx0 = (L % g.ncg) << g.cg_shift;
    const float ry0 = (float)(ty * 16) + 0.5f, ry1 = (float)min(ty * 16 + 16, g.H) - 0.5f;
    unsigned mask = 0u;
    for (int b = lo; b < hi; ++b) {
        const int tx = tx0 + b;
        const float rx0 = (float)(tx * 16) + 0.5f, rx1 = (float)min(tx * 16 + 16, g.W) - 0.5f;
        if (tile_keep(m.x, m.y, A, B, C, ge.w, rx0, ry0, rx1, ry1)) mask |= 1u << b;
    }
    return mask;
}

// ---- count: table[b * nch + c] = items of chunk c covering child b
template <int K>
__global__ void __launch_bounds__(32 * TL_WARPS)
k_level_count(const TlGeom g, int2 *__restrict__ in, const int32_t *__restrict__ order,
              const int2 *__restrict__ rects, const int32_t *__restrict__ n_vis, int nlists,
              const int32_t *__restrict__ list_off, const int32_t *__restrict__ chunk_off, int nch,
              int32_t *__restrict__ table, int32_t *__restrict__ slice_cnt /* [nch][TL_WARPS][32] */,
              unsigned *__restrict__ ticket, const float2 *__restrict__ means2d, const float4 *__restrict__ geo) {
    __shared__ int s_cnt[TL_WARPS][32];
    __shared__ int s_diff[TL_WARPS][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0u;  // armed for this level's prefix kernel
    if (g.overflow != nullptr && *g.overflow != 0) return;
    const int c = blockIdx.x;
    int L, begin, end;
    if (!tl_chunk<K>(c, nlists, list_off, chunk_off, n_vis, L, begin, end)) return;  // CTA-uniform
    const int NB = tl_nb<K>(g);
    const int sb = min(end, begin + warp * TlCh<K>::slice), se = min(end, sb + TlCh<K>::slice);
    int mine;
    if (K == 4 && g.exact) {
        // exact tile masks, stored in place of the item's column range for the fill kernel
        mine = 0;
        for (int i0 = sb; i0 < se; i0 += 32) {
            unsigned mask = 0u;
            if (i0 + lane < se) {
                mask = tl_tile_mask(g, L, in[i0 + lane], means2d, geo);
                in[i0 + lane].y = (int)mask;
            }
            for (int b = 0; b < NB; ++b) {
                const int c = __popc(__ballot_sync(0xffffffffu, (mask >> b) & 1u));
                if (lane == b) mine += c;
            }
        }
    } else {
        mine = tl_slice_count<K>(g, NB, L, sb, se, lane, s_diff[warp], in, order, rects);
    }
    s_cnt[warp][lane] = lane < NB ? mine : 0;
    __syncthreads();
    // exclusive prefix over the slices (read back by the fill kernel) and the chunk's total
    int pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < TL_WARPS; ++w) {
        const int v = s_cnt[w][lane];
        pre += (w < warp) ? v : 0;
        tot += v;
    }
    slice_cnt[((size_t)c * TL_WARPS + warp) * 32 + lane] = pre;
    if (warp == 0 && lane < NB) table[(size_t)lane * nch + c] = tot;
}

// ---- prefix: one CTA per child list (L, b): in-place exclusive scan of table[., b] over the chunks of L and the
// child's length; the last CTA to finish then builds the child offsets and (levels 1-3) the next level's chunk
// table, or (level 4) isect_offsets.
constexpr int TL_PT = 512;  // threads of a prefix CTA (levels 3-4: 16 child lists per CTA, so few CTAs contend for the ticket)

__device__ __forceinline__ int tl_block_excl_scan(int v, int &total, int *s_w /* TL_PT / 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = tl_warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int pre = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < TL_PT / 32; ++w) {
        const int x = s_w[w];
        pre += (w < warp) ? x : 0;
        total += x;
    }
    return pre + incl - v;
}

template <int K>
__global__ void __launch_bounds__(TL_PT)
k_level_prefix(const TlGeom g, const int32_t *__restrict__ n_vis, int nlists, const int32_t *__restrict__ chunk_off,
               int nch, int32_t *__restrict__ table, int32_t *__restrict__ out_len, int32_t *__restrict__ out_off,
               int32_t *__restrict__ out_chunk_off, int32_t *__restrict__ isect_offsets,
               unsigned *__restrict__ ticket, long long cap_out /* entries the output list of this level can hold */) {
    __shared__ int s_w[TL_PT / 32];
    __shared__ unsigned s_last;
    const int tid = threadIdx.x;
    if (g.overflow != nullptr && *g.overflow != 0) return;
    const int NB = tl_nb<K>(g);
    const int nout = nlists * NB;
    if (K >= 3) {  // short chunk tables: one warp per child list
        const int lane = tid & 31;
        const int o = blockIdx.x * (TL_PT / 32) + (tid >> 5);
        if (o < nout) {
            const int L = o / NB, b = o - L * NB;
            const int c0 = chunk_off[L], c1 = chunk_off[L + 1];
            int32_t *row = table + (size_t)b * nch;
            int carry = 0;
            for (int c = c0; c < c1; c += 32) {
                const bool ok = c + lane < c1;
                const int v = ok ? row[c + lane] : 0;
                const int incl = tl_warp_incl_scan(v, lane);
                if (ok) row[c + lane] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) out_len[o] = carry;
        }
    } else {
        const int o = blockIdx.x;  // child list index L * NB + b
        const int L = o / NB, b = o - L * NB;
        int c0, c1;
        if (K == 1) {
            c0 = 0;
            c1 = (*n_vis + TlCh<K>::v - 1) / TlCh<K>::v;
        } else {
            c0 = chunk_off[L];
            c1 = chunk_off[L + 1];
        }
        int32_t *row = table + (size_t)b * nch;  // this child's counts, contiguous over the chunks
        int carry = 0;
        for (int c = c0; c < c1; c += TL_PT) {
            const bool ok = c + tid < c1;
            const int v = ok ? row[c + tid] : 0;
            int total;
            const int ex = tl_block_excl_scan(v, total, s_w);
            if (ok) row[c + tid] = carry + ex;
            carry += total;
        }
        if (tid == 0) out_len[o] = carry;
    }
    // last CTA: exclusive scans over all child lists (16 consecutive entries per thread and round)
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    constexpr int IPT = 16;
    constexpr int CHN = TlCh<(K < 4 ? K + 1 : 4)>::v;  // chunk size of the level that consumes the children
    int carry_e = 0, carry_c = 0;
    for (int base = 0; base < nout; base += TL_PT * IPT) {
        const int i0 = base + tid * IPT;
        int len[IPT], se = 0, sc = 0;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            len[k] = i0 + k < nout ? out_len[i0 + k] : 0;
            se += len[k];
            sc += (len[k] + CHN - 1) / CHN;
        }
        int te, tc;
        int pe = carry_e + tl_block_excl_scan(se, te, s_w);
        int pc = carry_c + tl_block_excl_scan(sc, tc, s_w);
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int i = i0 + k;
            if (i < nout) {
                out_off[i] = pe;
                if (K < 4) {
                    out_chunk_off[i] = pc;
                } else {  // child (L, b) of level 4 is tile (y, x)
                    const int L = i / NB, b = i - L * NB;
                    const int y = L / g.ncg, x = ((L % g.ncg) << g.cg_shift) + b;
                    if (y < g.tile_h && x < g.tile_w) isect_offsets[y * g.tile_w + x] = pe;
                }
            }
            pe += len[k];
            pc += (len[k] + CHN - 1) / CHN;
        }
        carry_e += te;
        carry_c += tc;
    }
    if (tid == 0) {
        out_off[nout] = carry_e;
        if (K < 4) out_chunk_off[nout] = carry_c;
        if (K == 4 && g.offs_total) isect_offsets[g.tile_w * g.tile_h] = carry_e;
        if (g.overflow != nullptr && (long long)carry_e > cap_out) *g.overflow = K;  // the fill of this level has not run yet
    }
}

// ---- fill: the items of chunk c are appended, in order, to every child list they cover
template <int K>
__global__ void __launch_bounds__(32 * TL_WARPS)
k_level_fill(const TlGeom g, const int2 *__restrict__ in, const int32_t *__restrict__ order,
             const int2 *__restrict__ rects, const int32_t *__restrict__ n_vis, int nlists,
             const int32_t *__restrict__ list_off, const int32_t *__restrict__ chunk_off, int nch,
             const int32_t *__restrict__ table, const int32_t *__restrict__ slice_cnt,
             const int32_t *__restrict__ out_off, int2 *__restrict__ out2, int32_t *__restrict__ out1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    if (g.overflow != nullptr && *g.overflow != 0) return;
    const int c = blockIdx.x;
    int L, begin, end;
    if (!tl_chunk<K>(c, nlists, list_off, chunk_off, n_vis, L, begin, end)) return;  // CTA-uniform
    const int NB = tl_nb<K>(g);
    const int sb = min(end, begin + warp * TlCh<K>::slice), se = min(end, sb + TlCh<K>::slice);
    // the slices of the chunk are filled concurrently: slice w starts behind what slices 0..w-1 append
    int cur = 0;  // lane b holds the cursor of child b
    if (lane < NB)
        cur = out_off[L * NB + lane] + table[(size_t)lane * nch + c] +
              slice_cnt[((size_t)c * TL_WARPS + warp) * 32 + lane];
    int2 nxt = make_int2(0, 0);
    if (sb + lane < se) nxt = tl_item<K>(sb + lane, in, order, rects);
    for (int i0 = sb; i0 < se; i0 += 32) {
        const int2 it = nxt;
        int lo = 0, hi = 0;
        const bool masked = K == 4 && g.exact;  // it.y holds the exact tile mask written by the count kernel
        if (i0 + lane < se && !masked) tl_bins<K>(it.y, L, g, lo, hi);
        const unsigned mask = (masked && i0 + lane < se) ? (unsigned)it.y : 0u;
        if (i0 + 32 + lane < se) nxt = tl_item<K>(i0 + 32 + lane, in, order, rects);  // prefetch
        int2 pay = it;
        if (K == 2 && hi > lo) pay.y = rects[it.x].x;  // rows are resolved: carry the column range from here on
        for (int b = 0; b < NB; ++b) {
            const bool hit = masked ? ((mask >> b) & 1u) != 0u : (lo <= b && b < hi);
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (bal == 0u) continue;
            const int base = __shfl_sync(0xffffffffu, cur, b);
            if (hit) {
                const int pos = base + __popc(bal & lt);
                if (K == 4) out1[pos] = pay.x;
                else out2[pos] = pay;
            }
            if (lane == b) cur += __popc(bal);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static inline size_t tl_align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TlLayout {
    int nl[5];         // lists entering level k (k = 1..4)
    int nb[5];         // children per list at level k
    long long nch[5];  // chunks at level k (upper bound)
    size_t table[5], slice[5], out_len[5], out_off[5], chunk_off[5], out[4], ticket, total;
    long long cap_out[5];  // entries the output list of level k can hold
};

// totals = {M, S, E1, E3, n_vis} as written by b2s_bin_sort_depth
static bool tl_layout(const long long *t, int tile_w, int tile_h, TlGeom &g, TlLayout &L) {
    if (!tl_geom(tile_w, tile_h, g)) return false;
    const long long M = t[0], S = t[1], E1 = t[2], E3 = t[3], nv = t[4];
    const long long in_items[5] = {0, nv, E1, S, E3};
    const long long out_items[5] = {0, E1, S, E3, M};
    L.nb[1] = g.nrg; L.nb[2] = 1 << g.rg_shift; L.nb[3] = g.ncg; L.nb[4] = 1 << g.cg_shift;
    L.nl[1] = 1;
    for (int k = 2; k <= 4; ++k) L.nl[k] = L.nl[k - 1] * L.nb[k - 1];
    size_t o = 0;
    for (int k = 1; k <= 4; ++k) {
        L.nch[k] = in_items[k] / tl_ch(k) + L.nl[k];
        const size_t nout = (size_t)L.nl[k] * L.nb[k];
        L.cap_out[k] = out_items[k];
        L.table[k] = o; o += tl_align256((size_t)L.nch[k] * L.nb[k] * 4);
        L.slice[k] = o; o += tl_align256((size_t)L.nch[k] * TL_WARPS * 32 * 4);
        L.out_len[k] = o; o += tl_align256(nout * 4);
        L.out_off[k] = o; o += tl_align256((nout + 1) * 4);
        L.chunk_off[k] = o; o += tl_align256((nout + 1) * 4);  // chunk table of level k + 1
        if (k < 4) { L.out[k] = o; o += tl_align256((size_t)(out_items[k] > 0 ? out_items[k] : 1) * 8); }
    }
    L.ticket = o; o += 256;
    L.total = o + 1024;
    return true;
}

extern "C" size_t b2s_bin_tiles_workspace_bytes(const long long *totals_host, int tile_w, int tile_h) {
    TlGeom g;
    TlLayout L;
    if (!totals_host) return 0;
    for (int k = 0; k < 5; ++k)
        if (totals_host[k] < 0 || totals_host[k] >= (1LL << 31)) return 0;
    if (!tl_layout(totals_host, tile_w, tile_h, g, L)) return 0;
    return L.total;
}

template <int K>
static int tl_run_level(const TlGeom &g, const TlLayout &L, char *w, const int32_t *order, const int2 *rects,
                        const int32_t *n_vis, int32_t *flatten_ids, int32_t *isect_offsets, const float2 *means2d,
                        const float4 *geo, cudaStream_t st) {
    int2 *in = K == 1 ? nullptr : (int2 *)(w + L.out[K - 1]);
    const int32_t *list_off = K == 1 ? nullptr : (const int32_t *)(w + L.out_off[K - 1]);
    const int32_t *chunk_off = K == 1 ? nullptr : (const int32_t *)(w + L.chunk_off[K - 1]);
    int32_t *table = (int32_t *)(w + L.table[K]);
    int32_t *slice_cnt = (int32_t *)(w + L.slice[K]);
    int32_t *out_len = (int32_t *)(w + L.out_len[K]);
    int32_t *out_off = (int32_t *)(w + L.out_off[K]);
    int32_t *out_chunk_off = (int32_t *)(w + L.chunk_off[K]);
    unsigned *ticket = (unsigned *)(w + L.ticket);
    int2 *out2 = K < 4 ? (int2 *)(w + L.out[K < 4 ? K : 1]) : nullptr;
    const int nlists = L.nl[K];
    const int nch = (int)L.nch[K];
    k_level_count<K><<<nch, 32 * TL_WARPS, 0, st>>>(g, in, order, rects, n_vis, nlists, list_off, chunk_off, nch, table,
                                                    slice_cnt, ticket, means2d, geo);
    B2S_LAUNCH_CHECK();
    const int nout = nlists * L.nb[K];
    k_level_prefix<K><<<K >= 3 ? b2s_div_up(nout, TL_PT / 32) : nout, TL_PT, 0, st>>>(g, n_vis, nlists, chunk_off, nch, table, out_len, out_off,
                                                          out_chunk_off, isect_offsets, ticket, L.cap_out[K]);
    B2S_LAUNCH_CHECK();
    k_level_fill<K><<<nch, 32 * TL_WARPS, 0, st>>>(g, in, order, rects, n_vis, nlists, list_off, chunk_off, nch, table,
                                                   slice_cnt, out_off, out2, flatten_ids);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_bin_tiles(const int32_t *tile_rects, const int32_t *order, const int32_t *n_vis,
                             const long long *totals_host, int N, int tile_size, int tile_w, int tile_h, int W, int H,
                             const float *means2d, const float *geo, int offsets_with_total, int32_t *overflow,
                             int levels, int32_t *flatten_ids, int32_t *isect_offsets, void *workspace,
                             size_t workspace_bytes, b2s_stream_t stream) {
    if (levels != 3 && levels != 4) return B2S_ERR_ARG;
    if (N < 0 || !totals_host || tile_w <= 0 || tile_h <= 0) return B2S_ERR_ARG;
    if ((means2d == nullptr) != (geo == nullptr)) return B2S_ERR_ARG;
    if (means2d != nullptr && (W <= 0 || H <= 0 || tile_w * 16 < W || tile_h * 16 < H)) return B2S_ERR_ARG;
    const long long M = totals_host[0];
    for (int k = 0; k < 5; ++k)
        if (totals_host[k] < 0 || totals_host[k] >= (1LL << 31)) return B2S_ERR_ARG;
    if (tile_size != 16) return B2S_ERR_UNSUPPORTED;
    TlGeom g;
    TlLayout L;
    if (!tl_layout(totals_host, tile_w, tile_h, g, L)) return B2S_ERR_UNSUPPORTED;
    if (workspace_bytes < L.total) return B2S_ERR_WORKSPACE;
    g.W = W;
    g.H = H;
    g.exact = means2d != nullptr;
    g.offs_total = offsets_with_total != 0;
    g.overflow = overflow;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = tile_w * tile_h;
    if (M == 0 || N == 0) {
        if (levels == 3)  // every (row, column-group) list is empty
            cudaMemsetAsync((char *)workspace + L.out_off[3], 0, sizeof(int32_t) * ((size_t)L.nl[3] * L.nb[3] + 1), st);
        else
            cudaMemsetAsync(isect_offsets, 0, sizeof(int32_t) * (size_t)(T + (g.offs_total ? 1 : 0)), st);
        return B2S_OK;
    }
    const float2 *m2 = (const float2 *)means2d;
    const float4 *ge = (const float4 *)geo;
    char *w = (char *)workspace;
    const int2 *rects = (const int2 *)tile_rects;
    int rc;
    if ((rc = tl_run_level<1>(g, L, w, order, rects, n_vis, flatten_ids, isect_offsets, m2, ge, st)) != B2S_OK) return rc;
    if ((rc = tl_run_level<2>(g, L, w, order, rects, n_vis, flatten_ids, isect_offsets, m2, ge, st)) != B2S_OK) return rc;
    if ((rc = tl_run_level<3>(g, L, w, order, rects, n_vis, flatten_ids, isect_offsets, m2, ge, st)) != B2S_OK) return rc;
    if (levels == 3) return B2S_OK;  // the caller walks the (row, column-group) lists itself (b2s_bin_tiles_l3_view)
    if ((rc = tl_run_level<4>(g, L, w, order, rects, n_vis, flatten_ids, isect_offsets, m2, ge, st)) != B2S_OK) return rc;
    return B2S_OK;
}

// Where the output of level 3 lives inside the workspace of a build with the same totals: items = int2 (Gaussian id,
// tile-column range x0 | x1 << 16) of the (tile row y, column group c) lists, list index y * ncg + c; offsets = int32
// [nlists + 1].  Tile (y, x) belongs to list y * ncg + (x >> cg_shift) and is covered by an item iff x0 <= x < x1.
extern "C" int b2s_bin_tiles_l3_view(const long long *totals_host, int tile_w, int tile_h, size_t *items_byte_offset,
                                     size_t *offsets_byte_offset, int *nlists, int *ncg, int *cg_shift) {
    TlGeom g = {};
    TlLayout L;
    if (!totals_host || !items_byte_offset || !offsets_byte_offset || !nlists || !ncg || !cg_shift) return B2S_ERR_ARG;
    if (!tl_layout(totals_host, tile_w, tile_h, g, L)) return B2S_ERR_UNSUPPORTED;
    *items_byte_offset = L.out[3];
    *offsets_byte_offset = L.out_off[3];
    *nlists = L.nl[3] * L.nb[3];
    *ncg = g.ncg;
    *cg_shift = g.cg_shift;
    return B2S_OK;
}

// Row-group / column-group geometry used by b2s_bin_sort_depth for the E1 / E3 totals (same rule as tl_geom).
int b2s_tl_shifts(int tile_w, int tile_h, int *rg_shift, int *cg_shift) {
    TlGeom g = {};
    if (!tl_geom(tile_w, tile_h, g)) return B2S_ERR_UNSUPPORTED;
    *rg_shift = g.rg_shift;
    *cg_shift = g.cg_shift;
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
