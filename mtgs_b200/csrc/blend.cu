// Front-to-back alpha blending of depth-sorted tile lists, forward and backward (sm_100a).
//
// Replaces upstream gsplat v1.4.0 rasterize_to_pixels_{fwd,bwd}_kernel (kernels K8/K9)
// as reached from mtgs/scene_model/mtgs_scene_graph.py:641-662, plus the torch glue upstream runs around
// them: the expected-depth division render[..., -1] / alpha.clamp(1e-10) and its VJP.
//
// These kernels are FP32-ALU / MUFU.EX2 / shuffle bound, not HBM bound: no tensor cores.
// Blackwell-first choices:
//   * the lists they walk are built from TIGHT rectangles (project.cu: upstream's 3-sigma rectangle intersected with
//     the extents of the footprint {alpha >= 1/255}), and while staging a batch the forward applies the exact
//     per-tile test (tile_keep: can the pair reach alpha >= 1/255 at any pixel centre of the tile?) and compacts the
//     survivors with warp ballots.  Upstream's lists hold ~3x more entries that contribute to no pixel; dropping
//     them changes no output bit.  The test runs lazily, only for the part of a list that is walked before the
//     tile saturates (running it inside the tile-list build for all pairs was measured 90 us slower);
//   * forward: one CTA of 128 threads per 16x16 tile, TWO pixels per thread; a batch of 128 list entries is
//     gathered (id -> projected record), compacted into a shared-memory block in SoA form, and the finished block
//     is handed to the TMA engine (cp.async.bulk shared -> global, SASS UBLKCP) as a "walk record" while the CTA
//     blends it; the gathers of the NEXT batch are in flight during the blend of the current one;
//   * backward: 64 threads per tile, FOUR pixels per thread; it never touches the id lists or the per-Gaussian
//     arrays: it replays the forward's walk records back to front, each 6-8 KB block arriving by ONE TMA bulk copy
//     (cp.async.bulk global -> shared on an mbarrier), double buffered, so the next block lands while the
//     current one is being differentiated -- no gather latency, no staging instructions, no staging barriers;
//   * exponent evaluated in log2 units (log2e folded into the conic while staging) -> one ex2.approx per pair;
//   * both kernels are instruction-issue bound, so the per-pixel arithmetic runs on Blackwell's packed fp32x2
//     instructions (FFMA2 / FMUL2 / FADD2: two IEEE-rn fp32 operations per issue slot, scalar operands broadcast
//     for free): a thread's vertically adjacent pixels travel as the two halves of a float2;
//   * backward: the per-Gaussian partial sums (xy, |xy|, conic, opacity, <=8 colour channels) are reduced
//     with a transposing butterfly (16 shuffles instead of 16 x 5), parked per warp in shared memory, summed
//     over the CTA's two warps and flushed with 16-byte vector reductions (red.global.add.v4.f32): at most
//     2 + CDIM/4 L2 atomic operations per (tile, Gaussian) instead of upstream's (9 + CDIM) per (warp, Gaussian).
#include "common.cuh"

// compile-time tunables (tools/ab_lib.py measures variants; the defaults are the measured best)
#ifndef B2S_FWD_MINB
#define B2S_FWD_MINB 1
#endif
#ifndef B2S_FWD_UNROLL
#define B2S_FWD_UNROLL 2
#endif
#ifndef B2S_FWD_WARP_EXIT
#define B2S_FWD_WARP_EXIT 1
#endif
#ifndef B2S_FWD_EXIT_CHUNK
#define B2S_FWD_EXIT_CHUNK 16
#endif
#ifndef B2S_BWD_WARP_SKIP
#define B2S_BWD_WARP_SKIP 1
#endif
#ifndef B2S_BWD_PAIR_SKIP
#define B2S_BWD_PAIR_SKIP 0
#endif
#ifndef B2S_BWD_FL
#define B2S_BWD_FL 64
#endif
#ifndef B2S_BWD_MINB4
#define B2S_BWD_MINB4 10
#endif
constexpr int BL_FWD_UNROLL = B2S_FWD_UNROLL;
constexpr int BL_THREADS = 128;
constexpr int BL_BATCH = 128;  // list entries per walk-record block
#ifndef B2S_BWD_PX
#define B2S_BWD_PX 4  // default pixels per thread of the backward (both variants are built)
#endif

// float4s per walk-record block: q[128] = (mx, my, A, B), c[128] = (C, opacity, Gaussian id bits, 0), col[128][CDIM/4],
// then one header float4 whose .x holds the number of entries of the block (as int bits)
template <int CDIM> struct BlkLayout {
    static constexpr int CQ = CDIM / 4;
    static constexpr int HDR = BL_BATCH * (2 + CQ);  // index of the header
    static constexpr int F4 = HDR + 1;
    static constexpr unsigned BYTES = F4 * 16;
};

// exponent in log2 units: s = A dx^2 + B dx dy + C dy^2 with A = a/2*log2e, B = b*log2e, C = c/2*log2e.
// Written with explicit roundings so that forward and backward evaluate identical bits.
// ---- packed fp32x2 helpers (sm_100a FFMA2 / FMUL2 / FADD2) ----
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 abs2(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
// two pixels of one column (same dx)
__device__ __forceinline__ float2 splat_power2(float A, float B, float C, float dx, float2 dy) {
    const float2 u = fma2(bc2(B), dy, bc2(__fmul_rn(A, dx)));
    return fma2(mul2(bc2(C), dy), dy, mul2(u, bc2(dx)));
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct FwdRec {  // one list entry gathered into registers
    float4 q, c;
    float4 col[2];
    bool keep;
};
// List item -> projected record.  The lists walked here are the (tile row, column group) lists of the tile-list
// hierarchy (level 3): an item = (Gaussian id, tile-column range); the LAST filter level (does the item cover this
// tile's column?) and the exact per-tile test (tile_keep: can the pair reach alpha >= 1/255 at any pixel centre of the
// tile?) both run here, lazily, only for the part of the list that is walked before the tile saturates -- the blend
// reads ~16 % of each list, so building per-tile lists for everything was 90 us of mostly unread output.
template <int CQ>
__device__ __forceinline__ void fwd_gather(FwdRec &r, int idx, int end, const int2 *__restrict__ items, int tx,
                                           const float2 *__restrict__ means2d, const float4 *__restrict__ geo,
                                           const float4 *__restrict__ colpack, float rx0, float ry0, float rx1,
                                           float ry1) {
    r.keep = false;
    if (idx < end) {
        const int2 it = items[idx];
        if (tx >= (it.y & 0xffff) && tx < ((it.y >> 16) & 0xffff)) {
            const int g = it.x;
            const float2 m = means2d[g];
            const float4 ge = geo[g];
            r.q = make_float4(m.x, m.y, 0.5f * B2S_LOG2E * ge.x, B2S_LOG2E * ge.y);
            r.c = make_float4(0.5f * B2S_LOG2E * ge.z, ge.w, __int_as_float(g), 0.f);
            r.keep = tile_keep(m.x, m.y, r.q.z, r.q.w, r.c.x, ge.w, rx0, ry0, rx1, ry1);
            if (r.keep) {
#pragma unroll
                for (int k = 0; k < CQ; ++k) r.col[k] = colpack[(size_t)g * CQ + k];
            }
        }
    }
}

template <int CDIM, int DOUT, bool ED>
__global__ void __launch_bounds__(BL_THREADS, B2S_FWD_MINB)
k_blend_fwd(const float2 *__restrict__ means2d, const float4 *__restrict__ geo, const float4 *__restrict__ colpack,
            const int32_t *__restrict__ list_off /* [lists + 1] */, const int2 *__restrict__ items, int ncg, int cg_shift,
            int W, int H, int tile_w, float *__restrict__ render, float *__restrict__ alpha_out,
            int32_t *__restrict__ last_ids, float4 *__restrict__ records /* walk-record blocks, or null */,
            unsigned *__restrict__ blk_counter /* bump allocator of record blocks (zeroed by the caller) */,
            unsigned block_cap /* record blocks allocated */,
            int2 *__restrict__ tile_blocks /* [tiles]: (last block written, blocks written) */,
            int32_t *__restrict__ skip /* or null: non-zero = the lists are invalid (capacity overflow); set here when
                                          the record blocks run out */) {
    if (skip != nullptr && *(volatile int32_t *)skip != 0) return;
    constexpr int CQ = CDIM / 4;
    using BL = BlkLayout<CDIM>;
    // two block-shaped buffers: kept entries are packed densely (slot s lives in buffer (s >> 7) & 1), a block is
    // blended -- and handed to the TMA engine as a walk record -- as soon as its 128 slots are full
    __shared__ __align__(128) float4 s_blk[2][BL::F4];
    __shared__ int s_wcnt[BL_THREADS / 32];

    const int tile = blockIdx.x;
    const int ti = tile / tile_w, tj = tile - ti * tile_w;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = tj * 16 + (lane & 15);
    const int y0 = ti * 16 + 4 * warp + 2 * (lane >> 4);  // the thread's pixels: rows y0 and y0 + 1 of column x
    const bool in0 = x < W && y0 < H, in1 = x < W && (y0 + 1) < H;
    const float px = (float)x + 0.5f;
    const int L = ti * ncg + (tj >> cg_shift);
    const int start = list_off[L], end = list_off[L + 1];
    // pixel-centre rectangle of the tile clipped to the image
    const float rx0 = (float)(tj * 16) + 0.5f, ry0 = (float)(ti * 16) + 0.5f;
    const float rx1 = (float)min(tj * 16 + 16, W) - 0.5f, ry1 = (float)min(ti * 16 + 16, H) - 0.5f;
    const unsigned lt = lanemask_lt();

    // the thread's two pixels are the halves of every float2 below
    float2 T = make_float2(in0 ? 1.f : 0.f, in1 ? 1.f : 0.f);  // 0 = finished (terminated, or outside the image)
    float2 Tout = make_float2(1.f, 1.f);
    float2 acc[CDIM];
#pragma unroll
    for (int k = 0; k < CDIM; ++k) acc[k] = make_float2(0.f, 0.f);
    const float2 npy = make_float2(-((float)y0 + 0.5f), -((float)y0 + 1.5f));
    int cur0 = 0, cur1 = 0;
    int npend = 0;    // kept entries appended so far (CTA-uniform)
    int nblend = 0;   // entries blended so far = 128 x blocks finished
    int prev_blk = -1, nblocks = 0;  // chain of this tile's record blocks (meaningful in thread 0)

    // blend entries [0, cnt) of block buffer `bufi`; they carry the tile-local ids nblend .. nblend + cnt - 1
    auto process_block = [&](int bufi, int cnt) {
        float4 *s_q = s_blk[bufi], *s_c = s_blk[bufi] + BL_BATCH, *s_col = s_blk[bufi] + 2 * BL_BATCH;
        if (records != nullptr) {
            bool stored = false;  // (thread 0)
            if (threadIdx.x == 0) {
                const unsigned blk = atomicAdd(blk_counter, 1u);
                if (blk < block_cap) {
                    s_blk[bufi][BL::HDR] = make_float4(__int_as_float(cnt), __int_as_float(prev_blk), 0.f, 0.f);
                    prev_blk = (int)blk;
                    ++nblocks;
                    stored = true;
                } else if (skip != nullptr) {
                    // more pairs than the capacity this launch was sized for: nothing is written out of bounds, the
                    // frame is flagged and rebuilt (or the graph replay reported) by the host
                    atomicExch(skip, B2S_OVERFLOW_RECORDS);
                }
            }
            b2s_fence_async_smem();
            __syncthreads();
            if (stored) {
                b2s_bulk_s2g(records + (size_t)prev_blk * BL::F4, s_blk[bufi], BL::BYTES);
                b2s_bulk_commit();
            }
        }
        // Branch-free inner loop.  A pixel that has terminated (T' <= 1e-4 at some entry) carries T = 0 from then on:
        // every later entry gives T' = 0, is "not applied" and adds exact zeros, so no per-pixel done flag is tested
        // here; Tout keeps the transmittance behind the last applied entry (what alpha is computed from).
#if B2S_FWD_WARP_EXIT
        // a warp whose 64 pixels have all terminated leaves the block early (checked every 32 entries); the entries it
        // skips would have added exact zeros
        for (int tb = 0; tb < cnt; tb += B2S_FWD_EXIT_CHUNK) {
        if (__all_sync(0xffffffffu, T.x == 0.f && T.y == 0.f)) break;
        const int te = min(cnt, tb + B2S_FWD_EXIT_CHUNK);
#pragma unroll BL_FWD_UNROLL
        for (int t = tb; t < te; ++t) {
#else
        {
#pragma unroll BL_FWD_UNROLL
        for (int t = 0; t < cnt; ++t) {
#endif
            const float4 sq = s_q[t];
            const float4 sc = s_c[t];
            const float dx = sq.x - px;
            const float2 dy = add2(bc2(sq.y), npy);
            const float2 p = splat_power2(sq.z, sq.w, sc.x, dx, dy);
            const float2 ov = mul2(bc2(sc.y), make_float2(ex2_approx(-p.x), ex2_approx(-p.y)));
            const float2 al = make_float2(fminf(B2S_ALPHA_MAX, ov.x), fminf(B2S_ALPHA_MAX, ov.y));
            const bool ok0 = p.x >= 0.f && al.x >= B2S_ALPHA_MIN;
            const bool ok1 = p.y >= 0.f && al.y >= B2S_ALPHA_MIN;
            const float2 nT = mul2(T, fma2(al, bc2(-1.0f), bc2(1.0f)));  // T (1 - alpha)
            const bool ap0 = ok0 && nT.x > B2S_T_EPS, ap1 = ok1 && nT.y > B2S_T_EPS;  // Gaussian applied
            float2 vis = mul2(al, T);
            vis.x = ap0 ? vis.x : 0.f;
            vis.y = ap1 ? vis.y : 0.f;
            T.x = ok0 ? (ap0 ? nT.x : 0.f) : T.x;  // applied: T'; hit but T' <= eps: terminate; untouched: keep
            T.y = ok1 ? (ap1 ? nT.y : 0.f) : T.y;
            Tout.x = ap0 ? nT.x : Tout.x;
            Tout.y = ap1 ? nT.y : Tout.y;
            const int id = nblend + t;
            cur0 = ap0 ? id : cur0;
            cur1 = ap1 ? id : cur1;
#pragma unroll
            for (int j = 0; j < CQ; ++j) {
                const float4 v = s_col[t * CQ + j];
                acc[4 * j] = fma2(bc2(v.x), vis, acc[4 * j]);
                acc[4 * j + 1] = fma2(bc2(v.y), vis, acc[4 * j + 1]);
                acc[4 * j + 2] = fma2(bc2(v.z), vis, acc[4 * j + 2]);
                acc[4 * j + 3] = fma2(bc2(v.w), vis, acc[4 * j + 3]);
            }
        }
        }
        nblend += cnt;
    };

    bool all_done = false;
    FwdRec nxt;
    fwd_gather<CQ>(nxt, start + (int)threadIdx.x, end, items, tj, means2d, geo, colpack, rx0, ry0, rx1, ry1);
    for (int base = start; base < end; base += BL_BATCH) {
        // a TMA store still reading a block buffer must finish before new entries are appended to that buffer
        if (records != nullptr && threadIdx.x == 0) b2s_bulk_wait_read();
        if (__syncthreads_and(T.x == 0.f && T.y == 0.f)) { all_done = true; break; }
        // ballot compaction of the kept entries of this batch: slot order == list order
        const unsigned bal = __ballot_sync(0xffffffffu, nxt.keep);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < BL_THREADS / 32; ++w) {
            const int nw = s_wcnt[w];
            off += (w < warp) ? nw : 0;
            total += nw;
        }
        if (nxt.keep) {
            const int slot = npend + off + __popc(bal & lt);
            float4 *bufp = s_blk[(slot >> 7) & 1];
            const int e = slot & (BL_BATCH - 1);
            bufp[e] = nxt.q;
            bufp[BL_BATCH + e] = nxt.c;
#pragma unroll
            for (int j = 0; j < CQ; ++j) bufp[2 * BL_BATCH + e * CQ + j] = nxt.col[j];
        }
        npend += total;
        // gathers of the next batch fly while full blocks are blended
        fwd_gather<CQ>(nxt, base + BL_BATCH + (int)threadIdx.x, end, items, tj, means2d, geo, colpack, rx0, ry0, rx1, ry1);
        __syncthreads();
        if (npend - nblend >= BL_BATCH) process_block((nblend >> 7) & 1, BL_BATCH);  // at most one block fills per batch
    }
    if (!all_done && npend > nblend) {  // the partial block at the end of the list
        if (records != nullptr && threadIdx.x == 0) b2s_bulk_wait_read();
        if (!__syncthreads_and(T.x == 0.f && T.y == 0.f)) process_block((nblend >> 7) & 1, npend - nblend);
    }
    if (threadIdx.x == 0) {
        if (records != nullptr) b2s_bulk_wait_read();  // shared memory must outlive the last store
        if (tile_blocks != nullptr) tile_blocks[tile] = make_int2(prev_blk, nblocks);
    }

    const float T0 = Tout.x, T1 = Tout.y;
    float acc0[CDIM], acc1[CDIM];
#pragma unroll
    for (int k = 0; k < CDIM; ++k) {
        acc0[k] = acc[k].x;
        acc1[k] = acc[k].y;
    }

    // epilogue: optional expected-depth normalisation of the last written channel
    if (in0) {
        const size_t pid = (size_t)y0 * W + x;
        const float al = 1.0f - T0;
        if (ED) acc0[DOUT - 1] = acc0[DOUT - 1] / fmaxf(al, 1e-10f);
        alpha_out[pid] = al;
        last_ids[pid] = cur0;
        if (DOUT == 4) {
            reinterpret_cast<float4 *>(render)[pid] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
        } else {
#pragma unroll
            for (int k = 0; k < DOUT; ++k) render[pid * DOUT + k] = acc0[k];
        }
    }
    if (in1) {
        const size_t pid = (size_t)(y0 + 1) * W + x;
        const float al = 1.0f - T1;
        if (ED) acc1[DOUT - 1] = acc1[DOUT - 1] / fmaxf(al, 1e-10f);
        alpha_out[pid] = al;
        last_ids[pid] = cur1;
        if (DOUT == 4) {
            reinterpret_cast<float4 *>(render)[pid] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
        } else {
#pragma unroll
            for (int k = 0; k < DOUT; ++k) render[pid * DOUT + k] = acc1[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Gradient contribution of one Gaussian at a PAIR of pixels (upstream rasterize_to_pixels_bwd inner body), the two
// pixels in the halves of every float2.  A pixel the Gaussian does not touch has al = 0 and g = 0, which leaves its
// T / buf unchanged and adds exact zeros.  Accumulators (summed over the halves and the warp afterwards):
//   v[0..1] = -v_mean2d, v[2..3] = |v_mean2d|, v[4..6] = -(2 v_conic_a, v_conic_b, 2 v_conic_c), v[7] = v_opacity,
//   v[8..] = v_colour.  (signs and the factors 1/2 are applied once per Gaussian in the flush)
//   al = alpha (0 if masked), g = e^-sigma where the alpha clamp is inactive (else 0), ca/cb/cc = raw conic.
template <int CDIM>
__device__ __forceinline__ void pair_grad(float2 (&v)[16], float2 &T, float2 &Bv, const float2 (&vrc)[CDIM],
                                          const float2 Tfvra, const float (&col)[CDIM], const float ca, const float cb,
                                          const float cc, const float opac, const float dx, const float2 dy,
                                          const float2 al, const float2 g) {
    const float2 om = fma2(al, bc2(-1.0f), bc2(1.0f));
    const float2 ra = make_float2(rcp_approx(om.x), rcp_approx(om.y));  // alpha <= 0.999: well conditioned
    T = mul2(T, ra);                                                     // transmittance in front of this Gaussian
    const float2 fac = mul2(al, T);
    // Upstream's sum_k (c_k T - S_k ra) v_out_k, with S the un-normalised colour accumulated behind this Gaussian,
    // equals T sum_k (c_k - R_k) v_out_k where R = S / (transmittance behind).  R follows the convex update
    // R <- R + alpha (c - R), and only its projection Bv = <R, v_out> is ever needed, which follows the SAME update:
    // Bv <- Bv + alpha (<c, v_out> - Bv).  One scalar per pixel instead of CDIM, 2 + CDIM ops instead of 3 CDIM.
    float2 cv = mul2(bc2(col[0]), vrc[0]);
    v[8] = fma2(fac, vrc[0], v[8]);
#pragma unroll
    for (int k = 1; k < CDIM; ++k) {
        cv = fma2(bc2(col[k]), vrc[k], cv);
        v[8 + k] = fma2(fac, vrc[k], v[8 + k]);
    }
    const float2 acc = add2(cv, neg2(Bv));
    Bv = fma2(al, acc, Bv);
    const float2 va = fma2(T, acc, mul2(Tfvra, ra));  // v_alpha
    v[7] = fma2(g, va, v[7]);
    const float2 vs = mul2(mul2(bc2(opac), g), va);   // -v_sigma
    const float2 u = mul2(vs, bc2(dx)), w = mul2(vs, dy);
    v[4] = fma2(u, bc2(dx), v[4]);
    v[5] = fma2(u, dy, v[5]);
    v[6] = fma2(w, dy, v[6]);
    const float2 vx = fma2(bc2(ca), u, mul2(bc2(cb), w));
    const float2 vy = fma2(bc2(cb), u, mul2(bc2(cc), w));
    v[0] = add2(v[0], vx);
    v[1] = add2(v[1], vy);
    v[2] = add2(v[2], abs2(vx));
    v[3] = add2(v[3], abs2(vy));
}

// Transposing butterfly: NV (<= 16) values per lane -> value k summed over the warp lands in lane 2k (and 2k+1).
// Values NV..15 do not exist (CDIM = 4 uses 12): their partners in the first exchange are reduced without the
// lane-dependent selects.
template <int NV>
__device__ __forceinline__ float warp_reduce16_transposed(float (&v)[16], const int lane) {
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i + 8 < NV) {
            const float send = hi ? v[i] : v[i + 8];
            const float keep = hi ? v[i + 8] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        } else {
            v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);  // lanes 16..31 end up with a copy (slot 8 + i is unused)
        }
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = hi ? v[i] : v[i + 4];
        const float keep = hi ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = hi ? v[i] : v[i + 2];
        const float keep = hi ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    hi = lane & 2;
    {
        const float send = hi ? v[0] : v[1];
        const float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

template <int CDIM, int DOUT, bool ED>
__device__ __forceinline__ void load_pixel_cotangent(size_t pid, const float *__restrict__ render,
                                                     const float *__restrict__ alpha_in,
                                                     const float *__restrict__ v_render,
                                                     const float *__restrict__ v_alpha, float (&vrc)[CDIM],
                                                     float &vra, float &Tf) {
    const float al = alpha_in[pid];
    Tf = 1.0f - al;
#pragma unroll
    for (int k = 0; k < CDIM; ++k) vrc[k] = 0.f;
#pragma unroll
    for (int k = 0; k < DOUT; ++k) vrc[k] = v_render[pid * DOUT + k];
    vra = v_alpha[pid];
    if (ED) {
        // out_d = acc_d / max(alpha, 1e-10): v_acc_d = v_out_d / a ; v_alpha -= v_out_d * out_d / a (alpha >= 1e-10)
        const float a = fmaxf(al, 1e-10f);
        const float vd = vrc[DOUT - 1];
        const float outd = render[pid * DOUT + DOUT - 1];
        vrc[DOUT - 1] = vd / a;
        if (al >= 1e-10f) vra -= vd * outd / a;
    }
}

// PX pixels per thread (same column, PX adjacent rows) = PX / 2 packed pixel pairs; 256 / PX threads per tile.
// PX = 8: ONE warp per tile -- one butterfly and no cross-warp sum per (tile, Gaussian); PX = 4: two warps.
template <int CDIM, int DOUT, bool ED, int PX>
__global__ void __launch_bounds__(256 / PX, PX == 8 ? (CDIM == 4 ? 14 : 10) : (CDIM == 4 ? B2S_BWD_MINB4 : 7))
k_blend_bwd(const int32_t *__restrict__ skip /* or null: non-zero = the forward of this frame was abandoned */,
            const int2 *__restrict__ tile_blocks /* [tiles]: (last record block, blocks written) */,
            const float4 *__restrict__ records, int W, int H, int tile_w, const float *__restrict__ render, const float *__restrict__ alpha_in,
            const int32_t *__restrict__ last_ids, const float *__restrict__ v_render,
            const float *__restrict__ v_alpha, float *__restrict__ v_xyabs, float *__restrict__ v_geo,
            float *__restrict__ v_colpack) {
    if (skip != nullptr && *skip != 0) return;
    constexpr int THREADS = 256 / PX, WARPS = THREADS / 32, NP = PX / 2;
    constexpr int CQ = CDIM / 4;
    constexpr int NV = 8 + CDIM;   // partial sums per Gaussian
    constexpr int NQUAD = 2 + CQ;  // float4 groups flushed per Gaussian
    using BL = BlkLayout<CDIM>;
    __shared__ __align__(128) float4 s_blk[2][BL::F4];       // two walk-record blocks in flight
    constexpr int FL = B2S_BWD_FL; // list entries per flush round (half a block: keeps s_acc small, 10+ CTAs per SM)
    __shared__ __align__(16) float s_acc[WARPS][FL][NV];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ int s_max[WARPS];
    static_assert(PX == 4 || PX == 8, "pixels travel in pairs; 4 or 8 per thread");

    const int tile = blockIdx.x;
    const int ti = tile / tile_w, tj = tile - ti * tile_w;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = tj * 16 + (lane & 15);
#if B2S_BWD_PAIR_SKIP
    // pixel pair q of a lane = rows 4q + 2 (lane >> 4) + {0, 1} of the warp's strip: the warp's pairs q form BANDS of 4
    // contiguous rows, so a small Gaussian misses whole bands and their gradient terms are skipped (warp-uniform)
    const int ystrip = ti * 16 + 2 * PX * warp + 2 * (lane >> 4);
#define B2S_ROW(j) (ystrip + 4 * ((j) >> 1) + ((j) & 1))
#else
    const int ybase = ti * 16 + 2 * PX * warp + PX * (lane >> 4);
#define B2S_ROW(j) (ybase + (j))
#endif
    const float px = (float)x + 0.5f;
    const int2 tb = tile_blocks[tile];

    float2 T[NP], Tfvra[NP], npy[NP];
    float2 vrc[NP][CDIM], Bv[NP];
    int bin[PX];
    int maxbin = -1;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        float Tj = 1.f, Tfj = 1.f, vraj = 0.f;
        float vrcj[CDIM];
#pragma unroll
        for (int k = 0; k < CDIM; ++k) vrcj[k] = 0.f;
        bin[j] = -1;
        if (x < W && B2S_ROW(j) < H) {
            const size_t pid = (size_t)B2S_ROW(j) * W + x;
            load_pixel_cotangent<CDIM, DOUT, ED>(pid, render, alpha_in, v_render, v_alpha, vrcj, vraj, Tfj);
            Tj = Tfj;
            bin[j] = last_ids[pid];
        }
        maxbin = max(maxbin, bin[j]);
        const float npyj = -((float)B2S_ROW(j) + 0.5f);
        if (j & 1) {
            T[j / 2].y = Tj; Tfvra[j / 2].y = Tfj * vraj; npy[j / 2].y = npyj; Bv[j / 2].y = 0.f;
#pragma unroll
            for (int k = 0; k < CDIM; ++k) vrc[j / 2][k].y = vrcj[k];
        } else {
            T[j / 2].x = Tj; Tfvra[j / 2].x = Tfj * vraj; npy[j / 2].x = npyj; Bv[j / 2].x = 0.f;
#pragma unroll
            for (int k = 0; k < CDIM; ++k) vrc[j / 2][k].x = vrcj[k];
        }
    }
    // last list index any pixel of this tile blended
    maxbin = __reduce_max_sync(0xffffffffu, maxbin);
    const int wmax = maxbin;  // ... and any pixel of this warp
    if (lane == 0) s_max[warp] = maxbin;
    if (threadIdx.x == 0) {
        b2s_mbar_init(&s_bar[0], 1);
        b2s_mbar_init(&s_bar[1], 1);
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < WARPS; ++w) maxbin = max(maxbin, s_max[w]);
    // ids are tile-local and dense: entry id lives in the tile's block id >> 7.  The forward chained the tile's blocks
    // (header: entries, previous block); it may have written blocks behind the last contribution: skip them.
    const int hi0 = maxbin;
    if (hi0 < 0 || tb.y <= 0) return;  // nothing was blended in this tile (CTA-uniform)
    const int nblk = min((hi0 >> 7) + 1, tb.y);
    __shared__ int s_next;
    if (threadIdx.x == 0) {
        int blk = tb.x;
        for (int skip = tb.y - nblk; skip > 0; --skip)
            blk = __float_as_int(__ldg(&records[(size_t)blk * BL::F4 + BL::HDR]).y);
        s_next = blk;
        b2s_mbar_expect_tx(&s_bar[0], BL::BYTES);
        b2s_bulk_g2s(s_blk[0], records + (size_t)blk * BL::F4, BL::BYTES, &s_bar[0]);
    }

    for (int k = nblk - 1, it = 0; k >= 0; --k, ++it) {
        const int st = it & 1;
        b2s_mbar_wait(&s_bar[st], (it >> 1) & 1);
        // the header of the block that just landed names the previous block of the chain: fetch it into the other
        // buffer (last read in the previous iteration, which ended with a CTA barrier) while this one is processed
        if (threadIdx.x == 0 && k > 0) {
            const int prev = __float_as_int(s_blk[st][BL::HDR].y);
            b2s_mbar_expect_tx(&s_bar[st ^ 1], BL::BYTES);
            b2s_bulk_g2s(s_blk[st ^ 1], records + (size_t)prev * BL::F4, BL::BYTES, &s_bar[st ^ 1]);
        }
        const float4 *s_q = s_blk[st], *s_c = s_blk[st] + BL_BATCH, *s_col = s_blk[st] + 2 * BL_BATCH;
        const int base = k << 7;
        const int cnt = __float_as_int(s_blk[st][BL::HDR].x);  // entries of this block
        const int tmax = min(cnt - 1, hi0 - base);

        for (int t1 = tmax; t1 >= 0; t1 = (t1 & ~(FL - 1)) - 1) {  // flush rounds: entries [t0, t1] share s_acc
        const int t0 = t1 & ~(FL - 1);
        if (t1 != tmax) __syncthreads();  // the previous round's flush has read s_acc
        for (int t = t1; t >= t0; --t) {
            const int id = base + t;
#if B2S_BWD_WARP_SKIP
            if (id > wmax) {  // behind the last contribution of every pixel of this warp (warp-uniform)
                if (lane < NV) s_acc[warp][t - t0][lane] = 0.f;
                continue;
            }
#endif
            const float4 sq = s_q[t];
            const float4 sc = s_c[t];
            const float dx = sq.x - px;
            float2 dy[NP], al[NP], g[NP];
            bool okq[NP];
            bool any_ok = false;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                dy[q] = add2(bc2(sq.y), npy[q]);
                const float2 p = splat_power2(sq.z, sq.w, sc.x, dx, dy[q]);
                const float2 ev = make_float2(ex2_approx(-p.x), ex2_approx(-p.y));
                const float2 ov = mul2(bc2(sc.y), ev);  // opacity * e^-sigma before the 0.999 clamp
                const float a0 = fminf(B2S_ALPHA_MAX, ov.x), a1 = fminf(B2S_ALPHA_MAX, ov.y);
                const bool ok0 = id <= bin[2 * q] && p.x >= 0.f && a0 >= B2S_ALPHA_MIN;
                const bool ok1 = id <= bin[2 * q + 1] && p.y >= 0.f && a1 >= B2S_ALPHA_MIN;
                al[q] = make_float2(ok0 ? a0 : 0.f, ok1 ? a1 : 0.f);
                g[q] = make_float2((ok0 && ov.x <= B2S_ALPHA_MAX) ? ev.x : 0.f, (ok1 && ov.y <= B2S_ALPHA_MAX) ? ev.y : 0.f);
                okq[q] = ok0 || ok1;
                any_ok = any_ok || okq[q];
            }
            if (!__any_sync(0xffffffffu, any_ok)) {
                if (lane < NV) s_acc[warp][t - t0][lane] = 0.f;
                continue;
            }
            float2 v2[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) v2[k2] = make_float2(0.f, 0.f);
            float col[CDIM];
#pragma unroll
            for (int j = 0; j < CQ; ++j) {
                const float4 cv = s_col[t * CQ + j];
                col[4 * j] = cv.x; col[4 * j + 1] = cv.y; col[4 * j + 2] = cv.z; col[4 * j + 3] = cv.w;
            }
            // raw conic (a, b, c) from the log2-domain coefficients: a = 2 A ln2, b = B ln2, c = 2 C ln2
            const float ca = 2.f * B2S_LN2 * sq.z, cb = B2S_LN2 * sq.w, cc = 2.f * B2S_LN2 * sc.x;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
#if B2S_BWD_PAIR_SKIP
                if (!__any_sync(0xffffffffu, okq[q])) continue;  // no pixel of this band is touched: exact zeros
#endif
                pair_grad<CDIM>(v2, T[q], Bv[q], vrc[q], Tfvra[q], col, ca, cb, cc, sc.y, dx, dy[q], al[q], g[q]);
            }
            float v[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) v[k2] = k2 < NV ? v2[k2].x + v2[k2].y : 0.f;
            const float r = warp_reduce16_transposed<NV>(v, lane);
            if (!(lane & 1) && (lane >> 1) < NV) s_acc[warp][t - t0][lane >> 1] = r;
        }
        __syncthreads();

        // flush: item = (slot, quad); sum over the warps, one 16-byte vector reduction per non-zero quad
        const int nitem = (t1 - t0 + 1) * 4;
        for (int item = (int)threadIdx.x; item < nitem; item += THREADS) {
            const int slot = item >> 2, quad = item & 3;
            if (quad < NQUAD) {
                float4 s = *reinterpret_cast<const float4 *>(&s_acc[0][slot][4 * quad]);
#pragma unroll
                for (int w = 1; w < WARPS; ++w) {
                    const float4 o = *reinterpret_cast<const float4 *>(&s_acc[w][slot][4 * quad]);
                    s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
                }
                // undo the accumulation convention of pair_grad: quad 0 = (-v_x, -v_y, |v_x|, |v_y|),
                // quad 1 = (-2 v_conic_a, -v_conic_b, -2 v_conic_c, v_opacity), quads >= 2 = v_colour
                if (quad == 0) { s.x = -s.x; s.y = -s.y; }
                if (quad == 1) { s.x = -0.5f * s.x; s.y = -s.y; s.z = -0.5f * s.z; }
                if (s.x != 0.f || s.y != 0.f || s.z != 0.f || s.w != 0.f) {
                    const int gid = __float_as_int(s_c[t0 + slot].z);
                    float *dst = quad == 0 ? v_xyabs + (size_t)gid * 4
                               : quad == 1 ? v_geo + (size_t)gid * 4
                                           : v_colpack + (size_t)gid * CDIM + (quad - 2) * 4;
                    red_add_v4(dst, s.x, s.y, s.z, s.w);
                }
            }
        }
        }  // flush rounds
        __syncthreads();  // s_acc and this record buffer are free again
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
template <int CDIM, int DOUT, bool ED>
static int launch_fwd(const float *means2d, const float *geo, const float *colpack, const int32_t *list_off,
                      const int32_t *items, int ncg, int cg_shift, int W, int H, int tile_w, int tile_h, float *render,
                      float *alpha, int32_t *last_ids, float *records, unsigned *blk_counter, unsigned block_cap,
                      int32_t *tile_blocks, int32_t *skip, cudaStream_t st) {
    k_blend_fwd<CDIM, DOUT, ED><<<tile_w * tile_h, BL_THREADS, 0, st>>>(
        (const float2 *)means2d, (const float4 *)geo, (const float4 *)colpack, list_off, (const int2 *)items, ncg,
        cg_shift, W, H, tile_w, render, alpha, last_ids, (float4 *)records, blk_counter, block_cap, (int2 *)tile_blocks,
        skip);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
template <int CDIM, int DOUT, bool ED>
static int launch_bwd(const int32_t *skip, const int32_t *tile_blocks, const float *records, int W, int H, int tile_w, int tile_h,
                      const float *render, const float *alpha, const int32_t *last_ids, const float *v_render,
                      const float *v_alpha, float *v_xyabs, float *v_geo, float *v_colpack, int px, cudaStream_t st) {
    if (px == 4)
        k_blend_bwd<CDIM, DOUT, ED, 4><<<tile_w * tile_h, 64, 0, st>>>(skip, (const int2 *)tile_blocks, (const float4 *)records,
                                                                       W, H, tile_w, render, alpha, last_ids, v_render,
                                                                       v_alpha, v_xyabs, v_geo, v_colpack);
    else
        k_blend_bwd<CDIM, DOUT, ED, 8><<<tile_w * tile_h, 32, 0, st>>>(skip, (const int2 *)tile_blocks, (const float4 *)records,
                                                                       W, H, tile_w, render, alpha, last_ids, v_render,
                                                                       v_alpha, v_xyabs, v_geo, v_colpack);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

// (cdim, d_out) combinations MTGS reaches: RGB -> (4,3); RGB+ED -> (4,4);
// RGB+normals -> (8,6); RGB+normals+ED -> (8,7).  Others in 1..cdim are instantiated for completeness.
#define B2S_DISPATCH(FN, ...)                                                                      \
    do {                                                                                           \
        const bool ed = expected_depth != 0;                                                       \
        if (cdim == 4) {                                                                           \
            switch (d_out) {                                                                       \
                case 1: return ed ? FN<4, 1, true>(__VA_ARGS__) : FN<4, 1, false>(__VA_ARGS__);    \
                case 2: return ed ? FN<4, 2, true>(__VA_ARGS__) : FN<4, 2, false>(__VA_ARGS__);    \
                case 3: return ed ? FN<4, 3, true>(__VA_ARGS__) : FN<4, 3, false>(__VA_ARGS__);    \
                case 4: return ed ? FN<4, 4, true>(__VA_ARGS__) : FN<4, 4, false>(__VA_ARGS__);    \
            }                                                                                      \
        } else if (cdim == 8) {                                                                    \
            switch (d_out) {                                                                       \
                case 5: return ed ? FN<8, 5, true>(__VA_ARGS__) : FN<8, 5, false>(__VA_ARGS__);    \
                case 6: return ed ? FN<8, 6, true>(__VA_ARGS__) : FN<8, 6, false>(__VA_ARGS__);    \
                case 7: return ed ? FN<8, 7, true>(__VA_ARGS__) : FN<8, 7, false>(__VA_ARGS__);    \
                case 8: return ed ? FN<8, 8, true>(__VA_ARGS__) : FN<8, 8, false>(__VA_ARGS__);    \
            }                                                                                      \
        }                                                                                          \
        return B2S_ERR_UNSUPPORTED;                                                                \
    } while (0)

extern "C" uint32_t b2s_blend_record_blocks(long long pair_capacity, int n_tiles) {
    if (pair_capacity < 0 || n_tiles < 0) return 0;
    const size_t blocks = (size_t)(pair_capacity >> 7) + (size_t)n_tiles + 1;
    return blocks > 0xffffffffull ? 0xffffffffu : (uint32_t)blocks;
}

extern "C" size_t b2s_blend_record_bytes(long long pair_capacity, int n_tiles, int cdim) {
    if (pair_capacity < 0 || n_tiles < 0 || (cdim != 4 && cdim != 8)) return 0;
    // blocks are packed densely per tile: at most ceil(kept / 128) per tile, kept <= (Gaussian, tile) pairs in total
    const size_t blocks = b2s_blend_record_blocks(pair_capacity, n_tiles);
    return blocks * ((size_t)BL_BATCH * (size_t)(2 + cdim / 4) + 1) * 16;  // + header
}

extern "C" int b2s_blend_fwd(const float *means2d, const float *geo, const float *colpack, const int32_t *list_offsets,
                             const int32_t *list_items, int ncg, int cg_shift, int W, int H, int tile_w, int tile_h,
                             int cdim, int d_out, int expected_depth, float *render, float *alpha, int32_t *last_ids,
                             float *records, uint32_t record_blocks, uint32_t *block_counter, int32_t *tile_blocks,
                             int32_t *skip_flag, b2s_stream_t stream) {
    if (W <= 0 || H <= 0 || tile_w * 16 < W || tile_h * 16 < H || ncg < 1 || cg_shift < 0) return B2S_ERR_ARG;
    if (records != nullptr && (((uintptr_t)records & 127) || !block_counter || !tile_blocks)) return B2S_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    B2S_DISPATCH(launch_fwd, means2d, geo, colpack, list_offsets, list_items, ncg, cg_shift, W, H, tile_w, tile_h, render,
                 alpha, last_ids, records, block_counter, record_blocks, tile_blocks, skip_flag, st);
}

extern "C" int b2s_blend_bwd(const int32_t *tile_blocks, const float *records, int W, int H, int tile_w, int tile_h,
                             int cdim, int d_out, int expected_depth, const float *render, const float *alpha,
                             const int32_t *last_ids, const float *v_render, const float *v_alpha, float *v_xyabs,
                             float *v_geo, float *v_colpack, int px_per_thread, const int32_t *skip_flag,
                             b2s_stream_t stream) {
    if (W <= 0 || H <= 0 || tile_w * 16 < W || tile_h * 16 < H || records == nullptr || tile_blocks == nullptr)
        return B2S_ERR_ARG;
    if ((uintptr_t)records & 127) return B2S_ERR_ARG;
    if (px_per_thread != 0 && px_per_thread != 4 && px_per_thread != 8) return B2S_ERR_UNSUPPORTED;
    const int px = px_per_thread == 0 ? B2S_BWD_PX : px_per_thread;
    cudaStream_t st = (cudaStream_t)stream;
    B2S_DISPATCH(launch_bwd, skip_flag, tile_blocks, records, W, H, tile_w, tile_h, render, alpha, last_ids, v_render, v_alpha,
                 v_xyabs, v_geo, v_colpack, px, st);
}
