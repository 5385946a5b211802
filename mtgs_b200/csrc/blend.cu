// Front-to-back alpha blending of depth-sorted tile lists, forward and backward (sm_100a).
//
// Replaces upstream gsplat v1.4.0 rasterize_to_pixels_{fwd,bwd}_kernel (kernels K8/K9)
// as reached from mtgs/scene_model/mtgs_scene_graph.py:641-662, plus the torch glue upstream runs around
// them: the expected-depth division render[..., -1] / alpha.clamp(1e-10) and its VJP.
//
// These kernels are FP32-ALU / MUFU.EX2 / shuffle bound, not HBM bound: no tensor cores.
// Blackwell-first choices:
//   * one CTA of 128 threads per 16x16 tile, TWO pixels per thread (same column, adjacent rows): the staged
//     Gaussian record is read from shared memory once per two pixel-pairs and the per-(warp,Gaussian) gradient
//     reduction in the backward covers 64 pixels instead of 32;
//   * exact per-tile culling while staging: a (Gaussian, tile) pair whose minimum exponent over the tile's
//     pixel centres already gives alpha < 1/255 contributes to no pixel, so it is dropped from the staged batch
//     (ballot compaction).  The isect lists stay bit-identical to upstream; only dead work disappears;
//   * exponent evaluated in log2 units (log2e folded into the conic while staging) -> one ex2.approx per pair;
//   * both kernels are instruction-issue bound, so the per-pixel arithmetic runs on Blackwell's packed fp32x2
//     instructions (FFMA2 / FMUL2 / FADD2: two IEEE-rn fp32 operations per issue slot, scalar operands broadcast
//     for free): a thread's vertically adjacent pixels travel as the two halves of a float2.  Every packed op is
//     the same rn operation the scalar code performed, so the forward image is bit-identical to the scalar kernel;
//   * backward: the 16 per-Gaussian partial sums (xy, |xy|, conic, opacity, <=8 colour channels) are reduced
//     with a transposing butterfly (16 shuffles instead of 16 x 5), parked per warp in shared memory, summed
//     over the CTA's four warps and flushed with 16-byte vector reductions (red.global.add.v4.f32): at most
//     2 + CDIM/4 L2 atomic operations per (tile, Gaussian) instead of upstream's (9 + CDIM) per (warp, Gaussian).
#include <cstdlib>

#include "common.cuh"

constexpr int BL_THREADS = 128;
constexpr int BL_WARPS = BL_THREADS / 32;
constexpr int BL_BATCH = BL_THREADS;

// exponent in log2 units: s = A dx^2 + B dx dy + C dy^2 with A = a/2*log2e, B = b*log2e, C = c/2*log2e.
// Written with explicit roundings so that forward and backward evaluate identical bits.
__device__ __forceinline__ float splat_power(float A, float B, float C, float dx, float dy) {
    float u = fmaf(B, dy, __fmul_rn(A, dx));
    return fmaf(__fmul_rn(C, dy), dy, __fmul_rn(u, dx));
}

// ---- packed fp32x2 helpers (sm_100a FFMA2 / FMUL2 / FADD2) ----
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 abs2(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
// splat_power for two pixels of one column (same dx): identical bits per half as the scalar function
__device__ __forceinline__ float2 splat_power2(float A, float B, float C, float dx, float2 dy) {
    const float2 u = fma2(bc2(B), dy, bc2(__fmul_rn(A, dx)));
    return fma2(mul2(bc2(C), dy), dy, mul2(u, bc2(dx)));
}

// Can this Gaussian reach alpha >= 1/255 at any pixel centre of the rectangle [rx0,rx1]x[ry0,ry1]?
// Conservative (never drops a contributing pair): exact box-constrained minimum of the convex quadratic,
// compared with log2(255 * opacity) plus a slack that dominates fp32 evaluation error.
__device__ __forceinline__ bool tile_keep(float mx, float my, float A, float B, float C, float opac, float rx0,
                                          float ry0, float rx1, float ry1) {
    if (!(opac >= 0.0039f)) return false;  // opac * e^-sigma <= opac < 1/255 (also drops NaN)
    float ex = fminf(fmaxf(mx, rx0), rx1) - mx;  // nearest pixel-centre coordinate minus mean (0 if inside)
    float ey = fminf(fmaxf(my, ry0), ry1) - my;
    if (ex == 0.f && ey == 0.f) return true;
    float tau = __log2f(opac * 255.0f);
    float smin = 3.0e38f, mag = 0.f;
    if (ex != 0.f) {  // facing vertical edge: dx fixed, minimise over dy
        float dy = fminf(fmaxf(-B * ex / (2.f * C), ry0 - my), ry1 - my);
        float t0 = A * ex * ex, t1 = B * ex * dy, t2 = C * dy * dy;
        float s = t0 + t1 + t2;
        if (s < smin) { smin = s; mag = fabsf(t0) + fabsf(t1) + fabsf(t2); }
    }
    if (ey != 0.f) {  // facing horizontal edge
        float dx = fminf(fmaxf(-B * ey / (2.f * A), rx0 - mx), rx1 - mx);
        float t0 = A * dx * dx, t1 = B * dx * ey, t2 = C * ey * ey;
        float s = t0 + t1 + t2;
        if (s < smin) { smin = s; mag = fabsf(t0) + fabsf(t1) + fabsf(t2); }
    }
    return !(smin > tau + 0.02f + 2e-5f * mag);  // NaN-safe: keep on NaN
}

struct TileGeom {
    int x, y0;               // this thread's pixel column and first row (second row is y0 + 1)
    bool in0, in1;
    float px, py0, py1;
    float rx0, ry0, rx1, ry1;  // pixel-centre rectangle of the tile clipped to the image
    int start, end;
};

__device__ __forceinline__ TileGeom tile_geom(int tile_w, int tile_h, int W, int H, long long M,
                                              const int32_t *__restrict__ offsets) {
    TileGeom g;
    const int tile = blockIdx.x;
    const int ti = tile / tile_w, tj = tile - ti * tile_w;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    g.x = tj * 16 + (lane & 15);
    g.y0 = ti * 16 + 4 * warp + 2 * (lane >> 4);
    g.in0 = g.x < W && g.y0 < H;
    g.in1 = g.x < W && (g.y0 + 1) < H;
    g.px = (float)g.x + 0.5f;
    g.py0 = (float)g.y0 + 0.5f;
    g.py1 = g.py0 + 1.0f;
    g.rx0 = (float)(tj * 16) + 0.5f;
    g.ry0 = (float)(ti * 16) + 0.5f;
    g.rx1 = (float)min(tj * 16 + 16, W) - 0.5f;
    g.ry1 = (float)min(ti * 16 + 16, H) - 0.5f;
    g.start = offsets[tile];
    g.end = (tile == tile_w * tile_h - 1) ? (int)M : offsets[tile + 1];
    return g;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int CDIM, int DOUT, bool ED>
__global__ void __launch_bounds__(BL_THREADS)
k_blend_fwd(const float2 *__restrict__ means2d, const float4 *__restrict__ geo, const float4 *__restrict__ colpack,
            const int32_t *__restrict__ offsets, const int32_t *__restrict__ flatten_ids, long long M, int W, int H,
            int tile_w, int tile_h, float *__restrict__ render, float *__restrict__ alpha_out,
            int32_t *__restrict__ last_ids) {
    constexpr int CQ = CDIM / 4;
    __shared__ float4 s_q[BL_BATCH];        // mx, my, A, B
    __shared__ float2 s_c[BL_BATCH];        // C, opacity
    __shared__ int s_idx[BL_BATCH];         // index into the sorted list
    __shared__ float4 s_col[BL_BATCH][CQ];
    __shared__ int s_wcnt[BL_WARPS];

    const TileGeom tg = tile_geom(tile_w, tile_h, W, H, M, offsets);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();

    // the thread's two pixels (rows y0, y0 + 1) are the halves of every float2 below
    float2 T = make_float2(1.f, 1.f);
    float2 acc[CDIM];
#pragma unroll
    for (int k = 0; k < CDIM; ++k) acc[k] = make_float2(0.f, 0.f);
    const float2 npy = make_float2(-tg.py0, -tg.py1);
    int cur0 = 0, cur1 = 0;
    bool done0 = !tg.in0, done1 = !tg.in1;

    for (int base = tg.start; base < tg.end; base += BL_BATCH) {
        if (__syncthreads_and(done0 && done1)) break;
        const int idx = base + threadIdx.x;
        bool keep = false;
        int g = 0;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 c = make_float2(0.f, 0.f);
        if (idx < tg.end) {
            g = flatten_ids[idx];
            const float2 m = means2d[g];
            const float4 ge = geo[g];
            q = make_float4(m.x, m.y, 0.5f * B2S_LOG2E * ge.x, B2S_LOG2E * ge.y);
            c = make_float2(0.5f * B2S_LOG2E * ge.z, ge.w);
            keep = tile_keep(m.x, m.y, q.z, q.w, c.x, c.y, tg.rx0, tg.ry0, tg.rx1, tg.ry1);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < BL_WARPS; ++w) {
            int n = s_wcnt[w];
            off += (w < warp) ? n : 0;
            total += n;
        }
        if (keep) {
            const int slot = off + __popc(bal & lt);
            s_q[slot] = q;
            s_c[slot] = c;
            s_idx[slot] = idx;
#pragma unroll
            for (int k = 0; k < CQ; ++k) s_col[slot][k] = colpack[(size_t)g * CQ + k];
        }
        __syncthreads();

#pragma unroll 2
        for (int t = 0; t < total; ++t) {
            const float4 sq = s_q[t];
            const float2 sc = s_c[t];
            const float dx = sq.x - tg.px;
            const float2 dy = add2(bc2(sq.y), npy);
            const float2 p = splat_power2(sq.z, sq.w, sc.x, dx, dy);
            const float2 ov = mul2(bc2(sc.y), make_float2(ex2_approx(-p.x), ex2_approx(-p.y)));
            const float2 al = make_float2(fminf(B2S_ALPHA_MAX, ov.x), fminf(B2S_ALPHA_MAX, ov.y));
            const bool ok0 = !done0 && p.x >= 0.f && al.x >= B2S_ALPHA_MIN;
            const bool ok1 = !done1 && p.y >= 0.f && al.y >= B2S_ALPHA_MIN;
            if (ok0 || ok1) {
                const float2 nT = mul2(T, fma2(al, bc2(-1.0f), bc2(1.0f)));  // T (1 - alpha)
                float2 vis = mul2(al, T);
                const bool ap0 = ok0 && nT.x > B2S_T_EPS, ap1 = ok1 && nT.y > B2S_T_EPS;  // Gaussian applied
                done0 = done0 || (ok0 && !ap0);
                done1 = done1 || (ok1 && !ap1);
                vis.x = ap0 ? vis.x : 0.f;
                vis.y = ap1 ? vis.y : 0.f;
                T.x = ap0 ? nT.x : T.x;
                T.y = ap1 ? nT.y : T.y;
                const int id = s_idx[t];
                cur0 = ap0 ? id : cur0;
                cur1 = ap1 ? id : cur1;
#pragma unroll
                for (int k = 0; k < CQ; ++k) {
                    const float4 v = s_col[t][k];
                    acc[4 * k] = fma2(bc2(v.x), vis, acc[4 * k]);
                    acc[4 * k + 1] = fma2(bc2(v.y), vis, acc[4 * k + 1]);
                    acc[4 * k + 2] = fma2(bc2(v.z), vis, acc[4 * k + 2]);
                    acc[4 * k + 3] = fma2(bc2(v.w), vis, acc[4 * k + 3]);
                }
            }
        }
    }
    const float T0 = T.x, T1 = T.y;
    float acc0[CDIM], acc1[CDIM];
#pragma unroll
    for (int k = 0; k < CDIM; ++k) {
        acc0[k] = acc[k].x;
        acc1[k] = acc[k].y;
    }

    // epilogue: optional expected-depth normalisation of the last written channel
    if (tg.in0) {
        const size_t pid = (size_t)tg.y0 * W + tg.x;
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
    if (tg.in1) {
        const size_t pid = (size_t)(tg.y0 + 1) * W + tg.x;
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
__device__ __forceinline__ void pair_grad(float2 (&v)[16], float2 &T, float2 (&buf)[CDIM], const float2 (&vrc)[CDIM],
                                          const float2 Tfvra, const float (&col)[CDIM], const float ca, const float cb,
                                          const float cc, const float opac, const float dx, const float2 dy,
                                          const float2 al, const float2 g) {
    const float2 om = fma2(al, bc2(-1.0f), bc2(1.0f));
    const float2 ra = make_float2(rcp_approx(om.x), rcp_approx(om.y));  // alpha <= 0.999: well conditioned
    T = mul2(T, ra);                                                     // transmittance in front of this Gaussian
    const float2 fac = mul2(al, T);
    // buf[] holds R = (colour accumulated behind this Gaussian) / (transmittance behind it).  Upstream's
    // (c T - S ra) with S the un-normalised sum equals T (c - R); R is updated as a convex combination
    // R <- R + alpha (c - R), which needs no division and one multiply less per channel.
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < CDIM; ++k) {
        const float2 d = add2(bc2(col[k]), neg2(buf[k]));
        v[8 + k] = fma2(fac, vrc[k], v[8 + k]);
        acc = fma2(d, vrc[k], acc);
        buf[k] = fma2(al, d, buf[k]);
    }
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

// PX pixels per thread (same column, PX adjacent rows); 256 / PX threads per tile.  PX = 4 halves the number of
// (warp, Gaussian) reductions and staged-record reads per pixel compared with PX = 2.
template <int CDIM, int DOUT, bool ED, int PX>
__global__ void __launch_bounds__(256 / PX, PX == 4 ? (CDIM == 4 ? 9 : 8) : 1)
k_blend_bwd(const float2 *__restrict__ means2d, const float4 *__restrict__ geo, const float4 *__restrict__ colpack,
            const int32_t *__restrict__ offsets, const int32_t *__restrict__ flatten_ids, long long M, int W, int H,
            int tile_w, int tile_h, const float *__restrict__ render, const float *__restrict__ alpha_in,
            const int32_t *__restrict__ last_ids, const float *__restrict__ v_render,
            const float *__restrict__ v_alpha, float *__restrict__ v_xyabs, float *__restrict__ v_geo,
            float *__restrict__ v_colpack) {
    constexpr int THREADS = 256 / PX;
    constexpr int WARPS = THREADS / 32;
    constexpr int BATCH = 128;            // staged Gaussians per round
    constexpr int SPT = BATCH / THREADS;  // staged per thread
    constexpr int CQ = CDIM / 4;
    constexpr int NQUAD = 2 + CQ;  // float4 groups flushed per Gaussian
    __shared__ float4 s_q[BATCH];
    __shared__ float2 s_c[BATCH];
    __shared__ int s_idx[BATCH];
    __shared__ int s_gid[BATCH];
    __shared__ float4 s_col[BATCH][CQ];
    __shared__ int s_wcnt[SPT][WARPS];
    __shared__ __align__(16) float s_acc[WARPS][BATCH][16];

    const int tile = blockIdx.x;
    const int ti = tile / tile_w, tj = tile - ti * tile_w;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int x = tj * 16 + (lane & 15);
    const int ybase = ti * 16 + 2 * PX * warp + PX * (lane >> 4);
    const float px = (float)x + 0.5f;
    const float rx0 = (float)(tj * 16) + 0.5f, ry0 = (float)(ti * 16) + 0.5f;
    const float rx1 = (float)min(tj * 16 + 16, W) - 0.5f, ry1 = (float)min(ti * 16 + 16, H) - 0.5f;
    const int start = offsets[tile];
    const int end = (tile == tile_w * tile_h - 1) ? (int)M : offsets[tile + 1];

    static_assert(PX % 2 == 0, "pixels travel in pairs");
    constexpr int NP = PX / 2;  // pixel pairs per thread: rows (ybase + 2 q, ybase + 2 q + 1) are the halves of a float2
    float2 T[NP], Tfvra[NP], npy[NP];
    float2 vrc[NP][CDIM], buf[NP][CDIM];
    int bin[PX];
    int maxbin = -1;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        float Tj = 1.f, Tfj = 1.f, vraj = 0.f;
        float vrcj[CDIM];
#pragma unroll
        for (int k = 0; k < CDIM; ++k) vrcj[k] = 0.f;
        bin[j] = -1;
        if (x < W && ybase + j < H) {
            const size_t pid = (size_t)(ybase + j) * W + x;
            load_pixel_cotangent<CDIM, DOUT, ED>(pid, render, alpha_in, v_render, v_alpha, vrcj, vraj, Tfj);
            Tj = Tfj;
            bin[j] = last_ids[pid];
        }
        maxbin = max(maxbin, bin[j]);
        const float npyj = -((float)(ybase + j) + 0.5f);
        if (j & 1) {
            T[j / 2].y = Tj; Tfvra[j / 2].y = Tfj * vraj; npy[j / 2].y = npyj;
#pragma unroll
            for (int k = 0; k < CDIM; ++k) { vrc[j / 2][k].y = vrcj[k]; buf[j / 2][k].y = 0.f; }
        } else {
            T[j / 2].x = Tj; Tfvra[j / 2].x = Tfj * vraj; npy[j / 2].x = npyj;
#pragma unroll
            for (int k = 0; k < CDIM; ++k) { vrc[j / 2][k].x = vrcj[k]; buf[j / 2][k].x = 0.f; }
        }
    }
    // last sorted index any pixel of this tile blended
    maxbin = __reduce_max_sync(0xffffffffu, maxbin);
    if (lane == 0) s_wcnt[0][warp] = maxbin;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < WARPS; ++w) maxbin = max(maxbin, s_wcnt[0][w]);
    const int hi0 = min(end - 1, maxbin);

    for (int hi = hi0; hi >= start; hi -= BATCH) {
        __syncthreads();  // previous batch fully flushed before its buffers are reused
        // back to front: slot order == descending sorted index; thread stages entries hi - (q * THREADS + tid)
        bool keep[SPT];
        int g[SPT], idx[SPT];
        float4 q[SPT];
        float2 c[SPT];
        unsigned bal[SPT];
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            idx[u] = hi - (u * THREADS + (int)threadIdx.x);
            keep[u] = false;
            g[u] = 0;
            q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            c[u] = make_float2(0.f, 0.f);
            if (idx[u] >= start) {
                g[u] = flatten_ids[idx[u]];
                const float2 m = means2d[g[u]];
                const float4 ge = geo[g[u]];
                q[u] = make_float4(m.x, m.y, 0.5f * B2S_LOG2E * ge.x, B2S_LOG2E * ge.y);
                c[u] = make_float2(0.5f * B2S_LOG2E * ge.z, ge.w);
                keep[u] = tile_keep(m.x, m.y, q[u].z, q[u].w, c[u].x, c[u].y, rx0, ry0, rx1, ry1);
            }
            bal[u] = __ballot_sync(0xffffffffu, keep[u]);
            if (lane == 0) s_wcnt[u][warp] = __popc(bal[u]);
        }
        __syncthreads();
        int total = 0;
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            int off = total;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const int n = s_wcnt[u][w];
                off += (w < warp) ? n : 0;
                total += n;
            }
            if (keep[u]) {
                const int slot = off + __popc(bal[u] & lt);
                s_q[slot] = q[u];
                s_c[slot] = c[u];
                s_idx[slot] = idx[u];
                s_gid[slot] = g[u];
#pragma unroll
                for (int k = 0; k < CQ; ++k) s_col[slot][k] = colpack[(size_t)g[u] * CQ + k];
            }
        }
        __syncthreads();

        for (int t = 0; t < total; ++t) {
            const float4 sq = s_q[t];
            const float2 sc = s_c[t];
            const int id = s_idx[t];
            const float dx = sq.x - px;
            float2 dy[NP], al[NP], g[NP];
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
                any_ok = any_ok || ok0 || ok1;
            }
            if (!__any_sync(0xffffffffu, any_ok)) {
                if (lane < 16) s_acc[warp][t][lane] = 0.f;
                continue;
            }
            float2 v2[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v2[k] = make_float2(0.f, 0.f);
            float col[CDIM];
#pragma unroll
            for (int k = 0; k < CQ; ++k) {
                const float4 cv = s_col[t][k];
                col[4 * k] = cv.x; col[4 * k + 1] = cv.y; col[4 * k + 2] = cv.z; col[4 * k + 3] = cv.w;
            }
            // raw conic (a, b, c) from the log2-domain coefficients: a = 2 A ln2, b = B ln2, c = 2 C ln2
            const float ca = 2.f * B2S_LN2 * sq.z, cb = B2S_LN2 * sq.w, cc = 2.f * B2S_LN2 * sc.x;
#pragma unroll
            for (int q = 0; q < NP; ++q)
                pair_grad<CDIM>(v2, T[q], buf[q], vrc[q], Tfvra[q], col, ca, cb, cc, sc.y, dx, dy[q], al[q], g[q]);
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = k < 8 + CDIM ? v2[k].x + v2[k].y : 0.f;
            const float r = warp_reduce16_transposed<8 + CDIM>(v, lane);
            if (!(lane & 1)) s_acc[warp][t][lane >> 1] = r;
        }
        __syncthreads();

        // flush: item = (slot, quad); consecutive threads read consecutive float4s of s_acc (conflict free)
#pragma unroll
        for (int j = 0; j < 4 * BATCH / THREADS; ++j) {
            const int item = (int)threadIdx.x + THREADS * j;
            const int slot = item >> 2, quad = item & 3;
            if (slot < total && quad < NQUAD) {
                float4 s = reinterpret_cast<const float4 *>(&s_acc[0][slot][0])[quad];
#pragma unroll
                for (int w = 1; w < WARPS; ++w) {
                    const float4 o = reinterpret_cast<const float4 *>(&s_acc[w][slot][0])[quad];
                    s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
                }
                // undo the accumulation convention of pair_grad: quad 0 = (-v_x, -v_y, |v_x|, |v_y|),
                // quad 1 = (-2 v_conic_a, -v_conic_b, -2 v_conic_c, v_opacity), quads >= 2 = v_colour
                if (quad == 0) { s.x = -s.x; s.y = -s.y; }
                if (quad == 1) { s.x = -0.5f * s.x; s.y = -s.y; s.z = -0.5f * s.z; }
                if (s.x != 0.f || s.y != 0.f || s.z != 0.f || s.w != 0.f) {
                    const int gid = s_gid[slot];
                    float *dst = quad == 0 ? v_xyabs + (size_t)gid * 4
                               : quad == 1 ? v_geo + (size_t)gid * 4
                                           : v_colpack + (size_t)gid * CDIM + (quad - 2) * 4;
                    red_add_v4(dst, s.x, s.y, s.z, s.w);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
template <int CDIM, int DOUT, bool ED>
static int launch_fwd(const float *means2d, const float *geo, const float *colpack, const int32_t *offsets,
                      const int32_t *flatten_ids, long long M, int W, int H, int tile_w, int tile_h, float *render,
                      float *alpha, int32_t *last_ids, cudaStream_t st) {
    k_blend_fwd<CDIM, DOUT, ED><<<tile_w * tile_h, BL_THREADS, 0, st>>>(
        (const float2 *)means2d, (const float4 *)geo, (const float4 *)colpack, offsets, flatten_ids, M, W, H, tile_w,
        tile_h, render, alpha, last_ids);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
template <int CDIM, int DOUT, bool ED>
static int launch_bwd(const float *means2d, const float *geo, const float *colpack, const int32_t *offsets,
                      const int32_t *flatten_ids, long long M, int W, int H, int tile_w, int tile_h,
                      const float *render, const float *alpha, const int32_t *last_ids, const float *v_render,
                      const float *v_alpha, float *v_xyabs, float *v_geo, float *v_colpack, cudaStream_t st) {
    // pixels per thread in the backward (see k_blend_bwd); B2S_BWD_PX=2 selects the 128-thread variant (tuning knob)
    static const int px = [] {
        const char *e = getenv("B2S_BWD_PX");
        const int v = e ? atoi(e) : 4;
        return (v == 2 || v == 8) ? v : 4;
    }();
    if (px == 2)
        k_blend_bwd<CDIM, DOUT, ED, 2><<<tile_w * tile_h, 128, 0, st>>>(
            (const float2 *)means2d, (const float4 *)geo, (const float4 *)colpack, offsets, flatten_ids, M, W, H, tile_w,
            tile_h, render, alpha, last_ids, v_render, v_alpha, v_xyabs, v_geo, v_colpack);
    else if (px == 8)
        k_blend_bwd<CDIM, DOUT, ED, 8><<<tile_w * tile_h, 32, 0, st>>>(
            (const float2 *)means2d, (const float4 *)geo, (const float4 *)colpack, offsets, flatten_ids, M, W, H, tile_w,
            tile_h, render, alpha, last_ids, v_render, v_alpha, v_xyabs, v_geo, v_colpack);
    else
        k_blend_bwd<CDIM, DOUT, ED, 4><<<tile_w * tile_h, 64, 0, st>>>(
            (const float2 *)means2d, (const float4 *)geo, (const float4 *)colpack, offsets, flatten_ids, M, W, H, tile_w,
            tile_h, render, alpha, last_ids, v_render, v_alpha, v_xyabs, v_geo, v_colpack);
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

extern "C" int b2s_blend_fwd(const float *means2d, const float *geo, const float *colpack,
                             const int32_t *isect_offsets, const int32_t *flatten_ids, long long M, int W, int H,
                             int tile_w, int tile_h, int cdim, int d_out, int expected_depth, float *render,
                             float *alpha, int32_t *last_ids, b2s_stream_t stream) {
    if (W <= 0 || H <= 0 || M < 0 || tile_w * 16 < W || tile_h * 16 < H) return B2S_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    B2S_DISPATCH(launch_fwd, means2d, geo, colpack, isect_offsets, flatten_ids, M, W, H, tile_w, tile_h, render,
                 alpha, last_ids, st);
}

extern "C" int b2s_blend_bwd(const float *means2d, const float *geo, const float *colpack,
                             const int32_t *isect_offsets, const int32_t *flatten_ids, long long M, int W, int H,
                             int tile_w, int tile_h, int cdim, int d_out, int expected_depth, const float *render,
                             const float *alpha, const int32_t *last_ids, const float *v_render,
                             const float *v_alpha, float *v_xyabs, float *v_geo, float *v_colpack,
                             b2s_stream_t stream) {
    if (W <= 0 || H <= 0 || M < 0 || tile_w * 16 < W || tile_h * 16 < H) return B2S_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    B2S_DISPATCH(launch_bwd, means2d, geo, colpack, isect_offsets, flatten_ids, M, W, H, tile_w, tile_h, render,
                 alpha, last_ids, v_render, v_alpha, v_xyabs, v_geo, v_colpack, st);
}
