// Projection + EWA covariance, forward and backward (sm_100a).
//
// Replaces upstream gsplat v1.4.0 fully_fused_projection_{fwd,bwd}_kernel as reached from
// mtgs/scene_model/mtgs_scene_graph.py:641-662, and fuses into the same pass
//   - upstream isect_tiles pass 1 (tile count per Gaussian, A.2),
//   - the depth sort key,
//   - torch glue that upstream does outside the kernel: opacities*compensations, cat(colors, depths),
//     zero padding of the colour channels.
//
// The forward is HBM-bound streaming work: one thread per Gaussian, SoA reads that a warp covers with
// fully used 32 B sectors (means 12 B, quats 16 B, scales 12 B per lane -> contiguous per warp).
//
// CANONICAL OP ORDER: every fp32 operation of the forward that feeds a discrete decision
// (radius, cull, tile rectangle, depth key) is a single IEEE round-to-nearest op written with
// __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn in a fixed order, never contracted to FMA, so
// radii / tile counts / keys are reproducible bit for bit by an independent fp32 implementation.
#include "common.cuh"

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define DIV(a, b) __fdiv_rn((a), (b))
#define SQRT(a) __fsqrt_rn((a))
#define DOT3(a0, b0, a1, b1, a2, b2) ADD(ADD(MUL(a0, b0), MUL(a1, b1)), MUL(a2, b2))

struct CamParams {
    float R[9];
    float t[3];
    float fx, fy, cx, cy;
    float lim_x_pos, lim_x_neg, lim_y_pos, lim_y_neg;
    float Wf, Hf;
};

__device__ __forceinline__ void load_camera(const float *__restrict__ viewmat, const float *__restrict__ K,
                                            int W, int H, CamParams &c) {
    c.R[0] = viewmat[0]; c.R[1] = viewmat[1]; c.R[2] = viewmat[2];
    c.R[3] = viewmat[4]; c.R[4] = viewmat[5]; c.R[5] = viewmat[6];
    c.R[6] = viewmat[8]; c.R[7] = viewmat[9]; c.R[8] = viewmat[10];
    c.t[0] = viewmat[3]; c.t[1] = viewmat[7]; c.t[2] = viewmat[11];
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    c.Wf = (float)W; c.Hf = (float)H;
    float tan_fovx = DIV(MUL(0.5f, c.Wf), c.fx);
    float tan_fovy = DIV(MUL(0.5f, c.Hf), c.fy);
    c.lim_x_pos = ADD(DIV(SUB(c.Wf, c.cx), c.fx), MUL(0.3f, tan_fovx));
    c.lim_x_neg = ADD(DIV(c.cx, c.fx), MUL(0.3f, tan_fovx));
    c.lim_y_pos = ADD(DIV(SUB(c.Hf, c.cy), c.fy), MUL(0.3f, tan_fovy));
    c.lim_y_neg = ADD(DIV(c.cy, c.fy), MUL(0.3f, tan_fovy));
}

// normalised-quaternion rotation matrix, row-major (upstream quat_to_rotmat)
__device__ __forceinline__ void quat_to_rotmat_rn(float w, float x, float y, float z, float *R) {
    float inv_norm = DIV(1.0f, SQRT(ADD(ADD(ADD(MUL(x, x), MUL(y, y)), MUL(z, z)), MUL(w, w))));
    x = MUL(x, inv_norm); y = MUL(y, inv_norm); z = MUL(z, inv_norm); w = MUL(w, inv_norm);
    float x2 = MUL(x, x), y2 = MUL(y, y), z2 = MUL(z, z);
    float xy = MUL(x, y), xz = MUL(x, z), yz = MUL(y, z);
    float wx = MUL(w, x), wy = MUL(w, y), wz = MUL(w, z);
    R[0] = SUB(1.f, MUL(2.f, ADD(y2, z2))); R[1] = MUL(2.f, SUB(xy, wz)); R[2] = MUL(2.f, ADD(xz, wy));
    R[3] = MUL(2.f, ADD(xy, wz)); R[4] = SUB(1.f, MUL(2.f, ADD(x2, z2))); R[5] = MUL(2.f, SUB(yz, wx));
    R[6] = MUL(2.f, SUB(xz, wy)); R[7] = MUL(2.f, ADD(yz, wx)); R[8] = SUB(1.f, MUL(2.f, ADD(x2, y2)));
}

// camera-space covariance Sc = R (Rq S)(Rq S)^T R^T, all nine entries in canonical order
__device__ __forceinline__ void covar_cam_rn(const float *R, const float *Rq, float s0, float s1, float s2,
                                             float *Sigma, float *Sc) {
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        M[i * 3 + 0] = MUL(Rq[i * 3 + 0], s0);
        M[i * 3 + 1] = MUL(Rq[i * 3 + 1], s1);
        M[i * 3 + 2] = MUL(Rq[i * 3 + 2], s2);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Sigma[i * 3 + j] = DOT3(M[i * 3 + 0], M[j * 3 + 0], M[i * 3 + 1], M[j * 3 + 1], M[i * 3 + 2], M[j * 3 + 2]);
    float A[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            A[i * 3 + j] = DOT3(R[i * 3 + 0], Sigma[0 * 3 + j], R[i * 3 + 1], Sigma[1 * 3 + j], R[i * 3 + 2], Sigma[2 * 3 + j]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Sc[i * 3 + j] = DOT3(A[i * 3 + 0], R[j * 3 + 0], A[i * 3 + 1], R[j * 3 + 1], A[i * 3 + 2], R[j * 3 + 2]);
}

template <int CDIM>
__global__ void __launch_bounds__(256, 6)
k_project_fwd(const float *__restrict__ means, const float *__restrict__ quats, const float *__restrict__ scales,
              const float *__restrict__ opacities, const float *__restrict__ colors_in,
              const float *__restrict__ viewmat, const float *__restrict__ K, int N, int W, int H, int tile_w,
              int tile_h, float eps2d, float near_plane, float far_plane, float radius_clip, int calc_comp,
              int d_in, int with_depth, int32_t *__restrict__ radii, float2 *__restrict__ means2d,
              float *__restrict__ depths, float4 *__restrict__ geo, float *__restrict__ comps,
              float *__restrict__ colpack, int32_t *__restrict__ tiles_per_gauss,
              uint32_t *__restrict__ sort_keys, int2 *__restrict__ tile_rects, int2 *__restrict__ tight_rects,
              int rg_shift, int cg_shift, unsigned long long *__restrict__ totals /* [B2S_N_TOTALS], pre-zeroed */,
              float4 *__restrict__ bwd_arena /* [N * (2 + CDIM / 4)] or null: zero-filled for the blend backward */,
              int use_tma /* all input arrays 16-byte aligned */) {
    // The CTA's 256 input rows (means | quats | scales | opacities | colours: contiguous runs of the SoA arrays) are
    // staged by TMA bulk copies issued by ONE thread onto an mbarrier (cp.async.bulk, SASS UBLKCP): every byte of
    // the tile is in flight before any thread needs it, instead of each thread waiting for its own strided loads
    // (long-scoreboard stalls were ~30 % of this kernel's samples).  The camera constants (six divisions) are
    // computed once per CTA by another warp meanwhile.
    __shared__ int s_tot[B2S_N_TOTALS];
    __shared__ CamParams s_cam;
    __shared__ __align__(8) unsigned long long s_bar;
    extern __shared__ __align__(128) float s_in[];  // [768 means | 1024 quats | 768 scales | 256 opac | 256 d_in colours]
    float *s_means = s_in, *s_quats = s_in + 768, *s_scales = s_in + 1792, *s_opac = s_in + 2560, *s_cols = s_in + 2816;
    const int g0 = blockIdx.x * 256;
    const bool staged = use_tma && g0 + 256 <= N;  // CTA-uniform; the last, partial CTA loads directly
    if (threadIdx.x < B2S_N_TOTALS) s_tot[threadIdx.x] = 0;
    if (threadIdx.x == 0 && staged) {
        b2s_mbar_init(&s_bar, 1);
        b2s_mbar_expect_tx(&s_bar, (unsigned)((11 + d_in) * 1024));
        b2s_bulk_g2s(s_means, means + (size_t)g0 * 3, 3072, &s_bar);
        b2s_bulk_g2s(s_quats, quats + (size_t)g0 * 4, 4096, &s_bar);
        b2s_bulk_g2s(s_scales, scales + (size_t)g0 * 3, 3072, &s_bar);
        b2s_bulk_g2s(s_opac, opacities + g0, 1024, &s_bar);
        if (d_in > 0) b2s_bulk_g2s(s_cols, colors_in + (size_t)g0 * d_in, (unsigned)(d_in * 1024), &s_bar);
    }
    if (threadIdx.x == 32) load_camera(viewmat, K, W, H, s_cam);
    __syncthreads();
    const int g_raw = g0 + threadIdx.x;
    const bool valid = g_raw < N;
    const int g = valid ? g_raw : N - 1;  // out-of-range threads recompute the last Gaussian and write nothing
    const CamParams &cam = s_cam;
    if (staged) b2s_mbar_wait(&s_bar, 0);
    const int t3 = 3 * threadIdx.x;

    float p0, p1, p2;
    if (staged) { p0 = s_means[t3]; p1 = s_means[t3 + 1]; p2 = s_means[t3 + 2]; }
    else { p0 = means[3 * g]; p1 = means[3 * g + 1]; p2 = means[3 * g + 2]; }
    float pc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        pc[i] = ADD(DOT3(cam.R[i * 3 + 0], p0, cam.R[i * 3 + 1], p1, cam.R[i * 3 + 2], p2), cam.t[i]);

    int32_t radius_i = 0;
    int32_t ntiles = 0;
    int2 rect = make_int2(0, 0);  // (x0 | x1 << 16, y0 | y1 << 16) in tiles, max exclusive
    int tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0;  // tight rectangle
    float op = 0.f;                          // opacity * compensation
    float mx = 0.f, my = 0.f, ca = 0.f, cb = 0.f, cc = 0.f, comp = 1.f;
    float z = pc[2];
    bool ok = !(z < near_plane || z > far_plane);
    if (ok) {
        const float4 q = staged ? reinterpret_cast<const float4 *>(s_quats)[threadIdx.x]
                                : reinterpret_cast<const float4 *>(quats)[g];
        float s0, s1, s2;
        if (staged) { s0 = s_scales[t3]; s1 = s_scales[t3 + 1]; s2 = s_scales[t3 + 2]; }
        else { s0 = scales[3 * g]; s1 = scales[3 * g + 1]; s2 = scales[3 * g + 2]; }
        float Rq[9], Sigma[9], Sc[9];
        quat_to_rotmat_rn(q.x, q.y, q.z, q.w, Rq);
        covar_cam_rn(cam.R, Rq, s0, s1, s2, Sigma, Sc);

        float x = pc[0], y = pc[1];
        float rz = DIV(1.0f, z);
        float rz2 = MUL(rz, rz);
        float tx = MUL(z, fminf(cam.lim_x_pos, fmaxf(-cam.lim_x_neg, MUL(x, rz))));
        float ty = MUL(z, fminf(cam.lim_y_pos, fmaxf(-cam.lim_y_neg, MUL(y, rz))));
        float j00 = MUL(cam.fx, rz), j02 = -MUL(MUL(cam.fx, tx), rz2);
        float j11 = MUL(cam.fy, rz), j12 = -MUL(MUL(cam.fy, ty), rz2);
        float B00 = ADD(MUL(j00, Sc[0]), MUL(j02, Sc[6]));
        float B01 = ADD(MUL(j00, Sc[1]), MUL(j02, Sc[7]));
        float B02 = ADD(MUL(j00, Sc[2]), MUL(j02, Sc[8]));
        float B11 = ADD(MUL(j11, Sc[4]), MUL(j12, Sc[7]));
        float B12 = ADD(MUL(j11, Sc[5]), MUL(j12, Sc[8]));
        float c00 = ADD(MUL(B00, j00), MUL(B02, j02));
        float c01 = ADD(MUL(B01, j11), MUL(B02, j12));
        float c11 = ADD(MUL(B11, j11), MUL(B12, j12));
        mx = ADD(MUL(MUL(cam.fx, x), rz), cam.cx);
        my = ADD(MUL(MUL(cam.fy, y), rz), cam.cy);

        float det_orig = SUB(MUL(c00, c11), MUL(c01, c01));
        c00 = ADD(c00, eps2d);
        c11 = ADD(c11, eps2d);
        float det = SUB(MUL(c00, c11), MUL(c01, c01));
        comp = SQRT(fmaxf(0.0f, DIV(det_orig, det)));
        ok = det > 0.0f;
        if (ok) {
            float inv_det = DIV(1.0f, det);
            ca = MUL(c11, inv_det);
            cb = MUL(-c01, inv_det);
            cc = MUL(c00, inv_det);
            float b = MUL(0.5f, ADD(c00, c11));
            float v1 = ADD(b, SQRT(fmaxf(0.01f, SUB(MUL(b, b), det))));
            float radius = ceilf(MUL(3.0f, SQRT(v1)));
            ok = !(radius <= radius_clip);
            ok = ok && !(ADD(mx, radius) <= 0.0f || SUB(mx, radius) >= cam.Wf || ADD(my, radius) <= 0.0f ||
                         SUB(my, radius) >= cam.Hf);
            if (ok) {
                radius_i = (int32_t)radius;
                // upstream isect_tiles pass 1 on the integer radius (tile size 16: divisions are exact)
                float tr = DIV((float)radius_i, 16.0f);
                float txc = DIV(mx, 16.0f), tyc = DIV(my, 16.0f);
                float fx0 = floorf(SUB(txc, tr)), fy0 = floorf(SUB(tyc, tr));
                float fx1 = ceilf(ADD(txc, tr)), fy1 = ceilf(ADD(tyc, tr));
                int x0 = fx0 <= 0.f ? 0 : (fx0 >= (float)tile_w ? tile_w : (int)fx0);
                int y0 = fy0 <= 0.f ? 0 : (fy0 >= (float)tile_h ? tile_h : (int)fy0);
                int x1 = fx1 <= 0.f ? 0 : (fx1 >= (float)tile_w ? tile_w : (int)fx1);
                int y1 = fy1 <= 0.f ? 0 : (fy1 >= (float)tile_h ? tile_h : (int)fy1);
                ntiles = (y1 - y0) * (x1 - x0);
                rect = make_int2(x0 | (x1 << 16), y0 | (y1 << 16));
                // ---- tight rectangle for the blend's own tile lists: tiles whose pixel centres the footprint
                // {alpha >= 1/255} = {sigma <= ln(255 opacity)} can reach, intersected with upstream's 3-sigma
                // rectangle.  Not an upstream output (info["tiles_per_gauss"] / flatten_ids stay upstream's): it only
                // removes (Gaussian, tile) pairs that contribute to no pixel.  Extents: |dx| <= sqrt(2 t c00), |dy| <=
                // sqrt(2 t c11) with c00, c11 the blurred 2-D covariance.  The blend evaluates the STORED conic, whose
                // ellipse differs from the ideal one by a relative ~2^-24 kappa in its extents (kappa = c00 c11 / det,
                // the cancellation in the inverse), and it evaluates sigma with an fp32 error ~2^-22 (|A|+|B|+|C|) R^2:
                // both are covered by the slack.  The exact per-tile test follows while the blend stages a batch.
                op = staged ? s_opac[threadIdx.x] : opacities[g];
                if (calc_comp) op = MUL(op, comp);
                if (op >= 0.0039f) {  // below 1/255 the Gaussian reaches no pixel (same gate as tile_keep)
                    const float Rb = radius + 16.0f;
                    const float t2 = __log2f(op * 255.0f) + 0.02f +
                                     4e-7f * B2S_LOG2E * (0.5f * ca + fabsf(cb) + 0.5f * cc) * Rb * Rb;  // log2 units
                    const float kappa = fminf(c00 * c11 * inv_det, 1e12f);
                    const float infl = (1.0f + 2.5e-7f * kappa) * (2.0f * B2S_LN2) * t2;  // 2 t (ln units), inflated
                    const float hx = fminf(sqrtf(infl * c00) * 1.0001f + 0.01f, 1e6f);
                    const float hy = fminf(sqrtf(infl * c11) * 1.0001f + 0.01f, 1e6f);
                    // tile k holds the pixel centres 16 k + 0.5 .. 16 k + 15.5
                    const float mxc = fminf(fmaxf(mx, -1e6f), 1e6f), myc = fminf(fmaxf(my, -1e6f), 1e6f);
                    tx0 = max(x0, (int)ceilf((mxc - hx - 15.5f) * 0.0625f));
                    tx1 = min(x1, (int)floorf((mxc + hx - 0.5f) * 0.0625f) + 1);
                    ty0 = max(y0, (int)ceilf((myc - hy - 15.5f) * 0.0625f));
                    ty1 = min(y1, (int)floorf((myc + hy - 0.5f) * 0.0625f) + 1);
                    if (tx1 <= tx0 || ty1 <= ty0) tx0 = tx1 = ty0 = ty1 = 0;
                }
            }
        }
    }
    const int2 trect = make_int2(tx0 | (tx1 << 16), ty0 | (ty1 << 16));
    // ---- list sizes of the blend's tile lists (tilelists.cu) over the tight rectangles + the number of visible
    // Gaussians: one integer warp reduction (redux.sync) per value, at most 5 atomics per CTA
    {
        int v[B2S_N_TOTALS];
        const int th_ = ty1 - ty0, tw_ = tx1 - tx0;
        v[0] = valid ? th_ * tw_ : 0;
        v[1] = valid ? th_ : 0;
        v[2] = (valid && th_ > 0) ? ((ty1 - 1) >> rg_shift) - (ty0 >> rg_shift) + 1 : 0;
        v[3] = (valid && tw_ > 0) ? th_ * (((tx1 - 1) >> cg_shift) - (tx0 >> cg_shift) + 1) : 0;
        v[4] = (valid && radius_i > 0) ? 1 : 0;
#pragma unroll
        for (int k = 0; k < B2S_N_TOTALS; ++k) {
            const int sum = __reduce_add_sync(0xffffffffu, v[k]);
            if ((threadIdx.x & 31) == 0 && sum != 0) atomicAdd(&s_tot[k], sum);
        }
    }
    __syncthreads();
    if (threadIdx.x < B2S_N_TOTALS && s_tot[threadIdx.x] != 0)
        atomicAdd(totals + threadIdx.x, (unsigned long long)s_tot[threadIdx.x]);
    if (!valid) return;
    if (bwd_arena != nullptr) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        bwd_arena[g] = z4;
        bwd_arena[(size_t)N + g] = z4;
#pragma unroll
        for (int k = 0; k < CDIM / 4; ++k) bwd_arena[2 * (size_t)N + (size_t)g * (CDIM / 4) + k] = z4;
    }
    radii[g] = radius_i;
    tiles_per_gauss[g] = ntiles;
    tile_rects[g] = rect;
    tight_rects[g] = trect;
    sort_keys[g] = radius_i > 0 ? __float_as_uint(z) : 0xFFFFFFFFu;
    if (radius_i > 0) {
        if (calc_comp) comps[g] = comp;
        means2d[g] = make_float2(mx, my);
        depths[g] = z;
        geo[g] = make_float4(ca, cb, cc, op);
        float cp[CDIM];
#pragma unroll
        for (int k = 0; k < CDIM; ++k) cp[k] = 0.f;
#pragma unroll
        for (int k = 0; k < CDIM; ++k)
            if (k < d_in) cp[k] = staged ? s_cols[threadIdx.x * d_in + k] : colors_in[(size_t)g * d_in + k];
        if (with_depth) {
#pragma unroll
            for (int k = 0; k < CDIM; ++k)
                if (k == d_in) cp[k] = z;
        }
        float4 *dst = reinterpret_cast<float4 *>(colpack + (size_t)g * CDIM);
#pragma unroll
        for (int k = 0; k < CDIM / 4; ++k) dst[k] = make_float4(cp[4 * k], cp[4 * k + 1], cp[4 * k + 2], cp[4 * k + 3]);
    } else {
        // culled rows are never read by the blend; define them (zeros) so callers need no memset pass
        if (calc_comp) comps[g] = 0.f;
        means2d[g] = make_float2(0.f, 0.f);
        depths[g] = 0.f;
        geo[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 *dst = reinterpret_cast<float4 *>(colpack + (size_t)g * CDIM);
#pragma unroll
        for (int k = 0; k < CDIM / 4; ++k) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// List sizes of a tile-list build over arbitrary rectangles (used for upstream's lists, which are only built when a
// caller reads info["flatten_ids"] / ["isect_offsets"] / ["isect_ids"]): totals = {M, S, E1, E3, n with tiles}.
__global__ void __launch_bounds__(256)
k_rect_totals(const int2 *__restrict__ rects, int N, int rg_shift, int cg_shift, unsigned long long *__restrict__ totals) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    long long v[5] = {0, 0, 0, 0, 0};
    if (g < N) {
        const int2 rc = rects[g];
        const int x0 = rc.x & 0xffff, x1 = (rc.x >> 16) & 0xffff, y0 = rc.y & 0xffff, y1 = (rc.y >> 16) & 0xffff;
        const int h = max(0, y1 - y0), w = max(0, x1 - x0);
        v[0] = (long long)h * w;
        v[1] = h;
        v[2] = h > 0 ? ((y1 - 1) >> rg_shift) - (y0 >> rg_shift) + 1 : 0;
        v[3] = w > 0 ? (long long)h * (((x1 - 1) >> cg_shift) - (x0 >> cg_shift) + 1) : 0;
        v[4] = (h > 0 && w > 0) ? 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0 && v[k] != 0) atomicAdd(totals + k, (unsigned long long)v[k]);
    }
}

extern "C" int b2s_bin_rect_totals(const int32_t *rects, int N, int tile_w, int tile_h, int64_t *totals,
                                   b2s_stream_t stream) {
    if (N < 0 || totals == nullptr) return B2S_ERR_ARG;
    int rg_shift = 0, cg_shift = 0;
    if (b2s_tl_shifts(tile_w, tile_h, &rg_shift, &cg_shift) != B2S_OK) return B2S_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(totals, 0, 5 * sizeof(int64_t), st);
    if (N == 0) return B2S_OK;
    k_rect_totals<<<b2s_div_up(N, 256), 256, 0, st>>>((const int2 *)rects, N, rg_shift, cg_shift,
                                                     (unsigned long long *)totals);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

// --------------------------------------------------------------------------------------------
// backward.  Tolerance-checked, so FMA contraction is allowed here.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const float *A, const float *B, float *C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}
__device__ __forceinline__ void mat3_mul_bt(const float *A, const float *B, float *C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
__device__ __forceinline__ void mat3_mul_at(const float *A, const float *B, float *C) {  // A^T B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[0 * 3 + i] * B[0 * 3 + j] + A[1 * 3 + i] * B[1 * 3 + j] + A[2 * 3 + i] * B[2 * 3 + j];
}

// EXCH = false: gradients are written to this GPU's v_means / v_quats / v_scales / v_opacities.
// EXCH = true (multi-GPU): the CTA's 256 gradient rows (means 3, quats 4, scales 3, opacity 1, colours
// d_in floats per Gaussian) are transposed through shared memory and stored with 16-byte coalesced stores straight
// into the staging slot of the rank that OWNS these rows (peer memory over NVLink) -- the reduce-scatter half of
// the shared-gradient all-reduce happens while the projection backward is still computing, and this GPU never
// writes (or later re-reads) a local copy of its partial gradient.  The peer stores are fire-and-forget (no
// per-CTA system fence: that exposed one NVLink round trip per CTA); the kernel boundary completes them and a
// one-warp kernel then raises this rank's flag at every peer (exchange.cu, which also holds the reduce +
// all-gather half).
template <int CDIM, bool EXCH>
__global__ void __launch_bounds__(256, 3)
k_project_bwd(const float *__restrict__ means, const float *__restrict__ quats, const float *__restrict__ scales,
              const float *__restrict__ opacities, const float *__restrict__ viewmat, const float *__restrict__ K,
              int N, int W, int H, float eps2d, int calc_comp, int d_in, int with_depth,
              const int32_t *__restrict__ radii, const float4 *__restrict__ geo, const float *__restrict__ comps,
              const float *__restrict__ v_means2d, int v_m2d_stride, const float4 *__restrict__ v_geo,
              const float *__restrict__ v_colpack, float *__restrict__ v_means, float4 *__restrict__ v_quats,
              float *__restrict__ v_scales, float *__restrict__ v_opacities, float *__restrict__ v_colors_out,
              float *__restrict__ v_viewmat, const B2sExchange ex, int use_tma /* every staged array 16-byte aligned */) {
    // Inputs of the CTA's 256 rows staged by TMA bulk copies on an mbarrier, camera constants once per CTA (see
    // k_project_fwd).  Staged: means, quats, scales, geo, the blend's three gradient rows; the 4-byte-per-row arrays
    // (radii, opacities, compensations) are read directly (one coalesced 128-byte line per warp).
    __shared__ CamParams s_cam;
    __shared__ __align__(8) unsigned long long s_bar;
    extern __shared__ __align__(128) float s_dyn[];
    // layout (floats): means 768 | quats 1024 | scales 768 | geo 1024 | v_geo 1024 | v_m2d 256 * stride | v_colpack 256 * CDIM
    float *s_means = s_dyn, *s_quats = s_dyn + 768, *s_scales = s_dyn + 1792, *s_geo = s_dyn + 2560, *s_vgeo = s_dyn + 3584,
          *s_vm2d = s_dyn + 4608, *s_vcol = s_dyn + 4608 + 256 * v_m2d_stride;
    // EXCH: every rank walks the owners' shards starting behind its own (rank r begins with the rows rank r + 1 owns),
    // so at any moment the ranks store into DIFFERENT peers instead of all into the same one (NVLink ingress of one GPU)
    const int cta = EXCH ? (int)((blockIdx.x + (unsigned)ex.cta_rot) % gridDim.x) : (int)blockIdx.x;
    const int g0 = cta * 256;
    const bool staged = use_tma && g0 + 256 <= N;  // CTA-uniform
    if (threadIdx.x == 0 && staged) {
        b2s_mbar_init(&s_bar, 1);
        b2s_mbar_expect_tx(&s_bar, (unsigned)((18 + v_m2d_stride + CDIM) * 1024));
        b2s_bulk_g2s(s_means, means + (size_t)g0 * 3, 3072, &s_bar);
        b2s_bulk_g2s(s_quats, quats + (size_t)g0 * 4, 4096, &s_bar);
        b2s_bulk_g2s(s_scales, scales + (size_t)g0 * 3, 3072, &s_bar);
        b2s_bulk_g2s(s_geo, geo + g0, 4096, &s_bar);
        b2s_bulk_g2s(s_vgeo, v_geo + g0, 4096, &s_bar);
        b2s_bulk_g2s(s_vm2d, v_means2d + (size_t)g0 * v_m2d_stride, (unsigned)(v_m2d_stride * 1024), &s_bar);
        b2s_bulk_g2s(s_vcol, v_colpack + (size_t)g0 * CDIM, CDIM * 1024, &s_bar);
    }
    if (threadIdx.x == 32) load_camera(viewmat, K, W, H, s_cam);
    __syncthreads();
    float vR[9], vt[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) vR[k] = 0.f;
    vt[0] = vt[1] = vt[2] = 0.f;
    // The visible rows are COMPACTED over the CTA (see k_project_fwd): the VJP chain below runs on ceil(n_live / 32)
    // full warps; a culled row only gets its zeros written by its own thread.
    __shared__ int s_wlive[8];
    __shared__ short s_rows[256];
    const int g_own = g0 + threadIdx.x;
    const bool own_live = g_own < N && radii[g_own] > 0;
    {
        const unsigned bal = __ballot_sync(0xffffffffu, own_live);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) s_wlive[warp] = __popc(bal);
        __syncthreads();
        int off = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) off += (w < warp) ? s_wlive[w] : 0;
        if (own_live) s_rows[off + __popc(bal & ((1u << lane) - 1u))] = (short)threadIdx.x;
        __syncthreads();
    }
    int n_live = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) n_live += s_wlive[w];
    const bool live = (int)threadIdx.x < n_live;
    const int row = live ? (int)s_rows[threadIdx.x] : 0;
    const int g = g0 + row;
    if (staged) b2s_mbar_wait(&s_bar, 0);
    const int t3 = 3 * row;
    // this Gaussian's gradient row (zeros when it is culled)
    float o_means[3] = {0.f, 0.f, 0.f}, o_scales[3] = {0.f, 0.f, 0.f}, o_opac = 0.f;
    float4 o_quat = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        const CamParams &cam = s_cam;
        const float *R = cam.R;
        const float4 cg = staged ? reinterpret_cast<const float4 *>(s_geo)[row] : geo[g];
        const float2 gxy = staged ? *reinterpret_cast<const float2 *>(s_vm2d + row * v_m2d_stride)
                                  : *reinterpret_cast<const float2 *>(v_means2d + (size_t)g * v_m2d_stride);
        const float4 gge = staged ? reinterpret_cast<const float4 *>(s_vgeo)[row] : v_geo[g];
        float ia = cg.x, ib = cg.y, ic = cg.z;
        // inverse VJP: v_cov2d = -Minv * G * Minv with G = [[ga, gb/2],[gb/2, gc]]
        float ga = gge.x, gb = 0.5f * gge.y, gc = gge.z;
        float p00 = ia * ga + ib * gb, p01 = ia * gb + ib * gc;
        float p10 = ib * ga + ic * gb, p11 = ib * gb + ic * gc;
        float G[4];
        G[0] = -(p00 * ia + p01 * ib); G[1] = -(p00 * ib + p01 * ic);
        G[2] = -(p10 * ia + p11 * ib); G[3] = -(p10 * ib + p11 * ic);
        // opacity_eff = opacity * comp
        float v_op_eff = gge.w;
        float op = opacities[g];
        if (calc_comp) {
            float comp = comps[g];
            float v_comp = v_op_eff * op;
            o_opac = v_op_eff * comp;
            if (!EXCH) v_opacities[g] = o_opac;
            float det_conic = ia * ic - ib * ib;
            float v_sqr_comp = v_comp * 0.5f / (comp + 1e-6f);
            float one_minus = 1.0f - comp * comp;
            G[0] += v_sqr_comp * (one_minus * ia - eps2d * det_conic);
            G[1] += v_sqr_comp * (one_minus * ib);
            G[2] += v_sqr_comp * (one_minus * ib);
            G[3] += v_sqr_comp * (one_minus * ic - eps2d * det_conic);
        } else {
            o_opac = v_op_eff;
            if (!EXCH) v_opacities[g] = o_opac;
        }
        float v_depth = with_depth ? (staged ? s_vcol[row * CDIM + d_in] : v_colpack[(size_t)g * CDIM + d_in]) : 0.f;

        float p[3];
        if (staged) { p[0] = s_means[t3]; p[1] = s_means[t3 + 1]; p[2] = s_means[t3 + 2]; }
        else { p[0] = means[3 * g]; p[1] = means[3 * g + 1]; p[2] = means[3 * g + 2]; }
        float pc[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) pc[i] = R[i * 3 + 0] * p[0] + R[i * 3 + 1] * p[1] + R[i * 3 + 2] * p[2] + cam.t[i];
        const float4 q = staged ? reinterpret_cast<const float4 *>(s_quats)[row]
                                : reinterpret_cast<const float4 *>(quats)[g];
        float s[3];
        if (staged) { s[0] = s_scales[t3]; s[1] = s_scales[t3 + 1]; s[2] = s_scales[t3 + 2]; }
        else { s[0] = scales[3 * g]; s[1] = scales[3 * g + 1]; s[2] = scales[3 * g + 2]; }
        float inv_norm = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        float qw = q.x * inv_norm, qx = q.y * inv_norm, qy = q.z * inv_norm, qz = q.w * inv_norm;
        float Rq[9];
        {
            float x2 = qx * qx, y2 = qy * qy, z2 = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
            float wx = qw * qx, wy = qw * qy, wz = qw * qz;
            Rq[0] = 1.f - 2.f * (y2 + z2); Rq[1] = 2.f * (xy - wz); Rq[2] = 2.f * (xz + wy);
            Rq[3] = 2.f * (xy + wz); Rq[4] = 1.f - 2.f * (x2 + z2); Rq[5] = 2.f * (yz - wx);
            Rq[6] = 2.f * (xz - wy); Rq[7] = 2.f * (yz + wx); Rq[8] = 1.f - 2.f * (x2 + y2);
        }
        float Mm[9], Sigma[9], RS[9], Sc[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Mm[i * 3 + j] = Rq[i * 3 + j] * s[j];
        mat3_mul_bt(Mm, Mm, Sigma);
        mat3_mul(R, Sigma, RS);     // R Sigma  (Sigma symmetric => also R Sigma^T)
        mat3_mul_bt(RS, R, Sc);

        // perspective projection VJP
        float x = pc[0], y = pc[1], z = pc[2];
        float rz = 1.0f / z, rz2 = rz * rz, rz3 = rz2 * rz;
        float xr = x * rz, yr = y * rz;
        float tx = z * fminf(cam.lim_x_pos, fmaxf(-cam.lim_x_neg, xr));
        float ty = z * fminf(cam.lim_y_pos, fmaxf(-cam.lim_y_neg, yr));
        float J[6] = {cam.fx * rz, 0.f, -cam.fx * tx * rz2, 0.f, cam.fy * rz, -cam.fy * ty * rz2};
        float GJ[6], GtJ[6];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                GJ[i * 3 + j] = G[i * 2 + 0] * J[0 * 3 + j] + G[i * 2 + 1] * J[1 * 3 + j];
                GtJ[i * 3 + j] = G[0 * 2 + i] * J[0 * 3 + j] + G[1 * 2 + i] * J[1 * 3 + j];
            }
        float vSc[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) vSc[i * 3 + j] = J[0 * 3 + i] * GJ[0 * 3 + j] + J[1 * 3 + i] * GJ[1 * 3 + j];
        float vpc[3];
        vpc[0] = cam.fx * rz * gxy.x;
        vpc[1] = cam.fy * rz * gxy.y;
        vpc[2] = -(cam.fx * x * gxy.x + cam.fy * y * gxy.y) * rz2;
        float vJ[6];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) a += GJ[i * 3 + k] * Sc[j * 3 + k] + GtJ[i * 3 + k] * Sc[k * 3 + j];
                vJ[i * 3 + j] = a;
            }
        if (xr <= cam.lim_x_pos && xr >= -cam.lim_x_neg) vpc[0] += -cam.fx * rz2 * vJ[2];
        else vpc[2] += -cam.fx * rz3 * vJ[2] * tx;
        if (yr <= cam.lim_y_pos && yr >= -cam.lim_y_neg) vpc[1] += -cam.fy * rz2 * vJ[5];
        else vpc[2] += -cam.fy * rz3 * vJ[5] * ty;
        vpc[2] += -cam.fx * rz2 * vJ[0] - cam.fy * rz2 * vJ[4] + 2.f * cam.fx * tx * rz3 * vJ[2] +
                  2.f * cam.fy * ty * rz3 * vJ[5];
        vpc[2] += v_depth;

        // world->camera VJPs
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            vt[i] = vpc[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) vR[i * 3 + j] = vpc[i] * p[j];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            o_means[j] = R[0 * 3 + j] * vpc[0] + R[1 * 3 + j] * vpc[1] + R[2 * 3 + j] * vpc[2];
            if (!EXCH) v_means[3 * g + j] = o_means[j];
        }
        // v_R += vSc (R Sigma^T) + vSc^T (R Sigma)
        float tmp[9];
        mat3_mul(vSc, RS, tmp);
#pragma unroll
        for (int k = 0; k < 9; ++k) vR[k] += tmp[k];
        mat3_mul_at(vSc, RS, tmp);
#pragma unroll
        for (int k = 0; k < 9; ++k) vR[k] += tmp[k];
        // v_Sigma = R^T vSc R
        float vSigma[9];
        mat3_mul_at(R, vSc, tmp);
        mat3_mul(tmp, R, vSigma);

        // quat/scale VJP: Sigma = M M^T, M = Rq S
        float Sym[9], vM[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Sym[i * 3 + j] = vSigma[i * 3 + j] + vSigma[j * 3 + i];
        mat3_mul(Sym, Mm, vM);
        float Gq[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Gq[i * 3 + j] = vM[i * 3 + j] * s[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            o_scales[j] = Rq[0 * 3 + j] * vM[0 * 3 + j] + Rq[1 * 3 + j] * vM[1 * 3 + j] + Rq[2 * 3 + j] * vM[2 * 3 + j];
            if (!EXCH) v_scales[3 * g + j] = o_scales[j];
        }
        float vqn[4];
        vqn[0] = 2.f * (qx * (Gq[7] - Gq[5]) + qy * (Gq[2] - Gq[6]) + qz * (Gq[3] - Gq[1]));
        vqn[1] = 2.f * (-2.f * qx * (Gq[4] + Gq[8]) + qy * (Gq[1] + Gq[3]) + qz * (Gq[2] + Gq[6]) + qw * (Gq[7] - Gq[5]));
        vqn[2] = 2.f * (qx * (Gq[1] + Gq[3]) - 2.f * qy * (Gq[0] + Gq[8]) + qz * (Gq[5] + Gq[7]) + qw * (Gq[2] - Gq[6]));
        vqn[3] = 2.f * (qx * (Gq[2] + Gq[6]) + qy * (Gq[5] + Gq[7]) - 2.f * qz * (Gq[0] + Gq[4]) + qw * (Gq[3] - Gq[1]));
        float dotp = vqn[0] * qw + vqn[1] * qx + vqn[2] * qy + vqn[3] * qz;
        o_quat = make_float4((vqn[0] - dotp * qw) * inv_norm, (vqn[1] - dotp * qx) * inv_norm,
                             (vqn[2] - dotp * qy) * inv_norm, (vqn[3] - dotp * qz) * inv_norm);
    }
    if (!EXCH) {
        if (v_colors_out != nullptr && g_own < N) {  // contiguous [N, d_in] colour gradient (zeros for culled rows)
            for (int k = 0; k < d_in; ++k)
                v_colors_out[(size_t)g_own * d_in + k] =
                    own_live ? (staged ? s_vcol[threadIdx.x * CDIM + k] : v_colpack[(size_t)g_own * CDIM + k]) : 0.f;
        }
        if (live) v_quats[g] = o_quat;
        if (!own_live && g_own < N) {  // culled: define the row (zeros) so callers need no memset pass
            v_means[3 * g_own] = v_means[3 * g_own + 1] = v_means[3 * g_own + 2] = 0.f;
            v_scales[3 * g_own] = v_scales[3 * g_own + 1] = v_scales[3 * g_own + 2] = 0.f;
            v_quats[g_own] = make_float4(0.f, 0.f, 0.f, 0.f);
            v_opacities[g_own] = 0.f;
        }
    } else {
        // ---- transpose the CTA's rows through shared memory, then 16-byte coalesced stores into the owner's slot
        float *s_ex = s_dyn;  // [768 means | 1024 quats | 768 scales | 256 opac | d_col * 256 colours]
        const int d_col = ex.d_col;
        const int t = threadIdx.x;
        float o_col[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            o_col[k] = (live && k < d_col) ? (staged ? s_vcol[row * CDIM + k] : v_colpack[(size_t)g * CDIM + k]) : 0.f;
        __syncthreads();      // every thread has read its staged inputs: the buffer is reused for the output rows
        if (live) {  // the row this thread worked on
            s_ex[3 * row] = o_means[0]; s_ex[3 * row + 1] = o_means[1]; s_ex[3 * row + 2] = o_means[2];
            reinterpret_cast<float4 *>(s_ex + 768)[row] = o_quat;
            s_ex[1792 + 3 * row] = o_scales[0]; s_ex[1792 + 3 * row + 1] = o_scales[1]; s_ex[1792 + 3 * row + 2] = o_scales[2];
            s_ex[2560 + row] = o_opac;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < d_col) s_ex[2816 + d_col * row + k] = o_col[k];
        }
        if (!own_live) {  // this thread's own row is culled (or beyond N): zeros
            s_ex[3 * t] = 0.f; s_ex[3 * t + 1] = 0.f; s_ex[3 * t + 2] = 0.f;
            reinterpret_cast<float4 *>(s_ex + 768)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            s_ex[1792 + 3 * t] = 0.f; s_ex[1792 + 3 * t + 1] = 0.f; s_ex[1792 + 3 * t + 2] = 0.f;
            s_ex[2560 + t] = 0.f;
            for (int k = 0; k < d_col; ++k) s_ex[2816 + d_col * t + k] = 0.f;
        }
        __syncthreads();
        const int owner = g0 / ex.shard, l0 = g0 - owner * ex.shard;
        const int rows = min(256, N - g0);
        float *slot = ex.stage[owner] + (size_t)ex.rank * ex.slot_floats;
        const int widths[5] = {3, 4, 3, 1, d_col};
        int s_off = 0, a_p = 0;
#pragma unroll
        for (int p = 0; p < 5; ++p) {
            const int wp = widths[p];
            float *dst = slot + (size_t)a_p * ex.shard + (size_t)wp * l0;
            const float *src = s_ex + s_off;
            const int nvalid = wp * rows;
            for (int e = 4 * t; e < nvalid; e += 1024) {
                if (e + 3 < nvalid) {
                    *reinterpret_cast<float4 *>(dst + e) = *reinterpret_cast<const float4 *>(src + e);
                } else {
                    for (int k = e; k < nvalid; ++k) dst[k] = src[k];
                }
            }
            s_off += wp * 256;
            a_p += wp;
        }
    }
    // viewmat gradient: 12 sums over all Gaussians -> warp butterfly, smem across warps, 12 atomics / CTA
    if (v_viewmat != nullptr) {
        __shared__ float s_part[8][12];
        float vals[12];
#pragma unroll
        for (int k = 0; k < 9; ++k) vals[k] = vR[k];
        vals[9] = vt[0]; vals[10] = vt[1]; vals[11] = vt[2];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], o);
        }
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) s_part[warp][k] = vals[k];
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += s_part[w][threadIdx.x];
            int k = threadIdx.x;
            int dst = k < 9 ? (k / 3) * 4 + (k % 3) : (k - 9) * 4 + 3;
            if (a != 0.f) atomicAdd(v_viewmat + dst, a);
        }
    }
}

extern "C" int b2s_project_fwd(const float *means, const float *quats, const float *scales,
                               const float *opacities, const float *colors_in, const float *viewmat,
                               const float *K, int N, int W, int H, int tile_size, int tile_w, int tile_h,
                               float eps2d, float near_plane, float far_plane, float radius_clip, int calc_comp,
                               int d_in, int with_depth, int cdim, int32_t *radii, float *means2d, float *depths,
                               float *geo, float *comps, float *colpack, int32_t *tiles_per_gauss,
                               uint32_t *sort_keys, int32_t *tile_rects, int32_t *tight_rects, int64_t *totals,
                               float *bwd_arena, b2s_stream_t stream) {
    if (N < 0 || W <= 0 || H <= 0 || totals == nullptr) return B2S_ERR_ARG;
    if (N > 0 && tight_rects == nullptr) return B2S_ERR_ARG;
    int rg_shift = 0, cg_shift = 0;
    if (b2s_tl_shifts(tile_w, tile_h, &rg_shift, &cg_shift) != B2S_OK) return B2S_ERR_UNSUPPORTED;
    if (tile_size != 16 || tile_w > 32767 || tile_h > 32767) return B2S_ERR_UNSUPPORTED;
    if (cdim != 4 && cdim != 8) return B2S_ERR_UNSUPPORTED;
    if (d_in < 0 || d_in + (with_depth ? 1 : 0) > cdim) return B2S_ERR_ARG;
    if (calc_comp && comps == nullptr) return B2S_ERR_ARG;
    if (N == 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(b2s_div_up(N, 256)), block(256);
    const int use_tma = ((((uintptr_t)means | (uintptr_t)quats | (uintptr_t)scales | (uintptr_t)opacities |
                           (uintptr_t)colors_in) & 15) == 0) ? 1 : 0;
    const size_t smem = (size_t)(11 + d_in) * 256 * sizeof(float);
#define LAUNCH(CD)                                                                                           \
    k_project_fwd<CD><<<grid, block, smem, st>>>(means, quats, scales, opacities, colors_in, viewmat, K, N, W, H, \
                                              tile_w, tile_h, eps2d, near_plane, far_plane, radius_clip,     \
                                              calc_comp, d_in, with_depth, radii, (float2 *)means2d, depths, \
                                              (float4 *)geo, comps, colpack, tiles_per_gauss, sort_keys, \
                                              (int2 *)tile_rects, (int2 *)tight_rects, rg_shift, cg_shift,  \
                                              (unsigned long long *)totals, (float4 *)bwd_arena, use_tma)
    if (cdim == 4) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

extern "C" int b2s_project_bwd(const float *means, const float *quats, const float *scales,
                               const float *opacities, const float *viewmat, const float *K, int N, int W, int H,
                               float eps2d, int calc_comp, int d_in, int with_depth, int cdim,
                               const int32_t *radii, const float *geo, const float *comps, const float *v_means2d,
                               int v_means2d_stride, const float *v_geo, const float *v_colpack, float *v_means,
                               float *v_quats, float *v_scales, float *v_opacities, float *v_colors,
                               float *v_viewmat, b2s_stream_t stream) {
    if (N < 0) return B2S_ERR_ARG;
    if (cdim != 4 && cdim != 8) return B2S_ERR_UNSUPPORTED;
    if (calc_comp && comps == nullptr) return B2S_ERR_ARG;
    if (v_means2d_stride < 2 || (v_means2d_stride & 1)) return B2S_ERR_ARG;
    if (N == 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(b2s_div_up(N, 256)), block(256);
    B2sExchange none = {};
    const int use_tma = ((((uintptr_t)means | (uintptr_t)quats | (uintptr_t)scales | (uintptr_t)geo | (uintptr_t)v_geo |
                           (uintptr_t)v_means2d | (uintptr_t)v_colpack) & 15) == 0) ? 1 : 0;
    const size_t smem = (size_t)(18 + v_means2d_stride + cdim) * 256 * sizeof(float);
    if (v_means2d_stride > 8) return B2S_ERR_ARG;
#define LAUNCH(CD)                                                                                                \
    k_project_bwd<CD, false><<<grid, block, smem, st>>>(means, quats, scales, opacities, viewmat, K, N, W, H, eps2d,  \
                                                     calc_comp, d_in, with_depth, radii, (const float4 *)geo,     \
                                                     comps, v_means2d, v_means2d_stride, (const float4 *)v_geo,   \
                                                     v_colpack, v_means, (float4 *)v_quats, v_scales, v_opacities, \
                                                     v_colors, v_viewmat, none, use_tma)
    if (cdim == 4) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}

// Exchange variant (called by b2s_project_bwd_exchange in exchange.cu): rows [0, n_rows) go to their owners' slots.
int b2s_launch_project_bwd_exchange(const float *means, const float *quats, const float *scales, const float *opacities,
                                    const float *viewmat, const float *K, int n_rows, int W, int H, float eps2d,
                                    int calc_comp, int d_in, int with_depth, int cdim, const int32_t *radii,
                                    const float *geo, const float *comps, const float *v_means2d, int v_means2d_stride,
                                    const float *v_geo, const float *v_colpack, float *v_viewmat, const B2sExchange &ex,
                                    cudaStream_t st) {
    dim3 grid(b2s_div_up(n_rows, 256)), block(256);
    const int use_tma = ((((uintptr_t)means | (uintptr_t)quats | (uintptr_t)scales | (uintptr_t)geo | (uintptr_t)v_geo |
                           (uintptr_t)v_means2d | (uintptr_t)v_colpack) & 15) == 0) ? 1 : 0;
    if (v_means2d_stride > 8) return B2S_ERR_ARG;
    const size_t smem_in = (size_t)(18 + v_means2d_stride + cdim) * 256 * sizeof(float);
    const size_t smem_out = (size_t)(11 + ex.d_col) * 256 * sizeof(float);
    const size_t smem = smem_in > smem_out ? smem_in : smem_out;
#define LAUNCH(CD)                                                                                               \
    k_project_bwd<CD, true><<<grid, block, smem, st>>>(means, quats, scales, opacities, viewmat, K, n_rows, W, H, \
                                                       eps2d, calc_comp, d_in, with_depth, radii,                \
                                                       (const float4 *)geo, comps, v_means2d, v_means2d_stride,  \
                                                       (const float4 *)v_geo, v_colpack, nullptr, nullptr,       \
                                                       nullptr, nullptr, nullptr, v_viewmat, ex, use_tma)
    if (cdim == 4) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    B2S_LAUNCH_CHECK();
    return B2S_OK;
}
