/*
 * splat_oracle.c -- CPU restatement of the Gaussian-splat rasterization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mtgs_b200/ may import, link or call
 * this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the un-vendored pip
 * dependency gsplat (pinned v1.4.0, /root/reference/requirements.txt:12), which is
 * not installed in the build container, and the reference ships no golden
 * vectors or tests for it.  This file restates the
 * published algorithm of gsplat v1.4.0 as MTGS drives it:
 *   call site  mtgs/scene_model/mtgs_scene_graph.py:641-662  (rasterization)
 *   call site  mtgs/scene_model/gaussian_model/vanilla_gaussian_splatting.py:309-318 (SH)
 * Upstream function names are given next to each restated function.
 *
 * Numerics: every per-element quantity is computed in IEEE fp32 with one
 * rounding per operation, in the operation order written here
 * (compile with -ffp-contract=off, no -ffast-math).  The CUDA product follows
 * the same order for the projection/binning stages (DESIGN.md "canonical op
 * order"), so radii, tile counts, keys and offsets compare bit-exactly.
 * Reductions over many terms (gradient sums over pixels / Gaussians) are
 * accumulated in double here so the oracle is the more accurate side.
 *
 * Layout: row-major contiguous arrays, one camera (C = 1), quats (w,x,y,z).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ALPHA_MAX 0.999f
#define ALPHA_MIN (1.0f / 255.0f)
#define T_EPS 1e-4f

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* helpers                                                                    */
/* ------------------------------------------------------------------------- */

/* upstream: quat_to_rotmat (normalises inside). Row-major R[9]. */
static void quat_to_rotmat(const float *q, float *R) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    float inv_norm = 1.0f / sqrtf(((x * x + y * y) + z * z) + w * w);
    x *= inv_norm; y *= inv_norm; z *= inv_norm; w *= inv_norm;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    R[0] = 1.f - 2.f * (y2 + z2); R[1] = 2.f * (xy - wz);       R[2] = 2.f * (xz + wy);
    R[3] = 2.f * (xy + wz);       R[4] = 1.f - 2.f * (x2 + z2); R[5] = 2.f * (yz - wx);
    R[6] = 2.f * (xz - wy);       R[7] = 2.f * (yz + wx);       R[8] = 1.f - 2.f * (x2 + y2);
}

/* C = A * B, 3x3 row-major, sum left to right */
static void mat3_mul(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = (A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j]) + A[i * 3 + 2] * B[2 * 3 + j];
}
/* C = A * B^T */
static void mat3_mul_bt(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = (A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1]) + A[i * 3 + 2] * B[j * 3 + 2];
}

/* upstream: quat_scale_to_covar_preci (covariance only). Sigma = (R S)(R S)^T */
static void quat_scale_to_covar(const float *q, const float *s, float *Rq, float *Sigma) {
    float M[9];
    quat_to_rotmat(q, Rq);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = Rq[i * 3 + j] * s[j];
    mat3_mul_bt(M, M, Sigma);
}

typedef struct {
    float fx, fy, cx, cy;
    float lim_x_pos, lim_x_neg, lim_y_pos, lim_y_neg;
} Pinhole;

static Pinhole make_pinhole(const float *K, int W, int H) {
    Pinhole c;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    float tan_fovx = 0.5f * (float)W / c.fx;
    float tan_fovy = 0.5f * (float)H / c.fy;
    c.lim_x_pos = ((float)W - c.cx) / c.fx + 0.3f * tan_fovx;
    c.lim_x_neg = c.cx / c.fx + 0.3f * tan_fovx;
    c.lim_y_pos = ((float)H - c.cy) / c.fy + 0.3f * tan_fovy;
    c.lim_y_neg = c.cy / c.fy + 0.3f * tan_fovy;
    return c;
}

/* ------------------------------------------------------------------------- */
/* A.1 projection forward.  upstream: fully_fused_projection_fwd_kernel       */
/* ------------------------------------------------------------------------- */
void orc_project_fwd(const float *means, const float *quats, const float *scales,
                     const float *viewmat, const float *K, int N, int W, int H,
                     float eps2d, float near_plane, float far_plane, float radius_clip,
                     int calc_comp, int32_t *radii, float *means2d, float *depths,
                     float *conics, float *comps) {
    float R[9] = {viewmat[0], viewmat[1], viewmat[2], viewmat[4], viewmat[5], viewmat[6],
                  viewmat[8], viewmat[9], viewmat[10]};
    float t[3] = {viewmat[3], viewmat[7], viewmat[11]};
    Pinhole cam = make_pinhole(K, W, H);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < N; ++g) {
        const float *p = means + 3 * g;
        float pc[3];
        for (int i = 0; i < 3; ++i)
            pc[i] = ((R[i * 3 + 0] * p[0] + R[i * 3 + 1] * p[1]) + R[i * 3 + 2] * p[2]) + t[i];
        radii[g] = 0;
        if (pc[2] < near_plane || pc[2] > far_plane) continue;

        float Rq[9], Sigma[9], A[9], Sc[9];
        quat_scale_to_covar(quats + 4 * g, scales + 3 * g, Rq, Sigma);
        mat3_mul(R, Sigma, A);
        mat3_mul_bt(A, R, Sc);

        /* upstream: persp_proj */
        float x = pc[0], y = pc[1], z = pc[2];
        float rz = 1.0f / z;
        float rz2 = rz * rz;
        float tx = z * fminf(cam.lim_x_pos, fmaxf(-cam.lim_x_neg, x * rz));
        float ty = z * fminf(cam.lim_y_pos, fmaxf(-cam.lim_y_neg, y * rz));
        float j00 = cam.fx * rz, j02 = -((cam.fx * tx) * rz2);
        float j11 = cam.fy * rz, j12 = -((cam.fy * ty) * rz2);
        float B00 = j00 * Sc[0] + j02 * Sc[6];
        float B01 = j00 * Sc[1] + j02 * Sc[7];
        float B02 = j00 * Sc[2] + j02 * Sc[8];
        float B11 = j11 * Sc[4] + j12 * Sc[7];
        float B12 = j11 * Sc[5] + j12 * Sc[8];
        float c00 = B00 * j00 + B02 * j02;
        float c01 = B01 * j11 + B02 * j12;
        float c11 = B11 * j11 + B12 * j12;
        float mx = (cam.fx * x) * rz + cam.cx;
        float my = (cam.fy * y) * rz + cam.cy;

        /* upstream: add_blur */
        float det_orig = c00 * c11 - c01 * c01;
        c00 += eps2d;
        c11 += eps2d;
        float det = c00 * c11 - c01 * c01;
        float comp = sqrtf(fmaxf(0.0f, det_orig / det));
        if (!(det > 0.0f)) continue;

        float inv_det = 1.0f / det;
        float ca = c11 * inv_det, cb = -c01 * inv_det, cc = c00 * inv_det;

        float b = 0.5f * (c00 + c11);
        float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
        float radius = ceilf(3.0f * sqrtf(v1));
        if (radius <= radius_clip) continue;
        if (mx + radius <= 0.0f || mx - radius >= (float)W || my + radius <= 0.0f ||
            my - radius >= (float)H)
            continue;

        radii[g] = (int32_t)radius;
        means2d[2 * g + 0] = mx;
        means2d[2 * g + 1] = my;
        depths[g] = z;
        conics[3 * g + 0] = ca;
        conics[3 * g + 1] = cb;
        conics[3 * g + 2] = cc;
        if (calc_comp) comps[g] = comp;
    }
}

/* ------------------------------------------------------------------------- */
/* A.2 tile binning.  upstream: isect_tiles (two passes) + cumsum             */
/* ------------------------------------------------------------------------- */
static inline void tile_rect(float mx, float my, int32_t radius, int tile_size, int tile_w,
                             int tile_h, int *x0, int *y0, int *x1, int *y1) {
    float tr = (float)radius / (float)tile_size;
    float tx = mx / (float)tile_size, ty = my / (float)tile_size;
    float fx0 = floorf(tx - tr), fy0 = floorf(ty - tr);
    float fx1 = ceilf(tx + tr), fy1 = ceilf(ty + tr);
    /* device float->uint32 cast saturates negatives to 0, then min(., grid) */
    *x0 = fx0 <= 0.f ? 0 : (fx0 >= (float)tile_w ? tile_w : (int)fx0);
    *y0 = fy0 <= 0.f ? 0 : (fy0 >= (float)tile_h ? tile_h : (int)fy0);
    *x1 = fx1 <= 0.f ? 0 : (fx1 >= (float)tile_w ? tile_w : (int)fx1);
    *y1 = fy1 <= 0.f ? 0 : (fy1 >= (float)tile_h ? tile_h : (int)fy1);
}

/* returns total number of intersections M; fills tiles_per_gauss[N] */
int64_t orc_isect_count(const float *means2d, const int32_t *radii, int N, int tile_size,
                        int tile_w, int tile_h, int32_t *tiles_per_gauss) {
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int g = 0; g < N; ++g) {
        if (radii[g] <= 0) { tiles_per_gauss[g] = 0; continue; }
        int x0, y0, x1, y1;
        tile_rect(means2d[2 * g], means2d[2 * g + 1], radii[g], tile_size, tile_w, tile_h, &x0,
                  &y0, &x1, &y1);
        tiles_per_gauss[g] = (y1 - y0) * (x1 - x0);
        total += tiles_per_gauss[g];
    }
    return total;
}

int orc_tile_bits(int n_tiles) {
    /* upstream: (uint32_t)floor(log2(n_tiles)) + 1 */
    int b = 0;
    while ((1LL << (b + 1)) <= (int64_t)n_tiles) ++b;
    return b + 1;
}

/* emit unsorted (key, value) pairs in Gaussian order, row-major over the rect */
void orc_isect_emit(const float *means2d, const int32_t *radii, const float *depths, int N,
                    int tile_size, int tile_w, int tile_h, int64_t *isect_ids,
                    int32_t *flatten_ids) {
    int tile_n_bits = orc_tile_bits(tile_w * tile_h);
    (void)tile_n_bits; /* camera id is 0 for C = 1, so its field is all zeros */
    /* start of every Gaussian's run = exclusive scan of its rectangle area (emission order = Gaussian order) */
    int64_t *first = (int64_t *)malloc(sizeof(int64_t) * ((size_t)N + 1));
    first[0] = 0;
    for (int g = 0; g < N; ++g) {
        int64_t n = 0;
        if (radii[g] > 0) {
            int x0, y0, x1, y1;
            tile_rect(means2d[2 * g], means2d[2 * g + 1], radii[g], tile_size, tile_w, tile_h, &x0,
                      &y0, &x1, &y1);
            n = (int64_t)(y1 - y0) * (x1 - x0);
        }
        first[g + 1] = first[g] + n;
    }
#pragma omp parallel for schedule(dynamic, 1024)
    for (int g = 0; g < N; ++g) {
        if (radii[g] <= 0) continue;
        int x0, y0, x1, y1;
        tile_rect(means2d[2 * g], means2d[2 * g + 1], radii[g], tile_size, tile_w, tile_h, &x0,
                  &y0, &x1, &y1);
        int32_t dbits;
        memcpy(&dbits, depths + g, 4);
        int64_t depth_enc = (int64_t)dbits; /* upstream sign-extends an int32 view */
        int64_t cur = first[g];
        for (int i = y0; i < y1; ++i)
            for (int j = x0; j < x1; ++j) {
                int64_t tile_id = (int64_t)i * tile_w + j;
                isect_ids[cur] = (tile_id << 32) | depth_enc;
                flatten_ids[cur] = g;
                ++cur;
            }
    }
    free(first);
}

/* stable LSD radix sort of (int64 key, int32 value) on bits [0, end_bit).
 * upstream: cub::DeviceRadixSort::SortPairs(..., 0, 32 + tile_n_bits + cam_n_bits) */
void orc_sort_pairs(int64_t M, int end_bit, int64_t *keys, int32_t *vals) {
    if (M <= 1) return;
    int64_t *k2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)M);
    int32_t *v2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
    int64_t *ka = keys, *kb = k2;
    int32_t *va = vals, *vb = v2;
    int nt = orc_num_threads();
    if (nt < 1) nt = 1;
    if ((int64_t)nt > M) nt = (int)M;
    const int NB = 1 << 11;
    int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)nt * NB);
    for (int shift = 0; shift < end_bit; shift += 11) {
        int bits = end_bit - shift < 11 ? end_bit - shift : 11;
        int nb = 1 << bits;
        uint64_t mask = (uint64_t)nb - 1;
        /* slice t = [M t / nt, M (t+1) / nt): per-slice digit counts, then offsets in (digit, slice) order keep
         * the sort stable whatever the thread count */
#pragma omp parallel for schedule(static, 1) num_threads(nt)
        for (int t = 0; t < nt; ++t) {
            int64_t *c = cnt + (size_t)t * NB;
            memset(c, 0, sizeof(int64_t) * (size_t)nb);
            int64_t lo = M * t / nt, hi = M * (t + 1) / nt;
            for (int64_t i = lo; i < hi; ++i) c[((uint64_t)ka[i]) >> shift & mask]++;
        }
        int64_t run = 0;
        for (int d = 0; d < nb; ++d)
            for (int t = 0; t < nt; ++t) {
                int64_t n = cnt[(size_t)t * NB + d];
                cnt[(size_t)t * NB + d] = run;
                run += n;
            }
#pragma omp parallel for schedule(static, 1) num_threads(nt)
        for (int t = 0; t < nt; ++t) {
            int64_t *c = cnt + (size_t)t * NB;
            int64_t lo = M * t / nt, hi = M * (t + 1) / nt;
            for (int64_t i = lo; i < hi; ++i) {
                int64_t pos = c[((uint64_t)ka[i]) >> shift & mask]++;
                kb[pos] = ka[i];
                vb[pos] = va[i];
            }
        }
        int64_t *tk = ka; ka = kb; kb = tk;
        int32_t *tv = va; va = vb; vb = tv;
    }
    free(cnt);
    if (ka != keys) {
        memcpy(keys, ka, sizeof(int64_t) * (size_t)M);
        memcpy(vals, va, sizeof(int32_t) * (size_t)M);
    }
    free(k2);
    free(v2);
}

/* upstream: isect_offset_encode.  offsets[t] = first sorted index whose tile >= t */
void orc_isect_offsets(const int64_t *isect_ids, int64_t M, int n_tiles, int32_t *offsets) {
    int64_t idx = 0;
    for (int t = 0; t < n_tiles; ++t) {
        while (idx < M && (int64_t)(((uint64_t)isect_ids[idx]) >> 32) < (int64_t)t) ++idx;
        offsets[t] = (int32_t)idx;
    }
}

/* ------------------------------------------------------------------------- */
/* A.3 blend forward.  upstream: rasterize_to_pixels_fwd_kernel               */
/* ------------------------------------------------------------------------- */
void orc_blend_fwd(const float *means2d, const float *conics, const float *colors,
                   const float *opacities, const float *backgrounds, int W, int H, int tile_size,
                   int tile_w, int tile_h, const int32_t *offsets, const int32_t *flatten_ids,
                   int64_t M, int CD, float *render_colors, float *render_alphas,
                   int32_t *last_ids) {
    int n_tiles = tile_w * tile_h;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < n_tiles; ++tile) {
        int ti = tile / tile_w, tj = tile % tile_w;
        int64_t start = offsets[tile];
        int64_t end = (tile == n_tiles - 1) ? M : offsets[tile + 1];
        float pix[64];
        for (int li = 0; li < tile_size; ++li)
            for (int lj = 0; lj < tile_size; ++lj) {
                int i = ti * tile_size + li, j = tj * tile_size + lj;
                if (i >= H || j >= W) continue;
                float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.0f;
                int32_t cur = 0;
                for (int k = 0; k < CD; ++k) pix[k] = 0.f;
                for (int64_t idx = start; idx < end; ++idx) {
                    int g = flatten_ids[idx];
                    float dx = means2d[2 * g] - px, dy = means2d[2 * g + 1] - py;
                    float a = conics[3 * g], b = conics[3 * g + 1], c = conics[3 * g + 2];
                    float sigma = 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy;
                    float alpha = fminf(ALPHA_MAX, opacities[g] * expf(-sigma));
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    float next_T = T * (1.0f - alpha);
                    if (next_T <= T_EPS) break;
                    float vis = alpha * T;
                    const float *cp = colors + (size_t)g * CD;
                    for (int k = 0; k < CD; ++k) pix[k] += cp[k] * vis;
                    cur = (int32_t)idx;
                    T = next_T;
                }
                size_t pid = (size_t)i * W + j;
                render_alphas[pid] = 1.0f - T;
                for (int k = 0; k < CD; ++k)
                    render_colors[pid * CD + k] = backgrounds ? pix[k] + T * backgrounds[k] : pix[k];
                last_ids[pid] = cur;
            }
    }
}

/* ------------------------------------------------------------------------- */
/* A.4 blend backward.  upstream: rasterize_to_pixels_bwd_kernel              */
/* Gradients are accumulated in double per tile then added under a critical   */
/* section per Gaussian (order-independent to ~1e-16).                        */
/* ------------------------------------------------------------------------- */
void orc_blend_bwd(const float *means2d, const float *conics, const float *colors,
                   const float *opacities, const float *backgrounds, int N, int W, int H,
                   int tile_size, int tile_w, int tile_h, const int32_t *offsets,
                   const int32_t *flatten_ids, int64_t M, int CD, const float *render_alphas,
                   const int32_t *last_ids, const float *v_render_colors,
                   const float *v_render_alphas, double *v_means2d, double *v_means2d_abs,
                   double *v_conics, double *v_colors, double *v_opacities) {
    int n_tiles = tile_w * tile_h;
    memset(v_means2d, 0, sizeof(double) * 2 * (size_t)N);
    if (v_means2d_abs) memset(v_means2d_abs, 0, sizeof(double) * 2 * (size_t)N);
    memset(v_conics, 0, sizeof(double) * 3 * (size_t)N);
    memset(v_colors, 0, sizeof(double) * (size_t)CD * (size_t)N);
    memset(v_opacities, 0, sizeof(double) * (size_t)N);
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < n_tiles; ++tile) {
        int ti = tile / tile_w, tj = tile % tile_w;
        int64_t start = offsets[tile];
        int64_t end = (tile == n_tiles - 1) ? M : offsets[tile + 1];
        int64_t L = end - start;
        if (L <= 0) continue;
        int stride = 8 + CD;
        double *acc = (double *)calloc((size_t)L * stride, sizeof(double));
        float buffer[64];
        for (int li = 0; li < tile_size; ++li)
            for (int lj = 0; lj < tile_size; ++lj) {
                int i = ti * tile_size + li, j = tj * tile_size + lj;
                if (i >= H || j >= W) continue;
                size_t pid = (size_t)i * W + j;
                float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T_final = 1.0f - render_alphas[pid];
                float T = T_final;
                const float *v_rc = v_render_colors + pid * CD;
                float v_ra = v_render_alphas[pid];
                for (int k = 0; k < CD; ++k) buffer[k] = 0.f;
                int64_t bin_final = last_ids[pid];
                for (int64_t idx = bin_final < end - 1 ? bin_final : end - 1; idx >= start; --idx) {
                    int g = flatten_ids[idx];
                    float dx = means2d[2 * g] - px, dy = means2d[2 * g + 1] - py;
                    float a = conics[3 * g], b = conics[3 * g + 1], c = conics[3 * g + 2];
                    float opac = opacities[g];
                    float sigma = 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy;
                    float vis = expf(-sigma);
                    float alpha = fminf(ALPHA_MAX, opac * vis);
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    double *A = acc + (size_t)(idx - start) * stride;
                    float ra = 1.0f / (1.0f - alpha);
                    T *= ra;
                    float fac = alpha * T;
                    const float *cp = colors + (size_t)g * CD;
                    float v_alpha = 0.f;
                    for (int k = 0; k < CD; ++k) {
                        A[8 + k] += (double)(fac * v_rc[k]);
                        v_alpha += (cp[k] * T - buffer[k] * ra) * v_rc[k];
                    }
                    v_alpha += T_final * ra * v_ra;
                    if (backgrounds) {
                        float accum = 0.f;
                        for (int k = 0; k < CD; ++k) accum += backgrounds[k] * v_rc[k];
                        v_alpha += -T_final * ra * accum;
                    }
                    if (opac * vis <= ALPHA_MAX) {
                        float v_sigma = -opac * vis * v_alpha;
                        A[4] += (double)(0.5f * v_sigma * dx * dx);
                        A[5] += (double)(v_sigma * dx * dy);
                        A[6] += (double)(0.5f * v_sigma * dy * dy);
                        float vx = v_sigma * (a * dx + b * dy);
                        float vy = v_sigma * (b * dx + c * dy);
                        A[0] += (double)vx;
                        A[1] += (double)vy;
                        A[2] += (double)fabsf(vx);
                        A[3] += (double)fabsf(vy);
                        A[7] += (double)(vis * v_alpha);
                    }
                    for (int k = 0; k < CD; ++k) buffer[k] += cp[k] * fac;
                }
            }
        /* merge into the per-Gaussian sums (double atomics: tiles run in parallel) */
#define ORC_ADD(dst, val)            \
    do {                             \
        double v_ = (val);           \
        if (v_ != 0.0) {             \
            _Pragma("omp atomic")    \
            (dst) += v_;             \
        }                            \
    } while (0)
        for (int64_t e = 0; e < L; ++e) {
            int g = flatten_ids[start + e];
            const double *A = acc + (size_t)e * stride;
            ORC_ADD(v_means2d[2 * g], A[0]);
            ORC_ADD(v_means2d[2 * g + 1], A[1]);
            if (v_means2d_abs) {
                ORC_ADD(v_means2d_abs[2 * g], A[2]);
                ORC_ADD(v_means2d_abs[2 * g + 1], A[3]);
            }
            ORC_ADD(v_conics[3 * g], A[4]);
            ORC_ADD(v_conics[3 * g + 1], A[5]);
            ORC_ADD(v_conics[3 * g + 2], A[6]);
            ORC_ADD(v_opacities[g], A[7]);
            for (int k = 0; k < CD; ++k) ORC_ADD(v_colors[(size_t)g * CD + k], A[8 + k]);
        }
#undef ORC_ADD
        free(acc);
    }
}

/* ------------------------------------------------------------------------- */
/* A.5 projection backward.  upstream: fully_fused_projection_bwd_kernel      */
/* v_viewmat (16 floats, row 3 zero) accumulated in double.                   */
/* ------------------------------------------------------------------------- */
void orc_project_bwd(const float *means, const float *quats, const float *scales,
                     const float *viewmat, const float *K, int N, int W, int H, float eps2d,
                     const int32_t *radii, const float *conics, const float *comps,
                     const float *v_means2d, const float *v_depths, const float *v_conics,
                     const float *v_comps, float *v_means, float *v_quats, float *v_scales,
                     double *v_viewmat) {
    float R[9] = {viewmat[0], viewmat[1], viewmat[2], viewmat[4], viewmat[5], viewmat[6],
                  viewmat[8], viewmat[9], viewmat[10]};
    float t[3] = {viewmat[3], viewmat[7], viewmat[11]};
    Pinhole cam = make_pinhole(K, W, H);
    double vR_acc[9] = {0}, vt_acc[3] = {0};
#pragma omp parallel for schedule(static) reduction(+ : vR_acc[:9], vt_acc[:3])
    for (int g = 0; g < N; ++g) {
        for (int k = 0; k < 3; ++k) { v_means[3 * g + k] = 0.f; v_scales[3 * g + k] = 0.f; }
        for (int k = 0; k < 4; ++k) v_quats[4 * g + k] = 0.f;
        if (radii[g] <= 0) continue;

        /* inverse_vjp: v_cov2d = -Minv * v_Minv * Minv */
        float ia = conics[3 * g], ib = conics[3 * g + 1], ic = conics[3 * g + 2];
        float ga = v_conics[3 * g], gb = 0.5f * v_conics[3 * g + 1], gc = v_conics[3 * g + 2];
        /* P = Minv * G */
        float p00 = ia * ga + ib * gb, p01 = ia * gb + ib * gc;
        float p10 = ib * ga + ic * gb, p11 = ib * gb + ic * gc;
        float vc00 = -(p00 * ia + p01 * ib), vc01 = -(p00 * ib + p01 * ic);
        float vc10 = -(p10 * ia + p11 * ib), vc11 = -(p10 * ib + p11 * ic);

        if (v_comps && comps) { /* add_blur_vjp */
            float comp = comps[g], v_comp = v_comps[g];
            float det_conic = ia * ic - ib * ib;
            float v_sqr_comp = v_comp * 0.5f / (comp + 1e-6f);
            float one_minus = 1.0f - comp * comp;
            vc00 += v_sqr_comp * (one_minus * ia - eps2d * det_conic);
            vc01 += v_sqr_comp * (one_minus * ib);
            vc10 += v_sqr_comp * (one_minus * ib);
            vc11 += v_sqr_comp * (one_minus * ic - eps2d * det_conic);
        }

        const float *p = means + 3 * g;
        float pc[3];
        for (int i = 0; i < 3; ++i)
            pc[i] = ((R[i * 3 + 0] * p[0] + R[i * 3 + 1] * p[1]) + R[i * 3 + 2] * p[2]) + t[i];
        float Rq[9], Sigma[9], A[9], Sc[9];
        quat_scale_to_covar(quats + 4 * g, scales + 3 * g, Rq, Sigma);
        mat3_mul(R, Sigma, A);
        mat3_mul_bt(A, R, Sc);

        /* persp_proj_vjp */
        float x = pc[0], y = pc[1], z = pc[2];
        float rz = 1.0f / z, rz2 = rz * rz, rz3 = rz2 * rz;
        float xr = x * rz, yr = y * rz;
        float tx = z * fminf(cam.lim_x_pos, fmaxf(-cam.lim_x_neg, xr));
        float ty = z * fminf(cam.lim_y_pos, fmaxf(-cam.lim_y_neg, yr));
        float J[6] = {cam.fx * rz, 0.f, -cam.fx * tx * rz2, 0.f, cam.fy * rz, -cam.fy * ty * rz2};
        float G[4] = {vc00, vc01, vc10, vc11};
        /* v_Sc = J^T G J  (3x3) */
        float GJ[6]; /* 2x3 = G * J */
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j) GJ[i * 3 + j] = G[i * 2 + 0] * J[0 * 3 + j] + G[i * 2 + 1] * J[1 * 3 + j];
        float vSc[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) vSc[i * 3 + j] = J[0 * 3 + i] * GJ[0 * 3 + j] + J[1 * 3 + i] * GJ[1 * 3 + j];
        float vmx = v_means2d[2 * g], vmy = v_means2d[2 * g + 1];
        float vpc[3];
        vpc[0] = cam.fx * rz * vmx;
        vpc[1] = cam.fy * rz * vmy;
        vpc[2] = -(cam.fx * x * vmx + cam.fy * y * vmy) * rz2;
        /* v_J = G J Sc^T + G^T J Sc   (2x3) */
        float GtJ[6];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j) GtJ[i * 3 + j] = G[0 * 2 + i] * J[0 * 3 + j] + G[1 * 2 + i] * J[1 * 3 + j];
        float vJ[6];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j) {
                float s = 0.f;
                for (int k = 0; k < 3; ++k) s += GJ[i * 3 + k] * Sc[j * 3 + k] + GtJ[i * 3 + k] * Sc[k * 3 + j];
                vJ[i * 3 + j] = s;
            }
        if (xr <= cam.lim_x_pos && xr >= -cam.lim_x_neg) vpc[0] += -cam.fx * rz2 * vJ[2];
        else vpc[2] += -cam.fx * rz3 * vJ[2] * tx;
        if (yr <= cam.lim_y_pos && yr >= -cam.lim_y_neg) vpc[1] += -cam.fy * rz2 * vJ[5];
        else vpc[2] += -cam.fy * rz3 * vJ[5] * ty;
        vpc[2] += -cam.fx * rz2 * vJ[0] - cam.fy * rz2 * vJ[4] + 2.f * cam.fx * tx * rz3 * vJ[2] +
                  2.f * cam.fy * ty * rz3 * vJ[5];
        vpc[2] += v_depths[g];

        /* pos_world_to_cam_vjp */
        float vR[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) vR[i * 3 + j] = vpc[i] * p[j];
        for (int j = 0; j < 3; ++j)
            v_means[3 * g + j] = R[0 * 3 + j] * vpc[0] + R[1 * 3 + j] * vpc[1] + R[2 * 3 + j] * vpc[2];
        /* covar_world_to_cam_vjp: v_R += vSc R Sigma^T + vSc^T R Sigma ; v_Sigma = R^T vSc R */
        float RS[9], RSt[9], tmp[9];
        mat3_mul(R, Sigma, RS);      /* R Sigma */
        mat3_mul_bt(R, Sigma, RSt);  /* R Sigma^T */
        mat3_mul(vSc, RSt, tmp);
        for (int k = 0; k < 9; ++k) vR[k] += tmp[k];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                float s = 0.f;
                for (int k = 0; k < 3; ++k) s += vSc[k * 3 + i] * RS[k * 3 + j];
                vR[i * 3 + j] += s;
            }
        float vSigma[9], tmp2[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                float s = 0.f;
                for (int k = 0; k < 3; ++k) s += R[k * 3 + i] * vSc[k * 3 + j];
                tmp2[i * 3 + j] = s;
            }
        mat3_mul(tmp2, R, vSigma);

        /* quat_scale_to_covar_vjp */
        const float *s = scales + 3 * g;
        float Mm[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Mm[i * 3 + j] = Rq[i * 3 + j] * s[j];
        float Sym[9], vM[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Sym[i * 3 + j] = vSigma[i * 3 + j] + vSigma[j * 3 + i];
        mat3_mul(Sym, Mm, vM);
        float vRq[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) vRq[i * 3 + j] = vM[i * 3 + j] * s[j];
        for (int j = 0; j < 3; ++j)
            v_scales[3 * g + j] = Rq[0 * 3 + j] * vM[0 * 3 + j] + Rq[1 * 3 + j] * vM[1 * 3 + j] + Rq[2 * 3 + j] * vM[2 * 3 + j];
        /* quat_to_rotmat_vjp */
        const float *q = quats + 4 * g;
        float inv_norm = 1.0f / sqrtf(((q[1] * q[1] + q[2] * q[2]) + q[3] * q[3]) + q[0] * q[0]);
        float qw = q[0] * inv_norm, qx = q[1] * inv_norm, qy = q[2] * inv_norm, qz = q[3] * inv_norm;
        const float *Gq = vRq; /* row-major G[i][j] */
        float vqn[4];
        vqn[0] = 2.f * (qx * (Gq[7] - Gq[5]) + qy * (Gq[2] - Gq[6]) + qz * (Gq[3] - Gq[1]));
        vqn[1] = 2.f * (-2.f * qx * (Gq[4] + Gq[8]) + qy * (Gq[1] + Gq[3]) + qz * (Gq[2] + Gq[6]) + qw * (Gq[7] - Gq[5]));
        vqn[2] = 2.f * (qx * (Gq[1] + Gq[3]) - 2.f * qy * (Gq[0] + Gq[8]) + qz * (Gq[5] + Gq[7]) + qw * (Gq[2] - Gq[6]));
        vqn[3] = 2.f * (qx * (Gq[2] + Gq[6]) + qy * (Gq[5] + Gq[7]) - 2.f * qz * (Gq[0] + Gq[4]) + qw * (Gq[3] - Gq[1]));
        float dotp = vqn[0] * qw + vqn[1] * qx + vqn[2] * qy + vqn[3] * qz;
        v_quats[4 * g + 0] = (vqn[0] - dotp * qw) * inv_norm;
        v_quats[4 * g + 1] = (vqn[1] - dotp * qx) * inv_norm;
        v_quats[4 * g + 2] = (vqn[2] - dotp * qy) * inv_norm;
        v_quats[4 * g + 3] = (vqn[3] - dotp * qz) * inv_norm;

        for (int k = 0; k < 9; ++k) vR_acc[k] += (double)vR[k];
        for (int k = 0; k < 3; ++k) vt_acc[k] += (double)vpc[k];
    }
    for (int k = 0; k < 16; ++k) v_viewmat[k] = 0.0;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) v_viewmat[i * 4 + j] = vR_acc[i * 3 + j];
        v_viewmat[i * 4 + 3] = vt_acc[i];
    }
}

/* ------------------------------------------------------------------------- */
/* A.6 spherical harmonics.  upstream: compute_sh_fwd/bwd_kernel              */
/* (sh_coeffs_to_color_fast / _vjp), degrees 0..4, coeffs [N,K,3], dirs [N,3] */
/* ------------------------------------------------------------------------- */
static void sh_basis(int degree, float x, float y, float z, float *B) {
    B[0] = 0.2820947917738781f;
    if (degree < 1) return;
    B[1] = -0.48860251190292f * y;
    B[2] = 0.48860251190292f * z;
    B[3] = -0.48860251190292f * x;
    if (degree < 2) return;
    float z2 = z * z;
    float fTmp0B = -1.092548430592079f * z;
    float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    B[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    B[7] = fTmp0B * x;
    B[5] = fTmp0B * y;
    B[8] = 0.5462742152960395f * fC1;
    B[4] = 0.5462742152960395f * fS1;
    if (degree < 3) return;
    float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    float fTmp1B = 1.445305721320277f * z;
    float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    B[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
    B[13] = fTmp0C * x;
    B[11] = fTmp0C * y;
    B[14] = fTmp1B * fC1;
    B[10] = fTmp1B * fS1;
    B[15] = -0.5900435899266435f * fC2;
    B[9] = -0.5900435899266435f * fS2;
    if (degree < 4) return;
    float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    float fTmp2B = -1.770130769779931f * z;
    float fC3 = x * fC2 - y * fS2, fS3 = x * fS2 + y * fC2;
    B[20] = 1.984313483298443f * z * B[12] + -1.006230589874905f * B[6];
    B[21] = fTmp0D * x;
    B[19] = fTmp0D * y;
    B[22] = fTmp1C * fC1;
    B[18] = fTmp1C * fS1;
    B[23] = fTmp2B * fC2;
    B[17] = fTmp2B * fS2;
    B[24] = 0.6258357354491763f * fC3;
    B[16] = 0.6258357354491763f * fS3;
}

void orc_sh_fwd(int degree, const float *dirs, const float *coeffs, const uint8_t *masks, int N,
                int K, float *colors) {
    int nb = (degree + 1) * (degree + 1);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < N; ++g) {
        if (masks && !masks[g]) continue;
        float x = dirs[3 * g], y = dirs[3 * g + 1], z = dirs[3 * g + 2];
        float inorm = 1.0f / sqrtf(x * x + y * y + z * z);
        if (degree >= 1) { x *= inorm; y *= inorm; z *= inorm; }
        float B[25];
        sh_basis(degree, x, y, z, B);
        for (int c = 0; c < 3; ++c) {
            float r = 0.f;
            for (int k = 0; k < nb; ++k) r += B[k] * coeffs[((size_t)g * K + k) * 3 + c];
            colors[3 * g + c] = r;
        }
    }
}

/* v_dirs computed by central differences of the analytic basis in double is
 * avoided: we differentiate the polynomial basis analytically through the
 * normalisation (same result as upstream's hand-written VJP). */
static void sh_basis_d(int degree, double x, double y, double z, double *B) {
    B[0] = 0.2820947917738781;
    if (degree < 1) return;
    B[1] = -0.48860251190292 * y; B[2] = 0.48860251190292 * z; B[3] = -0.48860251190292 * x;
    if (degree < 2) return;
    double z2 = z * z, fTmp0B = -1.092548430592079 * z, fC1 = x * x - y * y, fS1 = 2. * x * y;
    B[6] = 0.9461746957575601 * z2 - 0.3153915652525201; B[7] = fTmp0B * x; B[5] = fTmp0B * y;
    B[8] = 0.5462742152960395 * fC1; B[4] = 0.5462742152960395 * fS1;
    if (degree < 3) return;
    double fTmp0C = -2.285228997322329 * z2 + 0.4570457994644658, fTmp1B = 1.445305721320277 * z;
    double fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    B[12] = z * (1.865881662950577 * z2 - 1.119528997770346); B[13] = fTmp0C * x; B[11] = fTmp0C * y;
    B[14] = fTmp1B * fC1; B[10] = fTmp1B * fS1; B[15] = -0.5900435899266435 * fC2; B[9] = -0.5900435899266435 * fS2;
    if (degree < 4) return;
    double fTmp0D = z * (-4.683325804901025 * z2 + 2.007139630671868);
    double fTmp1C = 3.31161143515146 * z2 - 0.47308734787878, fTmp2B = -1.770130769779931 * z;
    double fC3 = x * fC2 - y * fS2, fS3 = x * fS2 + y * fC2;
    B[20] = 1.984313483298443 * z * B[12] + -1.006230589874905 * B[6];
    B[21] = fTmp0D * x; B[19] = fTmp0D * y; B[22] = fTmp1C * fC1; B[18] = fTmp1C * fS1;
    B[23] = fTmp2B * fC2; B[17] = fTmp2B * fS2; B[24] = 0.6258357354491763 * fC3; B[16] = 0.6258357354491763 * fS3;
}

void orc_sh_bwd(int degree, const float *dirs, const float *coeffs, const uint8_t *masks, int N,
                int K, const float *v_colors, float *v_coeffs, float *v_dirs) {
    int nb = (degree + 1) * (degree + 1);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < N; ++g) {
        for (int k = 0; k < K * 3; ++k) v_coeffs[(size_t)g * K * 3 + k] = 0.f;
        if (v_dirs) { v_dirs[3 * g] = v_dirs[3 * g + 1] = v_dirs[3 * g + 2] = 0.f; }
        if (masks && !masks[g]) continue;
        double x = dirs[3 * g], y = dirs[3 * g + 1], z = dirs[3 * g + 2];
        double inorm = 1.0 / sqrt(x * x + y * y + z * z);
        double nx = x, ny = y, nz = z;
        if (degree >= 1) { nx *= inorm; ny *= inorm; nz *= inorm; }
        double B[25];
        sh_basis_d(degree, nx, ny, nz, B);
        for (int k = 0; k < nb; ++k)
            for (int c = 0; c < 3; ++c) v_coeffs[((size_t)g * K + k) * 3 + c] = (float)(B[k] * v_colors[3 * g + c]);
        if (!v_dirs || degree < 1) continue;
        /* dL/d(n) by analytic differentiation of the polynomial basis: use the
         * exactness of central differences on polynomials of degree <= 4 with a
         * 5-point stencil in double (error O(h^4 * d5) = 0 for degree <= 4). */
        double gn[3];
        const double h = 1e-2;
        for (int a = 0; a < 3; ++a) {
            double acc = 0.0;
            static const double wts[4] = {1.0 / 12.0, -8.0 / 12.0, 8.0 / 12.0, -1.0 / 12.0};
            static const double off[4] = {-2.0, -1.0, 1.0, 2.0};
            for (int s = 0; s < 4; ++s) {
                double v[3] = {nx, ny, nz};
                v[a] += off[s] * h;
                double Bs[25];
                sh_basis_d(degree, v[0], v[1], v[2], Bs);
                double f = 0.0;
                for (int k = 0; k < nb; ++k)
                    for (int c = 0; c < 3; ++c) f += Bs[k] * (double)coeffs[((size_t)g * K + k) * 3 + c] * (double)v_colors[3 * g + c];
                acc += wts[s] * f;
            }
            gn[a] = acc / h;
        }
        /* through n = d / |d| : v_d = (gn - (gn . n) n) / |d| */
        double dotp = gn[0] * nx + gn[1] * ny + gn[2] * nz;
        v_dirs[3 * g + 0] = (float)((gn[0] - dotp * nx) * inorm);
        v_dirs[3 * g + 1] = (float)((gn[1] - dotp * ny) * inorm);
        v_dirs[3 * g + 2] = (float)((gn[2] - dotp * nz) * inorm);
    }
}

/* exposed for the golden-vector pin against the reference's own quat_to_rotmat
 * (mtgs/scene_model/gaussian_model/utils.py:14-41, same (w,x,y,z) convention) */
void orc_quat_to_rotmat(const float *quats, int N, float *R) {
    for (int g = 0; g < N; ++g) quat_to_rotmat(quats + 4 * g, R + 9 * g);
}
