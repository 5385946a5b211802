/*
 * b200splat.h -- flat C ABI of libb200splat.so (hand-written sm_100a CUDA kernels).
 *
 * Drop-in boundary for the one hot path MTGS delegates to gsplat v1.4.0
 * (SURVEY.md section 8b).  Each entry point replaces an upstream operator that
 * MTGS reaches through
 *     gsplat.rendering.rasterization      mtgs/scene_model/mtgs_scene_graph.py:21, 641-662
 *     gsplat.cuda._wrapper.spherical_harmonics
 *                                         mtgs/scene_model/gaussian_model/vanilla_gaussian_splatting.py:16, 317
 *                                         (same call: multi_color_gaussian_splatting.py:96,
 *                                          rigid_node.py:248, deformable_node.py:125)
 * The reference-side binding is Python (ctypes); see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - row-major contiguous fp32 / int32 / int64 buffers, one camera (C = 1);
 *   - quaternions (w,x,y,z); viewmat = world->camera 4x4 (OpenCV axes); K = 3x3;
 *   - the library never allocates, frees or synchronises: all buffers and the
 *     scratch workspace are caller-owned; work is enqueued on `stream`;
 *   - return value: 0 on success, B2S_ERR_* (< 0) on argument errors,
 *     -(int)cudaError_t - 1000 on a CUDA launch error; b2s_error_string() names it;
 *   - thread-compatible: the only mutable global is the atomic diagnostic counter behind b2s_launch_count();
 *     one in-flight call per stream.
 */
#ifndef B200SPLAT_H
#define B200SPLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *b2s_stream_t; /* cudaStream_t */

#define B2S_OK 0
#define B2S_ERR_ARG (-1)         /* bad size / null pointer */
#define B2S_ERR_UNSUPPORTED (-2) /* channel count / tile size / degree not built */
#define B2S_ERR_WORKSPACE (-3)   /* workspace too small */

int b2s_version(void);
const char *b2s_error_string(int code);
/* number of kernels launched by this library since load (bench.py "gpu_launches") */
long long b2s_launch_count(void);

/* ---- projection + EWA covariance (upstream fully_fused_projection fwd; SURVEY A.1, A.2 pass 1) ----
 * Also emits, per Gaussian: the tile count of upstream isect_tiles pass 1 and its tile rectangle (tile_rects), the
 * depth sort key (float bits of depth, 0xFFFFFFFF when culled), the packed blend record
 *   geo[g]     = (conic a, conic b, conic c, opacity * compensation)
 *   colpack[g] = (colors_in[0..d_in), depth if with_depth, zero pad) , cdim floats
 * and tight_rects: tile_rects intersected with the tiles whose pixel centres the footprint {alpha >= 1/255} of the
 * Gaussian can reach -- the rectangles the blend's own tile lists are built from (not an upstream output).
 * totals (int64[5], zero-filled by the caller) receives the list sizes of the tile-list build over tight_rects:
 *   [0] = list length, [1] = S (tile-row hits), [2] = E1 (row-group hits), [3] = E3 ((row, column-group) hits),
 *   [4] = number of visible Gaussians (radii > 0).
 * The caller reads them back ONCE (the one device->host read of the path, overlapped with the depth sort) to size
 * the lists and the workspace of b2s_bin_tiles.  (b2s_bin_rect_totals computes the same five numbers for any other
 * rectangle array, e.g. tile_rects when upstream's lists are wanted.)  bwd_arena (may be NULL) = the N * (8 + cdim) floats of the blend
 * backward's accumulation buffers (v_xyabs | v_geo | v_colpack), zero-filled here so that no memset pass is needed.
 * comps may be NULL when calc_comp == 0.  colors_in is [N, d_in]. */
int b2s_project_fwd(const float *means, const float *quats, const float *scales,
                    const float *opacities, const float *colors_in, const float *viewmat,
                    const float *K, int N, int W, int H, int tile_size, int tile_w, int tile_h,
                    float eps2d, float near_plane, float far_plane, float radius_clip,
                    int calc_comp, int d_in, int with_depth, int cdim, int32_t *radii,
                    float *means2d, float *depths, float *geo, float *comps, float *colpack,
                    int32_t *tiles_per_gauss, uint32_t *sort_keys,
                    int32_t *tile_rects /* [N,2]: x0 | x1 << 16, y0 | y1 << 16 */,
                    int32_t *tight_rects /* [N,2], same packing */, int64_t *totals /* [5] */,
                    float *bwd_arena, b2s_stream_t stream);

/* upstream fully_fused_projection bwd fused with the opacity*compensation and
 * colour/depth un-packing VJPs.  v_means2d[g * v_means2d_stride + {0,1}] (stride 2, or 4 when it aliases
 * the blend's (xy, |xy|) arena), v_geo[g] = (v_conic a,b,c, v_opacity_eff), v_colpack[g][cdim].
 * v_colors (may be NULL): contiguous [N, d_in] copy of the colour part of v_colpack (zeros for culled rows).
 * v_viewmat (16 floats, pre-zeroed) may be NULL. */
int b2s_project_bwd(const float *means, const float *quats, const float *scales,
                    const float *opacities, const float *viewmat, const float *K, int N, int W,
                    int H, float eps2d, int calc_comp, int d_in, int with_depth, int cdim,
                    const int32_t *radii, const float *geo, const float *comps,
                    const float *v_means2d, int v_means2d_stride, const float *v_geo,
                    const float *v_colpack,
                    float *v_means, float *v_quats, float *v_scales, float *v_opacities, float *v_colors,
                    float *v_viewmat, b2s_stream_t stream);

/* ---- multi-GPU: projection backward fused with the exchange of the shared-node gradients ----
 * One traversal camera per GPU over replicated shared Gaussians (mtgs_scene_graph.py:548 renders one camera per
 * call); rows [0, n_shared) of the rasterizer inputs are the replicated ones, rows [n_shared, N) are rank-local.
 * Each rank owns shard = b2s_exchange_shard_rows(n_shared, world) consecutive rows.  d_col = d_in when
 * exchange_colors != 0 (the colours are replicated view-independent inputs), else 0: colour gradients stay in
 * v_colpack on the GPU (view-dependent colours, e.g. SH evaluated per camera, must be reduced at their own leaves
 * because their Jacobian differs from rank to rank).  Buffers (peer-mapped through CUDA IPC, allocated with
 * b2s_peer_alloc so that they are exportable):
 *   stage  [world][(11 + d_col) * shard] floats  partial gradient rows of MY shard, one slot per source rank
 *   arena  [(11 + d_col) * rows_cap]     floats  reduced gradient, SoA blocks: means at 0, quats at 3 rows_cap,
 *                                                scales at 7 rows_cap, opacities at 10 rows_cap, colours at 11 rows_cap
 *   flags  [2][8] uint32 (zero-initialised)     phase-0 / phase-1 arrival flags, written by the peers
 * The call enqueues: projection backward storing its rows straight into the owners' stage slots (peer stores), the
 * plain projection backward for the rank-local rows, the reduce of my shard + store of the result into every rank's
 * arena, and a stream-ordered wait.  When it has run, arena holds scale * sum over ranks for the shared rows and the
 * local gradient for the others.  *_ptrs_host are HOST arrays of `world` device pointers (this rank's own buffers at
 * index `rank`); epoch must increase by one per call on every rank, or be 0 on every rank and call: the library then
 * counts the steps itself in word 32 of this rank's flag buffer (no launch argument changes from step to step, so the
 * whole step can be captured in a CUDA graph; flag buffers are 256 zero-initialised bytes).  The waits block like a library collective (ranks
 * must stay in lock-step) and give up after timeout_s seconds (<= 0: 120 s): then *status (a word the device can write;
 * pinned host memory lets the host poll it without synchronising) becomes non-zero AND the shared rows of the arena are
 * filled with NaN -- never stale or partial gradients.  Colour gradients of the rank-local rows are left in v_colpack.
 * phases: bit 0 = projection backward + peer stores + "partials delivered" flags, bit 1 = reduce + broadcast,
 * bit 2 = "result delivered" flags, bit 3 = final wait (15 = all; the split exists so that a test can play several
 * ranks on one GPU from one stream). */
int b2s_exchange_shard_rows(int n_shared, int world);
int b2s_peer_alloc(size_t bytes, void **dev_ptr); /* cudaMalloc + zero fill (set-up only) */
int b2s_peer_free(void *dev_ptr);
int b2s_ipc_export(void *dev_ptr, unsigned char handle_out[64]);
int b2s_ipc_import(const unsigned char handle[64], void **peer_ptr);
int b2s_ipc_close(void *peer_ptr);
int b2s_project_bwd_exchange(const float *means, const float *quats, const float *scales,
                             const float *opacities, const float *viewmat, const float *K, int N, int W,
                             int H, float eps2d, int calc_comp, int d_in, int with_depth, int cdim,
                             const int32_t *radii, const float *geo, const float *comps,
                             const float *v_means2d, int v_means2d_stride, const float *v_geo,
                             const float *v_colpack, float *v_viewmat, int n_shared, int exchange_colors,
                             int world, int rank, long long rows_cap, float scale, unsigned epoch, int phases,
                             float timeout_s, const unsigned long long *stage_ptrs_host,
                             const unsigned long long *arena_ptrs_host,
                             const unsigned long long *flag_ptrs_host, unsigned *status,
                             b2s_stream_t stream);

/* ---- tile binning + depth sort (upstream isect_tiles + radix sort + isect_offset_encode; A.2) ----
 * Two-level formulation with identical results to the stable 64-bit sort:
 *   (1) b2s_bin_sort_depth : stable sort of the visible Gaussians by depth key (one cooperative kernel) ->
 *       order[0..n_vis), *n_vis (device int32).  Entries of order beyond n_vis are undefined.
 *   (2) b2s_bin_tiles      : hierarchy of order-preserving filters (Gaussians -> row groups -> tile rows ->
 *       column groups -> tiles) walking the Gaussians in depth order -> flatten_ids, isect_offsets (no sort of the
 *       intersections).  totals_host = {list length, S, E1, E3, n_vis} for the rectangles passed in (b2s_project_fwd's
 *       totals for tight_rects, b2s_bin_rect_totals for any other array); they size flatten_ids (list length
 *       entries) and the workspace.
 *       means2d == geo == NULL: upstream's lists (every tile of the rectangle), bit-identical flatten_ids /
 *       isect_offsets.  means2d, geo given ("exact" mode, with tight_rects): the last level keeps only the (Gaussian,
 *       tile) pairs that can reach alpha >= 1/255 at a pixel centre of the tile -- the lists the blend kernels walk;
 *       the list is then at most `list length` long.  offsets_with_total != 0: isect_offsets has tile_w * tile_h + 1
 *       entries, the last one = the final list length (so the blend needs no host-side count).
 *       overflow (device int32, zero-filled by the caller; may be NULL): CAPACITY MODE -- totals_host are then not the
 *       exact sizes but capacities chosen by the caller (e.g. the previous frame's sizes plus headroom), so that the
 *       build can be enqueued without waiting for this frame's sizes.  If a level needs more room than provided, the
 *       word becomes non-zero, every later kernel of the build returns at once (nothing is written out of bounds)
 *       and the lists are invalid: the caller must rebuild with exact sizes (b2s_blend_fwd takes the same word as
 *       skip_flag and does nothing when it is set).
 *   (3) b2s_bin_isect_ids  : optional, rebuilds upstream's int64 isect_ids for inspection. */
int b2s_bin_rect_totals(const int32_t *rects, int N, int tile_w, int tile_h, int64_t *totals /* [5] */,
                        b2s_stream_t stream);
size_t b2s_bin_depth_workspace_bytes(int N);
int b2s_bin_sort_depth(const uint32_t *sort_keys, int N, int32_t *order, int32_t *n_vis, void *workspace,
                       size_t workspace_bytes, b2s_stream_t stream);
/* diagnostic: as b2s_bin_sort_depth, plus 13 %globaltimer stamps of the kernel's phases (device uint64[13]) */
int b2s_debug_sort_depth_phases(const uint32_t *sort_keys, int N, int32_t *order, int32_t *n_vis, void *workspace,
                                size_t workspace_bytes, unsigned long long *phase_ns, b2s_stream_t stream);
size_t b2s_bin_tiles_workspace_bytes(const long long *totals_host, int tile_w, int tile_h);
int b2s_bin_tiles(const int32_t *rects, const int32_t *order, const int32_t *n_vis,
                  const long long *totals_host, int N, int tile_size, int tile_w, int tile_h, int W, int H,
                  const float *means2d, const float *geo, int offsets_with_total, int32_t *overflow,
                  int levels /* 4: per-tile lists; 3: stop at the (tile row, column group) lists */,
                  int32_t *flatten_ids, int32_t *isect_offsets, void *workspace,
                  size_t workspace_bytes, b2s_stream_t stream);
/* levels == 3: where the (tile row y, column group c) lists live inside `workspace` (same totals): items = int2
 * (Gaussian id, tile-column range x0 | x1 << 16), list index y * ncg + c, offsets = int32 [nlists + 1].  Tile (y, x)
 * belongs to list y * ncg + (x >> cg_shift) and is covered by an item iff x0 <= x < x1 -- the last filter level, which
 * b2s_blend_fwd applies lazily while it walks (only ~16 % of every list is read before a tile saturates). */
int b2s_bin_tiles_l3_view(const long long *totals_host, int tile_w, int tile_h, size_t *items_byte_offset,
                          size_t *offsets_byte_offset, int *nlists, int *ncg, int *cg_shift);
int b2s_bin_isect_ids(const int32_t *isect_offsets, int n_tiles, const int32_t *flatten_ids,
                      const float *depths, long long M, int64_t *isect_ids, b2s_stream_t stream);

/* ---- alpha blending (upstream rasterize_to_pixels fwd / bwd; A.3, A.4) ----
 * cdim in {4, 8}; d_out = channels written per pixel (<= cdim); expected_depth != 0 divides channel
 * d_out-1 by max(alpha, 1e-10) in the epilogue (upstream does that in torch).  16x16 tiles only.
 * The forward walks the depth-ordered (tile row, column group) lists of b2s_bin_tiles(levels = 3) /
 * b2s_bin_tiles_l3_view: per tile it keeps the items that cover the tile's column AND can reach alpha >= 1/255 at a
 * pixel centre of the tile (exact test; dropping the others changes no output bit), packs them densely into blocks of
 * 128 and blends block by block.  last_ids [H, W]: tile-local dense index of the last entry each pixel blended.
 * records (may be NULL for a forward without backward; 128-byte aligned, b2s_blend_record_bytes(pairs, tiles, cdim)
 * bytes with pairs = the tight list length of b2s_project_fwd's totals): every blended block is stored as a "walk
 * record" (projected mean, log2-domain conic, opacity, Gaussian id, colours of its entries, SoA; header: entries,
 * previous block of the tile) through one TMA bulk store at an index drawn from *block_counter (device uint32, zeroed
 * by the caller); tile_blocks [tiles, 2] receives (last block, number of blocks) of every tile.  The backward replays
 * each tile's chain back to front through double-buffered TMA bulk loads and never touches the lists or the
 * per-Gaussian arrays.  v_xyabs [N,4] = (v_mean2d xy, |v_mean2d| xy), v_geo [N,4] = (v_conic abc, v_opacity_eff),
 * v_colpack [N,cdim]: accumulated with 16-byte vector reductions, zero-filled by the caller (b2s_project_fwd does it).
 * skip_flag: the overflow word of a capacity-mode b2s_bin_tiles (may be NULL): when it is set neither kernel does
 * anything.  record_blocks = b2s_blend_record_blocks(pairs, tiles), the blocks `records` has room for: a forward that
 * needs more (pairs was a capacity, not this frame's size) writes nothing out of bounds and sets *skip_flag = 9
 * (exact sizes can never run out). */
uint32_t b2s_blend_record_blocks(long long pair_capacity, int n_tiles);
size_t b2s_blend_record_bytes(long long pair_capacity, int n_tiles, int cdim);
int b2s_blend_fwd(const float *means2d, const float *geo, const float *colpack, const int32_t *list_offsets,
                  const int32_t *list_items, int ncg, int cg_shift, int W, int H, int tile_w, int tile_h,
                  int cdim, int d_out, int expected_depth, float *render, float *alpha, int32_t *last_ids,
                  float *records, uint32_t record_blocks, uint32_t *block_counter, int32_t *tile_blocks,
                  int32_t *skip_flag, b2s_stream_t stream);
int b2s_blend_bwd(const int32_t *tile_blocks, const float *records, int W, int H, int tile_w, int tile_h,
                  int cdim, int d_out, int expected_depth, const float *render, const float *alpha,
                  const int32_t *last_ids, const float *v_render, const float *v_alpha, float *v_xyabs,
                  float *v_geo, float *v_colpack, int px_per_thread /* 0 = default; 4 or 8 (tuning / tests) */,
                  const int32_t *skip_flag, b2s_stream_t stream);

/* ---- spherical harmonics (upstream compute_sh fwd / bwd; A.6) ----
 * dirs [N,3], coeffs [N,K,3], masks (uint8, may be NULL), degree in 0..4 with (degree+1)^2 <= K. */
int b2s_sh_fwd(int degree, const float *dirs, const float *coeffs, const uint8_t *masks, int N,
               int K, float *colors, b2s_stream_t stream);
int b2s_sh_bwd(int degree, const float *dirs, const float *coeffs, const uint8_t *masks, int N,
               int K, const float *v_colors, float *v_coeffs, float *v_dirs /* may be NULL */,
               b2s_stream_t stream);

/* ---- masked SSIM (reference mtgs/utils/ssim.py:56-108, called at mtgs_scene_graph.py:822-840) ----
 * X, Y: [N, C, H, W] contiguous fp32; win: the 1-D Gaussian window (device, win_size odd <= 15, H and W >= win_size);
 * "valid" filtering as the reference: the SSIM map is [N, C, H - win_size + 1, W - win_size + 1].
 * mask: uint8 over the FULL image (the kernel reads it at the window centres, i.e. cropped by win_size / 2 as the
 *   reference does), addressed mask[n * mask_n_stride + c * mask_c_stride + y * W + x] (c stride 0 = shared by the
 *   channels); NULL = no mask.
 * fwd: acc[plane] = (sum of mask * ssim, sum of mask) as doubles, pre-zeroed by the caller, plane = n * C + c; map_a,
 *   map_b, map_c (SSIM-map shaped) receive mask * dS/dmu_Y, mask * dS/dE[YY], mask * dS/dE[XY]; map_ax (may be NULL)
 *   mask * dS/dmu_X, needed only for the gradient w.r.t. X.
 * bwd: grad = plane_scale[plane] * (F^T map_a + 2 self F^T map_b + other F^T map_c) with F^T the adjoint of the
 *   filter; gradient w.r.t. Y: (self, other, map_a) = (Y, X, map_a); w.r.t. X: (X, Y, map_ax). */
int b2s_ssim_fwd(const float *X, const float *Y, const uint8_t *mask, long long mask_n_stride,
                 long long mask_c_stride, int N, int C, int H, int W, const float *win, int win_size, float C1,
                 float C2, float *map_a, float *map_b, float *map_c, float *map_ax, double *acc,
                 b2s_stream_t stream);
int b2s_ssim_bwd(const float *self, const float *other, const float *map_a, const float *map_b,
                 const float *map_c, const float *plane_scale, int N, int C, int H, int W, const float *win,
                 int win_size, float *grad, b2s_stream_t stream);

/* ---- other image-space losses of the training step ----
 * Reference op sequences: masked L1 on RGB / normals mtgs_scene_graph.py:825-828, 929; LiDAR depth loss (L1 /
 * InverseL1) :875-884; TVLoss mtgs/utils/geometric_loss.py:287-303; calculate_depth_ncc_loss :322-348;
 * pcd_to_normal / normal_from_depth_image :350-388 (+ camera_utils.py:74-148).  Every forward accumulates (sum, count)
 * -- or the two directional sums for TV -- into acc (device double[2], zero-filled by the caller); the mean is formed
 * on the device by the caller, nothing synchronises.  Backward passes take the upstream gradient as a DEVICE scalar.
 * masked L1: pred, gt [P, C]; mask uint8 [P] or NULL; mode 0 = |gt - pred|, mode 1 = |1/(gt+eps) - 1/(pred+eps)|.
 * TV: pred [B, H, W, C].  NCC: depth maps [H, W], mask uint8 [H, W]; patches patch x patch at `stride`, zero padding
 * patch / 2 as F.unfold; stats [npy * npx, 6] (b2s_ncc_patch_grid gives npy, npx) is written by the forward and read
 * by the backward.  normal_from_depth: A_t = 12 device floats, inv(c2w[:3,:3]) row-major then c2w[:3,3]. */
int b2s_masked_l1_fwd(const float *pred, const float *gt, const uint8_t *mask, long long P, int C, int mode,
                      float eps, double *acc, b2s_stream_t stream);
int b2s_masked_l1_bwd(const float *pred, const float *gt, const uint8_t *mask, long long P, int C, int mode,
                      float eps, const double *acc, const float *grad_out, float *grad_pred, b2s_stream_t stream);
int b2s_tv_fwd(const float *pred, int B, int H, int W, int C, double *acc, b2s_stream_t stream);
int b2s_tv_bwd(const float *pred, int B, int H, int W, int C, const float *grad_out, float *grad_pred,
               b2s_stream_t stream);
int b2s_ncc_patch_grid(int H, int W, int patch, int stride, int *npy, int *npx);
int b2s_ncc_fwd(const float *pred, const float *gt, const uint8_t *mask, int H, int W, int patch, int stride,
                float *stats, double *acc, b2s_stream_t stream);
int b2s_ncc_bwd(const float *pred, const float *gt, int H, int W, int patch, int stride, const float *stats,
                const double *acc, const float *grad_out, float *grad_pred, b2s_stream_t stream);
int b2s_normal_from_depth(const float *depth, int H, int W, float fx, float fy, float cx, float cy,
                          const float *A_t, float *normals, b2s_stream_t stream);

/* ---- Gaussian arena: node gather + activations + SH colour in one kernel each way ----
 * Reference: per-node chains of torch ops (gaussian_model/vanilla_gaussian_splatting.py:299-322, rigid_node.py:206-215,
 * 243-252) followed by a per-attribute torch.cat over the nodes (mtgs_scene_graph.py:408-461).  All nodes live in one
 * structure-of-arrays arena: means [N,3], scales_raw [N,3] (log), quats_raw [N,4] (wxyz, any norm), opac_raw [N]
 * (logit), sh [N,K,3] (features_dc then features_rest); node_of [N] (int32, may be NULL = one node) indexes poses
 * [n_nodes,16] = R (9, row-major) | t (3) | q (4, wxyz) of each node for this frame; campos [3] = camera centre.
 * fwd: means_w = means R^T + t, quats_w = q_node (x) quats / ||quats||, scales = exp, opac = sigmoid, colors [N,3] =
 *   clamp(SH_degree(normalise(means_w - campos), sh) + 0.5, 0, 1) (degree 0: sigmoid(sh[:,0])); clamp_mask [N] (uint8)
 *   records which channels lie strictly inside the clamp.  bwd: gradients of the raw arrays from those of the five
 *   outputs (view directions are detached, as in the reference; no gradient to the node poses). */
int b2s_arena_fwd(const float *means, const float *scales_raw, const float *quats_raw, const float *opac_raw,
                  const float *sh, const int32_t *node_of, const float *poses, const float *campos, int N, int K,
                  int degree, float *means_w, float *quats_w, float *scales, float *opac, float *colors,
                  uint8_t *clamp_mask, b2s_stream_t stream);
int b2s_arena_bwd(const float *means, const float *scales_raw, const float *quats_raw, const float *opac_raw,
                  const float *sh, const int32_t *node_of, const float *poses, const float *campos, int N, int K,
                  int degree, const uint8_t *clamp_mask, const float *v_means_w, const float *v_quats_w,
                  const float *v_scales, const float *v_opac, const float *v_colors, float *g_means, float *g_quats,
                  float *g_scales, float *g_opac, float *g_sh, b2s_stream_t stream);

/* ---- fused multi-tensor Adam + densification primitives ----
 * Reference: one torch.optim.Adam per (node, attribute) group (mtgs/scene_model/custom_trainer.py:115-136), the
 * after_train statistics and optimizer-state surgery of gaussian_model/vanilla_gaussian_splatting.py:392-474 and the
 * cull / split / duplicate steps of :476-699.
 * b2s_adam_multi: ONE launch steps every tensor: tensors_dev = device array of descriptors, chunk c of
 *   b2s_adam_chunk() elements belongs to tensor chunk_tensor_dev[c] and starts at element chunk_start_dev[c].
 *   Arithmetic = torch.optim.Adam (no amsgrad); the step-dependent scalars are computed by the host in double.
 * b2s_densify_stats: for radii > 0: xys_grad_norm += ||grad2d * (W, H) / 2||, vis_counts += 1, max_2dsize =
 *   max(max_2dsize, radii)   (grad2d rows have grad_stride floats: 2, or 4 for the blend's (xy, |xy|) arena).
 * b2s_mask_scan + b2s_mask_gather_rows: order-preserving compaction of [N, row_floats] rows by a uint8 keep mask;
 *   scan once (workspace = b2s_mask_workspace_bytes(N); *total = rows kept, device int32), then gather every tensor
 *   that shares the mask (parameters and their Adam moments). */
typedef struct B2sAdamTensor {
    float *p;
    const float *g;
    float *m;
    float *v;
    long long n;
    float lr, beta1, beta2, eps, weight_decay;
    /* scalars torch.optim.Adam forms in double on the host: lr / (1 - beta1^t), sqrt(1 - beta2^t), 1 - beta1, 1 - beta2 */
    float step_size, bias_correction2_sqrt, one_minus_beta1, one_minus_beta2, _pad;
} B2sAdamTensor;
int b2s_adam_chunk(void);
int b2s_adam_multi(const B2sAdamTensor *tensors_dev, const int32_t *chunk_tensor_dev,
                   const long long *chunk_start_dev, int n_chunks, b2s_stream_t stream);
int b2s_densify_stats(const float *grad2d, int grad_stride, const int32_t *radii, int N, int W, int H,
                      float *xys_grad_norm, float *vis_counts, float *max_2dsize, b2s_stream_t stream);
size_t b2s_mask_workspace_bytes(long long N);
int b2s_mask_scan(const uint8_t *keep, long long N, void *workspace, size_t workspace_bytes, int32_t *total,
                  b2s_stream_t stream);
int b2s_mask_gather_rows(const uint8_t *keep, long long N, const void *workspace, const float *src, float *dst,
                         int row_floats, b2s_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SPLAT_H */
