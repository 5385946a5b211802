"""``rasterization`` -- drop-in for ``gsplat.rendering.rasterization`` as MTGS calls it.

Reference call site: mtgs/scene_model/mtgs_scene_graph.py:641-662 (kwargs), consumers of the returns at
:663-690 (``render``/``alpha``), :666-670 and :1157-1183 (``info["means2d"]`` with ``.retain_grad()`` /
``.grad`` / ``.absgrad``, ``info["radii"]``).  Signature, argument meaning, return layout and error
behaviour follow upstream gsplat v1.4.0 (requirements.txt:12); the arithmetic runs in hand-written
sm_100a kernels behind the C ABI of ``include/b200splat.h``.  PyTorch is used for device memory, the
autograd graph and the current stream only.  There is no CPU / PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Optional per-stage CUDA-event timing (bench.py sets PROFILE = {} to collect (start, end) event pairs per
# stage on the launching stream; None = zero overhead).
PROFILE = None
# pixels per thread of the blend backward: 0 = the library's default; tools / tests may set 4 or 8
BWD_PX = 0


class _timed:
    def __init__(self, stage: str):
        self.stage = stage

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record()
            PROFILE.setdefault(self.stage, []).append((self.a, self.b))
        return False


def _need_cuda(*ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mtgs_b200 kernels need CUDA tensors (no CPU fallback path exists)")


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32 tensor, got {t.dtype}")
    return t.contiguous()


def padded_channels(ch: int) -> int:
    """Blend width for ``ch`` payload channels (4 or 8; upstream pads to its own compile-time set)."""
    if ch < 1 or ch > 8:
        raise NotImplementedError(f"{ch} blended channels: this build instantiates 1..8 (MTGS uses 3, 4, 6, 7)")
    return 4 if ch <= 4 else 8


class Meta(dict):
    """``info`` dict.  ``flatten_ids`` / ``isect_offsets`` / ``isect_ids`` (upstream's lists over the full 3-sigma
    rectangles) are built on first access: the blend kernels walk their own, exactly culled lists, and MTGS never
    reads these keys.  When read they are bit-identical to upstream's."""

    _LAZY = ("flatten_ids", "isect_offsets", "isect_ids")
    _lazy = None

    def __missing__(self, key):
        if key in self._LAZY and self._lazy is not None:
            self.update(self._lazy(key))
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or (key in self._LAZY and self._lazy is not None)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


# ------------------------------------------------------------------------------------------------
class _Project(torch.autograd.Function):
    """Projection + EWA (+ tile count, sort key, blend record packing).  One kernel each way."""

    @staticmethod
    def forward(ctx, means, quats, scales, opacities, colors, viewmat, K, W, H, tile_w, tile_h, eps2d, near, far,
                radius_clip, calc_comp, with_depth, cdim, want_arena):
        lib = _lib.load()
        N = means.shape[0]
        d_in = 0 if colors is None else colors.shape[1]
        dev = means.device
        radii = torch.empty(N, dtype=torch.int32, device=dev)
        # every row is written by the kernel (zeros for culled Gaussians): no memset passes
        means2d = torch.empty(1, N, 2, dtype=torch.float32, device=dev)
        depths = torch.empty(N, dtype=torch.float32, device=dev)
        geo = torch.empty(N, 4, dtype=torch.float32, device=dev)
        comps = torch.empty(N, dtype=torch.float32, device=dev) if calc_comp else None
        colpack = torch.empty(N, cdim, dtype=torch.float32, device=dev)
        tiles = torch.empty(N, dtype=torch.int32, device=dev)
        keys = torch.empty(N, dtype=torch.int32, device=dev)
        rects = torch.empty(N, 2, dtype=torch.int32, device=dev)
        tight = torch.empty(N, 2, dtype=torch.int32, device=dev)
        totals = torch.zeros(8, dtype=torch.int64, device=dev)  # 5 list sizes, capacity-overflow word, record-block counter
        # accumulation buffers of the blend backward (v_xyabs | v_geo | v_colpack), zero-filled by the kernel
        arena = torch.empty(N * (8 + cdim) if want_arena else 0, dtype=torch.float32, device=dev)
        with _timed("project_fwd"):
            _lib.check(lib.b2s_project_fwd(_ptr(means), _ptr(quats), _ptr(scales), _ptr(opacities), _ptr(colors),
                                           _ptr(viewmat), _ptr(K), N, W, H, 16, tile_w, tile_h, eps2d, near, far,
                                           radius_clip, int(calc_comp), d_in, int(with_depth), cdim, _ptr(radii),
                                           _ptr(means2d), _ptr(depths), _ptr(geo), _ptr(comps), _ptr(colpack),
                                           _ptr(tiles), _ptr(keys), _ptr(rects), _ptr(tight), _ptr(totals),
                                           _ptr(arena) if want_arena and N > 0 else None, _stream()),
                       "b2s_project_fwd")
        ctx.save_for_backward(means, quats, scales, opacities, viewmat, K, radii, geo, comps)
        ctx.set_materialize_grads(False)  # no zero tensors for the gradients of the integer / scratch outputs
        ctx.cfg = (W, H, eps2d, calc_comp, d_in, with_depth, cdim)
        ctx.has_colors = colors is not None
        ctx.mark_non_differentiable(radii, depths, tiles, keys, rects, tight, totals, arena)
        return means2d, geo, colpack, radii, depths, tiles, keys, rects, tight, totals, arena

    @staticmethod
    def backward(ctx, v_means2d, v_geo, v_colpack, *_unused):
        lib = _lib.load()
        means, quats, scales, opacities, viewmat, K, radii, geo, comps = ctx.saved_tensors
        W, H, eps2d, calc_comp, d_in, with_depth, cdim = ctx.cfg
        N = means.shape[0]
        dev = means.device
        if v_means2d is None:
            v_means2d = torch.zeros(N, 2, dtype=torch.float32, device=dev)
        v_means2d = v_means2d.reshape(N, 2)
        if not (v_means2d.stride(1) == 1 and v_means2d.stride(0) in (2, 4) and v_means2d.data_ptr() % 8 == 0):
            v_means2d = v_means2d.contiguous()
        v_geo = torch.zeros(N, 4, dtype=torch.float32, device=dev) if v_geo is None else v_geo.contiguous()
        v_colpack = (torch.zeros(N, cdim, dtype=torch.float32, device=dev) if v_colpack is None
                     else v_colpack.contiguous())
        need_view = ctx.needs_input_grad[5]
        v_view = torch.zeros(4, 4, dtype=torch.float32, device=dev) if need_view else None
        from . import parallel
        ex = parallel.current_exchange()
        if ex is not None:
            # multi-GPU: gradients leave the kernel straight into the owners' peer memory and come back reduced
            if ex.d_in != d_in or N > ex.rows_cap or N < ex.n_shared or not ctx.has_colors:
                raise RuntimeError(f"GradExchange(n_shared={ex.n_shared}, d_in={ex.d_in}, rows_cap={ex.rows_cap}) "
                                   f"does not match this call (N={N}, d_in={d_in})")
            args = (_ptr(means), _ptr(quats), _ptr(scales), _ptr(opacities), _ptr(viewmat), _ptr(K), N, W, H, eps2d,
                    int(calc_comp), d_in, int(with_depth), cdim, _ptr(radii), _ptr(geo), _ptr(comps), _ptr(v_means2d),
                    int(v_means2d.stride(0)), _ptr(v_geo), _ptr(v_colpack), _ptr(v_view))
            if PROFILE is not None and ex._phases == 15:  # per-phase timing for bench.py
                with _timed("project_bwd_exchange"):
                    with _timed("exch_k1_project_bwd_peer_stores"):
                        ex.launch(1, args)
                    with _timed("exch_k2_reduce_bcast"):
                        ex.launch(2)
                    with _timed("exch_k3_wait"):
                        ex.launch(4 | 8)
            else:
                ex.launch(ex._phases, args)
            gv = ex.grad_views(N)
            if ex.exchange_colors and N > ex.n_shared:  # rank-local rows: colour gradients join the arena as they are
                gv["colors"][ex.n_shared:] = v_colpack[ex.n_shared:, :d_in]
            if not ex.zero_copy:
                gv = {k: v.clone() for k, v in gv.items()}
            # view-dependent colours (exchange_colors=False) are reduced at their own leaves by the caller
            v_colors = gv["colors"] if ex.exchange_colors else v_colpack[:, :d_in]
            return (gv["means"], gv["quats"], gv["scales"], gv["opacities"], v_colors, v_view) + (None,) * 13
        v_means = torch.empty_like(means)
        v_quats = torch.empty_like(quats)
        v_scales = torch.empty_like(scales)
        v_opac = torch.empty_like(opacities)
        want_colors = ctx.has_colors and ctx.needs_input_grad[4]
        v_colors = torch.empty(N, d_in, dtype=torch.float32, device=dev) if want_colors else None
        with _timed("project_bwd"):
            _lib.check(lib.b2s_project_bwd(_ptr(means), _ptr(quats), _ptr(scales), _ptr(opacities), _ptr(viewmat),
                                           _ptr(K), N, W, H, eps2d, int(calc_comp), d_in, int(with_depth), cdim,
                                           _ptr(radii), _ptr(geo), _ptr(comps), _ptr(v_means2d),
                                           int(v_means2d.stride(0)), _ptr(v_geo), _ptr(v_colpack), _ptr(v_means),
                                           _ptr(v_quats), _ptr(v_scales), _ptr(v_opac), _ptr(v_colors), _ptr(v_view),
                                           _stream()),
                       "b2s_project_bwd")
        return (v_means, v_quats, v_scales, v_opac, v_colors, v_view) + (None,) * 13


class _Blend(torch.autograd.Function):
    """Per-tile front-to-back blend.  ``means2d`` is a formal input so ``retain_grad()`` / ``.absgrad`` work
    exactly as with upstream (mtgs_scene_graph.py:666-667, 1171-1174)."""

    @staticmethod
    def forward(ctx, means2d, geo, colpack, lists, arena, W, H, tile_w, tile_h, cdim, d_out, ed, absgrad, pair_cap,
                totals, use_flag=False):
        lib = _lib.load()
        # words of the projection's `totals`: [6] = record-block bump allocator (zeroed with the totals),
        # [5] = the capacity-overflow word
        counter_ptr = C.c_void_p(totals.data_ptr() + 48)
        skip_ptr = C.c_void_p(totals.data_ptr() + 40) if use_flag else None
        ctx.totals = totals  # the backward reads the overflow word: keep the tensor alive
        dev = means2d.device
        ws, items_ptr, offs_ptr, ncg, cg_shift = lists  # (row, column-group) lists inside the tile-list workspace
        render = torch.empty(1, H, W, d_out, dtype=torch.float32, device=dev)
        alpha = torch.empty(1, H, W, 1, dtype=torch.float32, device=dev)
        last_ids = torch.empty(H, W, dtype=torch.int32, device=dev)
        records = tile_blocks = None
        nblocks = 0
        if arena.numel() > 0:  # a backward may follow: keep the walk records
            nbytes = int(lib.b2s_blend_record_bytes(pair_cap, tile_w * tile_h, cdim))
            nblocks = int(lib.b2s_blend_record_blocks(pair_cap, tile_w * tile_h))
            records = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
            tile_blocks = torch.empty(tile_h * tile_w, 2, dtype=torch.int32, device=dev)
        with _timed("blend_fwd"):
            _lib.check(lib.b2s_blend_fwd(_ptr(means2d), _ptr(geo), _ptr(colpack), offs_ptr, items_ptr, ncg, cg_shift, W, H,
                                         tile_w, tile_h, cdim, d_out, int(ed), _ptr(render), _ptr(alpha),
                                         _ptr(last_ids), _ptr(records), nblocks, counter_ptr, _ptr(tile_blocks),
                                         skip_ptr, _stream()), "b2s_blend_fwd")
        ctx.save_for_backward(means2d, render, alpha, last_ids, records, tile_blocks, arena)
        ctx.cfg = (W, H, tile_w, tile_h, cdim, d_out, ed, absgrad, geo.shape[0])
        ctx.skip_ptr = skip_ptr
        ctx.set_materialize_grads(False)
        ctx.arena_clean = True
        ctx.mark_non_differentiable(last_ids)
        return render, alpha, last_ids

    @staticmethod
    def backward(ctx, v_render, v_alpha, _v_last=None):
        lib = _lib.load()
        means2d, render, alpha, last_ids, records, tile_blocks, arena = ctx.saved_tensors
        W, H, tile_w, tile_h, cdim, d_out, ed, absgrad, N = ctx.cfg
        if records is None:
            raise RuntimeError("rasterization was run without gradient tracking; no backward is possible")
        v_render = torch.zeros_like(render) if v_render is None else v_render.contiguous()
        v_alpha = torch.zeros_like(alpha) if v_alpha is None else v_alpha.contiguous()
        if not ctx.arena_clean:  # a second backward through the same graph (retain_graph=True)
            arena = torch.zeros_like(arena)
        ctx.arena_clean = False
        # one arena holding the three 16-byte-row accumulation buffers, zero-filled by the projection forward
        v_xyabs = arena[: 4 * N].view(N, 4)
        v_geo = arena[4 * N: 8 * N].view(N, 4)
        v_colpack = arena[8 * N:].view(N, cdim)
        if N > 0:
            with _timed("blend_bwd"):
                _lib.check(lib.b2s_blend_bwd(_ptr(tile_blocks), _ptr(records), W, H, tile_w, tile_h, cdim, d_out, int(ed),
                                             _ptr(render), _ptr(alpha), _ptr(last_ids), _ptr(v_render), _ptr(v_alpha),
                                             _ptr(v_xyabs), _ptr(v_geo), _ptr(v_colpack), BWD_PX, ctx.skip_ptr,
                                             _stream()),
                           "b2s_blend_bwd")
        if absgrad:
            # upstream: `means2d.absgrad = v_means2d_abs` on the tensor object handed in by the caller
            means2d.absgrad = v_xyabs[:, 2:4].unsqueeze(0)
        return (v_xyabs[:, 0:2].unsqueeze(0), v_geo, v_colpack) + (None,) * 13


# ------------------------------------------------------------------------------------------------
_PINNED_TOTALS: Dict[int, tuple] = {}
# Capacities of the tile-list build per (device, tile grid): the largest list sizes seen so far plus headroom.  With
# them the whole forward is enqueued WITHOUT waiting for this frame's sizes (they are read back, asynchronously, and
# checked once everything is queued); a frame that needs more room is rebuilt with its exact sizes.
_CAPACITY: Dict[tuple, list] = {}
CAPACITY_HEADROOM = 1.25
SYNC_SIZES = False  # True: always wait for the exact sizes before building the lists (debugging / tests)
# While a CUDA graph is being captured (mtgs_b200.graph.GraphedStep) nothing may wait for the device: the forward
# then runs in capacity mode without the host-side check, and every call leaves (totals tensor, capacities) here so
# that the owner of the graph can verify after a replay that no level overflowed.
_CAPTURED: list = []


def _start_totals_readback(totals: Tensor):
    """Asynchronous device->host copy of the list sizes summed by the projection kernel (the one device->host read
    of the path) into a pinned buffer; returns (host tensor, event)."""
    dev = totals.device
    slot = _PINNED_TOTALS.get(dev.index)
    if slot is None:
        slot = _PINNED_TOTALS[dev.index] = (torch.zeros(8, dtype=torch.int64).pin_memory(), torch.cuda.Event())
    host, ev = slot
    host.copy_(totals, non_blocking=True)
    ev.record()
    return host, ev


def _sort_depth(keys: Tensor):
    """Depth order of the visible Gaussians (runs while the list sizes travel to the host)."""
    lib = _lib.load()
    dev = keys.device
    N = keys.shape[0]
    order = torch.empty(N, dtype=torch.int32, device=dev)
    n_vis = torch.empty(1, dtype=torch.int32, device=dev)
    wsb = int(lib.b2s_bin_depth_workspace_bytes(N))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with _timed("bin_sort_depth"):
        _lib.check(lib.b2s_bin_sort_depth(_ptr(keys), N, _ptr(order), _ptr(n_vis), _ptr(ws), wsb, _stream()),
                   "b2s_bin_sort_depth")
    return order, n_vis


def _tile_lists(rects: Tensor, order: Tensor, n_vis: Tensor, sizes, tile_w: int, tile_h: int, W: int, H: int,
                walk: bool, overflow_ptr=None):
    """Depth-ordered lists over ``rects``.  ``walk``: the lists the blend walks -- the build stops at the (tile row,
    column group) level and the blend applies the last filter level lazily; returns (workspace, items pointer, offsets
    pointer, ncg, cg_shift).  Otherwise upstream's per-tile lists over the 3-sigma rectangles: (flatten_ids,
    isect_offsets).  ``sizes`` = (list length, S, E1, E3, n_vis): the exact sizes summed for ``rects``, or -- with
    ``overflow_ptr`` (capacity mode) -- capacities."""
    lib = _lib.load()
    dev = rects.device
    N = rects.shape[0]
    tot = (C.c_longlong * 5)(*sizes)
    wsb = int(lib.b2s_bin_tiles_workspace_bytes(tot, tile_w, tile_h))
    if wsb == 0:
        raise NotImplementedError(f"tile grid {tile_w}x{tile_h} not supported")
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    if walk:
        ids = offsets = None
    else:
        ids = torch.empty(sizes[0], dtype=torch.int32, device=dev)
        offsets = torch.empty(tile_h * tile_w, dtype=torch.int32, device=dev)
    with _timed("bin_tiles" if walk else "bin_tiles_upstream_lists"):
        _lib.check(lib.b2s_bin_tiles(_ptr(rects), _ptr(order), _ptr(n_vis), tot, N, 16, tile_w, tile_h, W, H,
                                     None, None, 0, overflow_ptr, 3 if walk else 4, _ptr(ids), _ptr(offsets), _ptr(ws),
                                     wsb, _stream()), "b2s_bin_tiles")
    if not walk:
        return ids, offsets
    io, oo = C.c_size_t(), C.c_size_t()
    nl, ncg, cgs = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.b2s_bin_tiles_l3_view(tot, tile_w, tile_h, C.byref(io), C.byref(oo), C.byref(nl), C.byref(ncg),
                                         C.byref(cgs)), "b2s_bin_tiles_l3_view")
    base = ws.data_ptr()
    return ws, C.c_void_p(base + io.value), C.c_void_p(base + oo.value), ncg.value, cgs.value


def _isect_ids(offsets: Tensor, flatten_ids: Tensor, depths: Tensor) -> Tensor:
    lib = _lib.load()
    M = flatten_ids.shape[0]
    out = torch.empty(M, dtype=torch.int64, device=flatten_ids.device)
    with torch.cuda.device(flatten_ids.device):
        _lib.check(lib.b2s_bin_isect_ids(_ptr(offsets), offsets.numel(), _ptr(flatten_ids), _ptr(depths), M, _ptr(out),
                                         _stream()), "b2s_bin_isect_ids")
    return out


def _rasterize_one(means, quats, scales, opacities, colors, viewmat, K, width, height, near_plane, far_plane,
                   radius_clip, eps2d, render_mode, absgrad, rasterize_mode):
    N = means.shape[0]
    with_depth = render_mode in ("RGB+D", "RGB+ED", "D", "ED")
    ed = render_mode in ("RGB+ED", "ED")
    cols = None if render_mode in ("D", "ED") else colors
    d_in = 0 if cols is None else cols.shape[1]
    d_out = d_in + (1 if with_depth else 0)
    cdim = padded_channels(d_out)
    tile_w = math.ceil(width / 16.0)
    tile_h = math.ceil(height / 16.0)
    calc_comp = rasterize_mode == "antialiased"
    want_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (means, quats, scales, opacities, cols, viewmat))
    means2d, geo, colpack, radii, depths, tiles, keys, rects, tight, totals, arena = _Project.apply(
        means, quats, scales, opacities, cols, viewmat, K, width, height, tile_w, tile_h, float(eps2d),
        float(near_plane), float(far_plane), float(radius_clip), calc_comp, with_depth, cdim, want_grad)
    capturing = torch.cuda.is_current_stream_capturing()
    if not capturing:
        host, ev = _start_totals_readback(totals)
    order, n_vis = _sort_depth(keys)
    key = (means.device.index, tile_w, tile_h)
    caps = _CAPACITY.get(key)

    def build(sizes, overflow_ptr):
        # (tile row, column group) lists over the tight rectangles; the blend applies the last filter level and the
        # exact per-tile test lazily while it walks
        lists = _tile_lists(tight, order, n_vis, sizes, tile_w, tile_h, width, height, True, overflow_ptr)
        return _Blend.apply(means2d, geo, colpack, lists, arena, width, height, tile_w, tile_h, cdim, d_out, ed,
                            bool(absgrad), sizes[0], totals, overflow_ptr is not None)

    tot = None
    if capturing:
        if caps is None:
            raise RuntimeError("run this call eagerly at least once before capturing it in a CUDA graph: the list "
                               "capacities are learnt from earlier frames")
        flag = C.c_void_p(totals.data_ptr() + 40)
        render, alpha, last_ids = build((caps[0], caps[1], caps[2], caps[3], N), flag)
        _CAPTURED.append((totals, list(caps), key))
    elif caps is None or SYNC_SIZES or N == 0:
        ev.synchronize()
        tot = [int(v) for v in host.tolist()[:5]]
        render, alpha, last_ids = build(tuple(tot), None)
    else:
        # capacity mode: everything is enqueued before this frame's sizes are known; they are checked afterwards
        flag = C.c_void_p(totals.data_ptr() + 40)
        render, alpha, last_ids = build((caps[0], caps[1], caps[2], caps[3], N), flag)
        ev.synchronize()
        tot = [int(v) for v in host.tolist()[:5]]
        if any(tot[i] > caps[i] for i in range(4)):  # rare: rebuild with the exact sizes
            totals[6] = 0  # (stream-ordered) reset of the record-block counter
            render, alpha, last_ids = build(tuple(tot), None)
    if tot is not None:
        if caps is None:
            caps = _CAPACITY[key] = [0, 0, 0, 0]
        for i in range(4):
            caps[i] = max(caps[i], int(tot[i] * CAPACITY_HEADROOM) + 4096)
    # keys with a leading underscore are not part of upstream's info dict (bench.py reads them for K_pairs)
    meta = dict(radii=radii.unsqueeze(0), means2d=means2d, depths=depths.unsqueeze(0),
                conics=geo.detach()[:, :3].unsqueeze(0), opacities=geo.detach()[:, 3].unsqueeze(0),
                tiles_per_gauss=tiles.unsqueeze(0), _last_ids=last_ids, _tight_rects=tight)

    def upstream_lists(_key):
        """upstream's flatten_ids / isect_offsets / isect_ids, on demand (bit-identical to the 64-bit sort)."""
        with torch.cuda.device(rects.device):
            lib = _lib.load()
            up = torch.empty(5, dtype=torch.int64, device=rects.device)
            _lib.check(lib.b2s_bin_rect_totals(_ptr(rects), N, tile_w, tile_h, _ptr(up), _stream()), "b2s_bin_rect_totals")
            sizes = [int(v) for v in up.tolist()]
            if sizes[0] >= 2 ** 31:
                raise RuntimeError(f"{sizes[0]} tile intersections exceed the int32 offset range (same limit as upstream)")
            sizes[4] = int(n_vis.item())  # the depth order holds every visible Gaussian
            flat, offs = _tile_lists(rects, order, n_vis, tuple(sizes), tile_w, tile_h, width, height, False)
            return dict(flatten_ids=flat, isect_offsets=offs.view(1, tile_h, tile_w),
                        isect_ids=_isect_ids(offs, flat, depths))

    return render, alpha, meta, upstream_lists


def rasterization(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]  (w, x, y, z)
    scales: Tensor,  # [N, 3]
    opacities: Tensor,  # [N]
    colors: Tensor,  # [N, D] (or [N, K, 3] SH coefficients when sh_degree is given)
    viewmats: Tensor,  # [C, 4, 4] world -> camera
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
    covars: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    """Rasterize 3D Gaussians to ``(render_colors [C,H,W,D(+1)], render_alphas [C,H,W,1], meta)``.

    ``info["depths"]`` is not differentiable here (gradient reaches the depths through the blended depth channel of
    ``RGB+D`` / ``RGB+ED`` only, which is the only way MTGS uses them).  With ``C > 1`` cameras ``info["means2d"]``
    is a concatenation of the per-camera tensors and carries no ``.grad``.

    Same contract as upstream for the argument combinations MTGS uses:
    ``packed=False, tile_size=16, sparse_grad=False``, ``render_mode`` in RGB / RGB+D / RGB+ED / D / ED,
    ``rasterize_mode`` classic / antialiased, optional ``absgrad`` and ``backgrounds``.  Other combinations
    raise ``NotImplementedError`` (never a silent fallback).
    """
    N = means.shape[0]
    C_ = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C_, 4, 4), viewmats.shape
    assert Ks.shape == (C_, 3, 3), Ks.shape
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED"), render_mode
    assert rasterize_mode in ("classic", "antialiased"), rasterize_mode
    if packed:
        raise NotImplementedError("packed=True is not built (MTGS passes packed=False, mtgs_scene_graph.py:652)")
    if tile_size != 16:
        raise NotImplementedError("only tile_size=16 is built (MTGS: BLOCK_WIDTH = 16, mtgs_scene_graph.py:640)")
    if sparse_grad or distributed or covars is not None or camera_model != "pinhole":
        raise NotImplementedError("sparse_grad / distributed / covars / non-pinhole cameras are not built")
    if near_plane < 0:
        # the depth sort key is the fp32 bit pattern of z, which orders correctly for z >= 0 only (upstream compares
        # the bits as a signed integer); MTGS passes near_plane = 0.01 (mtgs_scene_graph.py:653)
        raise NotImplementedError("near_plane < 0 is not built")
    if C_ > 1 and absgrad:
        # upstream attaches .absgrad to ONE [C, N, 2] means2d tensor; here cameras are rendered one by one and
        # info["means2d"] is a concatenation outside the graph.  MTGS renders one camera per call (:548).
        raise NotImplementedError("absgrad with more than one camera is not built (MTGS uses C = 1)")
    _need_cuda(means, quats, scales, opacities, colors, viewmats, Ks)
    if sh_degree is None:
        assert (colors.dim() == 2 and colors.shape[0] == N) or (colors.dim() == 3 and colors.shape[:2] == (C_, N)), \
            colors.shape
    else:
        assert colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3, colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[1], colors.shape

    means, quats, scales, opacities = _f32c(means), _f32c(quats), _f32c(scales), _f32c(opacities)
    viewmats, Ks = _f32c(viewmats), _f32c(Ks)
    renders, alphas, metas, lazies = [], [], [], []
    with torch.cuda.device(means.device):
        for c in range(C_):
            if sh_degree is not None:
                from .cuda._wrapper import spherical_harmonics
                campos = torch.inverse(viewmats[c])[:3, 3]
                dirs = means - campos
                cols = torch.clamp_min(spherical_harmonics(sh_degree, dirs, colors) + 0.5, 0.0)
            else:
                cols = colors if colors.dim() == 2 else colors[c]
            cols = _f32c(cols)
            r, a, m, lz = _rasterize_one(means, quats, scales, opacities, cols, viewmats[c], Ks[c], int(width),
                                         int(height), near_plane, far_plane, radius_clip, eps2d, render_mode, absgrad,
                                         rasterize_mode)
            lazies.append(lz)
            if backgrounds is not None and render_mode not in ("D", "ED"):
                nb = backgrounds.shape[-1]
                r = torch.cat([r[..., :nb] + (1.0 - a) * backgrounds[c].reshape(1, 1, 1, nb), r[..., nb:]], dim=-1)
            renders.append(r)
            alphas.append(a)
            metas.append(m)

    tile_w = math.ceil(width / 16.0)
    tile_h = math.ceil(height / 16.0)
    meta = Meta(tile_width=tile_w, tile_height=tile_h, width=width, height=height, tile_size=tile_size,
                n_cameras=C_, camera_ids=None, gaussian_ids=None)
    if C_ == 1:
        meta.update(metas[0])
        meta._lazy = lazies[0]
        return renders[0], alphas[0], meta
    # C > 1: per-camera results stacked; flatten_ids / offsets follow upstream's camera-major numbering
    for k in ("radii", "means2d", "depths", "conics", "opacities", "tiles_per_gauss"):
        meta[k] = torch.cat([m[k] for m in metas], dim=0)

    def stacked_lists(_key):
        per_cam = [lz(_key) for lz in lazies]
        m_before, offs = 0, []
        for d in per_cam:
            offs.append(d["isect_offsets"] + m_before)
            m_before += d["flatten_ids"].shape[0]
        n_tiles = tile_w * tile_h
        tile_bits = int(math.floor(math.log2(n_tiles))) + 1
        return dict(isect_offsets=torch.cat(offs, dim=0),
                    flatten_ids=torch.cat([d["flatten_ids"] + c * N for c, d in enumerate(per_cam)], dim=0),
                    isect_ids=torch.cat([d["isect_ids"] | (c << (32 + tile_bits)) for c, d in enumerate(per_cam)], dim=0))

    meta._lazy = stacked_lists
    return torch.cat(renders, dim=0), torch.cat(alphas, dim=0), meta
