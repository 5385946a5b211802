"""ctypes/numpy front-end of the CPU oracle (``splat_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``splat_oracle.c``.  PARITY UNPINNED:
gsplat v1.4.0 (the un-vendored dependency that owns this arithmetic,
/root/reference/requirements.txt:12) is not installable here and the reference has no
golden vectors for this path; this is a restatement of its published algorithm as MTGS
calls it (mtgs/scene_model/mtgs_scene_graph.py:641-662).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsplat_oracle.so")
_lib = None

_f = C.POINTER(C.c_float)
_d = C.POINTER(C.c_double)
_i32 = C.POINTER(C.c_int32)
_i64 = C.POINTER(C.c_int64)
_u8 = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile the oracle with the Makefile next to it (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        r = subprocess.run(["make", "-C", _HERE, "-B"], capture_output=True, text=True)
        if r.returncode != 0:
            r = subprocess.run(["make", "-C", _HERE, "-B", "noomp"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_isect_count.restype = C.c_int64
        _lib.orc_tile_bits.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))


def _p(a: Optional[np.ndarray], t):
    if a is None:
        return None
    return a.ctypes.data_as(t)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# --------------------------------------------------------------------------- stages
def project_fwd(means, quats, scales, viewmat, K, W, H, eps2d=0.3, near_plane=0.01,
                far_plane=1e10, radius_clip=0.0, calc_comp=False) -> Dict[str, np.ndarray]:
    means, quats, scales = _f32(means), _f32(quats), _f32(scales)
    viewmat, K = _f32(viewmat).reshape(16), _f32(K).reshape(9)
    N = means.shape[0]
    radii = np.zeros(N, np.int32)
    means2d = np.zeros((N, 2), np.float32)
    depths = np.zeros(N, np.float32)
    conics = np.zeros((N, 3), np.float32)
    comps = np.zeros(N, np.float32) if calc_comp else None
    lib().orc_project_fwd(_p(means, _f), _p(quats, _f), _p(scales, _f), _p(viewmat, _f), _p(K, _f),
                          C.c_int(N), C.c_int(W), C.c_int(H), C.c_float(eps2d), C.c_float(near_plane),
                          C.c_float(far_plane), C.c_float(radius_clip), C.c_int(int(calc_comp)),
                          _p(radii, _i32), _p(means2d, _f), _p(depths, _f), _p(conics, _f), _p(comps, _f))
    return dict(radii=radii, means2d=means2d, depths=depths, conics=conics, compensations=comps)


def tile_bits(n_tiles: int) -> int:
    return int(lib().orc_tile_bits(C.c_int(n_tiles)))


def isect_tiles(means2d, radii, depths, tile_size, tile_w, tile_h, sort=True
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    means2d, depths = _f32(means2d), _f32(depths)
    radii = np.ascontiguousarray(radii, dtype=np.int32)
    N = radii.shape[0]
    tpg = np.zeros(N, np.int32)
    M = int(lib().orc_isect_count(_p(means2d, _f), _p(radii, _i32), C.c_int(N), C.c_int(tile_size),
                                  C.c_int(tile_w), C.c_int(tile_h), _p(tpg, _i32)))
    ids = np.zeros(M, np.int64)
    flat = np.zeros(M, np.int32)
    lib().orc_isect_emit(_p(means2d, _f), _p(radii, _i32), _p(depths, _f), C.c_int(N), C.c_int(tile_size),
                         C.c_int(tile_w), C.c_int(tile_h), _p(ids, _i64), _p(flat, _i32))
    if sort and M > 0:
        end_bit = 32 + tile_bits(tile_w * tile_h) + 1  # cam bits = floor(log2(1)) + 1 = 1
        lib().orc_sort_pairs(C.c_int64(M), C.c_int(end_bit), _p(ids, _i64), _p(flat, _i32))
    return tpg, ids, flat


def isect_offsets(isect_ids, tile_w, tile_h) -> np.ndarray:
    ids = np.ascontiguousarray(isect_ids, dtype=np.int64)
    off = np.zeros(tile_w * tile_h, np.int32)
    lib().orc_isect_offsets(_p(ids, _i64), C.c_int64(ids.shape[0]), C.c_int(tile_w * tile_h), _p(off, _i32))
    return off.reshape(tile_h, tile_w)


def blend_fwd(means2d, conics, colors, opacities, W, H, tile_size, offsets, flatten_ids, backgrounds=None):
    means2d, conics, colors, opacities = _f32(means2d), _f32(conics), _f32(colors), _f32(opacities)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    flat = np.ascontiguousarray(flatten_ids, dtype=np.int32)
    tile_h, tile_w = offsets.shape
    CD = colors.shape[1]
    assert CD <= 64
    bg = _f32(backgrounds) if backgrounds is not None else None
    out_c = np.zeros((H, W, CD), np.float32)
    out_a = np.zeros((H, W, 1), np.float32)
    last = np.zeros((H, W), np.int32)
    lib().orc_blend_fwd(_p(means2d, _f), _p(conics, _f), _p(colors, _f), _p(opacities, _f), _p(bg, _f),
                        C.c_int(W), C.c_int(H), C.c_int(tile_size), C.c_int(tile_w), C.c_int(tile_h),
                        _p(offsets, _i32), _p(flat, _i32), C.c_int64(flat.shape[0]), C.c_int(CD),
                        _p(out_c, _f), _p(out_a, _f), _p(last, _i32))
    return out_c, out_a, last


def blend_bwd(means2d, conics, colors, opacities, W, H, tile_size, offsets, flatten_ids, render_alphas,
              last_ids, v_render_colors, v_render_alphas, absgrad=False, backgrounds=None):
    means2d, conics, colors, opacities = _f32(means2d), _f32(conics), _f32(colors), _f32(opacities)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    flat = np.ascontiguousarray(flatten_ids, dtype=np.int32)
    tile_h, tile_w = offsets.shape
    N, CD = colors.shape
    ra, vrc, vra = _f32(render_alphas), _f32(v_render_colors), _f32(v_render_alphas)
    last = np.ascontiguousarray(last_ids, dtype=np.int32)
    bg = _f32(backgrounds) if backgrounds is not None else None
    v_m = np.zeros((N, 2), np.float64)
    v_abs = np.zeros((N, 2), np.float64) if absgrad else None
    v_con = np.zeros((N, 3), np.float64)
    v_col = np.zeros((N, CD), np.float64)
    v_op = np.zeros(N, np.float64)
    lib().orc_blend_bwd(_p(means2d, _f), _p(conics, _f), _p(colors, _f), _p(opacities, _f), _p(bg, _f),
                        C.c_int(N), C.c_int(W), C.c_int(H), C.c_int(tile_size), C.c_int(tile_w),
                        C.c_int(tile_h), _p(offsets, _i32), _p(flat, _i32), C.c_int64(flat.shape[0]),
                        C.c_int(CD), _p(ra, _f), _p(last, _i32), _p(vrc, _f), _p(vra, _f),
                        _p(v_m, _d), _p(v_abs, _d), _p(v_con, _d), _p(v_col, _d), _p(v_op, _d))
    return dict(v_means2d=v_m, v_means2d_abs=v_abs, v_conics=v_con, v_colors=v_col, v_opacities=v_op)


def project_bwd(means, quats, scales, viewmat, K, W, H, eps2d, radii, conics, comps, v_means2d, v_depths,
                v_conics, v_comps=None):
    means, quats, scales = _f32(means), _f32(quats), _f32(scales)
    viewmat, K = _f32(viewmat).reshape(16), _f32(K).reshape(9)
    N = means.shape[0]
    radii = np.ascontiguousarray(radii, dtype=np.int32)
    conics = _f32(conics)
    comps = _f32(comps) if comps is not None else None
    v_means2d, v_depths, v_conics = _f32(v_means2d), _f32(v_depths), _f32(v_conics)
    v_comps = _f32(v_comps) if v_comps is not None else None
    v_means = np.zeros((N, 3), np.float32)
    v_quats = np.zeros((N, 4), np.float32)
    v_scales = np.zeros((N, 3), np.float32)
    v_view = np.zeros(16, np.float64)
    lib().orc_project_bwd(_p(means, _f), _p(quats, _f), _p(scales, _f), _p(viewmat, _f), _p(K, _f),
                          C.c_int(N), C.c_int(W), C.c_int(H), C.c_float(eps2d), _p(radii, _i32),
                          _p(conics, _f), _p(comps, _f), _p(v_means2d, _f), _p(v_depths, _f),
                          _p(v_conics, _f), _p(v_comps, _f), _p(v_means, _f), _p(v_quats, _f),
                          _p(v_scales, _f), _p(v_view, _d))
    return dict(v_means=v_means, v_quats=v_quats, v_scales=v_scales, v_viewmat=v_view.reshape(4, 4))


def sh_fwd(degree, dirs, coeffs, masks=None):
    dirs, coeffs = _f32(dirs), _f32(coeffs)
    N, K = coeffs.shape[0], coeffs.shape[1]
    assert (degree + 1) ** 2 <= K
    m = np.ascontiguousarray(masks, dtype=np.uint8) if masks is not None else None
    out = np.zeros((N, 3), np.float32)
    lib().orc_sh_fwd(C.c_int(degree), _p(dirs, _f), _p(coeffs, _f), _p(m, _u8), C.c_int(N), C.c_int(K),
                     _p(out, _f))
    return out


def sh_bwd(degree, dirs, coeffs, v_colors, masks=None, need_dirs=True):
    dirs, coeffs, v_colors = _f32(dirs), _f32(coeffs), _f32(v_colors)
    N, K = coeffs.shape[0], coeffs.shape[1]
    m = np.ascontiguousarray(masks, dtype=np.uint8) if masks is not None else None
    v_coeffs = np.zeros((N, K, 3), np.float32)
    v_dirs = np.zeros((N, 3), np.float32) if need_dirs else None
    lib().orc_sh_bwd(C.c_int(degree), _p(dirs, _f), _p(coeffs, _f), _p(m, _u8), C.c_int(N), C.c_int(K),
                     _p(v_colors, _f), _p(v_coeffs, _f), _p(v_dirs, _f))
    return v_coeffs, v_dirs


# --------------------------------------------------------------------------- pipeline
_SUPPORTED_CH = (1, 2, 3, 4, 5, 8, 9, 16, 17, 32, 33, 64)


def padded_channels(ch: int) -> int:
    """upstream rasterize_to_pixels: pad to the next supported compile-time width."""
    if ch in _SUPPORTED_CH:
        return ch
    return 1 << (ch - 1).bit_length()


def rasterization(means, quats, scales, opacities, colors, viewmat, K, width, height, near_plane=0.01,
                  far_plane=1e10, radius_clip=0.0, eps2d=0.3, tile_size=16, render_mode="RGB",
                  rasterize_mode="classic"):
    """Full forward, one camera.  Mirrors upstream gsplat.rendering.rasterization for the
    argument subset MTGS passes (mtgs_scene_graph.py:641-661).  Returns (render, alpha, meta, ctx)."""
    assert render_mode in ("RGB", "RGB+ED", "RGB+D", "D", "ED")
    aa = rasterize_mode == "antialiased"
    pr = project_fwd(means, quats, scales, viewmat, K, width, height, eps2d, near_plane, far_plane,
                     radius_clip, calc_comp=aa)
    opac = _f32(opacities)
    if aa:
        opac = (opac * pr["compensations"]).astype(np.float32)
    cols = _f32(colors)
    if render_mode in ("RGB+D", "RGB+ED"):
        cols = np.concatenate([cols, pr["depths"][:, None]], axis=1)
    elif render_mode in ("D", "ED"):
        cols = pr["depths"][:, None].copy()
    ch = cols.shape[1]
    CD = padded_channels(ch)
    if CD != ch:
        cols = np.concatenate([cols, np.zeros((cols.shape[0], CD - ch), np.float32)], axis=1)
    tile_w = math.ceil(width / float(tile_size))
    tile_h = math.ceil(height / float(tile_size))
    tpg, ids, flat = isect_tiles(pr["means2d"], pr["radii"], pr["depths"], tile_size, tile_w, tile_h)
    offs = isect_offsets(ids, tile_w, tile_h)
    rc, ra, last = blend_fwd(pr["means2d"], pr["conics"], cols, opac, width, height, tile_size, offs, flat)
    rc_raw = rc
    rc = rc[..., :ch]
    if render_mode in ("ED", "RGB+ED"):
        rc = np.concatenate([rc[..., :-1], rc[..., -1:] / np.maximum(ra, np.float32(1e-10))], axis=-1)
    meta = dict(radii=pr["radii"], means2d=pr["means2d"], depths=pr["depths"], conics=pr["conics"],
                opacities=opac, tiles_per_gauss=tpg, isect_ids=ids, flatten_ids=flat, isect_offsets=offs,
                tile_width=tile_w, tile_height=tile_h, width=width, height=height, tile_size=tile_size,
                n_cameras=1)
    ctx = dict(pr=pr, cols=cols, ch=ch, CD=CD, opac=opac, last=last, ra=ra, rc_raw=rc_raw, aa=aa,
               render_mode=render_mode, args=(means, quats, scales, opacities, colors, viewmat, K, width,
                                              height, eps2d, tile_size))
    return rc.astype(np.float32), ra, meta, ctx


def rasterization_bwd(ctx, v_render, v_alpha, absgrad=False):
    """Analytic backward of :func:`rasterization` (upstream hand-written VJPs chained the way
    upstream autograd chains them)."""
    means, quats, scales, opacities, colors, viewmat, K, W, H, eps2d, tile_size = ctx["args"]
    pr, cols, ch, CD, opac, last, ra = (ctx[k] for k in ("pr", "cols", "ch", "CD", "opac", "last", "ra"))
    meta_offs, flat = ctx["meta_offs"], ctx["meta_flat"]
    v_render = _f32(v_render).reshape(H, W, ch)
    v_alpha = _f32(v_alpha).reshape(H, W, 1).copy()
    v_rc = np.zeros((H, W, CD), np.float32)
    v_rc[..., :ch] = v_render
    if ctx["render_mode"] in ("ED", "RGB+ED"):
        # out_d = acc_d / clamp(alpha, 1e-10)
        a = np.maximum(ra, np.float32(1e-10))
        acc_d = ctx["rc_raw"][..., ch - 1:ch]
        v_rc[..., ch - 1:ch] = v_render[..., -1:] / a
        v_alpha = v_alpha + np.where(ra > 1e-10, -v_render[..., -1:] * acc_d / (a * a), 0.0).astype(np.float32)
    g = blend_bwd(pr["means2d"], pr["conics"], cols, opac, W, H, tile_size, meta_offs, flat, ra, last, v_rc,
                  v_alpha, absgrad=absgrad)
    v_cols = g["v_colors"]
    v_depths = np.zeros(pr["depths"].shape[0], np.float64)
    if ctx["render_mode"] in ("RGB+D", "RGB+ED"):
        v_depths = v_cols[:, ch - 1].copy()
        v_colors_in = v_cols[:, :ch - 1]
    elif ctx["render_mode"] in ("D", "ED"):
        v_depths = v_cols[:, 0].copy()
        v_colors_in = None
    else:
        v_colors_in = v_cols[:, :ch]
    v_op = g["v_opacities"]
    v_comps = None
    if ctx["aa"]:
        v_comps = (v_op * _f32(opacities)).astype(np.float32)
        v_opacities_in = v_op * pr["compensations"]
    else:
        v_opacities_in = v_op
    pb = project_bwd(means, quats, scales, viewmat, K, W, H, eps2d, pr["radii"], pr["conics"],
                     pr["compensations"], g["v_means2d"], v_depths, g["v_conics"], v_comps)
    return dict(v_means=pb["v_means"], v_quats=pb["v_quats"], v_scales=pb["v_scales"],
                v_viewmat=pb["v_viewmat"], v_opacities=v_opacities_in, v_colors=v_colors_in,
                v_means2d=g["v_means2d"], v_means2d_abs=g["v_means2d_abs"], v_conics=g["v_conics"])


def rasterization_fwd_bwd(*args, v_render=None, v_alpha=None, absgrad=False, **kw):
    rc, ra, meta, ctx = rasterization(*args, **kw)
    ctx["meta_offs"], ctx["meta_flat"] = meta["isect_offsets"], meta["flatten_ids"]
    grads = rasterization_bwd(ctx, v_render, v_alpha, absgrad=absgrad) if v_render is not None else None
    return rc, ra, meta, grads


def quat_to_rotmat(quats) -> np.ndarray:
    q = _f32(quats)
    out = np.zeros((q.shape[0], 3, 3), np.float32)
    lib().orc_quat_to_rotmat(_p(q, _f), C.c_int(q.shape[0]), _p(out, _f))
    return out
