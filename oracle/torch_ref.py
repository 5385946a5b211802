"""Differentiable float64 torch-CPU restatement of the same path (TEST INFRASTRUCTURE ONLY).

Purpose: pin the *analytic* backward of ``splat_oracle.c`` (a restatement of upstream's
hand-written VJPs) against ``torch.autograd`` of the forward, in float64, on small scenes.
Discrete decisions that carry no gradient (cull mask, tile lists and their order) are taken
from the fp32 C oracle so both sides differentiate the same piecewise-smooth function
(SURVEY.md Appendix A.7).  PARITY UNPINNED with respect to gsplat itself (see cpu_ref.py).
"""
from __future__ import annotations

import torch

ALPHA_MAX = 0.999
ALPHA_MIN = 1.0 / 255.0
T_EPS = 1e-4


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


def project(means, quats, scales, viewmat, K, W, H, eps2d=0.3):
    """A.1 without the culls (caller masks by the oracle's radii)."""
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    pc = means @ R.T + t
    Rq = quat_to_rotmat(quats)
    M = Rq * scales[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    Sc = R @ Sigma @ R.T
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    x, y, z = pc.unbind(-1)
    tanx, tany = 0.5 * W / fx, 0.5 * H / fy
    lxp, lxn = (W - cx) / fx + 0.3 * tanx, cx / fx + 0.3 * tanx
    lyp, lyn = (H - cy) / fy + 0.3 * tany, cy / fy + 0.3 * tany
    tx = z * torch.minimum(lxp, torch.maximum(-lxn, x / z))
    ty = z * torch.minimum(lyp, torch.maximum(-lyn, y / z))
    zero = torch.zeros_like(z)
    J = torch.stack([fx / z, zero, -fx * tx / (z * z), zero, fy / z, -fy * ty / (z * z)], dim=-1).reshape(-1, 2, 3)
    cov2 = J @ Sc @ J.transpose(1, 2)
    means2d = torch.stack([fx * x / z + cx, fy * y / z + cy], dim=-1)
    det0 = cov2[:, 0, 0] * cov2[:, 1, 1] - cov2[:, 0, 1] * cov2[:, 1, 0]
    c00, c11, c01 = cov2[:, 0, 0] + eps2d, cov2[:, 1, 1] + eps2d, 0.5 * (cov2[:, 0, 1] + cov2[:, 1, 0])
    det1 = c00 * c11 - c01 * c01
    comp = torch.sqrt(torch.clamp(det0 / det1, min=0.0))
    conics = torch.stack([c11 / det1, -c01 / det1, c00 / det1], dim=-1)
    return means2d, z, conics, comp


def blend(means2d, conics, colors, opac, W, H, tile_size, offsets, flatten_ids):
    """A.3, vectorised per tile; returns (render [H,W,CD], alpha [H,W,1])."""
    tile_h, tile_w = offsets.shape
    M = flatten_ids.shape[0]
    CD = colors.shape[1]
    out_c = torch.zeros(H, W, CD, dtype=means2d.dtype)
    out_a = torch.zeros(H, W, 1, dtype=means2d.dtype)
    offs = offsets.reshape(-1).tolist() + [M]
    for tile in range(tile_h * tile_w):
        s, e = offs[tile], offs[tile + 1] if tile < tile_h * tile_w - 1 else M
        if e <= s:
            continue
        ti, tj = divmod(tile, tile_w)
        ys = torch.arange(ti * tile_size, min((ti + 1) * tile_size, H))
        xs = torch.arange(tj * tile_size, min((tj + 1) * tile_size, W))
        py, px = torch.meshgrid(ys.to(means2d.dtype) + 0.5, xs.to(means2d.dtype) + 0.5, indexing="ij")
        px, py = px.reshape(-1, 1), py.reshape(-1, 1)
        g = torch.as_tensor(flatten_ids[s:e], dtype=torch.long)
        dx = means2d[g, 0][None] - px
        dy = means2d[g, 1][None] - py
        a, b, c = conics[g, 0][None], conics[g, 1][None], conics[g, 2][None]
        sigma = 0.5 * (a * dx * dx + c * dy * dy) + b * dx * dy
        alpha = torch.clamp(opac[g][None] * torch.exp(-sigma), max=ALPHA_MAX)
        valid = (sigma >= 0) & (alpha >= ALPHA_MIN)
        aeff = torch.where(valid, alpha, torch.zeros_like(alpha))
        one_m = 1 - aeff
        T_incl = torch.cumprod(one_m, dim=1)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
        term = valid & (T_incl <= T_EPS)
        done = torch.cumsum(term.to(torch.int64), dim=1) > 0
        wgt = torch.where(done, torch.zeros_like(aeff), aeff * T_excl)
        col = wgt @ colors[g]
        T_fin = torch.prod(torch.where(done, torch.ones_like(one_m), one_m), dim=1, keepdim=True)
        yy = ys.repeat_interleave(xs.numel())
        xx = xs.repeat(ys.numel())
        out_c[yy, xx] = col
        out_a[yy, xx] = 1 - T_fin
    return out_c, out_a


def rasterization(means, quats, scales, opacities, colors, viewmat, K, W, H, meta, eps2d=0.3, tile_size=16,
                  render_mode="RGB", rasterize_mode="classic"):
    """Differentiable forward using the discrete structures in ``meta`` (from cpu_ref.rasterization)."""
    vis = torch.as_tensor(meta["radii"] > 0)
    m2d, z, con, comp = project(means, quats, scales, viewmat, K, W, H, eps2d)
    # culled rows must not leak gradient (nor NaNs): replace by detached constants
    safe = lambda t: torch.where(vis.reshape((-1,) + (1,) * (t.dim() - 1)), t, torch.zeros_like(t).detach())
    m2d, z, con, comp = safe(m2d), safe(z), safe(con), safe(comp)
    opac = opacities * comp if rasterize_mode == "antialiased" else opacities
    cols = colors
    if render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, z[:, None]], dim=1)
    elif render_mode in ("D", "ED"):
        cols = z[:, None]
    rc, ra = blend(m2d, con, cols, opac, W, H, tile_size, meta["isect_offsets"], meta["flatten_ids"])
    if render_mode in ("ED", "RGB+ED"):
        rc = torch.cat([rc[..., :-1], rc[..., -1:] / ra.clamp(min=1e-10)], dim=-1)
    return rc, ra, m2d


def sh_eval(degree, dirs, coeffs):
    """A.6 in float64 (autograd gives v_coeffs, v_dirs)."""
    d = dirs / dirs.norm(dim=-1, keepdim=True) if degree >= 1 else dirs
    x, y, z = d.unbind(-1)
    B = [torch.full_like(x, 0.2820947917738781)]
    if degree >= 1:
        B += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
    if degree >= 2:
        z2 = z * z
        t0b = -1.092548430592079 * z
        c1, s1 = x * x - y * y, 2 * x * y
        B += [0.5462742152960395 * s1, t0b * y, 0.9461746957575601 * z2 - 0.3153915652525201, t0b * x,
              0.5462742152960395 * c1]
    if degree >= 3:
        t0c = -2.285228997322329 * z2 + 0.4570457994644658
        t1b = 1.445305721320277 * z
        c2, s2 = x * c1 - y * s1, x * s1 + y * c1
        sh12 = z * (1.865881662950577 * z2 - 1.119528997770346)
        B += [-0.5900435899266435 * s2, t1b * s1, t0c * y, sh12, t0c * x, t1b * c1, -0.5900435899266435 * c2]
    if degree >= 4:
        t0d = z * (-4.683325804901025 * z2 + 2.007139630671868)
        t1c = 3.31161143515146 * z2 - 0.47308734787878
        t2b = -1.770130769779931 * z
        c3, s3 = x * c2 - y * s2, x * s2 + y * c2
        sh6 = 0.9461746957575601 * z2 - 0.3153915652525201
        B += [0.6258357354491763 * s3, t2b * s2, t1c * s1, t0d * y,
              1.984313483298443 * z * sh12 - 1.006230589874905 * sh6, t0d * x, t1c * c1, t2b * c2,
              0.6258357354491763 * c3]
    Bm = torch.stack(B, dim=-1)  # [N, nb]
    nb = Bm.shape[-1]
    return torch.einsum("nk,nkc->nc", Bm, coeffs[:, :nb, :])
