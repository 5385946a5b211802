"""GPU parity tests: hand-written sm_100a kernels (through the C ABI of libb200splat.so, as bound by
mtgs_b200._lib / mtgs_b200.rendering) vs. the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * bit-exact: radii, cull mask, tiles_per_gauss, sorted isect_ids, flatten_ids, isect_offsets, depths
  * fp32 tolerance, stated per assert: images rtol 1e-4 / atol 2e-5 (threshold flips bounded and counted),
    gradients rtol 5e-3 (+2e-4 of the tensor's max; upstream fast-math level)
PARITY UNPINNED with respect to gsplat itself (not installable; see oracle/cpu_ref.py).
"""
import os

import numpy as np
import pytest
import torch

from mtgs_b200 import scenes
from tests.util import assert_grad_close, assert_image_close, masked_psnr, psnr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _to_dev(s, dev, grad=False):
    t = {}
    for k in ("means", "quats", "scales", "opacities", "colors", "viewmat", "K"):
        t[k] = torch.tensor(s[k], dtype=torch.float32, device=dev)
    if grad:
        for k in ("means", "quats", "scales", "opacities", "colors", "viewmat"):
            t[k].requires_grad_(True)
    return t


def _gpu_raster(t, s, **kw):
    from mtgs_b200.rendering import rasterization
    return rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], t["viewmat"][None],
                         t["K"][None], s["width"], s["height"], packed=False, **kw)


def _cpu_raster(oracle, s, **kw):
    return oracle.rasterization(s["means"], s["quats"], s["scales"], s["opacities"], s["colors"], s["viewmat"],
                                s["K"], s["width"], s["height"], **kw)


def _check_images(r, a, rc, ra, rmode):
    """alpha and colour channels: tolerance + bounded threshold flips.  The expected-depth channel (ED) is
    acc_depth / alpha, which is ill-conditioned where alpha ~ 1/255 (one flipped Gaussian changes it by O(depth
    range)), so its flip bound is applied to the accumulated depth (channel * alpha) instead."""
    assert_image_close(a, ra, "alpha")
    if rmode in ("RGB+ED", "ED"):
        assert_image_close(r[..., :-1], rc[..., :-1], "render") if r.shape[-1] > 1 else None
        assert_image_close(r[..., -1:], rc[..., -1:], "expected depth", flip_bound=float("inf"))
        assert_image_close(r[..., -1:] * np.maximum(a, 1e-10), rc[..., -1:] * np.maximum(ra, 1e-10), "acc depth",
                           rtol=2e-4, atol=1e-4)
    else:
        assert_image_close(r, rc, "render")


SCENES = {
    "tiny": lambda: scenes.tiny(n=300, seed=3, width=64, height=48),
    "tiny_ragged": lambda: scenes.tiny(n=257, seed=9, width=77, height=53),  # W,H not multiples of 16
    "config1": lambda: scenes.config1(),
    "street20k": lambda: scenes.street(n=20_000, seed=1, width=1920, height=1080),
    # BASELINE config 5 resolutions: 4K exercises two column bands / seven row bands of the tile-list kernels
    "street20k_4k": lambda: scenes.street(n=20_000, seed=2, width=3840, height=2160),
    "street20k_540p": lambda: scenes.street(n=20_000, seed=3, width=960, height=540),
}
# more than 256 tile columns / rows: the interval multisplit of tilelists.cu needs a second pass over each chunk
BIN_ONLY_SCENES = {
    "wide_4400x304": lambda: scenes.street(n=6_000, seed=4, width=4400, height=304),
    "tall_304x4400": lambda: scenes.street(n=6_000, seed=5, width=304, height=4400),
}
ALL_SCENES = {**SCENES, **BIN_ONLY_SCENES}


@pytest.mark.parametrize("scene", list(ALL_SCENES))
@pytest.mark.parametrize("mode,rmode", [("classic", "RGB"), ("antialiased", "RGB+ED")])
def test_projection_and_binning_bit_exact(oracle, cuda_device, scene, mode, rmode):
    s = ALL_SCENES[scene]()
    t = _to_dev(s, cuda_device)
    with torch.no_grad():
        _, _, meta = _gpu_raster(t, s, render_mode=rmode, rasterize_mode=mode)
    _, _, ref, _ = _cpu_raster(oracle, s, render_mode=rmode, rasterize_mode=mode)
    radii = meta["radii"][0].cpu().numpy()
    np.testing.assert_array_equal(radii, ref["radii"], err_msg="radii")
    vis = ref["radii"] > 0
    assert vis.sum() > 0
    np.testing.assert_array_equal(meta["tiles_per_gauss"][0].cpu().numpy(), ref["tiles_per_gauss"])
    # canonical op order => identical fp32 bits for the projected quantities of visible Gaussians
    np.testing.assert_array_equal(meta["depths"][0].cpu().numpy()[vis], ref["depths"][vis], err_msg="depths")
    np.testing.assert_array_equal(meta["means2d"][0].detach().cpu().numpy()[vis], ref["means2d"][vis])
    np.testing.assert_array_equal(meta["conics"][0].cpu().numpy()[vis], ref["conics"][vis])
    np.testing.assert_array_equal(meta["opacities"][0].cpu().numpy()[vis], ref["opacities"][vis])
    np.testing.assert_array_equal(meta["flatten_ids"].cpu().numpy(), ref["flatten_ids"], err_msg="flatten_ids")
    np.testing.assert_array_equal(meta["isect_offsets"][0].cpu().numpy(), ref["isect_offsets"])
    np.testing.assert_array_equal(meta["isect_ids"].cpu().numpy(), ref["isect_ids"], err_msg="isect_ids")
    assert meta["tile_width"] == ref["tile_width"] and meta["tile_height"] == ref["tile_height"]


@pytest.mark.parametrize("scene", list(SCENES))
@pytest.mark.parametrize("mode,rmode,d_in", [("classic", "RGB", 3), ("antialiased", "RGB+ED", 3),
                                              ("antialiased", "RGB+ED", 6), ("classic", "RGB", 6),
                                              ("classic", "ED", 3), ("classic", "RGB+D", 3)])
def test_forward_image_parity(oracle, cuda_device, scene, mode, rmode, d_in):
    s = SCENES[scene]()
    if d_in != 3:
        rng = np.random.default_rng(42)
        s["colors"] = rng.uniform(-1, 1, (s["means"].shape[0], d_in)).astype(np.float32)
    t = _to_dev(s, cuda_device)
    with torch.no_grad():
        r, a, _ = _gpu_raster(t, s, render_mode=rmode, rasterize_mode=mode)
    rc, ra, _, _ = _cpu_raster(oracle, s, render_mode=rmode, rasterize_mode=mode)
    assert r.shape == (1,) + rc.shape and a.shape == (1,) + ra.shape
    _check_images(r[0].cpu().numpy(), a[0].cpu().numpy(), rc, ra, rmode)
    if rmode == "RGB":
        assert psnr(r[0, ..., :3].cpu().numpy(), rc[..., :3]) > 60.0  # >> the 0.05 dB PSNR-delta bar


@pytest.mark.parametrize("scene", ["tiny", "tiny_ragged", "config1", "street20k", "street20k_4k"])
@pytest.mark.parametrize("mode,rmode,d_in", [("classic", "RGB", 3), ("antialiased", "RGB+ED", 3),
                                              ("antialiased", "RGB+ED", 6)])
def test_backward_parity(oracle, cuda_device, scene, mode, rmode, d_in):
    s = SCENES[scene]()
    rng = np.random.default_rng(7)
    if d_in != 3:
        s["colors"] = rng.uniform(-1, 1, (s["means"].shape[0], d_in)).astype(np.float32)
    t = _to_dev(s, cuda_device, grad=True)
    r, a, meta = _gpu_raster(t, s, render_mode=rmode, rasterize_mode=mode, absgrad=True)
    meta["means2d"].retain_grad()  # what MTGS does (mtgs_scene_graph.py:666-667)
    v_r = rng.standard_normal(tuple(r.shape[1:])).astype(np.float32)
    v_a = rng.standard_normal(tuple(a.shape[1:])).astype(np.float32)
    loss = (r[0] * torch.tensor(v_r, device=cuda_device)).sum() + (a[0] * torch.tensor(v_a, device=cuda_device)).sum()
    loss.backward()
    rc, ra, ref, ctx = _cpu_raster(oracle, s, render_mode=rmode, rasterize_mode=mode)
    ctx["meta_offs"], ctx["meta_flat"] = ref["isect_offsets"], ref["flatten_ids"]
    g = oracle.rasterization_bwd(ctx, v_r, v_a, absgrad=True)
    assert_grad_close(meta["means2d"].grad[0].cpu().numpy(), g["v_means2d"], "means2d.grad")
    assert hasattr(meta["means2d"], "absgrad"), "absgrad attribute not set on info['means2d']"
    assert_grad_close(meta["means2d"].absgrad[0].cpu().numpy(), g["v_means2d_abs"], "means2d.absgrad")
    assert_grad_close(t["colors"].grad.cpu().numpy(), g["v_colors"], "v_colors")
    assert_grad_close(t["opacities"].grad.cpu().numpy(), g["v_opacities"], "v_opacities")
    assert_grad_close(t["means"].grad.cpu().numpy(), g["v_means"], "v_means")
    assert_grad_close(t["quats"].grad.cpu().numpy(), g["v_quats"], "v_quats")
    assert_grad_close(t["scales"].grad.cpu().numpy(), g["v_scales"], "v_scales")
    assert_grad_close(t["viewmat"].grad.cpu().numpy(), g["v_viewmat"], "v_viewmat", rtol=5e-3, scale_atol=1e-3)
    culled = ref["radii"] <= 0
    assert torch.all(t["means"].grad[torch.tensor(culled, device=cuda_device)] == 0)
    assert torch.all(meta["means2d"].grad[0][torch.tensor(culled, device=cuda_device)] == 0)


@pytest.mark.parametrize("scene", ["tiny_ragged", "config1", "street20k", "street20k_540p"])
def test_tight_rectangles_hold_every_contributing_pair(cuda_device, scene):
    """The blend's lists are built from TIGHT rectangles (upstream's 3-sigma rectangle intersected with the extents of
    the footprint alpha >= 1/255).  Every (Gaussian, tile) pair of upstream's lists that reaches alpha >= 1/255 at some
    pixel centre of its tile (brute force in float64 over all 256 pixels) must lie inside the Gaussian's tight
    rectangle; and the tight rectangles must actually be tighter."""
    s = SCENES[scene]()
    t = _to_dev(s, cuda_device)
    W, H, N = s["width"], s["height"], s["means"].shape[0]
    with torch.no_grad():
        _, _, meta = _gpu_raster(t, s, render_mode="RGB+ED", rasterize_mode="antialiased")
    dev = cuda_device
    tw = meta["tile_width"]
    flat, offs = meta["flatten_ids"].long(), meta["isect_offsets"].reshape(-1).long()
    M = flat.numel()
    tile_up = torch.searchsorted(offs, torch.arange(M, device=dev), right=True) - 1
    tr = meta["_tight_rects"].long()
    x0, x1, y0, y1 = tr[:, 0] & 0xffff, (tr[:, 0] >> 16) & 0xffff, tr[:, 1] & 0xffff, (tr[:, 1] >> 16) & 0xffff
    tx, ty = tile_up % tw, tile_up // tw
    inside = (tx >= x0[flat]) & (tx < x1[flat]) & (ty >= y0[flat]) & (ty < y1[flat])
    # brute force: best alpha of every upstream pair over the pixel centres of its tile
    m2 = meta["means2d"][0].double()
    con = meta["conics"][0].double()
    op = meta["opacities"][0].double()
    best = torch.empty(M, dtype=torch.float64, device=dev)
    jj, ii = torch.meshgrid(torch.arange(16, device=dev), torch.arange(16, device=dev), indexing="ij")
    for lo in range(0, M, 65536):
        hi = min(M, lo + 65536)
        g, tl = flat[lo:hi], tile_up[lo:hi]
        px = ((tl % tw) * 16)[:, None] + ii.reshape(1, -1)
        py = ((tl // tw) * 16)[:, None] + jj.reshape(1, -1)
        valid = (px < W) & (py < H)
        dx = m2[g, 0:1] - (px + 0.5)
        dy = m2[g, 1:2] - (py + 0.5)
        sig = 0.5 * (con[g, 0:1] * dx * dx + con[g, 2:3] * dy * dy) + con[g, 1:2] * dx * dy
        al = torch.clamp(op[g][:, None] * torch.exp(-sig), max=0.999)
        al = torch.where(valid & (sig >= 0), al, torch.zeros_like(al))
        best[lo:hi] = al.max(dim=1).values
    must = best >= (1.0 / 255.0) * (1 + 1e-3)
    assert int(must.sum()) > 0
    missing = must & ~inside
    assert int(missing.sum()) == 0, f"{int(missing.sum())} contributing pairs fall outside the tight rectangles (best alpha up to {float(best[missing].max()):.4f})"
    n_in = int(inside.sum())
    assert n_in < M, "tight rectangles are not tighter than upstream's"
    dead_kept = inside & (best < (1.0 / 255.0) * (1 - 2e-2))
    assert float(dead_kept.sum()) <= 0.6 * n_in + 8, (int(dead_kept.sum()), n_in)


def test_capacity_mode_equals_exact_sizes_and_survives_overflow(cuda_device):
    """The forward is enqueued with list capacities from earlier frames (no wait for this frame's sizes).  It must give
    the same bits as the build with exact sizes, and a frame that outgrows the capacities (here: forced) must be rebuilt,
    not corrupted."""
    from mtgs_b200 import rendering
    s = scenes.street(n=20_000, seed=7, width=960, height=540)
    t = _to_dev(s, cuda_device)
    kw = dict(render_mode="RGB+ED", rasterize_mode="antialiased")
    rendering._CAPACITY.clear()
    old = rendering.SYNC_SIZES
    try:
        rendering.SYNC_SIZES = True
        with torch.no_grad():
            r0, a0, m0 = _gpu_raster(t, s, **kw)
        rendering.SYNC_SIZES = False
        key = (cuda_device.index or 0, m0["tile_width"], m0["tile_height"])
        assert key in rendering._CAPACITY
        with torch.no_grad():
            r1, a1, m1 = _gpu_raster(t, s, **kw)          # capacity mode, capacities sufficient
        assert torch.equal(r0, r1) and torch.equal(a0, a1)
        for lvl in range(4):                              # every level's capacity too small in turn
            caps = list(rendering._CAPACITY[key])
            small = list(caps)
            small[lvl] = 64
            rendering._CAPACITY[key] = small
            t2 = _to_dev(s, cuda_device, grad=True)
            r2, a2, m2 = _gpu_raster(t2, s, absgrad=True, **kw)
            assert torch.equal(r0, r2.detach()) and torch.equal(a0, a2.detach()), f"level {lvl}"
            (r2.sum() + a2.sum()).backward()
            assert torch.isfinite(t2["means"].grad).all() and float(t2["means"].grad.abs().sum()) > 0
            assert rendering._CAPACITY[key][lvl] >= caps[lvl] - 1   # grown back
    finally:
        rendering.SYNC_SIZES = old


def test_graphed_step_replays_the_eager_step(cuda_device):
    """A whole step (forward + loss + backward) captured as one CUDA graph (mtgs_b200.graph.GraphedStep) gives the
    eager step's image bit for bit and its gradients to accumulation-order noise; a replay that outgrows the captured
    capacities is reported."""
    from mtgs_b200 import rendering
    from mtgs_b200.graph import GraphedStep
    s = scenes.street(n=30_000, seed=9, width=640, height=360)
    t = _to_dev(s, cuda_device, grad=True)
    kw = dict(render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)
    g = torch.Generator(device=cuda_device).manual_seed(2)
    w = torch.randn(1, 360, 640, 4, device=cuda_device, generator=g)
    names = ("means", "quats", "scales", "opacities", "colors")

    def step():
        r, a, _ = _gpu_raster(t, s, **kw)
        for k in names + ("viewmat",):
            t[k].grad = None
        ((r * w).sum() + a.sum()).backward()
        return r, a

    # nothing of the eager step may stay alive into the capture: a tensor with a grad_fn pins the autograd graph and
    # its AccumulateGrad nodes, which belong to the default stream (GraphedStep docstring)
    r0, a0 = (x.detach().clone() for x in step())
    g0 = {k: t[k].grad.clone() for k in names}
    gs = GraphedStep(step, warmup=2)
    for _ in range(3):
        r1, a1 = gs.replay()
    gs.check()
    assert torch.equal(r0.detach(), r1.detach()) and torch.equal(a0.detach(), a1.detach())
    for k in names:
        tol = 1e-4 * float(g0[k].abs().max()) + 1e-9
        assert float((t[k].grad - g0[k]).abs().max()) <= tol, k
    # a scene change that needs more room than was captured: the replay is flagged, nothing crashes
    with torch.no_grad():
        t["scales"].mul_(3.0)
    gs.replay()
    assert not gs.ok()
    gs.recapture()
    gs.replay()
    gs.check()


def test_golden_fixture_through_c_abi(cuda_device):
    """Committed fixture (tests/golden/oracle_tiny_golden.npz): no oracle code runs in this test."""
    g = np.load(os.path.join(GOLD, "oracle_tiny_golden.npz"))
    s = {k: g["in_" + k] for k in ("means", "quats", "scales", "opacities", "colors", "viewmat", "K")}
    s["width"], s["height"] = int(g["width"]), int(g["height"])
    t = _to_dev(s, cuda_device, grad=True)
    r, a, meta = _gpu_raster(t, s, render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)
    meta["means2d"].retain_grad()
    for k in ("radii", "tiles_per_gauss", "isect_offsets"):
        np.testing.assert_array_equal(meta[k][0].cpu().numpy(), g[k], err_msg=k)
    np.testing.assert_array_equal(meta["flatten_ids"].cpu().numpy(), g["flatten_ids"])
    np.testing.assert_array_equal(meta["isect_ids"].cpu().numpy(), g["isect_ids"])
    _check_images(r[0].detach().cpu().numpy(), a[0].detach().cpu().numpy(), g["render"], g["alpha"], "RGB+ED")
    loss = (r[0] * torch.tensor(g["v_render"], device=cuda_device)).sum() + \
           (a[0] * torch.tensor(g["v_alpha"], device=cuda_device)).sum()
    loss.backward()
    for k, tk in (("v_means", "means"), ("v_quats", "quats"), ("v_scales", "scales"), ("v_opacities", "opacities"),
                  ("v_colors", "colors")):
        assert_grad_close(t[tk].grad.cpu().numpy(), g[k], k)
    assert_grad_close(t["viewmat"].grad.cpu().numpy(), g["v_viewmat"], "v_viewmat", scale_atol=1e-3)
    assert_grad_close(meta["means2d"].absgrad[0].cpu().numpy(), g["v_means2d_abs"], "absgrad")


@pytest.mark.parametrize("degree,K", [(0, 16), (1, 16), (2, 16), (3, 16), (4, 25), (3, 25), (1, 4)])
def test_spherical_harmonics_parity(oracle, cuda_device, degree, K):
    from mtgs_b200.cuda._wrapper import spherical_harmonics
    rng = np.random.default_rng(100 + degree)
    n = 5000 + degree  # not a multiple of the CTA size
    dirs = (rng.standard_normal((n, 3)) * rng.uniform(0.1, 30, (n, 1))).astype(np.float32)
    coeffs = rng.standard_normal((n, K, 3)).astype(np.float32)
    v = rng.standard_normal((n, 3)).astype(np.float32)
    td = torch.tensor(dirs, device=cuda_device, requires_grad=True)
    tc = torch.tensor(coeffs, device=cuda_device, requires_grad=True)
    out = spherical_harmonics(degree, td, tc)
    ref = oracle.sh_fwd(degree, dirs, coeffs)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol=2e-5, atol=2e-5)
    (out * torch.tensor(v, device=cuda_device)).sum().backward()
    v_c, v_d = oracle.sh_bwd(degree, dirs, coeffs, v)
    np.testing.assert_allclose(tc.grad.cpu().numpy(), v_c, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(td.grad.cpu().numpy(), v_d, rtol=2e-3, atol=2e-4 * max(1e-6, np.abs(v_d).max()))
    # masks: masked rows produce zero colour and zero gradients
    mask = torch.tensor(rng.random(n) < 0.5, device=cuda_device)
    tc2 = tc.detach().clone().requires_grad_(True)
    out2 = spherical_harmonics(degree, td.detach(), tc2, masks=mask)
    assert torch.all(out2[~mask] == 0)
    np.testing.assert_allclose(out2[mask].detach().cpu().numpy(), ref[mask.cpu().numpy()], rtol=2e-5, atol=2e-5)
    out2.sum().backward()
    assert torch.all(tc2.grad[~mask] == 0)


def test_mtgs_call_pattern(cuda_device):
    """The exact kwargs of mtgs_scene_graph.py:641-661 through the gsplat module alias, plus the consumers
    at :663-690 and :1171-1178 (absgrad statistic, radii mask)."""
    import mtgs_b200
    mtgs_b200.install_as_gsplat()
    from gsplat.rendering import rasterization
    from gsplat.cuda._wrapper import spherical_harmonics
    s = scenes.street(n=30_000, seed=5, width=960, height=540, d_in=3)
    dev = cuda_device
    n = s["means"].shape[0]
    means = torch.tensor(s["means"], device=dev, requires_grad=True)
    raw_scales = torch.tensor(np.log(s["scales"]), device=dev, requires_grad=True)
    raw_quats = torch.tensor(s["quats"] * 1.7, device=dev, requires_grad=True)
    raw_opac = torch.logit(torch.tensor(s["opacities"], device=dev)).requires_grad_(True)
    feats = (torch.randn(n, 16, 3, device=dev) * 0.2).requires_grad_(True)
    c2w = torch.inverse(torch.tensor(s["viewmat"], device=dev))
    viewdirs = means.detach() - c2w[:3, 3]
    viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
    rgbs = torch.clamp(spherical_harmonics(3, viewdirs, feats) + 0.5, 0.0, 1.0)
    normals = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
    colors = torch.cat([rgbs, normals], dim=-1)
    viewmat = torch.tensor(s["viewmat"], device=dev).unsqueeze(0).requires_grad_(True)
    render, alpha, info = rasterization(
        means=means, quats=raw_quats / raw_quats.norm(dim=-1, keepdim=True), scales=torch.exp(raw_scales),
        opacities=torch.sigmoid(raw_opac), colors=colors, viewmats=viewmat,
        Ks=torch.tensor(s["K"], device=dev).unsqueeze(0), width=s["width"], height=s["height"], tile_size=16,
        packed=False, near_plane=0.01, far_plane=1e10, render_mode="RGB+ED", sparse_grad=False, absgrad=True,
        rasterize_mode="antialiased")
    assert render.shape == (1, 540, 960, 7) and alpha.shape == (1, 540, 960, 1)
    if info["radii"].ndim == 3:
        info["radii"] = (info["radii"][..., 0] * info["radii"][..., 1]).sqrt().int()
    assert info["means2d"].requires_grad
    info["means2d"].retain_grad()
    rgb = torch.clamp(render[..., :3] + (1 - alpha) * torch.ones(3, device=dev), 0.0, 1.0)
    depth = torch.where(alpha > 0, render[..., -1:], render[..., -1:].detach().max())
    nrm = render[..., 3:6]
    loss = rgb.mean() + 0.01 * depth.mean() + nrm.abs().mean()
    loss.backward()
    assert info["means2d"].grad is not None
    grads = info["means2d"].absgrad[0].detach()
    stat = (grads * grads.new_tensor([960, 540]).unsqueeze(0) * 0.5).norm(dim=-1)
    vis = info["radii"][0] > 0
    assert torch.isfinite(stat).all() and stat[vis].sum() > 0 and torch.all(stat[~vis] == 0)
    for p in (means, raw_scales, raw_quats, raw_opac, feats, viewmat):
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert feats.grad.abs().sum() > 0 and viewmat.grad.abs().sum() > 0


def test_error_behaviour(cuda_device):
    from mtgs_b200.rendering import rasterization
    s = scenes.tiny()
    t = _to_dev(s, cuda_device)
    args = (t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], t["viewmat"][None], t["K"][None], 64, 48)
    with pytest.raises(NotImplementedError):
        rasterization(*args)  # upstream default packed=True is not built
    with pytest.raises(NotImplementedError):
        rasterization(*args, packed=False, tile_size=8)
    with pytest.raises(AssertionError):
        rasterization(*args, packed=False, render_mode="XYZ")
    with pytest.raises(RuntimeError):
        rasterization(*(x.cpu() if torch.is_tensor(x) else x for x in args), packed=False)
    # N == 0 and M == 0 are tolerated (MTGS guards N == 0 itself, mtgs_scene_graph.py:595-598)
    e = torch.zeros
    r, a, m = rasterization(e(0, 3, device=cuda_device), e(0, 4, device=cuda_device), e(0, 3, device=cuda_device),
                            e(0, device=cuda_device), e(0, 3, device=cuda_device), t["viewmat"][None], t["K"][None],
                            64, 48, packed=False)
    assert r.shape == (1, 48, 64, 3) and float(a.abs().max()) == 0.0
    far = t["means"].clone()
    far[:, 2] = -5.0  # everything behind the camera: M == 0
    r, a, m = rasterization(far, t["quats"], t["scales"], t["opacities"], t["colors"], t["viewmat"][None],
                            t["K"][None], 64, 48, packed=False)
    assert m["flatten_ids"].numel() == 0 and float(a.abs().max()) == 0.0 and int(m["radii"].max()) == 0


@pytest.mark.parametrize("n,d_in", [(500_000, 3), (500_000, 6), (2_000_000, 3), (2_000_000, 6)])
def test_full_size_oracle_parity(oracle, cuda_device, n, d_in):
    """BASELINE config 2 itself (500 k @1920x1080) and the bench workload (2 M), CDIM 4 (RGB+ED) and CDIM 8
    (RGB+normals+ED, mtgs/config/MTGS.py:101-111), against the oracle: bit-exact bins, forward images, every
    gradient.  These are the deep-tile sizes (2 M: ~2350 entries per tile) where the transmittance reconstruction
    by division and the per-tile cull are exercised hardest."""
    s = scenes.street(n=n, seed=1, width=1920, height=1080, d_in=d_in)
    rng = np.random.default_rng(11)
    t = _to_dev(s, cuda_device, grad=True)
    kw = dict(render_mode="RGB+ED", rasterize_mode="antialiased")
    r, a, meta = _gpu_raster(t, s, absgrad=True, **kw)
    meta["means2d"].retain_grad()
    v_r = rng.standard_normal(tuple(r.shape[1:])).astype(np.float32)
    v_a = rng.standard_normal(tuple(a.shape[1:])).astype(np.float32)
    loss = (r[0] * torch.tensor(v_r, device=cuda_device)).sum() + (a[0] * torch.tensor(v_a, device=cuda_device)).sum()
    loss.backward()
    rc, ra, ref, ctx = _cpu_raster(oracle, s, **kw)
    # bins: bit-exact
    np.testing.assert_array_equal(meta["radii"][0].cpu().numpy(), ref["radii"], err_msg="radii")
    np.testing.assert_array_equal(meta["tiles_per_gauss"][0].cpu().numpy(), ref["tiles_per_gauss"])
    np.testing.assert_array_equal(meta["flatten_ids"].cpu().numpy(), ref["flatten_ids"], err_msg="flatten_ids")
    np.testing.assert_array_equal(meta["isect_offsets"][0].cpu().numpy(), ref["isect_offsets"])
    np.testing.assert_array_equal(meta["isect_ids"].cpu().numpy(), ref["isect_ids"], err_msg="isect_ids")
    # images
    rg, ag = r[0].detach().cpu().numpy(), a[0].detach().cpu().numpy()
    _check_images(rg, ag, rc, ra, "RGB+ED")
    assert masked_psnr(rg[..., :3], rc[..., :3]) > 60.0
    # gradients
    ctx["meta_offs"], ctx["meta_flat"] = ref["isect_offsets"], ref["flatten_ids"]
    g = oracle.rasterization_bwd(ctx, v_r, v_a, absgrad=True)
    assert_grad_close(meta["means2d"].grad[0].cpu().numpy(), g["v_means2d"], "means2d.grad")
    assert_grad_close(meta["means2d"].absgrad[0].cpu().numpy(), g["v_means2d_abs"], "means2d.absgrad")
    assert_grad_close(t["colors"].grad.cpu().numpy(), g["v_colors"], "v_colors")
    assert_grad_close(t["opacities"].grad.cpu().numpy(), g["v_opacities"], "v_opacities")
    assert_grad_close(t["means"].grad.cpu().numpy(), g["v_means"], "v_means")
    assert_grad_close(t["quats"].grad.cpu().numpy(), g["v_quats"], "v_quats")
    assert_grad_close(t["scales"].grad.cpu().numpy(), g["v_scales"], "v_scales")
    assert_grad_close(t["viewmat"].grad.cpu().numpy(), g["v_viewmat"], "v_viewmat", rtol=5e-3, scale_atol=1e-3)


def test_full_size_properties(cuda_device):
    """BASELINE workload size (2M Gaussians, 1920x1080): size-independent properties (no oracle)."""
    from mtgs_b200.rendering import rasterization
    s = scenes.street(n=2_000_000, seed=1)
    dev = cuda_device
    t = _to_dev(s, dev, grad=True)
    r, a, meta = _gpu_raster(t, s, render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)
    M = meta["flatten_ids"].numel()
    assert M == int(meta["tiles_per_gauss"].sum())
    ids = meta["isect_ids"]
    assert bool((ids[1:] >= ids[:-1]).all()), "intersections not sorted by (tile, depth)"
    tie = ids[1:] == ids[:-1]
    assert bool((meta["flatten_ids"][1:][tie] > meta["flatten_ids"][:-1][tie]).all()), "sort not stable"
    offs = meta["isect_offsets"].reshape(-1)
    assert bool((offs[1:] >= offs[:-1]).all()) and int(offs[0]) == 0 and int(offs[-1]) <= M
    tiles = (ids >> 32).to(torch.int64)
    expect = torch.searchsorted(tiles, torch.arange(offs.numel(), device=dev))
    assert torch.equal(expect.to(torch.int32), offs)
    cnt = torch.bincount(meta["flatten_ids"].long(), minlength=2_000_000)
    assert torch.equal(cnt.to(torch.int32), meta["tiles_per_gauss"][0])
    assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0 and bool(torch.isfinite(r).all())
    # determinism of the forward (no atomics on the forward path)
    with torch.no_grad():
        r2, a2, _ = _gpu_raster(t, s, render_mode="RGB+ED", rasterize_mode="antialiased")
    assert torch.equal(r2, r.detach()) and torch.equal(a2, a.detach())
    # adjoint identity of the colour path: the blend is linear in the colours, so
    #   <render_rgb, w> == <colors, d<render_rgb, w>/d colors>   at any size
    w = torch.randn(1, s["height"], s["width"], 3, device=dev)
    lhs = (r[..., :3] * w).sum()
    (gc,) = torch.autograd.grad(lhs, t["colors"], retain_graph=True)
    rhs = (gc * t["colors"]).sum()
    assert abs(float(lhs) - float(rhs)) <= 2e-3 * max(1.0, abs(float(lhs))), (float(lhs), float(rhs))
    # linearity of the VJP in the cotangent and zero cotangent -> zero gradient
    loss = (r[..., :3] * w).sum() + a.sum()
    g1 = torch.autograd.grad(loss, [t["means"], t["opacities"]], retain_graph=True)
    g2 = torch.autograd.grad(2.0 * loss, [t["means"], t["opacities"]], retain_graph=True)
    for x, y in zip(g1, g2):
        assert torch.isfinite(x).all()
        assert float((2 * x - y).abs().max()) <= 1e-3 * float(y.abs().max()) + 1e-6
    g0 = torch.autograd.grad(0.0 * loss, [t["means"]])[0]
    assert float(g0.abs().max()) == 0.0
