"""Stage-by-stage GPU vs oracle diagnostics (development aid; prints, never asserts).
    python tools/gpu_diag.py [scene] [n]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtgs_b200 import scenes, rendering  # noqa: E402
from mtgs_b200.rendering import rasterization  # noqa: E402
from oracle import cpu_ref  # noqa: E402


def cmp_exact(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}")
        return
    bad = np.flatnonzero(a.reshape(-1) != b.reshape(-1))
    print(f"  {name}: {'EXACT' if bad.size == 0 else f'{bad.size}/{a.size} differ, first idx {bad[:5]} got {a.reshape(-1)[bad[:5]]} want {b.reshape(-1)[bad[:5]]}'}")


def cmp_close(name, a, b, rtol=1e-4, atol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}")
        return
    d = np.abs(a - b)
    bad = d > atol + rtol * np.abs(b)
    sc = np.abs(b).max() if b.size else 0
    print(f"  {name}: max|ref| {sc:.3e} maxabs {d.max() if d.size else 0:.3e} rel-to-max {(d.max() / sc if sc else 0):.2e} "
          f"frac_bad {bad.mean() if bad.size else 0:.2e} finite {np.isfinite(a).all()}")


def run(s, rmode, mode, tag):
    dev = torch.device("cuda:0")
    print(f"== {tag} N={s['means'].shape[0]} {s['width']}x{s['height']} {rmode} {mode}")
    t = {k: torch.tensor(s[k], device=dev, requires_grad=k in ("means", "quats", "scales", "opacities", "colors", "viewmat"))
         for k in ("means", "quats", "scales", "opacities", "colors", "viewmat", "K")}
    rendering.PROFILE = {}
    r, a, meta = rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], t["viewmat"][None],
                               t["K"][None], s["width"], s["height"], packed=False, render_mode=rmode,
                               rasterize_mode=mode, absgrad=True)
    meta["means2d"].retain_grad()
    rng = np.random.default_rng(7)
    v_r = rng.standard_normal(tuple(r.shape[1:])).astype(np.float32)
    v_a = rng.standard_normal(tuple(a.shape[1:])).astype(np.float32)
    ((r[0] * torch.tensor(v_r, device=dev)).sum() + (a[0] * torch.tensor(v_a, device=dev)).sum()).backward()
    torch.cuda.synchronize()
    print("  stage ms:", {k: round(float(np.mean([x.elapsed_time(y) for x, y in v])), 3) for k, v in rendering.PROFILE.items()})
    rendering.PROFILE = None
    t0 = time.time()
    rc, ra, ref, ctx = cpu_ref.rasterization(s["means"], s["quats"], s["scales"], s["opacities"], s["colors"],
                                             s["viewmat"], s["K"], s["width"], s["height"], render_mode=rmode,
                                             rasterize_mode=mode)
    ctx["meta_offs"], ctx["meta_flat"] = ref["isect_offsets"], ref["flatten_ids"]
    g = cpu_ref.rasterization_bwd(ctx, v_r, v_a, absgrad=True)
    print(f"  oracle {time.time() - t0:.2f}s  N_vis {(ref['radii'] > 0).sum()}  M {ref['flatten_ids'].shape[0]}")
    vis = ref["radii"] > 0
    cmp_exact("radii", meta["radii"][0].cpu().numpy(), ref["radii"])
    cmp_exact("tiles_per_gauss", meta["tiles_per_gauss"][0].cpu().numpy(), ref["tiles_per_gauss"])
    cmp_exact("depths[vis]", meta["depths"][0].cpu().numpy()[vis], ref["depths"][vis])
    cmp_exact("means2d[vis]", meta["means2d"][0].detach().cpu().numpy()[vis], ref["means2d"][vis])
    cmp_exact("conics[vis]", meta["conics"][0].cpu().numpy()[vis], ref["conics"][vis])
    cmp_exact("opacities[vis]", meta["opacities"][0].cpu().numpy()[vis], ref["opacities"][vis])
    cmp_exact("flatten_ids", meta["flatten_ids"].cpu().numpy(), ref["flatten_ids"])
    cmp_exact("isect_offsets", meta["isect_offsets"][0].cpu().numpy(), ref["isect_offsets"])
    cmp_exact("isect_ids", meta["isect_ids"].cpu().numpy(), ref["isect_ids"])
    cmp_close("alpha", a[0].detach().cpu().numpy(), ra)
    cmp_close("render", r[0].detach().cpu().numpy(), rc)
    cmp_close("means2d.grad", meta["means2d"].grad[0].cpu().numpy(), g["v_means2d"], 5e-3, 2e-4 * np.abs(g["v_means2d"]).max())
    cmp_close("absgrad", meta["means2d"].absgrad[0].cpu().numpy(), g["v_means2d_abs"], 5e-3, 2e-4 * np.abs(g["v_means2d_abs"]).max())
    for k, gk in (("colors", "v_colors"), ("opacities", "v_opacities"), ("means", "v_means"), ("quats", "v_quats"),
                  ("scales", "v_scales"), ("viewmat", "v_viewmat")):
        ref_g = np.asarray(g[gk], np.float64)
        cmp_close(gk, t[k].grad.cpu().numpy(), ref_g, 5e-3, 2e-4 * np.abs(ref_g).max())


if __name__ == "__main__":
    torch.zeros(1, device="cuda")
    print(torch.cuda.get_device_name(0), "cpu cores", os.cpu_count(), "oracle threads", cpu_ref.num_threads())
    run(scenes.tiny(n=300, seed=3, width=64, height=48), "RGB+ED", "antialiased", "tiny")
    run(scenes.tiny(n=257, seed=9, width=77, height=53), "RGB", "classic", "tiny_ragged")
    run(scenes.config1(), "RGB+ED", "antialiased", "config1")
    s = scenes.street(n=100_000, seed=1)
    run(s, "RGB+ED", "antialiased", "street100k")
    s6 = scenes.street(n=100_000, seed=1, d_in=6)
    run(s6, "RGB+ED", "antialiased", "street100k-cdim8")
