"""Procedural stand-in for BASELINE configs 3-4 (multi-traversal road block; the nuPlan data is not available).

A "true" scene (shared geometry + per-traversal colour residuals, the structure of the reference's
MultiColorGaussianSplattingModel: features_dc + features_adapters[:, t], features_rest; reference
mtgs/scene_model/gaussian_model/multi_color_gaussian_splatting.py:53-101) renders target images for T traversals;
a perturbed copy is then optimised against them through the SAME public API MTGS uses:

    spherical_harmonics(n, viewdirs, coeffs) -> clamp(+0.5) -> rasterization(..., render_mode="RGB+ED",
    rasterize_mode="antialiased", absgrad=True) -> masked L1 + inverse-depth L1 (mtgs_b200.losses, the fused forms of
    mtgs_scene_graph.py:825-828, 875-879) -> backward -> ONE fused Adam launch over every parameter group
    (mtgs_b200.optim.FusedAdam) -> densification statistics from info["means2d"].absgrad in one pass
    (accumulate_densify_stats; mtgs_scene_graph.py:1171-1178 + vanilla_gaussian_splatting.py:455-474).

Single GPU:   python examples/train_multitraversal.py --iters 300
Multi GPU :   torchrun --nproc-per-node 4 --master-addr 127.0.0.1 examples/train_multitraversal.py --traversals 4
              (rank r owns traversal r; the gradients of the shared geometry are exchanged INSIDE the projection
               backward (GradExchange, exchange_colors=False because the SH colours are view dependent), the SH
               coefficient gradients are all-reduced at their leaves (SharedGradArena), the per-traversal
               adapters stay rank-local -- SURVEY.md 8e)
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mtgs_b200 import scenes  # noqa: E402
from mtgs_b200.cuda._wrapper import spherical_harmonics  # noqa: E402
from mtgs_b200.losses import masked_l1  # noqa: E402
from mtgs_b200.optim import FusedAdam, accumulate_densify_stats  # noqa: E402
from mtgs_b200.parallel import GradExchange, SharedGradArena, traversal_of_rank  # noqa: E402
from mtgs_b200.rendering import rasterization  # noqa: E402

C0 = 0.28209479177387814


def make_model(n, n_trav, width, height, seed, dev):
    s = scenes.street(n=n, seed=seed, width=width, height=height)
    g = torch.Generator().manual_seed(seed)
    p = {
        "means": torch.tensor(s["means"]),
        "scales": torch.log(torch.tensor(s["scales"])),                       # raw (exp activation)
        "quats": torch.tensor(s["quats"]),                                     # raw (normalised on use)
        "opacities": torch.logit(torch.tensor(s["opacities"]).clamp(1e-4, 1 - 1e-4)),
        "features_dc": (torch.tensor(s["colors"]) - 0.5) / C0,
        "features_rest": 0.05 * torch.randn(n, 15, 3, generator=g),
        "features_adapters": 0.3 * torch.randn(n, n_trav, 3, generator=g),     # per-traversal colour residual
    }
    return {k: v.to(dev).float() for k, v in p.items()}, s


def render(p, trav, viewmat, K, W, H, sh_degree, absgrad=True):
    means = p["means"]
    scales = torch.exp(p["scales"])
    quats = p["quats"] / p["quats"].norm(dim=-1, keepdim=True)
    opac = torch.sigmoid(p["opacities"])
    dc = p["features_dc"] + p["features_adapters"][:, trav, :]
    coeffs = torch.cat([dc[:, None, :], p["features_rest"]], dim=1)
    campos = torch.inverse(viewmat)[:3, 3]
    dirs = means.detach() - campos
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    rgbs = torch.clamp(spherical_harmonics(sh_degree, dirs, coeffs) + 0.5, 0.0, 1.0)
    out, alpha, info = rasterization(means=means, quats=quats, scales=scales, opacities=opac, colors=rgbs,
                                     viewmats=viewmat[None], Ks=K[None], width=W, height=H, tile_size=16,
                                     packed=False, near_plane=0.01, far_plane=1e10, render_mode="RGB+ED",
                                     sparse_grad=False, absgrad=absgrad, rasterize_mode="antialiased")
    rgb = torch.clamp(out[..., :3] + (1 - alpha) * 1.0, 0.0, 1.0)
    return rgb[0], out[0, ..., 3:4], alpha[0], info


def psnr(a, b):
    return float(-10.0 * torch.log10(torch.mean((a - b) ** 2).clamp_min(1e-12)))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-gauss", dest="n", type=int, default=100_000)
    ap.add_argument("--traversals", type=int, default=3)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--log-every", type=int, default=50)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--json-out", default=None, help="write per-traversal PSNR before / after and the iteration rate here")
    args = ap.parse_args(argv)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    W, H, T = args.width, args.height, args.traversals
    truth, s = make_model(args.n, T, W, H, args.seed, dev)
    K = torch.tensor(s["K"], device=dev)
    cams = [torch.tensor(scenes.street(n=8, seed=args.seed, width=W, height=H, camera=c)["viewmat"], device=dev)
            for c in range(T)]
    my_travs = traversal_of_rank(rank, world, T)
    with torch.no_grad():
        targets = {t: render(truth, t, cams[t], K, W, H, 3, absgrad=False)[:2] for t in my_travs}

    # perturbed model (identical on every rank: same seed)
    g = torch.Generator().manual_seed(args.seed + 1)
    model = {}
    for k, v in truth.items():
        noise = {"means": 0.02, "scales": 0.15, "quats": 0.05, "opacities": 0.5, "features_dc": 0.5,
                 "features_rest": 0.05, "features_adapters": 0.3}[k]
        model[k] = (v + noise * torch.randn(v.shape, generator=g).to(dev)).requires_grad_(True)
    # multi-GPU: geometry gradients (means / scales / quats / opacities: replicated elementwise activations) are
    # averaged inside the projection backward; the colour path goes through per-camera SH, so its shared leaves
    # (features_dc, features_rest) are all-reduced where they live
    arena = SharedGradArena([model[k] for k in ("features_dc", "features_rest")], average=True) if world > 1 else None
    exch = GradExchange(n_shared=args.n, d_in=3, rows_cap=args.n, average=True, exchange_colors=False) if world > 1 else None
    lrs = {"means": 1.6e-4, "scales": 5e-3, "quats": 1e-3, "opacities": 5e-2, "features_dc": 2.5e-3,
           "features_rest": 1.25e-4, "features_adapters": 2.5e-3}
    opt = FusedAdam([{"params": [model[k]], "lr": lrs[k]} for k in model], eps=1e-15)
    grad_norm_acc = torch.zeros(args.n, device=dev)
    first = {}
    vis_count = torch.zeros(args.n, device=dev)
    max_2dsize = torch.zeros(args.n, device=dev)
    t_start = None

    for it in range(args.iters + 1):
        if it == min(20, args.iters):  # iteration rate without the first-use costs
            torch.cuda.synchronize()
            t_start = (it, __import__("time").perf_counter())
        t = my_travs[it % len(my_travs)]
        sh_degree = min(it // max(1, args.iters // 4), 3)  # progressive SH degree (reference: sh_degree_interval)
        rgb, depth, alpha, info = render(model, t, cams[t], K, W, H, sh_degree)
        info["means2d"].retain_grad()
        tgt_rgb, tgt_depth = targets[t]
        loss = masked_l1(rgb, tgt_rgb) + 0.05 * masked_l1(depth, tgt_depth, inverse=True)
        if it < len(my_travs):
            first.setdefault(t, psnr(rgb.detach(), tgt_rgb))
        if arena is not None:
            arena.zero_()
            for k in ("means", "scales", "quats", "opacities", "features_adapters"):
                model[k].grad = None
            with exch.active():
                loss.backward()
            arena.all_reduce()
        else:
            opt.zero_grad(set_to_none=True)
            loss.backward()
        # densification statistics (mtgs_scene_graph.py:1171-1178 + vanilla_gaussian_splatting.py:455-474), one pass
        accumulate_densify_stats(info["means2d"].absgrad[0], info["radii"], W, H, grad_norm_acc, vis_count, max_2dsize)
        opt.step()
        if it % args.log_every == 0 and rank == 0:
            print(f"iter {it:5d}  traversal {t}  sh {sh_degree}  loss {float(loss):.5f}  "
                  f"psnr {psnr(rgb.detach(), tgt_rgb):.2f} dB  visible {int((info['radii'][0] > 0).sum())}  "
                  f"mean absgrad stat {float(grad_norm_acc.sum() / vis_count.sum().clamp_min(1)):.4e}", flush=True)
    torch.cuda.synchronize()
    its = (args.iters + 1 - t_start[0]) / max(1e-9, __import__("time").perf_counter() - t_start[1]) if t_start else None
    with torch.no_grad():
        final = {t: psnr(render(model, t, cams[t], K, W, H, 3, absgrad=False)[0].detach(), targets[t][0]) for t in my_travs}
    if rank == 0:
        print("final PSNR per traversal:", {k: round(v, 2) for k, v in final.items()}, "iterations/s:", its)
        if args.json_out:
            import json
            json.dump({"config": vars(args), "world": world, "psnr_first_db": first, "psnr_final_db": final,
                       "iterations_per_s": its,
                       "note": "procedural stand-in for BASELINE config 3 (the nuPlan road block is not available): a "
                               "perturbed multi-traversal model optimised against renders of the true one through the "
                               "public API"}, open(args.json_out, "w"), indent=1)
    if world > 1:
        import torch.distributed as dist
        exch.check()
        dist.barrier()
        exch.close()
        dist.destroy_process_group()
    return first, final


if __name__ == "__main__":
    main()
