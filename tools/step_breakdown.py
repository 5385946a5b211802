"""Where does a step's wall time go?  GPU time (events) vs CPU issue time, fwd / loss / bwd, per variant."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from mtgs_b200 import scenes
from mtgs_b200.rendering import rasterization

dev = torch.device("cuda:0")
for variant, d_in in (("rgbed", 3), ("mtgs", 6)):
    s = scenes.street(n=2_000_000, seed=1, d_in=d_in)
    p = {k: torch.tensor(s[k], device=dev).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")}
    vm = torch.tensor(s["viewmat"], device=dev)[None]; K = torch.tensor(s["K"], device=dev)[None]
    w_c = torch.randn(1, 1080, 1920, d_in + 1, device=dev); w_a = torch.randn(1, 1080, 1920, 1, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    acc = np.zeros(3); cpu = np.zeros(3)
    for it in range(13):
        for t in p.values(): t.grad = None
        torch.cuda.synchronize()
        t0 = time.perf_counter(); ev[0].record()
        r, a, m = rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], vm, K, 1920, 1080,
                                packed=False, render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)
        t1 = time.perf_counter(); ev[1].record()
        loss = (r * w_c).sum() + (a * w_a).sum()
        t2 = time.perf_counter(); ev[2].record()
        loss.backward()
        t3 = time.perf_counter(); ev[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
            cpu += [(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3]
    print(variant, "GPU ms fwd/loss/bwd", (acc / 10).round(3), " CPU-issue ms", (cpu / 10).round(3))
