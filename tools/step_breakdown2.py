import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from mtgs_b200 import scenes, rendering
from mtgs_b200.rendering import rasterization
dev = torch.device("cuda:0")
d_in = 6
s = scenes.street(n=2_000_000, seed=1, d_in=d_in)
names = ("means", "quats", "scales", "opacities", "colors")
for mode in ("tensor", "pinned_to"):
    if mode == "tensor":
        p = {k: torch.tensor(s[k], device=dev).requires_grad_(True) for k in names}
    else:
        host = {k: torch.from_numpy(s[k]).pin_memory() for k in names}
        p = {k: host[k].to(dev).requires_grad_(True) for k in names}
    vm = torch.from_numpy(s["viewmat"]).to(dev)[None]; K = torch.from_numpy(s["K"]).to(dev)[None]
    for wmode in ("plain", "generator"):
        if wmode == "plain":
            w_c = torch.randn(1, 1080, 1920, d_in + 1, device=dev); w_a = torch.randn(1, 1080, 1920, 1, device=dev)
        else:
            gen = torch.Generator(device=dev); gen.manual_seed(1234)
            w_c = torch.randn(1, 1080, 1920, d_in + 1, device=dev, generator=gen); w_a = torch.randn(1, 1080, 1920, 1, device=dev, generator=gen)
        for sync in (True, False):
            def step():
                r, a, m = rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], vm, K, 1920, 1080,
                                        packed=False, render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)
                loss = (r * w_c).sum() + (a * w_a).sum()
                for t in p.values(): t.grad = None
                loss.backward()
                return loss, m
            for _ in range(3): step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(10):
                loss, m = step()
                if sync: torch.cuda.synchronize()
            e1.record(); torch.cuda.synchronize()
            print(mode, wmode, "sync" if sync else "nosync", "gpu ms/step", round(e0.elapsed_time(e1) / 10, 3), "wall", round((time.perf_counter() - t0) * 100, 3),
                  "w_c stride", w_c.stride(), "colors stride", p["colors"].stride(), flush=True)
