"""Which part of the step can be captured into a CUDA graph?  (debug aid; run on the GPU box)"""
import sys
import traceback

import torch

sys.path.insert(0, ".")
from mtgs_b200 import rendering, scenes  # noqa: E402
from mtgs_b200.rendering import rasterization  # noqa: E402

dev = torch.device("cuda:0")
s = scenes.street(n=50_000, seed=1, width=640, height=360)
names = ("means", "quats", "scales", "opacities", "colors")
p = {k: torch.tensor(s[k], device=dev).requires_grad_(True) for k in names}
vm = torch.tensor(s["viewmat"], device=dev)[None]
Ks = torch.tensor(s["K"], device=dev)[None]
kw = dict(packed=False, render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)


def fwd_nograd():
    with torch.no_grad():
        return rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], vm, Ks, 640, 360, **kw)[:2]


def fwd_grad():
    return rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], vm, Ks, 640, 360, **kw)[:2]


def full():
    r, a = fwd_grad()
    for t in p.values():
        t.grad = None
    (r.sum() + a.sum()).backward()
    return r, a


def sort_only():
    keys = torch.randint(0, 2 ** 30, (100_000,), device=dev, dtype=torch.int32)
    return rendering._sort_depth(keys)


for name, fn in (("sort_only", sort_only), ("fwd_nograd", fwd_nograd), ("fwd_grad", fwd_grad), ("full", full)):
    try:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g):
                out = fn()
        except Exception as e:
            print(name, "CAPTURE FAILED:", type(e).__name__, str(e).splitlines()[0])
            c = e.__context__
            while c is not None:
                print("   caused by:", type(c).__name__, str(c).splitlines()[0])
                c = c.__context__
            traceback.print_exc(limit=6)
            continue
        g.replay()
        torch.cuda.synchronize()
        print(name, "captured and replayed OK")
    except Exception as e:
        print(name, "ERROR outside capture:", type(e).__name__, str(e).splitlines()[0])
