"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): forward + backward on two small scenes, the
multi-rank gradient exchange played on one GPU, and the SH operator."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mtgs_b200 import scenes  # noqa: E402
from mtgs_b200.cuda._wrapper import spherical_harmonics  # noqa: E402
from mtgs_b200.parallel import GradExchange  # noqa: E402
from mtgs_b200.rendering import rasterization  # noqa: E402

dev = torch.device("cuda:0")
NAMES = ("means", "quats", "scales", "opacities", "colors")


def run(s, ex=None, mode="antialiased", rmode="RGB+ED"):
    t = {k: torch.tensor(s[k], device=dev).requires_grad_(True) for k in NAMES}
    r, a, m = rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"],
                            torch.tensor(s["viewmat"], device=dev)[None], torch.tensor(s["K"], device=dev)[None],
                            s["width"], s["height"], packed=False, render_mode=rmode, rasterize_mode=mode, absgrad=True)
    loss = r.sum() + a.sum()
    if ex is None:
        loss.backward()
    else:
        with ex.active():
            loss.backward()
    return m


run(scenes.tiny(n=257, seed=9, width=77, height=53))
run(scenes.street(n=3000, seed=1, width=640, height=360), mode="classic", rmode="RGB")
run(scenes.street(n=1500, seed=4, width=4400, height=304))
exs = GradExchange.local_ranks(2, 2900, 3, rows_cap=3000, device=dev)
for r in range(2):
    run(scenes.street(n=3000, seed=1, width=640, height=360, camera=r), ex=exs[r])
GradExchange.finish_all(exs)
torch.cuda.synchronize()
exs[0].check()
dirs = torch.randn(1000, 3, device=dev)
co = torch.randn(1000, 16, 3, device=dev, requires_grad=True)
spherical_harmonics(3, dirs, co).sum().backward()
torch.cuda.synchronize()
print("sanitize_small done")
