"""GPU timeline of ONE steady-state bench step (forward + loss + backward) from the torch profiler: every kernel /
memcpy / memset with its start offset, duration and the idle gap in front of it.  Run on the GPU box:
    python tools/step_timeline.py [n_gauss] > gpurun_out/timeline.txt"""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from mtgs_b200 import scenes  # noqa: E402
from mtgs_b200.rendering import rasterization  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dev = torch.device("cuda:0")
s = scenes.street(n=n, seed=1)
names = ("means", "quats", "scales", "opacities", "colors")
p = {k: torch.tensor(s[k], device=dev).requires_grad_(True) for k in names}
viewmat = torch.tensor(s["viewmat"], device=dev)[None].requires_grad_(True)
Ks = torch.tensor(s["K"], device=dev)[None]
W, H = 1920, 1080
w_c = torch.randn(1, H, W, 4, device=dev)
w_a = torch.randn(1, H, W, 1, device=dev)


def step():
    r, a, _ = rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], viewmat, Ks, W, H,
                            packed=False, render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)
    loss = torch.dot(r.reshape(-1), w_c.reshape(-1)) + torch.dot(a.reshape(-1), w_a.reshape(-1))
    for t in p.values():
        t.grad = None
    viewmat.grad = None
    loss.backward()


for _ in range(10):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
evs.sort(key=lambda e: e.time_range.start)
# steps are delimited by the projection forward kernel; print the third one
starts = [i for i, e in enumerate(evs) if "k_project_fwd" in e.name]
lo, hi = starts[2], starts[3]
t0 = evs[lo].time_range.start
prev_end = t0
busy = 0.0
print(f"{'start_us':>9s} {'gap_us':>7s} {'dur_us':>8s}  kernel")
for e in evs[lo:hi]:
    st, en = e.time_range.start, e.time_range.end
    print(f"{st - t0:9.1f} {st - prev_end:7.1f} {en - st:8.1f}  {e.name[:100]}")
    busy += en - st
    prev_end = max(prev_end, en)
print(f"step span {evs[hi].time_range.start - t0:.1f} us, busy {busy:.1f} us, idle {evs[hi].time_range.start - t0 - busy:.1f} us")
