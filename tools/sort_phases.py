"""Phase timing of the cooperative depth sort (b2s_debug_sort_depth_phases): run on the GPU box.
    python tools/sort_phases.py [n_gauss]
Prints the duration of every phase (histogram + barrier, row scan + barrier, rank + scatter + barrier) per pass."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mtgs_b200 import _lib, scenes  # noqa: E402
from mtgs_b200.rendering import rasterization  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dev = torch.device("cuda:0")
lib = _lib.load()
s = scenes.street(n=n, seed=1)
t = {k: torch.tensor(s[k], device=dev) for k in ("means", "quats", "scales", "opacities", "colors", "viewmat", "K")}
# keys as the projection produces them
from mtgs_b200 import rendering as R  # noqa: E402
with torch.no_grad():
    out = R._Project.apply(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"], t["viewmat"], t["K"], 1920, 1080,
                           120, 68, 0.3, 0.01, 1e10, 0.0, True, True, 4, False)
keys = out[6]
order = torch.empty(n, dtype=torch.int32, device=dev)
nvis = torch.empty(1, dtype=torch.int32, device=dev)
wsb = int(lib.b2s_bin_depth_workspace_bytes(n))
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
ph = torch.zeros(13, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda x: C.c_void_p(x.data_ptr())
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.b2s_debug_sort_depth_phases(p(keys), n, p(order), p(nvis), p(ws), wsb, p(ph), st), "sort")
    e1.record()
    torch.cuda.synchronize()
    v = ph.cpu().numpy().astype(np.int64)
    d = np.diff(v) / 1e3
    print(f"run {it}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us, n_vis {int(nvis)}; phases (us) hist/scan/scatter per pass:",
          " | ".join(f"{d[3 * k]:.1f} {d[3 * k + 1]:.1f} {d[3 * k + 2]:.1f}" for k in range(4)), f"total {np.sum(d):.1f}")
ref = torch.sort(keys.view(torch.int32).long() & 0xFFFFFFFF, stable=True)
nv = int(nvis)
assert torch.equal(order[:nv].long(), ref.indices[:nv]), "order mismatch vs torch stable sort"
print("order == torch stable sort")
