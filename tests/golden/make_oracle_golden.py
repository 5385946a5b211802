"""Regression fixture of the CPU oracle's OWN outputs (not a reference golden: parity unpinned).

    python tests/golden/make_oracle_golden.py   ->  tests/golden/oracle_tiny_golden.npz

Used (a) by the CPU suite to detect accidental changes of the oracle and (b) by the GPU suite as a
fixture that travels to the GPU box (inputs + expected outputs, antialiased RGB+ED with absgrad).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mtgs_b200 import scenes  # noqa: E402
from oracle import cpu_ref  # noqa: E402

s = scenes.tiny(n=300, seed=3, width=64, height=48)
rc, ra, meta, ctx = cpu_ref.rasterization(s["means"], s["quats"], s["scales"], s["opacities"], s["colors"],
                                          s["viewmat"], s["K"], s["width"], s["height"], render_mode="RGB+ED",
                                          rasterize_mode="antialiased")
ctx["meta_offs"], ctx["meta_flat"] = meta["isect_offsets"], meta["flatten_ids"]
rng = np.random.default_rng(17)
v_r = rng.standard_normal(rc.shape).astype(np.float32)
v_a = rng.standard_normal(ra.shape).astype(np.float32)
g = cpu_ref.rasterization_bwd(ctx, v_r, v_a, absgrad=True)
out = dict(width=s["width"], height=s["height"], render=rc, alpha=ra, v_render=v_r, v_alpha=v_a)
for k in ("means", "quats", "scales", "opacities", "colors", "viewmat", "K"):
    out["in_" + k] = s[k]
for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets", "means2d", "depths", "conics",
          "opacities"):
    out[k] = meta[k]
for k in ("v_means", "v_quats", "v_scales", "v_opacities", "v_colors", "v_viewmat", "v_means2d", "v_means2d_abs"):
    out[k] = np.asarray(g[k], np.float64)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_tiny_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path))
