"""Seeded synthetic scenes for the BASELINE.json configs.

Pure numpy on the host; shapes and value ranges follow what MTGS hands to the rasterizer
(post-activation scales/opacities, unit quaternions, OpenCV world->camera ``viewmat``;
mtgs/scene_model/mtgs_scene_graph.py:601-661, gaussian_model/vanilla_gaussian_splatting.py:299-307).
There is no nuPlan data in the build or GPU containers, so every workload is procedural.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def _unit_quats(rng, n):
    q = rng.standard_normal((n, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32)


def config1(n: int = 10_000, seed: int = 0, width: int = 256, height: int = 256) -> Dict[str, np.ndarray]:
    """BASELINE config 1: 10k random Gaussians, one 256x256 pinhole camera, viewmat = I."""
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-4, 4, (n, 2))
    z = rng.uniform(1, 12, n)
    behind = rng.random(n) < 0.10
    z[behind] = rng.uniform(-2, 0.005, behind.sum())
    means = np.concatenate([xy, z[:, None]], 1).astype(np.float32)
    scales = np.exp(rng.uniform(np.log(0.02), np.log(0.4), (n, 3))).astype(np.float32)
    opac = rng.uniform(0.02, 0.98, n).astype(np.float32)
    colors = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    K = np.array([[width, 0, width / 2], [0, width, height / 2], [0, 0, 1]], np.float32)
    return dict(means=means, quats=_unit_quats(rng, n), scales=scales, opacities=opac, colors=colors,
                viewmat=np.eye(4, dtype=np.float32), K=K, width=width, height=height)


def street(n: int = 500_000, seed: int = 1, width: int = 1920, height: int = 1080, d_in: int = 3,
           camera: int = 0) -> Dict[str, np.ndarray]:
    """BASELINE config 2 ("street slab", nuPlan-front-camera-like intrinsics) and its 2M/3M scalings.

    ``camera`` selects one of several poses looking down the same slab (used to give every rank of a
    multi-GPU run its own traversal camera over the shared Gaussians)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-40, 40, n)
    y = rng.uniform(-8, 1.8, n)
    z = rng.uniform(0.5, 120, n)
    behind = rng.random(n) < 0.15
    z[behind] = rng.uniform(-60, 0.0, behind.sum())
    means = np.stack([x, y, z], 1)
    scales = np.exp(rng.uniform(np.log(0.02), np.log(0.5), (n, 3)))
    thin = rng.integers(0, 3, n)
    scales[np.arange(n), thin] *= 0.1  # "two_d_gaussians"-like flat splats (mtgs/config/MTGS.py:116)
    sky = rng.random(n) < 0.05
    ns = int(sky.sum())
    d = rng.standard_normal((ns, 3))
    d[:, 2] = np.abs(d[:, 2]) + 0.2
    d[:, 1] = -np.abs(d[:, 1])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    means[sky] = d * 1000.0
    scales[sky] *= 100.0
    opac = 1.0 / (1.0 + np.exp(-rng.normal(0, 2, n)))
    colors = rng.uniform(0, 1, (n, d_in))
    fx = fy = 1545.0 * width / 1920.0
    K = np.array([[fx, 0, width / 2.0], [0, fy, 560.0 * height / 1080.0], [0, 0, 1]], np.float32)
    # camera poses: small yaw / lateral shifts about the slab axis (world == camera-0 frame)
    yaw = np.deg2rad([0.0, 4.0, -4.0, 8.0, -8.0, 2.0, -2.0, 6.0][camera % 8])
    shift = [0.0, 1.5, -1.5, 3.0, -3.0, 0.75, -0.75, 2.25][camera % 8]
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])
    viewmat = np.eye(4)
    viewmat[:3, :3] = R
    viewmat[:3, 3] = R @ np.array([-shift, 0.0, 0.0])
    return dict(means=means.astype(np.float32), quats=_unit_quats(rng, n), scales=scales.astype(np.float32),
                opacities=opac.astype(np.float32), colors=colors.astype(np.float32),
                viewmat=viewmat.astype(np.float32), K=K, width=width, height=height)


def tiny(n: int = 300, seed: int = 3, width: int = 64, height: int = 48, d_in: int = 3) -> Dict[str, np.ndarray]:
    """Small scene for autograd / brute-force checks; includes culled and off-screen Gaussians."""
    rng = np.random.default_rng(seed)
    means = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.2, 1.2, n), rng.uniform(0.8, 6, n)], 1)
    means[: n // 10, 2] = rng.uniform(-1, 0.005, n // 10)
    scales = np.exp(rng.uniform(np.log(0.03), np.log(0.35), (n, 3)))
    opac = rng.uniform(0.05, 0.95, n)
    colors = rng.uniform(0, 1, (n, d_in))
    K = np.array([[60.0, 0, width / 2 + 0.7], [0, 62.0, height / 2 - 1.3], [0, 0, 1]], np.float32)
    ang = 0.1
    c, s = np.cos(ang), np.sin(ang)
    viewmat = np.eye(4)
    viewmat[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    viewmat[:3, 3] = [0.05, -0.02, 0.1]
    return dict(means=means.astype(np.float32), quats=_unit_quats(rng, n), scales=scales.astype(np.float32),
                opacities=opac.astype(np.float32), colors=colors.astype(np.float32),
                viewmat=viewmat.astype(np.float32), K=K, width=width, height=height)
