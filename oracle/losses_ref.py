"""numpy restatement (float64) of the image-space losses of row f3, values and gradients w.r.t. the prediction.

TEST INFRASTRUCTURE ONLY (see oracle/splat_oracle.c).  PINNED: tests/test_losses_cpu.py checks every function against
tests/golden/losses_reference_golden.npz, which was produced by the reference's own code
(tests/golden/make_losses_golden.py).  Reference: mtgs/scene_model/mtgs_scene_graph.py:825-828, 875-884, 929 (masked
L1 / inverse L1), mtgs/utils/geometric_loss.py:287-303 (TVLoss), :322-348 (calculate_depth_ncc_loss), :350-388
(pcd_to_normal / normal_from_depth_image) with mtgs/utils/camera_utils.py:74-148.
"""
from __future__ import annotations

import numpy as np


def masked_l1(pred, gt, mask, inverse=False, eps=1e-5, grad_out=1.0):
    pred = np.asarray(pred, np.float64)
    gt = np.asarray(gt, np.float64)
    m = np.asarray(mask, bool).reshape(pred.shape[:-1])
    if inverse:
        u = 1.0 / (gt + eps) - 1.0 / (pred + eps)
        d = np.sign(u) / (pred + eps) ** 2
    else:
        u = gt - pred
        d = -np.sign(u)
    n = m.sum() * pred.shape[-1]
    val = np.abs(u)[m].sum() / n
    grad = np.where(m[..., None], d, 0.0) * (grad_out / n)
    return val, grad


def tv_loss(pred, grad_out=1.0):
    p = np.asarray(pred, np.float64)
    dw = p[..., :, :-1, :] - p[..., :, 1:, :]
    dh = p[..., :-1, :, :] - p[..., 1:, :, :]
    val = np.abs(dw).mean() + np.abs(dh).mean()
    g = np.zeros_like(p)
    g[..., :, :-1, :] += np.sign(dw) * (grad_out / dw.size)
    g[..., :, 1:, :] -= np.sign(dw) * (grad_out / dw.size)
    g[..., :-1, :, :] += np.sign(dh) * (grad_out / dh.size)
    g[..., 1:, :, :] -= np.sign(dh) * (grad_out / dh.size)
    return val, g


def depth_ncc_loss(pred, gt, patch, stride, mask, grad_out=1.0):
    p = np.asarray(pred, np.float64).reshape(np.asarray(pred).shape[:2])
    g = np.asarray(gt, np.float64).reshape(p.shape)
    m = np.asarray(mask, bool).reshape(p.shape)
    H, W = p.shape
    pad = patch // 2
    npy = (H + 2 * pad - patch) // stride + 1
    npx = (W + 2 * pad - patch) // stride + 1
    n = patch * patch
    nccs, grads = [], np.zeros_like(p)
    contrib = []
    for py in range(npy):
        for px in range(npx):
            y0, x0 = py * stride - pad, px * stride - pad
            if y0 < 0 or x0 < 0 or y0 + patch > H or x0 + patch > W:
                continue  # touches the zero padding: its mask is not all ones
            if not m[y0:y0 + patch, x0:x0 + patch].all():
                continue
            a = p[y0:y0 + patch, x0:x0 + patch]
            b = g[y0:y0 + patch, x0:x0 + patch]
            ac, bc = a - a.mean(), b - b.mean()
            sa, sb = np.sqrt((ac ** 2).mean() + 1e-8), np.sqrt((bc ** 2).mean() + 1e-8)
            ah, bh = ac / sa, bc / sb
            ncc = (ah * bh).mean()
            nccs.append(ncc)
            contrib.append((y0, x0, (bh - ah * ncc) / (n * sa)))
    val = 1.0 - (np.mean(nccs) if nccs else np.nan)
    for y0, x0, d in contrib:
        grads[y0:y0 + patch, x0:x0 + patch] += d * (-grad_out / len(nccs))
    return val, grads


def normal_from_depth(depth, fx, fy, cx, cy, c2w):
    d = np.asarray(depth, np.float64)
    d = d.reshape(d.shape[:2])
    H, W = d.shape
    u, v = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    pts = np.stack([(u - cx) * d / fx, (v - cy) * d / fy, d], -1)
    c2w = np.asarray(c2w, np.float64)
    pts = pts @ np.linalg.inv(c2w[:3, :3]) + c2w[:3, 3]
    out = np.zeros((H, W, 3))
    l2r = pts[1:-1, 2:] - pts[1:-1, :-2]
    b2t = pts[:-2, 1:-1] - pts[2:, 1:-1]
    nrm = np.cross(l2r, b2t)
    out[1:-1, 1:-1] = nrm / np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-12)
    return out
