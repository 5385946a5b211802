"""Image-space losses of the MTGS training step as fused B200 kernels.

Drop-ins for the reference's torch op sequences (same names / argument meaning where the reference has a function):

    masked_l1(pred, gt, mask)                      torch.abs(gt - pred)[mask].mean()
                                                   mtgs/scene_model/mtgs_scene_graph.py:825-828 (RGB), :929 (normals)
    masked_l1(pred, gt, mask, inverse=True)        the LiDAR InverseL1 depth loss, mtgs_scene_graph.py:875-879
    TVLoss()(pred)                                 mtgs/utils/geometric_loss.py:287-303
    calculate_depth_ncc_loss(pred, gt, patch, stride, mask)
                                                   mtgs/utils/geometric_loss.py:322-348
    pcd_to_normal(xyz), normal_from_depth_image(depths, fx, fy, cx, cy, img_size, c2w, device)
                                                   mtgs/utils/geometric_loss.py:350-388

Unlike the reference's boolean indexing none of these synchronises with the host: sums and counts accumulate on the
device and the means are formed there.  Kernels: csrc/losses.cu behind include/b200splat.h.  CUDA float32 tensors only
(no CPU / PyTorch fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from .rendering import _need_cuda, _ptr, _stream


def _mask_u8(mask: Optional[Tensor], n_pix: int) -> Optional[Tensor]:
    if mask is None:
        return None
    m = mask.reshape(-1)
    if m.numel() != n_pix:
        raise ValueError(f"mask has {m.numel()} entries for {n_pix} pixels")
    return (m if m.dtype == torch.uint8 else (m != 0).to(torch.uint8)).contiguous()


class _MaskedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, mask, mode, eps):
        lib = _lib.load()
        Cn = pred.shape[-1]
        P = pred.numel() // Cn
        acc = torch.zeros(2, dtype=torch.float64, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(lib.b2s_masked_l1_fwd(_ptr(pred), _ptr(gt), _ptr(mask), P, Cn, mode, eps, _ptr(acc), _stream()),
                       "b2s_masked_l1_fwd")
        ctx.save_for_backward(pred, gt, mask, acc)
        ctx.cfg = (P, Cn, mode, eps)
        return (acc[0] / acc[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        pred, gt, mask, acc = ctx.saved_tensors
        P, Cn, mode, eps = ctx.cfg
        grad = torch.empty_like(pred)
        g = g.contiguous().to(torch.float32)
        with torch.cuda.device(pred.device):
            _lib.check(lib.b2s_masked_l1_bwd(_ptr(pred), _ptr(gt), _ptr(mask), P, Cn, mode, eps, _ptr(acc), _ptr(g),
                                             _ptr(grad), _stream()), "b2s_masked_l1_bwd")
        return grad, None, None, None, None


def masked_l1(pred: Tensor, gt: Tensor, mask: Optional[Tensor] = None, inverse: bool = False, eps: float = 1e-5) -> Tensor:
    """``torch.abs(gt - pred)[mask].mean()`` (``inverse``: of ``1 / (x + eps)``) for ``pred``, ``gt`` [..., C] and a
    per-pixel ``mask`` (any shape with one entry per pixel, e.g. [H, W] or [H, W, 1]); gradient w.r.t. ``pred``."""
    _need_cuda(pred, gt, mask)
    if pred.shape != gt.shape:
        raise ValueError(f"pred {tuple(pred.shape)} and gt {tuple(gt.shape)} differ")
    pred = pred.contiguous().float()
    gt = gt.contiguous().float()
    m = _mask_u8(mask, pred.numel() // pred.shape[-1])
    return _MaskedL1.apply(pred, gt, m, 1 if inverse else 0, float(eps))


class _TV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred):
        lib = _lib.load()
        H, W, Cn = pred.shape[-3:]
        B = pred.numel() // (H * W * Cn)
        acc = torch.zeros(2, dtype=torch.float64, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(lib.b2s_tv_fwd(_ptr(pred), B, H, W, Cn, _ptr(acc), _stream()), "b2s_tv_fwd")
        ctx.save_for_backward(pred)
        ctx.cfg = (B, H, W, Cn)
        return (acc[0] / (B * H * (W - 1) * Cn) + acc[1] / (B * (H - 1) * W * Cn)).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (pred,) = ctx.saved_tensors
        B, H, W, Cn = ctx.cfg
        grad = torch.empty_like(pred)
        g = g.contiguous().to(torch.float32)
        with torch.cuda.device(pred.device):
            _lib.check(lib.b2s_tv_bwd(_ptr(pred), B, H, W, Cn, _ptr(g), _ptr(grad), _stream()), "b2s_tv_bwd")
        return grad


class TVLoss(torch.nn.Module):
    """Total-variation loss of a [..., H, W, C] map (reference class of the same name, geometric_loss.py:287-303)."""

    def forward(self, pred: Tensor) -> Tensor:
        _need_cuda(pred)
        if pred.dim() < 3:
            raise ValueError("pred must be [..., H, W, C]")
        return _TV.apply(pred.contiguous().float())


class _NCC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, mask, patch, stride):
        lib = _lib.load()
        H, W = pred.shape
        npy, npx = C.c_int(), C.c_int()
        _lib.check(lib.b2s_ncc_patch_grid(H, W, patch, stride, C.byref(npy), C.byref(npx)), "b2s_ncc_patch_grid")
        stats = torch.empty(npy.value * npx.value, 6, dtype=torch.float32, device=pred.device)
        acc = torch.zeros(2, dtype=torch.float64, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(lib.b2s_ncc_fwd(_ptr(pred), _ptr(gt), _ptr(mask), H, W, patch, stride, _ptr(stats), _ptr(acc),
                                       _stream()), "b2s_ncc_fwd")
        ctx.save_for_backward(pred, gt, stats, acc)
        ctx.cfg = (H, W, patch, stride)
        return (1.0 - acc[0] / acc[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        pred, gt, stats, acc = ctx.saved_tensors
        H, W, patch, stride = ctx.cfg
        grad = torch.empty_like(pred)
        g = g.contiguous().to(torch.float32)
        with torch.cuda.device(pred.device):
            _lib.check(lib.b2s_ncc_bwd(_ptr(pred), _ptr(gt), H, W, patch, stride, _ptr(stats), _ptr(acc), _ptr(g),
                                       _ptr(grad), _stream()), "b2s_ncc_bwd")
        return grad, None, None, None, None


def calculate_depth_ncc_loss(pred_depth: Tensor, gt_depth: Tensor, patch_size: int = 7, stride: int = 7,
                             mask: Optional[Tensor] = None) -> Tensor:
    """1 - mean normalised cross-correlation over the fully masked-in patches (reference signature,
    geometric_loss.py:322).  ``pred_depth``, ``gt_depth``, ``mask``: [H, W, 1] (or [H, W])."""
    _need_cuda(pred_depth, gt_depth, mask)
    if mask is None:
        raise AttributeError("'NoneType' object has no attribute 'squeeze'")  # what the reference raises
    pred = (pred_depth.squeeze(-1) if pred_depth.dim() == 3 else pred_depth).contiguous().float()
    gt = (gt_depth.squeeze(-1) if gt_depth.dim() == 3 else gt_depth).contiguous().float()
    if pred.dim() != 2 or pred.shape != gt.shape:
        raise ValueError(f"depth maps must be [H, W(, 1)] of one shape, got {tuple(pred_depth.shape)} / {tuple(gt_depth.shape)}")
    m = _mask_u8(mask, pred.numel())
    return _NCC.apply(pred, gt, m, int(patch_size), int(stride))


def normal_from_depth_image(depths: Tensor, fx: float, fy: float, cx: float, cy: float, img_size: tuple, c2w: Tensor,
                            device: torch.device = None, smooth: bool = False) -> Tensor:
    """Normals [H, W, 3] estimated from a depth map (reference signature, geometric_loss.py:365-388; ``img_size`` is
    (width, height)).  No gradient (the reference passes a detached ground-truth depth)."""
    if smooth:
        raise NotImplementedError("smooth=True (cv2.GaussianBlur on the host) is not built; MTGS passes smooth=False")
    _need_cuda(depths)
    lib = _lib.load()
    W, H = int(img_size[0]), int(img_size[1])
    d = depths.detach().reshape(H, W).contiguous().float()
    c2w = c2w.to(d.device, torch.float32)
    A_t = torch.cat([torch.linalg.inv(c2w[:3, :3]).reshape(-1), c2w[:3, 3].reshape(-1)]).contiguous()
    out = torch.empty(H, W, 3, dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(lib.b2s_normal_from_depth(_ptr(d), H, W, float(fx), float(fy), float(cx), float(cy), _ptr(A_t),
                                             _ptr(out), _stream()), "b2s_normal_from_depth")
    return out
