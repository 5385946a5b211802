"""Masked SSIM -- drop-in for ``mtgs/utils/ssim.py`` (``ssim`` and ``MaskedSSIM``).

Reference: mtgs/utils/ssim.py:110-190 (``ssim``), :56-108 (``_ssim``), :193-238 (``MaskedSSIM``); MTGS calls
``self.ssim(gt_img.permute(2, 0, 1)[None], pred_img.permute(2, 0, 1)[None], mask=combined_mask)`` at
mtgs/scene_model/mtgs_scene_graph.py:822-840 with ``data_range=1.0, size_average=True, channel=3``.
Same signatures, argument meaning, return shapes and error behaviour for 4-d ``(N, C, H, W)`` inputs; the arithmetic
(two-pass Gaussian filter of X, Y, XX, YY, XY, SSIM map, masked mean, and the whole backward) runs in two fused
kernels behind the C ABI (``b2s_ssim_fwd`` / ``b2s_ssim_bwd``, csrc/ssim.cu).  No CPU / PyTorch fallback.
Parity for this row is PINNED: tests/golden/ssim_reference_golden.npz was produced by importing the reference's own
ssim.py (tests/golden/make_ssim_golden.py).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from . import _lib


def _fspecial_gauss_1d(size: int, sigma: float) -> Tensor:
    """Normalised 1-D Gaussian window of ``size`` taps, shape (1, 1, size) -- same values as the reference helper of
    this name (ssim.py:11-25)."""
    offsets = torch.arange(size, dtype=torch.float32) - float(size // 2)
    weights = torch.exp(offsets.square().neg() / (2 * sigma ** 2))
    return (weights / weights.sum()).view(1, 1, size)


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class _SSIMMap(torch.autograd.Function):
    """(X, Y, mask, win) -> per-plane (masked sum of SSIM, masked count) as a [N*C, 2] float64 tensor."""

    @staticmethod
    def forward(ctx, X, Y, mask_u8, mask_strides, win1d, C1, C2):
        lib = _lib.load()
        N, Cc, H, W = X.shape
        R = win1d.numel()
        Ho, Wo = H - R + 1, W - R + 1
        dev = X.device
        need_x = ctx.needs_input_grad[0]
        maps = torch.empty(4 if need_x else 3, N * Cc, Ho, Wo, dtype=torch.float32, device=dev)
        acc = torch.zeros(N * Cc, 2, dtype=torch.float64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(lib.b2s_ssim_fwd(_ptr(X), _ptr(Y), _ptr(mask_u8), mask_strides[0], mask_strides[1], N, Cc, H, W,
                                        _ptr(win1d), R, float(C1), float(C2), _ptr(maps[0]), _ptr(maps[1]), _ptr(maps[2]),
                                        _ptr(maps[3]) if need_x else None, _ptr(acc), stream), "b2s_ssim_fwd")
        ctx.save_for_backward(X, Y, maps, win1d)
        return acc

    @staticmethod
    def backward(ctx, v_acc):
        lib = _lib.load()
        X, Y, maps, win1d = ctx.saved_tensors
        N, Cc, H, W = X.shape
        dev = X.device
        # d(result)/d(ssim_p) = v_acc[plane, 0] * mask_p (the count column has no dependence on the images)
        scale = v_acc[:, 0].to(torch.float32).contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        gX = gY = None
        with torch.cuda.device(dev):
            if ctx.needs_input_grad[1]:
                gY = torch.empty_like(Y)
                _lib.check(lib.b2s_ssim_bwd(_ptr(Y), _ptr(X), _ptr(maps[0]), _ptr(maps[1]), _ptr(maps[2]), _ptr(scale),
                                            N, Cc, H, W, _ptr(win1d), win1d.numel(), _ptr(gY), stream), "b2s_ssim_bwd")
            if ctx.needs_input_grad[0]:
                gX = torch.empty_like(X)
                _lib.check(lib.b2s_ssim_bwd(_ptr(X), _ptr(Y), _ptr(maps[3]), _ptr(maps[1]), _ptr(maps[2]), _ptr(scale),
                                            N, Cc, H, W, _ptr(win1d), win1d.numel(), _ptr(gX), stream), "b2s_ssim_bwd")
        return gX, gY, None, None, None, None, None


def ssim(
    X: Tensor,
    Y: Tensor,
    data_range: float = 255,
    size_average: bool = True,
    win_size: int = 11,
    win_sigma: float = 1.5,
    win: Optional[Tensor] = None,
    K: Union[Tuple[float, float], List[float]] = (0.01, 0.03),
    nonnegative_ssim: bool = False,
    mask: Optional[Tensor] = None,
) -> Tensor:
    """Same contract as the reference ``ssim`` (ssim.py:110-190) for ``(N, C, H, W)`` float32 CUDA images."""
    if not X.shape == Y.shape:
        raise ValueError(f"Input images should have the same dimensions, but got {X.shape} and {Y.shape}.")
    if mask is not None:
        if mask.dim() == 3:  # H, W, C
            mask = mask.permute(2, 0, 1).unsqueeze(0)
        elif mask.dim() != 4:
            raise ValueError(f"mask should be (H, W, C) or (N, C, H, W), got {mask.shape}")
        if tuple(torch.broadcast_shapes(mask.shape, X.shape)) != tuple(X.shape):
            raise AssertionError(f"Mask shape {mask.shape} should be the same as X shape {X.shape}")
    for d in range(len(X.shape) - 1, 1, -1):
        if X.shape[d] == 1 and len(X.shape) > 4:
            X, Y = X.squeeze(dim=d), Y.squeeze(dim=d)
    if len(X.shape) != 4:
        if len(X.shape) == 5:
            raise NotImplementedError("3-D (N, C, D, H, W) SSIM is not built (MTGS uses 2-D images)")
        raise ValueError(f"Input images should be 4-d or 5-d tensors, but got {X.shape}")
    if mask is not None and size_average is not True:
        raise AssertionError("per channel ssim is not available if mask exist")
    if win is not None:
        win_size = win.shape[-1]
    if not (win_size % 2 == 1):
        raise ValueError("Window size should be odd.")
    if not X.is_cuda or not Y.is_cuda:
        raise RuntimeError("mtgs_b200.ssim needs CUDA tensors (no CPU fallback path exists)")
    if X.dtype != torch.float32 or Y.dtype != torch.float32:
        raise TypeError("mtgs_b200.ssim is built for float32 images")
    N, Cc, H, W = X.shape
    if H < win_size or W < win_size or win_size > 15:
        raise NotImplementedError(f"window {win_size} on a {H}x{W} image is not built (needs win_size <= 15 <= H, W)")
    if win is None:
        win1d = _fspecial_gauss_1d(win_size, win_sigma).reshape(-1)
    else:
        win1d = win.reshape(-1, win.shape[-1])[0]  # the reference repeats one window over the channels
    win1d = win1d.to(device=X.device, dtype=torch.float32).contiguous()
    K1, K2 = K
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    Xc, Yc = X.contiguous(), Y.contiguous()
    mask_u8, strides = None, (0, 0)
    if mask is not None:
        m = mask.to(device=X.device)
        m = (m != 0) if m.dtype != torch.bool else m
        mn, mc = m.shape[0], m.shape[1]
        mask_u8 = m.to(torch.uint8).expand(mn, mc, H, W).contiguous()
        strides = (0 if mn == 1 and N > 1 else mc * H * W, 0 if mc == 1 and Cc > 1 else H * W)
    acc = _SSIMMap.apply(Xc, Yc, mask_u8, strides, win1d, C1, C2)
    Ho, Wo = H - win_size + 1, W - win_size + 1
    if mask is not None:
        tot = acc.sum(0)
        out = (tot[0] / tot[1]).to(torch.float32)  # masked_select(...).mean(); NaN for an empty mask, as the reference
        return torch.relu(out) if nonnegative_ssim else out
    per_channel = (acc[:, 0] / float(Ho * Wo)).to(torch.float32).view(N, Cc)
    if nonnegative_ssim:
        per_channel = torch.relu(per_channel)
    return per_channel.mean() if size_average else per_channel.mean(1)


class MaskedSSIM(torch.nn.Module):
    """Same constructor and ``forward(X, Y, mask=None)`` as the reference class (ssim.py:193-238)."""

    def __init__(self, data_range: float = 255, size_average: bool = True, win_size: int = 11, win_sigma: float = 1.5,
                 channel: int = 3, spatial_dims: int = 2, K: Union[Tuple[float, float], List[float]] = (0.01, 0.03),
                 nonnegative_ssim: bool = False) -> None:
        super().__init__()
        self.win_size = win_size
        self.win = _fspecial_gauss_1d(win_size, win_sigma).repeat([channel, 1] + [1] * spatial_dims)
        self.size_average = size_average
        self.data_range = data_range
        self.K = K
        self.nonnegative_ssim = nonnegative_ssim

    def forward(self, X: Tensor, Y: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        return ssim(X, Y, data_range=self.data_range, size_average=self.size_average, win=self.win, K=self.K,
                    nonnegative_ssim=self.nonnegative_ssim, mask=mask)
