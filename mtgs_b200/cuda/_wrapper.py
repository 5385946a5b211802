"""``spherical_harmonics`` -- drop-in for ``gsplat.cuda._wrapper.spherical_harmonics``.

Reference call sites: mtgs/scene_model/gaussian_model/vanilla_gaussian_splatting.py:16, 309-318;
multi_color_gaussian_splatting.py:12, 89-101; rigid_node.py:17, 238-253; deformable_node.py:16, 119-130.
Contract follows upstream gsplat v1.4.0: ``dirs [..., 3]`` (normalised inside), ``coeffs [..., K, 3]``,
optional boolean ``masks [...]``; returns ``[..., 3]`` WITHOUT the +0.5 (MTGS adds it and clamps).
Gradients: coeffs always, dirs when it requires grad.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from .. import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor]):
        lib = _lib.load()
        N, K = coeffs.shape[0], coeffs.shape[1]
        m8 = masks.to(torch.uint8).contiguous() if masks is not None else None
        alloc = torch.zeros if m8 is not None else torch.empty  # masked rows are not written by the kernel
        colors = alloc(N, 3, dtype=torch.float32, device=coeffs.device)
        with torch.cuda.device(coeffs.device):
            _lib.check(lib.b2s_sh_fwd(degree, _ptr(dirs), _ptr(coeffs), _ptr(m8), N, K, _ptr(colors), _stream()),
                       "b2s_sh_fwd")
        ctx.save_for_backward(dirs, coeffs, m8)
        ctx.degree = degree
        return colors

    @staticmethod
    def backward(ctx, v_colors: Tensor):
        lib = _lib.load()
        dirs, coeffs, m8 = ctx.saved_tensors
        N, K = coeffs.shape[0], coeffs.shape[1]
        v_colors = v_colors.contiguous()
        v_coeffs = torch.empty_like(coeffs)
        v_dirs = torch.empty_like(dirs) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(coeffs.device):
            _lib.check(lib.b2s_sh_bwd(ctx.degree, _ptr(dirs), _ptr(coeffs), _ptr(m8), N, K, _ptr(v_colors),
                                      _ptr(v_coeffs), _ptr(v_dirs), _stream()), "b2s_sh_bwd")
        return None, v_dirs, (v_coeffs if ctx.needs_input_grad[2] else None), None


def spherical_harmonics(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """Evaluate real spherical harmonics of degree ``degrees_to_use`` (0..4) for every direction."""
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    assert dirs.shape[:-1] == coeffs.shape[:-2], (dirs.shape, coeffs.shape)
    assert dirs.shape[-1] == 3, dirs.shape
    assert coeffs.shape[-1] == 3, coeffs.shape
    if masks is not None:
        assert masks.shape == dirs.shape[:-1], masks.shape
        masks = masks.reshape(-1)
    if not coeffs.is_cuda or not dirs.is_cuda:
        raise RuntimeError("mtgs_b200 kernels need CUDA tensors (no CPU fallback path exists)")
    if coeffs.dtype != torch.float32 or dirs.dtype != torch.float32:
        raise TypeError("spherical_harmonics expects float32 dirs and coeffs")
    batch = dirs.shape[:-1]
    K = coeffs.shape[-2]
    out = _SphericalHarmonics.apply(int(degrees_to_use), dirs.reshape(-1, 3).contiguous(),
                                    coeffs.reshape(-1, K, 3).contiguous(), masks)
    return out.reshape(batch + (3,))
