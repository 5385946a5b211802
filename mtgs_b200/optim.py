"""Fused multi-tensor Adam and densification primitives.

Reference: nerfstudio builds ONE ``torch.optim.Adam`` per parameter group and MTGS has one group per (node, attribute)
(``mtgs/scene_model/custom_trainer.py:115-136``), stepped one after the other; the refinement code edits those
optimizers' states tensor by tensor (``gaussian_model/vanilla_gaussian_splatting.py:392-446``) and updates the
densification statistics with boolean indexing (``:448-474``).  Here:

* ``FusedAdam`` -- same update rule and ``state`` / ``param_groups`` layout as ``torch.optim.Adam`` (so the reference's
  ``remove_from_optim`` / ``dup_in_optim`` style surgery keeps working), but ``step()`` is ONE kernel launch over every
  tensor of every group (``csrc/optim.cu::k_adam_multi``).
* ``accumulate_densify_stats`` -- the ``after_train`` update in one pass.
* ``compact_rows`` / ``append_rows`` -- order-preserving cull and split / duplicate appends for parameters and their
  Adam moments together.

CUDA float32 tensors only (no CPU / PyTorch fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .rendering import _need_cuda, _ptr, _stream


class _AdamTensor(C.Structure):  # mirrors B2sAdamTensor in include/b200splat.h
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_longlong),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("step_size", C.c_float), ("bias_correction2_sqrt", C.c_float),
                ("one_minus_beta1", C.c_float), ("one_minus_beta2", C.c_float), ("_pad", C.c_float)]


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` semantics (no amsgrad / maximize), one kernel launch per ``step()`` for all groups.

    ``state[p]`` holds ``step`` (python int), ``exp_avg``, ``exp_avg_sq`` like torch's Adam, and schedulers may change
    ``param_groups[i]["lr"]`` between steps (nerfstudio's ExponentialDecay does)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._chunk = int(_lib.load().b2s_adam_chunk())
        self._layout_key = None
        self._chunk_tensor = self._chunk_start = None
        self._ring, self._ring_pos = [], 0

    def _tensors(self):
        out = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                out.append((p, group))
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        items = self._tensors()
        if not items:
            return loss
        lib = _lib.load()
        dev = items[0][0].device
        _need_cuda(*[p for p, _ in items])
        descs = (_AdamTensor * len(items))()
        key = []
        for i, (p, group) in enumerate(items):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("FusedAdam needs contiguous float32 parameters")
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] = int(st["step"]) + 1
            b1, b2 = group["betas"]
            d = descs[i]
            d.p, d.g, d.m, d.v, d.n = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
            d.lr, d.beta1, d.beta2, d.eps, d.weight_decay = group["lr"], b1, b2, group["eps"], group["weight_decay"]
            d.step_size = group["lr"] / (1.0 - b1 ** st["step"])
            d.bias_correction2_sqrt = (1.0 - b2 ** st["step"]) ** 0.5
            d.one_minus_beta1, d.one_minus_beta2 = 1.0 - b1, 1.0 - b2
            key.append(p.numel())
            st["_grad_keepalive"] = g
        key = tuple(key)
        if key != self._layout_key:  # chunk table only changes when tensor sizes do (densification)
            ct, cs = [], []
            for i, n in enumerate(key):
                for s in range(0, n, self._chunk):
                    ct.append(i)
                    cs.append(s)
            self._chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
            self._chunk_start = torch.tensor(cs, dtype=torch.int64, device=dev)
            # the descriptor table changes every step (bias corrections, learning rates) and travels through pinned
            # memory asynchronously: a ring of staging slots, each guarded by an event, so that a slot is never
            # rewritten before its copy has been executed (the host runs ahead of the device)
            nbytes = C.sizeof(_AdamTensor) * len(key)
            self._ring = [(torch.empty(nbytes, dtype=torch.uint8).pin_memory(),
                           torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.cuda.Event()) for _ in range(4)]
            self._ring_pos = 0
            self._layout_key = key
        host, desc_dev, ev = self._ring[self._ring_pos]
        self._ring_pos = (self._ring_pos + 1) % len(self._ring)
        ev.synchronize()  # no-op unless the device is more than a ring's worth of steps behind
        C.memmove(host.data_ptr(), C.addressof(descs), C.sizeof(descs))
        with torch.cuda.device(dev):
            desc_dev.copy_(host, non_blocking=True)
            _lib.check(lib.b2s_adam_multi(_ptr(desc_dev), _ptr(self._chunk_tensor), _ptr(self._chunk_start),
                                          int(self._chunk_tensor.numel()), _stream()), "b2s_adam_multi")
            ev.record()
        for p, _ in items:
            self.state[p].pop("_grad_keepalive", None)
        return loss


def accumulate_densify_stats(grad2d: Tensor, radii: Tensor, width: int, height: int, xys_grad_norm: Tensor,
                             vis_counts: Tensor, max_2dsize: Tensor) -> None:
    """In place: for ``radii > 0``: ``xys_grad_norm += ||grad2d * (W, H) / 2||``, ``vis_counts += 1``, ``max_2dsize =
    max(max_2dsize, radii)`` -- mtgs_scene_graph.py:1171-1178 + vanilla_gaussian_splatting.py:455-474 in one pass.
    ``grad2d``: ``info["means2d"].absgrad[0]`` (or ``.grad[0]``), [N, 2] (row stride 2 or 4)."""
    _need_cuda(grad2d, radii, xys_grad_norm, vis_counts, max_2dsize)
    lib = _lib.load()
    N = radii.numel()
    g = grad2d.reshape(N, 2)
    if not (g.stride(1) == 1 and g.stride(0) in (2, 4)):
        g = g.contiguous()
    r = radii.reshape(-1)
    r = r if r.dtype == torch.int32 and r.is_contiguous() else r.to(torch.int32).contiguous()
    for t in (xys_grad_norm, vis_counts, max_2dsize):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != N:
            raise ValueError("statistics must be contiguous float32 tensors with one entry per Gaussian")
    with torch.cuda.device(g.device):
        _lib.check(lib.b2s_densify_stats(_ptr(g), int(g.stride(0)), _ptr(r), N, int(width), int(height),
                                         _ptr(xys_grad_norm), _ptr(vis_counts), _ptr(max_2dsize), _stream()),
                   "b2s_densify_stats")


def compact_rows(tensors: Sequence[Tensor], keep: Tensor) -> List[Tensor]:
    """``[t[keep] for t in tensors]`` for a boolean ``keep`` over dim 0, order preserved: one scan of the mask, one
    gather per tensor (parameters and their Adam moments share the scan).  One host read (the kept count)."""
    _need_cuda(keep, *tensors)
    lib = _lib.load()
    N = keep.numel()
    k8 = (keep.reshape(-1) != 0).to(torch.uint8).contiguous() if keep.dtype != torch.uint8 else keep.reshape(-1).contiguous()
    dev = k8.device
    wsb = int(lib.b2s_mask_workspace_bytes(N))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    total = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.b2s_mask_scan(_ptr(k8), N, _ptr(ws), wsb, _ptr(total), _stream()), "b2s_mask_scan")
        n_keep = int(total.item())
        out = []
        for t in tensors:
            if t.shape[0] != N:
                raise ValueError(f"tensor with {t.shape[0]} rows, mask with {N}")
            if t.dtype != torch.float32:
                out.append(t[keep.reshape(-1).bool()])
                continue
            src = t.contiguous()
            rf = src.numel() // max(N, 1)
            dst = torch.empty((n_keep,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
            if N > 0 and n_keep > 0 and rf > 0:
                _lib.check(lib.b2s_mask_gather_rows(_ptr(k8), N, _ptr(ws), _ptr(src), _ptr(dst), rf, _stream()),
                           "b2s_mask_gather_rows")
            out.append(dst)
    return out


def append_rows(tensor: Tensor, new_rows: Optional[Tensor] = None, n_zero_rows: int = 0) -> Tensor:
    """``cat([tensor, new_rows])`` or ``cat([tensor, zeros(n_zero_rows, ...)])`` (split / duplicate appends and the
    zero optimizer state of the new Gaussians, vanilla_gaussian_splatting.py:418-440, 512-515)."""
    if new_rows is None:
        new_rows = torch.zeros((n_zero_rows,) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    return torch.cat([tensor, new_rows], dim=0)


def cull_optimizer_rows(optimizer: torch.optim.Optimizer, params: Dict[str, torch.nn.Parameter], keep: Tensor
                        ) -> Dict[str, torch.nn.Parameter]:
    """``remove_from_all_optim`` + the parameter re-creation of ``cull_gaussians`` (vanilla_gaussian_splatting.py:
    392-412, 620-621) for one optimizer holding every attribute of a node: parameters and both Adam moments are
    compacted with ONE mask scan; returns the new parameters (already registered in ``optimizer``)."""
    names = list(params)
    tensors, slots = [], []
    for n in names:
        p = params[n]
        tensors.append(p.detach())
        slots.append((n, "p"))
        st = optimizer.state.get(p, {})
        for k in ("exp_avg", "exp_avg_sq"):
            if k in st:
                tensors.append(st[k])
                slots.append((n, k))
    outs = compact_rows(tensors, keep)
    new_params: Dict[str, torch.nn.Parameter] = {}
    new_state: Dict[str, dict] = {n: {} for n in names}
    for (n, k), t in zip(slots, outs):
        if k == "p":
            new_params[n] = torch.nn.Parameter(t, requires_grad=params[n].requires_grad)
        else:
            new_state[n][k] = t
    for group in optimizer.param_groups:
        for i, p in enumerate(group["params"]):
            for n in names:
                if p is params[n]:
                    st = optimizer.state.pop(p, {})
                    st.update(new_state[n])
                    group["params"][i] = new_params[n]
                    if st:
                        optimizer.state[new_params[n]] = st
    if isinstance(optimizer, FusedAdam):
        optimizer._layout_key = None
    return new_params
