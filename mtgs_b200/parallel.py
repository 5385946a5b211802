"""Multi-GPU sharding of the hot path: one traversal camera per rank, all-reduce of shared-node gradients.

The reference has no working multi-GPU data path (DDP wiring at mtgs/scene_model/custom_pipeline.py:86-89 is
unused).  The scheme here is the one BASELINE.json's north_star names:
rendering is independent per camera (mtgs_scene_graph.py:548 asserts one camera per call), so every rank renders
its own traversal's camera over the replicated shared nodes; the only exchange step is ONE sum all-reduce per
step over a contiguous fp32 arena holding the gradients of the replicated tensors
(means 3 + scales 3 + quats 4 + opacities 1 + features_dc 3 = 14 floats = 56 B per shared Gaussian).

``torch.distributed`` is the plumbing (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).
Parameters' ``.grad`` are views INTO the arena, so backward writes straight into the communication buffer (no
pack/unpack copies) and the all-reduce can start as soon as the projection backward has run.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def traversal_of_rank(rank: int, world_size: int, n_traversals: int) -> List[int]:
    """Static traversal -> rank map (round robin when there are more traversals than ranks).
    Per-traversal tensors (colour residuals features_adapters[:, t], rigid nodes of t; reference
    multi_color_gaussian_splatting.py:53-87, rigid_node.py:87) stay rank-local and are never communicated."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return [t for t in range(n_traversals) if t % world_size == rank]


class SharedGradArena:
    """Contiguous gradient arena for the replicated (shared-node) parameters."""

    def __init__(self, params: Sequence[Tensor], process_group=None, average: bool = True):
        params = list(params)
        if not params:
            raise ValueError("no shared parameters")
        dev, dt = params[0].device, params[0].dtype
        for p in params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("shared parameters must live on one device with one dtype")
            if not p.is_leaf or not p.requires_grad:
                raise ValueError("shared parameters must be leaf tensors that require grad")
        self.params = params
        self.group = process_group
        self.average = average
        self.numel = sum(p.numel() for p in params)
        # 16-byte aligned slices so vectorised kernels can write the views directly
        self._offsets = []
        off = 0
        for p in params:
            self._offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.arena = torch.zeros(off, dtype=dt, device=dev)
        self._work = None
        self._stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.bind()

    def bind(self) -> None:
        """(Re)point every parameter's .grad at its arena slice (call again after densification replaced
        the parameter objects, reference vanilla_gaussian_splatting.py:512-515)."""
        for p, off in zip(self.params, self._offsets):
            p.grad = self.arena[off:off + p.numel()].view_as(p)

    @property
    def nbytes(self) -> int:
        return self.arena.numel() * self.arena.element_size()

    def zero_(self) -> None:
        self.arena.zero_()

    def all_reduce_async(self) -> None:
        """Launch the sum all-reduce of the whole arena on a side stream (overlaps with the rank-local
        backward work and optimizer steps that follow)."""
        if not dist.is_available() or not dist.is_initialized():
            return
        # NCCL averages inside the collective (no extra pass over the arena); gloo has no AVG
        self._avg_in_op = self.average and self._stream is not None and dist.get_backend(self.group) == "nccl"
        op = dist.ReduceOp.AVG if self._avg_in_op else dist.ReduceOp.SUM
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.arena.device))
            with torch.cuda.stream(self._stream):
                self._work = dist.all_reduce(self.arena, op=op, group=self.group, async_op=True)
        else:
            self._work = dist.all_reduce(self.arena, op=op, group=self.group, async_op=True)

    def wait(self) -> None:
        if self._work is not None:
            self._work.wait()
            self._work = None
            if self._stream is not None:
                torch.cuda.current_stream(self.arena.device).wait_stream(self._stream)
        if self.average and not getattr(self, "_avg_in_op", False) and dist.is_available() and dist.is_initialized():
            self.arena.div_(dist.get_world_size(self.group))

    def all_reduce(self) -> None:
        self.all_reduce_async()
        self.wait()


def sync_densification_stats(grad_norm_sum: Tensor, vis_counts: Tensor, max_radii: Tensor, group=None) -> None:
    """Keep the per-Gaussian densification statistics of the shared nodes in lock-step across ranks before
    ``refinement_after`` (reference vanilla_gaussian_splatting.py:448-474, 476-577): sums for the accumulated
    screen-space gradient norm and visibility count, max for the largest 2-D radius."""
    if not dist.is_available() or not dist.is_initialized():
        return
    dist.all_reduce(grad_norm_sum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(vis_counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(max_radii, op=dist.ReduceOp.MAX, group=group)


def seed_everything_identically(seed: int, step: int) -> torch.Generator:
    """Identical RNG stream on every rank for the split/duplicate sampling of densification
    (reference vanilla_gaussian_splatting.py:642, 687 draw torch.randn unsynchronised)."""
    g = torch.Generator()
    g.manual_seed(int(seed) * 1_000_003 + int(step))
    return g
