"""Multi-GPU sharding of the hot path: one traversal camera per rank, exchange of the shared-node gradients.

Two mechanisms: ``GradExchange`` (bottom of this file) fuses the exchange INTO the projection backward over NVLink
peer memory and is what ``bench.py --gpus N`` measures; ``SharedGradArena`` is the library all-reduce after the
backward that it replaces for the geometry gradients -- still the right tool for leaves whose Jacobian differs per
rank (SH coefficients) and the baseline of ``bench.py --exchange nccl``.

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


# ------------------------------------------------------------------------------------------------
# Fused gradient exchange (projection backward + reduce-scatter / all-gather over NVLink peer memory)
# ------------------------------------------------------------------------------------------------
_ACTIVE_EXCHANGE = None


def current_exchange():
    """The GradExchange that ``rasterization``'s backward should route the gradients of its Gaussian inputs
    through (None = plain single-GPU backward)."""
    return _ACTIVE_EXCHANGE


class _DevMem:
    """``__cuda_array_interface__`` view of library-owned device memory (torch.as_tensor aliases it)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class GradExchange:
    """Sum (or mean) over ranks of the gradients w.r.t. the rasterizer's Gaussian inputs, produced INSIDE the
    projection backward instead of by an all-reduce after it (``include/b200splat.h``: b2s_project_bwd_exchange;
    kernels in csrc/project.cu and csrc/exchange.cu).

    Rows ``[0, n_shared)`` of means / quats / scales / opacities / colors are the replicated shared-node Gaussians
    (identical values on every rank); rows beyond are rank-local (the vehicles of the rank's own traversal,
    reference rigid_node.py:87, 259-261) and stay on the GPU.  Because the activations MTGS applies to means,
    scales, quats and opacities before the rasterizer (vanilla_gaussian_splatting.py:299-307) are elementwise
    functions of replicated parameters, reducing at the rasterizer inputs and then back-propagating the reduced
    gradient locally gives the same leaf gradients as reducing at the leaves.  That argument does NOT hold for
    colours evaluated from spherical harmonics per camera (vanilla_gaussian_splatting.py:309-318): their Jacobian
    (the SH basis of the rank's own view directions) differs from rank to rank.  Pass ``exchange_colors=False``
    then: colour gradients stay local and the SH coefficient gradients are reduced at their leaves
    (``SharedGradArena``); ``exchange_colors=True`` is for replicated, view-independent colour inputs.

    Usage (one process per GPU, ``torch.distributed`` initialised with NCCL)::

        ex = GradExchange(n_shared=N, d_in=3, rows_cap=N)
        render, alpha, info = rasterization(...)
        with ex.active():
            loss.backward()          # .grad of the inputs already holds the mean over all ranks

    Contract (the same as for a library collective): every rank must call the exchange once per step, in lock-step.
    The waits block (up to ``timeout_s``, default 120 s); a timeout is loud -- the shared rows come back as NaN, the
    next ``launch`` / ``check()`` raises.  ``n_shared`` may change between steps (densification / pruning of the
    shared nodes, reference vanilla_gaussian_splatting.py:476-577): call ``set_n_shared`` on every rank after
    ``refinement_after``; it validates that all ranks agree.  One exchange per ``active()`` context: a second
    ``rasterization`` backward under the same context raises instead of overwriting the first one's gradients.

    By default the gradients handed to autograd are CLONES of the arena slices.  ``zero_copy=True`` hands out views of
    the arena instead (saves one 56 B / Gaussian copy): they are valid until the next exchange only, so the caller
    must consume them (optimizer step) before the next ``backward`` and must not keep ``.grad`` alive across steps
    (``zero_grad(set_to_none=True)``).
    """

    MAX_WORLD = 8

    def __init__(self, n_shared: int, d_in: int, rows_cap: Optional[int] = None, group=None, average: bool = True,
                 device: Optional[torch.device] = None, exchange_colors: bool = True, zero_copy: bool = False,
                 timeout_s: float = 120.0, device_epoch: bool = True, _local: Optional[tuple] = None):
        from . import _lib
        import ctypes as C
        self._lib = lib = _lib.load()
        self._check = _lib.check
        if _local is not None:            # test harness: several ranks played by one process on one GPU
            self.world, self.rank = _local
        elif dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        if self.world > self.MAX_WORLD:
            raise NotImplementedError(f"GradExchange supports up to {self.MAX_WORLD} ranks (one NVSwitch domain)")
        self.group = group
        self.n_shared, self.d_in = int(n_shared), int(d_in)
        self.exchange_colors = bool(exchange_colors)
        self.rows_cap = (int(rows_cap if rows_cap is not None else n_shared) + 3) // 4 * 4
        if self.rows_cap < self.n_shared:
            raise ValueError("rows_cap < n_shared")
        self.scale = 1.0 / self.world if average else 1.0
        self.zero_copy = bool(zero_copy)
        self.timeout_s = float(timeout_s)
        # device_epoch: the step counter lives on the device (incremented by the step's first signal kernel), so no
        # launch argument changes between steps and a step can be replayed from a CUDA graph
        self.device_epoch = bool(device_epoch)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.shard = int(lib.b2s_exchange_shard_rows(self.n_shared, self.world))
        # the staging buffer is sized for the largest shard n_shared can grow to (rows_cap)
        self.shard_cap = int(lib.b2s_exchange_shard_rows(self.rows_cap, self.world))
        self.F = 11 + (self.d_in if self.exchange_colors else 0)
        self._sizes = (self.world * self.F * self.shard_cap * 4, self.F * self.rows_cap * 4, 256)
        self._own = []
        with torch.cuda.device(self.device):
            for nbytes in self._sizes:
                p = C.c_void_p()
                self._check(lib.b2s_peer_alloc(nbytes, C.byref(p)), "b2s_peer_alloc")
                self._own.append(int(p.value))
        self.stage_ptrs = [0] * self.world
        self.arena_ptrs = [0] * self.world
        self.flag_ptrs = [0] * self.world
        self.stage_ptrs[self.rank], self.arena_ptrs[self.rank], self.flag_ptrs[self.rank] = self._own
        self._imported = []
        self.arena = torch.as_tensor(_DevMem(self._own[1], self.F * self.rows_cap), device=self.device)
        # status word in pinned host memory: the kernels write it through the unified address space, the host reads
        # it without synchronising (checked at the start of every launch)
        self.status = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.epoch = 0
        self._launches_in_ctx = 0
        if _local is None and self.world > 1:
            self._rendezvous()

    # -- set-up ---------------------------------------------------------------------------------
    def _rendezvous(self) -> None:
        """Exchange CUDA IPC handles of the three buffers and map every peer's copies."""
        import ctypes as C
        mine = []
        with torch.cuda.device(self.device):
            for p in self._own:
                buf = C.create_string_buffer(64)
                self._check(self._lib.b2s_ipc_export(C.c_void_p(p), buf), "b2s_ipc_export")
                mine.append(bytes(buf.raw))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, mine, group=self.group)
        failure = None
        with torch.cuda.device(self.device):
            for r, handles in enumerate(gathered):
                if r == self.rank or failure is not None:
                    continue
                ptrs = []
                try:
                    for h in handles:
                        q = C.c_void_p()
                        self._check(self._lib.b2s_ipc_import(C.create_string_buffer(h, 64), C.byref(q)),
                                    "b2s_ipc_import")
                        ptrs.append(int(q.value))
                        self._imported.append(int(q.value))
                    self.stage_ptrs[r], self.arena_ptrs[r], self.flag_ptrs[r] = ptrs
                except RuntimeError as e:  # keep the collectives below matched on every rank, then report
                    failure = e
        dist.barrier(group=self.group)
        if failure is not None:
            raise failure

    @staticmethod
    def local_ranks(world: int, n_shared: int, d_in: int, rows_cap: Optional[int] = None, average: bool = True,
                    device: Optional[torch.device] = None, exchange_colors: bool = True) -> List["GradExchange"]:
        """``world`` exchanges living in ONE process on ONE GPU, wired to each other's buffers (tests only:
        every rank's backward runs phase 1, then ``finish_all`` plays the remaining phases, all on one stream)."""
        exs = [GradExchange(n_shared, d_in, rows_cap, average=average, device=device, exchange_colors=exchange_colors,
                            _local=(world, r))
               for r in range(world)]
        for a in exs:
            for b in exs:
                a.stage_ptrs[b.rank], a.arena_ptrs[b.rank], a.flag_ptrs[b.rank] = b._own
        for a in exs:
            a._phases = 1
        return exs

    _phases = 15

    def close(self) -> None:
        if not getattr(self, "_own", None) and not getattr(self, "_imported", None):
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._imported:
                self._lib.b2s_ipc_close(p)
            for p in self._own:
                self._lib.b2s_peer_free(p)
        self._imported, self._own = [], []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def set_n_shared(self, n_shared: int) -> None:
        """Number of replicated rows from now on (after densification / pruning of the shared nodes).  Collective:
        every rank must call it with the same value (validated here)."""
        n_shared = int(n_shared)
        if n_shared < 0 or n_shared > self.rows_cap:
            raise ValueError(f"n_shared={n_shared} outside [0, rows_cap={self.rows_cap}]; build a larger GradExchange")
        if self.world > 1 and dist.is_available() and dist.is_initialized() and self._phases == 15:
            t = torch.tensor([n_shared, -n_shared], device=self.device, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            if int(t[0]) != n_shared or int(-t[1]) != n_shared:
                raise RuntimeError(f"ranks disagree on n_shared (this rank {n_shared}, range "
                                   f"[{int(-t[1])}, {int(t[0])}])")
        self.n_shared = n_shared
        self.shard = int(self._lib.b2s_exchange_shard_rows(self.n_shared, self.world))

    # -- per step -------------------------------------------------------------------------------
    def active(self):
        ex = self

        class _Ctx:
            def __enter__(self_inner):
                global _ACTIVE_EXCHANGE
                self_inner.prev = _ACTIVE_EXCHANGE
                _ACTIVE_EXCHANGE = ex
                ex._launches_in_ctx = 0
                return ex

            def __exit__(self_inner, *exc):
                global _ACTIVE_EXCHANGE
                _ACTIVE_EXCHANGE = self_inner.prev
                return False

        return _Ctx()

    def grad_views(self, N: int) -> Dict[str, Tensor]:
        """Views of the arena holding the gradients of an N-row call (rows < n_shared: reduced over ranks)."""
        c, a = self.rows_cap, self.arena
        out = {"means": a[0:3 * N].view(N, 3), "quats": a[3 * c:3 * c + 4 * N].view(N, 4),
               "scales": a[7 * c:7 * c + 3 * N].view(N, 3), "opacities": a[10 * c:10 * c + N]}
        if self.exchange_colors:
            out["colors"] = a[11 * c:11 * c + self.d_in * N].view(N, self.d_in)
        return out

    def _ptr_array(self, ptrs):
        import ctypes as C
        return (C.c_ulonglong * self.MAX_WORLD)(*(list(ptrs) + [0] * (self.MAX_WORLD - len(ptrs))))

    def launch(self, phases: int, args: Optional[tuple] = None) -> None:
        """Enqueue the selected phases on the current stream (``args`` = the projection-backward operands)."""
        import ctypes as C
        if int(self.status[0]) != 0:  # pinned host word: no synchronisation
            raise RuntimeError(f"gradient exchange timed out waiting for a peer in an earlier step (phase code "
                               f"{int(self.status[0])}); its shared gradients were NaN-filled")
        if phases & 1:
            if self._phases == 15 and self._launches_in_ctx > 0:
                raise RuntimeError("a second rasterization backward ran under the same GradExchange.active() context: "
                                   "its gradients would overwrite the first one's (one exchange per step)")
            self._launches_in_ctx += 1
            self.epoch += 1
        if args is None:
            args = self._last_args
        self._last_args = args
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self._check(self._lib.b2s_project_bwd_exchange(
            *args, self.n_shared, int(self.exchange_colors), self.world, self.rank, self.rows_cap, self.scale,
            0 if self.device_epoch else self.epoch, phases, self.timeout_s,
            self._ptr_array(self.stage_ptrs), self._ptr_array(self.arena_ptrs), self._ptr_array(self.flag_ptrs),
            C.c_void_p(self.status.data_ptr()), stream),
            "b2s_project_bwd_exchange")

    @staticmethod
    def finish_all(exs: Sequence["GradExchange"]) -> None:
        """Test harness (see ``local_ranks``): the reduce / broadcast phase of every rank, then every rank's wait."""
        for e in exs:
            e.launch(2 | 4)
        for e in exs:
            e.launch(8)

    def check(self) -> None:
        """Synchronise and raise if a peer never arrived (a spin loop timed out)."""
        torch.cuda.synchronize(self.device)
        code = int(self.status[0])
        if code:
            raise RuntimeError(f"gradient exchange timed out waiting for a peer (phase code {code})")


# ------------------------------------------------------------------------------------------------
# Camera sampling per rank (SURVEY 8e "Partitioning")
# ------------------------------------------------------------------------------------------------
class RankTraversalSampler:
    """Balanced multi-traversal camera sampler restricted to the traversals owned by one rank.

    Reference: ``MultiTraversalBalancedSampler`` (mtgs/dataset/utils/sampler.py:27-58) first draws a not-yet-seen
    traversal, then a not-yet-seen image of that traversal, refilling either pool when it runs empty.  With one
    traversal camera per GPU every rank runs the same two-level draw over ITS traversals
    (``traversal_of_rank``), so the union over ranks still visits every traversal equally often and every image
    of a traversal once per pass.  With ``world_size == 1`` the draw order is the reference's, call for call, under
    the same ``random`` seed (tests/test_parallel_cpu.py checks that against indices the reference produced).

    ``travel_ids[i]`` is the traversal of image ``i`` (``dataparser_outputs.travel_ids`` in the reference)."""

    def __init__(self, travel_ids: Sequence[int], rank: int = 0, world_size: int = 1, rng=None):
        import random as _random
        self._rng = rng if rng is not None else _random
        ids = [int(t) for t in travel_ids]
        all_traversals = set(ids)
        ordered = list(all_traversals)  # same construction as the reference (set -> list)
        mine = set(traversal_of_rank(rank, world_size, len(ordered)))
        self.traversals = [t for k, t in enumerate(ordered) if k in mine]
        if not self.traversals:
            raise ValueError(f"rank {rank} of {world_size} owns no traversal ({len(ordered)} traversals in the data)")
        self.traversal_indices = {t: [i for i, x in enumerate(ids) if x == t] for t in self.traversals}
        self.traversal_counts = {t: len(v) for t, v in self.traversal_indices.items()}
        self.unseen_traversals = list(self.traversals)
        self.unseen_per_traversal_images = {t: list(v) for t, v in self.traversal_indices.items()}

    def get_next_traversal(self) -> int:
        t = self.unseen_traversals.pop(self._rng.randint(0, len(self.unseen_traversals) - 1))
        if not self.unseen_traversals:
            self.unseen_traversals = list(self.traversals)
        return t

    def get_next_image_idx(self) -> int:
        t = self.get_next_traversal()
        pool = self.unseen_per_traversal_images[t]
        idx = pool.pop(self._rng.randint(0, len(pool) - 1))
        if not pool:
            self.unseen_per_traversal_images[t] = list(self.traversal_indices[t])
        return idx
