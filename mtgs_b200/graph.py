"""A whole training step (forward + loss + backward) as ONE CUDA graph.

The forward of ``rasterization`` never waits for the host once list capacities are known (``rendering.py``, capacity
mode), so a step on static tensors is a fixed sequence of ~30 kernel launches and can be captured with
``torch.cuda.graph`` and replayed with a single launch: the ~2 us gaps between dependent kernels and all host-side launch
work disappear.  CUDA streams and graphs, not a tracing compiler: nothing is rewritten, the same kernels run.

    step = GraphedStep(lambda: run_one_step(params))   # eager warm-up (learns the capacities), then capture
    for _ in range(k):
        step.replay()
    step.check()            # synchronises; raises if a replay outgrew the captured list capacities

Inside the graph the host cannot compare this frame's list sizes with the capacities, so the tile-list build only has its
device-side overflow word (a level that does not fit makes every later kernel of that forward return at once -- nothing is
written out of bounds, but the step's result is invalid).  ``check()`` / ``ok()`` read those words after the fact; a caller
whose camera or Gaussian count changes recaptures (``recapture()``), which also grows the capacities.
"""
from __future__ import annotations

from typing import Callable, List

import torch

from . import rendering


class GraphedStep:
    def __init__(self, fn: Callable[[], object], warmup: int = 3, pool=None):
        self.fn = fn
        self.warmup = int(warmup)
        self.pool = pool
        self.graph = None
        self.out = None
        self._captured: List[tuple] = []
        self.recapture()

    def recapture(self) -> None:
        """Eager warm-up (lazy module loading, allocator steady state, list capacities), then capture -- both on ONE
        side stream: autograd's AccumulateGrad nodes remember the stream they were created on, and a node that lives
        on another stream than the capture stream invalidates the capture.  The caller must not keep tensors of earlier
        eager steps alive (they pin the old autograd graph and with it AccumulateGrad nodes of the default stream)."""
        import gc
        self.out = None
        gc.collect()
        torch.cuda.synchronize()
        if getattr(self, "stream", None) is None:
            self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, self.warmup)):
                out = self.fn()
                del out
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        gc.collect()
        rendering._CAPTURED.clear()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: only the capturing thread is policed; background threads of the process (NCCL watchdog,
        # pinned-memory allocator) keep making CUDA calls during the capture, which the default "global" mode forbids
        try:
            with torch.cuda.graph(self.graph, pool=self.pool, stream=self.stream, capture_error_mode="thread_local"):
                self.out = self.fn()
        except Exception as e:
            first = e
            while first.__context__ is not None:  # the error that invalidated the capture, not capture_end's
                first = first.__context__
            if first is not e:
                import traceback
                where = "".join(traceback.format_tb(first.__traceback__)[-6:])
                raise RuntimeError(f"CUDA graph capture failed: {type(first).__name__}: {first}\n{where}") from e
            raise
        self._captured = list(rendering._CAPTURED)
        rendering._CAPTURED.clear()

    def replay(self):
        self.graph.replay()
        return self.out

    def ok(self) -> bool:
        """Synchronises and tells whether the LAST replay stayed inside the captured capacities; the capacities
        remembered by ``rendering`` grow to what that replay needed, so a ``recapture()`` afterwards fits."""
        torch.cuda.synchronize()
        good = True
        for totals, caps, key in self._captured:
            t = [int(v) for v in totals.tolist()]
            if t[5] != 0 or any(t[i] > caps[i] for i in range(4)):
                good = False
            cur = rendering._CAPACITY.setdefault(key, [0, 0, 0, 0])
            for i in range(4):
                cur[i] = max(cur[i], int(t[i] * rendering.CAPACITY_HEADROOM) + 4096)
        return good

    def check(self) -> None:
        if not self.ok():
            raise RuntimeError("a replayed step outgrew the tile-list capacities it was captured with; its result is "
                               "invalid -- call recapture() (the capacities have been raised)")
