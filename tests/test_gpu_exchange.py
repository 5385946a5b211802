"""Fused projection-backward + cross-rank gradient exchange (b2s_project_bwd_exchange).

Single-GPU tests play `world` ranks from one process (GradExchange.local_ranks: every rank's buffers live on the one
device and the peer pointer tables are wired to each other), which exercises the same kernels, slot / shard / arena
addressing and flag protocol as the multi-GPU run.  The real 2-GPU run (CUDA IPC + NVLink peer stores) is
tests/exchange_worker.py, launched by test_two_gpus when the box has two devices."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from mtgs_b200 import scenes

pytestmark = pytest.mark.gpu
NAMES = ("means", "quats", "scales", "opacities", "colors")


def _inputs(s, dev):
    t = {k: torch.tensor(s[k], dtype=torch.float32, device=dev).requires_grad_(True) for k in NAMES}
    return t


def _loss(t, cam, W, H, w_c, w_a):
    from mtgs_b200.rendering import rasterization
    r, a, _ = rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"],
                            torch.tensor(cam["viewmat"], device=w_c.device)[None], torch.tensor(cam["K"], device=w_c.device)[None],
                            W, H, packed=False, render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)
    return (r * w_c).sum() + (a * w_a).sum()


@pytest.mark.parametrize("world,n,n_tail,d_in,average,xcol", [(1, 1000, 0, 3, True, True), (2, 3001, 0, 3, True, True),
                                                              (4, 5000, 37, 3, False, True), (3, 2049, 5, 6, True, True),
                                                              (8, 20000, 0, 3, True, True), (2, 3001, 0, 3, True, False),
                                                              (3, 2049, 5, 6, False, False)])
def test_exchange_matches_sum_of_plain_backwards(cuda_device, world, n, n_tail, d_in, average, xcol):
    from mtgs_b200.parallel import GradExchange
    dev = cuda_device
    W, H = 320, 192
    cams = [scenes.street(n=n, seed=11, width=W, height=H, d_in=d_in, camera=r) for r in range(world)]
    gen = torch.Generator(device=dev).manual_seed(5)
    w_c = torch.randn(1, H, W, d_in + 1, device=dev, generator=gen)
    w_a = torch.randn(1, H, W, 1, device=dev, generator=gen)
    n_shared = n - n_tail
    # plain backward per rank
    plain = []
    for r in range(world):
        t = _inputs(cams[0], dev)
        _loss(t, cams[r], W, H, w_c, w_a).backward()
        plain.append({k: t[k].grad.detach().clone() for k in NAMES})
    scale = 1.0 / world if average else 1.0
    # exchange path: phase 1 of every rank, then the reduce / broadcast / wait phases
    exs = GradExchange.local_ranks(world, n_shared, d_in, rows_cap=n, average=average, device=dev, exchange_colors=xcol)
    try:
        local_colors = []
        for r in range(world):
            t = _inputs(cams[0], dev)
            with exs[r].active():
                _loss(t, cams[r], W, H, w_c, w_a).backward()
            local_colors.append(t["colors"].grad.detach().clone())
        GradExchange.finish_all(exs)
        torch.cuda.synchronize()
        for r in range(world):
            exs[r].check()
            got = exs[r].grad_views(n)
            if not xcol:  # view-dependent colours: the gradient stays local and unscaled
                assert "colors" not in got
                tol = 2e-4 * float(plain[r]["colors"].abs().max()) + 1e-9
                assert float((local_colors[r] - plain[r]["colors"]).abs().max()) <= tol
            for k in NAMES:
                if k == "colors" and not xcol:
                    continue
                want = sum(p[k][:n_shared] for p in plain) * scale
                g = got[k][:n_shared]
                tol = 2e-4 * float(want.abs().max()) + 1e-9
                assert float((g - want).abs().max()) <= tol, (k, r, float((g - want).abs().max()), tol)
                if n_tail:
                    wt, gt = plain[r][k][n_shared:], got[k][n_shared:]
                    tol = 2e-4 * float(plain[r][k].abs().max()) + 1e-9
                    assert float((gt - wt).abs().max()) <= tol, (k, "tail", r)
    finally:
        for e in exs:
            e.close()


def test_exchange_is_repeatable_across_steps(cuda_device):
    """Epoch-stamped flags: several steps through the same buffers without any reset."""
    from mtgs_b200.parallel import GradExchange
    dev = cuda_device
    world, n, W, H = 2, 4000, 320, 192
    cams = [scenes.street(n=n, seed=3, width=W, height=H, camera=r) for r in range(world)]
    w_c = torch.ones(1, H, W, 4, device=dev)
    w_a = torch.ones(1, H, W, 1, device=dev)
    exs = GradExchange.local_ranks(world, n, 3, device=dev)
    try:
        ref = None
        for step in range(3):
            for r in range(world):
                t = _inputs(cams[0], dev)
                with exs[r].active():
                    _loss(t, cams[r], W, H, w_c, w_a).backward()
            GradExchange.finish_all(exs)
            torch.cuda.synchronize()
            exs[0].check()
            cur = exs[1].grad_views(n)["means"].clone()
            if ref is None:
                ref = cur
            assert float((cur - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
            assert torch.equal(exs[0].grad_views(n)["quats"], exs[1].grad_views(n)["quats"])
    finally:
        for e in exs:
            e.close()


def test_missing_peer_times_out_instead_of_hanging(cuda_device):
    from mtgs_b200.parallel import GradExchange
    dev = cuda_device
    n, W, H = 600, 128, 96
    cam = scenes.street(n=n, seed=3, width=W, height=H)
    exs = GradExchange.local_ranks(2, n, 3, device=dev)
    try:
        t = _inputs(cam, dev)
        with exs[0].active():
            _loss(t, cam, W, H, torch.ones(1, H, W, 4, device=dev), torch.ones(1, H, W, 1, device=dev)).backward()
        exs[0].launch(2)  # reduce phase of rank 0; rank 1 never ran its phase 1
        torch.cuda.synchronize()
        with pytest.raises(RuntimeError, match="timed out"):
            exs[0].check()
    finally:
        for e in exs:
            e.close()


def test_two_gpus(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(root, "tests", "exchange_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0 and "EXCHANGE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
