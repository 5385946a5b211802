"""One rank of the multi-GPU gradient-exchange check (launched by torchrun; see test_gpu_exchange.py::test_two_gpus).
Compares the fused exchange (peer stores inside the projection backward) with an NCCL all-reduce of plain backwards."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mtgs_b200 import scenes  # noqa: E402
from mtgs_b200.parallel import GradExchange  # noqa: E402
from mtgs_b200.rendering import rasterization  # noqa: E402

NAMES = ("means", "quats", "scales", "opacities", "colors")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, W, H, n_tail = 50_000, 640, 360, 123
    base = scenes.street(n=n, seed=21, width=W, height=H)
    cam = scenes.street(n=n, seed=21, width=W, height=H, camera=rank)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    w_c = torch.randn(1, H, W, 4, device=dev, generator=gen)
    w_a = torch.randn(1, H, W, 1, device=dev, generator=gen)

    def inputs():
        return {k: torch.tensor(base[k], device=dev).requires_grad_(True) for k in NAMES}

    def loss_of(t):
        r, a, _ = rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"],
                                torch.tensor(cam["viewmat"], device=dev)[None], torch.tensor(cam["K"], device=dev)[None],
                                W, H, packed=False, render_mode="RGB+ED", rasterize_mode="antialiased", absgrad=True)
        return (r * w_c).sum() + (a * w_a).sum()

    t = inputs()
    loss_of(t).backward()
    local_grads = {k: t[k].grad.clone() for k in NAMES}
    want = {k: v.clone() for k, v in local_grads.items()}
    n_shared = n - n_tail
    for k in NAMES:
        shared = want[k][:n_shared].contiguous()
        dist.all_reduce(shared, op=dist.ReduceOp.AVG)
        want[k][:n_shared] = shared

    ex = GradExchange(n_shared=n_shared, d_in=3, rows_cap=n)
    for step in range(3):
        t = inputs()
        with ex.active():
            loss_of(t).backward()
        ex.check()
        for k in NAMES:
            tol = 2e-4 * float(want[k].abs().max()) + 1e-9
            err = float((t[k].grad - want[k]).abs().max())
            assert err <= tol, (rank, step, k, err, tol)
    dist.barrier()
    ex.close()
    if rank == 0:
        print("EXCHANGE_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
