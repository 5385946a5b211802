"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shared-gradient arena all-reduce,
densification-statistic sync and traversal sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mtgs_b200.parallel import SharedGradArena, sync_densification_stats, traversal_of_rank
        torch.manual_seed(0)  # identical replicated parameters on both ranks
        n = 1001
        shared = [torch.randn(n, 3, requires_grad=True), torch.randn(n, 3, requires_grad=True),
                  torch.randn(n, 4, requires_grad=True), torch.randn(n, requires_grad=True),
                  torch.randn(n, 3, requires_grad=True)]
        local = torch.randn(n, 3, requires_grad=True)  # per-traversal residual: never communicated
        arena = SharedGradArena(shared, average=True)
        assert arena.arena.numel() >= 14 * n and arena.nbytes >= 56 * n
        # rank-dependent "camera": loss differs per rank
        w = float(rank + 1)
        loss = sum((p * w).sum() * (i + 1) for i, p in enumerate(shared)) + (local * w).pow(2).sum()
        loss.backward()
        for i, p in enumerate(shared):  # backward wrote INTO the arena views
            assert p.grad.data_ptr() >= arena.arena.data_ptr()
            assert torch.allclose(p.grad, torch.full_like(p, w * (i + 1)))
        local_grad = local.grad.clone()
        arena.all_reduce()
        mean_w = sum(range(1, world + 1)) / world
        for i, p in enumerate(shared):
            assert torch.allclose(p.grad, torch.full_like(p, mean_w * (i + 1))), (rank, i)
        assert torch.equal(local.grad, local_grad)
        # second step accumulates from zero again
        arena.zero_()
        (shared[0].sum() * w).backward()
        arena.all_reduce()
        assert torch.allclose(shared[0].grad, torch.full_like(shared[0], mean_w))
        assert float(shared[1].grad.abs().max()) == 0.0
        # densification statistics
        gsum = torch.full((n,), float(rank + 1))
        vis = torch.full((n,), 1.0)
        rad = torch.full((n,), float(10 * (rank + 1)))
        sync_densification_stats(gsum, vis, rad)
        assert float(gsum[0]) == sum(range(1, world + 1)) and float(vis[0]) == world and float(rad[0]) == 10 * world
        assert traversal_of_rank(rank, world, 3) == [t for t in range(3) if t % world == rank]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"FAIL {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_shared_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_arena_validation_and_single_process():
    from mtgs_b200.parallel import SharedGradArena, traversal_of_rank
    with pytest.raises(ValueError):
        SharedGradArena([])
    with pytest.raises(ValueError):
        SharedGradArena([torch.zeros(3)])  # does not require grad
    p = torch.randn(5, 3, requires_grad=True)
    a = SharedGradArena([p])
    (p * 2).sum().backward()
    a.all_reduce()  # no process group: no-op, no averaging
    assert torch.allclose(p.grad, torch.full_like(p, 2.0))
    assert traversal_of_rank(1, 4, 8) == [1, 5]
    with pytest.raises(ValueError):
        traversal_of_rank(4, 4, 8)


def test_rank_sampler_reproduces_the_reference_draw_order_at_world_1():
    """Pinned against sequences produced by the reference's own MultiTraversalBalancedSampler
    (tests/golden/make_sampler_golden.py)."""
    import json
    import os
    import random
    from mtgs_b200.parallel import RankTraversalSampler
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_reference_golden.json")))
    for name, case in gold.items():
        random.seed(case["seed"])
        s = RankTraversalSampler(case["travel_ids"])
        got = [s.get_next_image_idx() for _ in range(len(case["sequence"]))]
        assert got == case["sequence"], name


def test_rank_sampler_partitions_traversals_and_stays_balanced():
    import random
    from collections import Counter
    from mtgs_b200.parallel import RankTraversalSampler
    travel_ids = [0] * 5 + [1] * 3 + [2] * 8 + [3] * 4
    seen_traversals = []
    for rank in range(2):
        s = RankTraversalSampler(travel_ids, rank=rank, world_size=2, rng=random.Random(rank))
        seen_traversals.append(set(s.traversals))
        draws = [s.get_next_image_idx() for _ in range(2 * 40)]
        per_trav = Counter(travel_ids[i] for i in draws)
        assert set(per_trav) == set(s.traversals)
        assert len(set(per_trav.values())) == 1                      # every owned traversal drawn equally often
        for t in s.traversals:                                       # one pass over a traversal visits each image once
            n = s.traversal_counts[t]
            first_pass = [i for i in draws if travel_ids[i] == t][:n]
            assert sorted(first_pass) == s.traversal_indices[t]
    assert seen_traversals[0].isdisjoint(seen_traversals[1]) and set.union(*seen_traversals) == {0, 1, 2, 3}
    import pytest
    with pytest.raises(ValueError):
        RankTraversalSampler([0, 0, 1], rank=3, world_size=4)
