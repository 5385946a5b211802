"""Row f2: fused multi-tensor Adam against torch.optim.Adam (the optimizer nerfstudio builds per group,
mtgs/scene_model/custom_trainer.py:115-136), densification statistics and row compaction against the reference's torch
expressions (gaussian_model/vanilla_gaussian_splatting.py:392-474, 580-621)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam(cuda_device):
    from mtgs_b200.optim import FusedAdam
    g = torch.Generator(device=cuda_device).manual_seed(3)
    shapes = [(5000, 3), (5000, 4), (5000, 3), (5000, 1), (5000, 1, 3), (5000, 15, 3), (37,), (4097, 2), (1, 1)]
    lrs = [1.6e-4, 1e-3, 5e-3, 5e-2, 2.5e-3, 1.25e-4, 1e-2, 3e-4, 1e-1]
    ref_p = [torch.randn(s, device=cuda_device, generator=g).requires_grad_(True) for s in shapes]
    our_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    ref = [torch.optim.Adam([p], lr=lr, eps=1e-15) for p, lr in zip(ref_p, lrs)]  # one optimizer per group, as nerfstudio
    ours = FusedAdam([dict(params=[p], lr=lr) for p, lr in zip(our_p, lrs)], eps=1e-15)
    for step in range(25):
        for i, (a, b) in enumerate(zip(ref_p, our_p)):
            gr = torch.randn(a.shape, device=cuda_device, generator=g) * (10.0 ** ((i % 5) - 3))
            if step % 7 == 3 and i == 2:
                gr = torch.zeros_like(gr)
            a.grad, b.grad = gr.clone(), gr.clone()
        for o in ref:
            o.step()
        ours.step()
        if step == 10:  # a scheduler changed the learning rates
            for o, grp in zip(ref, ours.param_groups):
                o.param_groups[0]["lr"] *= 0.5
                grp["lr"] *= 0.5
    for a, b, o in zip(ref_p, our_p, ref):
        np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().cpu().numpy(), rtol=2e-5, atol=1e-7)
        st_r, st_o = o.state[a], ours.state[b]
        np.testing.assert_allclose(st_o["exp_avg"].cpu().numpy(), st_r["exp_avg"].cpu().numpy(), rtol=1e-5, atol=1e-10)
        np.testing.assert_allclose(st_o["exp_avg_sq"].cpu().numpy(), st_r["exp_avg_sq"].cpu().numpy(), rtol=1e-5, atol=1e-14)
        assert int(st_o["step"]) == int(st_r["step"])


def test_densify_stats_match_reference_expressions(cuda_device):
    from mtgs_b200.optim import accumulate_densify_stats
    g = torch.Generator(device=cuda_device).manual_seed(4)
    N, W, H = 10_001, 960, 540
    arena = torch.randn(N, 4, device=cuda_device, generator=g)
    absgrad = arena[:, 2:4]                      # a stride-4 view, as rasterization hands it out
    radii = torch.randint(-1, 40, (1, N), device=cuda_device, generator=g, dtype=torch.int32)
    radii[0, ::3] = 0
    norm = torch.rand(N, device=cuda_device, generator=g)
    cnt = torch.ones(N, device=cuda_device)
    mx = torch.rand(N, device=cuda_device, generator=g) * 20
    want_n, want_c, want_m = norm.clone(), cnt.clone(), mx.clone()
    # reference: mtgs_scene_graph.py:1171-1178 and vanilla_gaussian_splatting.py:455-474
    vis = (radii > 0).flatten()
    grads = (absgrad * absgrad.new_tensor([W, H]).unsqueeze(0) * 0.5).norm(dim=-1)
    want_c[vis] += 1
    want_n[vis] += grads[vis]
    want_m[vis] = torch.maximum(want_m[vis], radii.flatten()[vis].float())
    accumulate_densify_stats(absgrad, radii, W, H, norm, cnt, mx)
    np.testing.assert_allclose(norm.cpu().numpy(), want_n.cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert torch.equal(cnt, want_c) and torch.equal(mx, want_m)


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 100_003])
def test_compact_rows_is_boolean_indexing(cuda_device, n):
    from mtgs_b200.optim import compact_rows
    g = torch.Generator(device=cuda_device).manual_seed(n + 1)
    keep = torch.rand(n, device=cuda_device, generator=g) < 0.63
    ts = [torch.randn(n, 3, device=cuda_device, generator=g), torch.randn(n, device=cuda_device, generator=g),
          torch.randn(n, 15, 3, device=cuda_device, generator=g), torch.randn(n, 4, device=cuda_device, generator=g)]
    outs = compact_rows(ts, keep)
    for t, o in zip(ts, outs):
        assert torch.equal(o, t[keep])


def test_cull_carries_optimizer_state(cuda_device):
    from mtgs_b200.optim import FusedAdam, cull_optimizer_rows
    g = torch.Generator(device=cuda_device).manual_seed(9)
    n = 3000
    params = {"means": torch.nn.Parameter(torch.randn(n, 3, device=cuda_device, generator=g)),
              "opacities": torch.nn.Parameter(torch.randn(n, 1, device=cuda_device, generator=g))}
    opt = FusedAdam([dict(params=[params["means"]], lr=1e-3), dict(params=[params["opacities"]], lr=5e-2)], eps=1e-15)
    for p in params.values():
        p.grad = torch.randn(p.shape, device=cuda_device, generator=g)
    opt.step()
    keep = torch.sigmoid(params["opacities"].detach()).squeeze() >= 0.4   # cull_gaussians' alpha criterion
    old = {k: (v.detach().clone(), opt.state[v]["exp_avg"].clone(), opt.state[v]["exp_avg_sq"].clone()) for k, v in params.items()}
    new = cull_optimizer_rows(opt, params, keep)
    for k, p in new.items():
        assert torch.equal(p.detach(), old[k][0][keep])
        assert torch.equal(opt.state[p]["exp_avg"], old[k][1][keep]) and torch.equal(opt.state[p]["exp_avg_sq"], old[k][2][keep])
        p.grad = torch.randn(p.shape, device=cuda_device, generator=g)
    opt.step()  # the chunk table is rebuilt for the new sizes
    assert all(torch.isfinite(p).all() for p in new.values())
