"""Row f1: the Gaussian arena (one fused activation + rigid transform + SH kernel each way, csrc/arena.cu) against the
reference's per-node torch expressions (gaussian_model/vanilla_gaussian_splatting.py:299-322, rigid_node.py:206-215,
243-252) written out with torch ops, with the quaternion helpers pinned to vectors produced by the reference's own
quat_mult / quat_to_rotmat (tests/golden/make_reference_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_utils_golden.npz"))


def quat_mult(q1, q2):  # reference utils.py:60-70 (checked against its outputs below)
    w1, x1, y1, z1 = torch.unbind(q1, dim=-1)
    w2, x2, y2, z2 = torch.unbind(q2, dim=-1)
    return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], dim=-1)


def quat_to_rotmat(q):
    w, x, y, z = torch.unbind(q / q.norm(dim=-1, keepdim=True), dim=-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(q.shape[:-1] + (3, 3))


def test_helpers_are_the_references():
    out = quat_mult(torch.tensor(GOLD["qm_a"]), torch.tensor(GOLD["qm_b"]))
    np.testing.assert_allclose(out.numpy(), GOLD["qm_out"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(quat_to_rotmat(torch.tensor(GOLD["quats"])).numpy(), GOLD["rotmats"], rtol=1e-5, atol=1e-6)


def _nodes(dev, K=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    nodes = {}
    for name, n in (("background", 3001), ("object_car_a", 257), ("object_car_b", 130)):
        nodes[name] = dict(means=r(n, 3) * 5, scales=r(n, 3) * 0.5 - 2, quats=r(n, 4), opacities=r(n, 1),
                           features_dc=r(n, 3) * 0.5, features_rest=r(n, K - 1, 3) * 0.2)
    return {k: {a: t.to(dev) for a, t in v.items()} for k, v in nodes.items()}


@pytest.mark.parametrize("degree,K", [(3, 16), (0, 16), (1, 4), (2, 9), (4, 25)])
def test_arena_matches_per_node_reference(cuda_device, degree, K):
    from mtgs_b200.cuda._wrapper import spherical_harmonics
    from mtgs_b200.scene import GaussianArena
    dev = cuda_device
    nodes = _nodes(dev, K)
    poses = {"object_car_a": (torch.tensor([0.9, 0.1, -0.3, 0.2]), torch.tensor([4.0, -1.0, 0.5])),
             "object_car_b": (torch.tensor([0.3, -0.7, 0.2, 0.6]), torch.tensor([-6.0, 2.0, 1.0]))}
    c2w = torch.eye(4, device=dev)
    c2w[:3, 3] = torch.tensor([0.3, -0.2, 1.1], device=dev)
    arena = GaussianArena(nodes).to(dev)
    for n, (q, t) in poses.items():
        arena.set_pose(n, q, t)
    outs = arena.activated(c2w, degree)
    # ---- the reference's per-node expressions + the scene graph's per-attribute cat
    leaves, ref = [], {k: [] for k in ("means", "quats", "scales", "opacities", "colors")}
    for name, p in nodes.items():
        p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        leaves.append(p)
        qn = p["quats"] / p["quats"].norm(dim=-1, keepdim=True)
        if name in poses:
            q, t = (x.to(dev) for x in poses[name])
            q = q / q.norm()
            means = p["means"] @ quat_to_rotmat(q).T + t
            quats = quat_mult(q.expand_as(qn), qn)
        else:
            means, quats = p["means"], qn
        colors = torch.cat((p["features_dc"][:, None, :], p["features_rest"]), dim=1)
        if degree > 0:
            vd = means.detach() - c2w[:3, 3]
            vd = vd / vd.norm(dim=-1, keepdim=True)
            rgb = torch.clamp(spherical_harmonics(degree, vd, colors) + 0.5, 0.0, 1.0)
        else:
            rgb = torch.sigmoid(colors[:, 0, :])
        for k, v in (("means", means), ("quats", quats), ("scales", torch.exp(p["scales"])),
                     ("opacities", torch.sigmoid(p["opacities"]).squeeze(-1)), ("colors", rgb)):
            ref[k].append(v)
    ref = {k: torch.cat(v, dim=0) for k, v in ref.items()}
    for o, k in zip(outs, ("means", "quats", "scales", "opacities", "colors")):
        np.testing.assert_allclose(o.detach().cpu().numpy(), ref[k].detach().cpu().numpy(), rtol=2e-5, atol=2e-5, err_msg=k)
    # ---- gradients of a random linear functional of all five outputs
    g = torch.Generator(device=dev).manual_seed(5)
    ws = [torch.randn(o.shape, device=dev, generator=g) for o in outs]
    sum((o * w).sum() for o, w in zip(outs, ws)).backward()
    sum((ref[k] * w).sum() for k, w in zip(("means", "quats", "scales", "opacities", "colors"), ws)).backward()
    for name, p in zip(nodes, leaves):
        s = arena.slices[name]
        for attr, got in (("means", arena.means.grad[s]), ("scales", arena.scales.grad[s]), ("quats", arena.quats.grad[s]),
                          ("opacities", arena.opacities.grad[s]), ("features_dc", arena.sh.grad[s, 0]),
                          ("features_rest", arena.sh.grad[s, 1:])):
            want = p[attr].grad.reshape(got.shape)
            tol = 2e-5 * float(want.abs().max()) + 1e-7
            assert float((got - want).abs().max()) <= tol + 2e-4 * float(want.abs().max()), (name, attr)


def test_arena_renders_and_trains_with_one_optimizer(cuda_device):
    from mtgs_b200 import scenes
    from mtgs_b200.optim import FusedAdam
    from mtgs_b200.scene import GaussianArena
    dev = cuda_device
    s = scenes.street(n=20_000, seed=3, width=480, height=270)
    n = s["means"].shape[0]
    g = torch.Generator().manual_seed(1)
    node = dict(means=torch.tensor(s["means"]), scales=torch.log(torch.tensor(s["scales"])), quats=torch.tensor(s["quats"]),
                opacities=torch.logit(torch.tensor(s["opacities"]).clamp(1e-4, 1 - 1e-4))[:, None],
                features_dc=torch.randn(n, 3, generator=g) * 0.3, features_rest=torch.randn(n, 15, 3, generator=g) * 0.05)
    half = n // 2
    nodes = {"background": {k: v[:half] for k, v in node.items()}, "sky": {k: v[half:] for k, v in node.items()}}
    arena = GaussianArena(nodes, device=dev)
    viewmat, K = torch.tensor(s["viewmat"], device=dev), torch.tensor(s["K"], device=dev)
    with torch.no_grad():
        target, _, _ = arena.render(viewmat, K, 480, 270, 3, render_mode="RGB", rasterize_mode="antialiased")
    arena.sh.data[:, 0] += 0.3 * torch.randn(n, 3, device=dev)
    opt = FusedAdam([dict(params=[arena.sh], lr=2e-2), dict(params=[arena.opacities], lr=1e-2),
                     dict(params=[arena.means], lr=1e-4), dict(params=[arena.scales], lr=1e-3),
                     dict(params=[arena.quats], lr=1e-3)], eps=1e-15)
    losses = []
    for _ in range(30):
        r, _, _ = arena.render(viewmat, K, 480, 270, 3, render_mode="RGB", rasterize_mode="antialiased")
        loss = (r - target).abs().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.6 * losses[0], (losses[0], losses[-1])
