"""Fused image-space loss kernels (csrc/losses.cu through mtgs_b200.losses) against golden vectors produced by the
reference's own code (tests/golden/make_losses_golden.py).  Tolerances: values rtol 2e-5 (fp32 inputs, double
accumulation), gradients rtol 2e-3 (+ a floor relative to the tensor's largest entry for the NCC's fp32 statistics)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses_reference_golden.npz"))


def _t(a, dev, grad=False):
    t = torch.tensor(np.asarray(a), device=dev)
    return t.requires_grad_(True) if grad else t


def _close(a, b, rtol=2e-5, atol=1e-7):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def test_masked_l1(cuda_device):
    from mtgs_b200.losses import masked_l1
    pred = _t(G["l1_pred"], cuda_device, True)
    v = masked_l1(pred, _t(G["l1_gt"], cuda_device), _t(G["l1_mask"], cuda_device))
    (v * 1.7).backward()
    _close(v, G["l1_val"])
    _close(pred.grad, G["l1_grad"], rtol=1e-4, atol=1e-9)
    # [H, W] mask (what combined_mask.squeeze(-1) is) and no mask
    v2 = masked_l1(pred.detach(), _t(G["l1_gt"], cuda_device), _t(G["l1_mask"][..., 0], cuda_device))
    assert float(v2) == float(v)
    v3 = masked_l1(pred.detach(), _t(G["l1_gt"], cuda_device))
    _close(v3, np.abs(G["l1_gt"].astype(np.float64) - G["l1_pred"]).mean(), rtol=1e-5)


def test_lidar_depth_losses(cuda_device):
    from mtgs_b200.losses import masked_l1
    for inverse, tag, rt in ((True, "d_inv", 2e-3), (False, "d_l1", 1e-4)):
        pred = _t(G["d_pred"], cuda_device, True)
        v = masked_l1(pred, _t(G["d_gt"], cuda_device), _t(G["d_mask"], cuda_device), inverse=inverse)
        (v * 0.5).backward()
        _close(v, G[tag + "_val"], rtol=1e-4)
        _close(pred.grad, G[tag + "_grad"], rtol=rt, atol=1e-9)
    # empty selection -> NaN like the reference's mean over nothing
    z = masked_l1(_t(G["d_pred"], cuda_device), _t(G["d_gt"], cuda_device), torch.zeros(70, 101, 1, dtype=torch.bool, device=cuda_device))
    assert torch.isnan(z)


def test_tv(cuda_device):
    from mtgs_b200.losses import TVLoss
    x = _t(G["tv_in"], cuda_device, True)
    v = TVLoss()(x)
    (v * 0.3).backward()
    _close(v, G["tv_val"])
    _close(x.grad, G["tv_grad"], rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("tag", ["ncc32", "ncc7", "ncc9"])
def test_ncc(cuda_device, tag):
    from mtgs_b200.losses import calculate_depth_ncc_loss
    patch, stride = (int(x) for x in G[tag + "_cfg"])
    pred = _t(G[tag + "_pred"], cuda_device, True)
    v = calculate_depth_ncc_loss(pred, _t(G[tag + "_gt"], cuda_device), patch, stride, mask=_t(G[tag + "_mask"], cuda_device))
    (v * 0.1).backward()
    _close(v, G[tag + "_val"], rtol=2e-4, atol=2e-6)
    ref = G[tag + "_grad"]
    _close(pred.grad, ref, rtol=5e-3, atol=5e-4 * np.abs(ref).max())


def test_normal_from_depth(cuda_device):
    from mtgs_b200.losses import normal_from_depth_image
    fx, fy, cx, cy = (float(x) for x in G["nd_k"])
    d = _t(G["nd_depth"], cuda_device)
    n0 = normal_from_depth_image(d, fx, fy, cx, cy, (90, 64), torch.eye(4), cuda_device)
    _close(n0, G["nd_eye"], rtol=1e-3, atol=2e-4)
    n1 = normal_from_depth_image(d, fx, fy, cx, cy, (90, 64), _t(G["nd_c2w"], cuda_device), cuda_device)
    _close(n1, G["nd_pose"], rtol=1e-3, atol=2e-4)


def test_full_resolution_runs(cuda_device):
    """960x540 (the training resolution, mtgs/config/MTGS.py:43): all losses + backward, finite."""
    from mtgs_b200.losses import TVLoss, calculate_depth_ncc_loss, masked_l1, normal_from_depth_image
    H, W = 540, 960
    g = torch.Generator(device=cuda_device).manual_seed(1)
    rgb = torch.rand(H, W, 3, device=cuda_device, generator=g, requires_grad=True)
    dep = (torch.rand(H, W, 1, device=cuda_device, generator=g) * 60 + 1).requires_grad_(True)
    nrm = torch.rand(H, W, 3, device=cuda_device, generator=g, requires_grad=True)
    gt_rgb, gt_dep = torch.rand_like(rgb), torch.rand_like(dep) * 60 + 1
    mask = torch.rand(H, W, 1, device=cuda_device, generator=g) < 0.98
    gt_n = (1 + normal_from_depth_image(gt_dep, 772.5, 772.5, 480.0, 280.0, (W, H), torch.eye(4), cuda_device)
            @ torch.diag(torch.tensor([1.0, -1.0, -1.0], device=cuda_device))) / 2
    loss = (0.8 * masked_l1(rgb, gt_rgb, mask) + 0.5 * masked_l1(dep, gt_dep, mask, inverse=True)
            + 0.1 * calculate_depth_ncc_loss(dep, gt_dep, 32, 16, mask=torch.ones_like(mask))
            + 0.1 * (masked_l1(nrm, gt_n, mask) + TVLoss()(nrm)))
    loss.backward()
    for t in (rgb, dep, nrm):
        assert torch.isfinite(t.grad).all() and float(t.grad.abs().sum()) > 0
    assert torch.isfinite(loss)
