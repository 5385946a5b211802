"""Masked-SSIM kernels (b2s_ssim_fwd / b2s_ssim_bwd through mtgs_b200.ssim) against the reference's own golden vectors
and the numpy oracle.  Tolerances (fp32 kernels vs fp32 reference / float64 oracle): value rtol 2e-5, gradients
rtol 2e-3 + 2e-5 of the largest entry."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import ssim_ref

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssim_reference_golden.npz"))
CASES = ["mtgs_call", "nomask_avg", "nomask_per_image", "nchw_mask_nonneg", "range255_win7"]


@pytest.mark.parametrize("name", CASES)
def test_kernels_match_reference_golden(cuda_device, name):
    from mtgs_b200.ssim import ssim
    kw = ast.literal_eval(str(GOLD[f"{name}/kwargs"]))
    X = torch.tensor(GOLD[f"{name}/X"], device=cuda_device, requires_grad=True)
    Y = torch.tensor(GOLD[f"{name}/Y"], device=cuda_device, requires_grad=True)
    mask = torch.tensor(GOLD[f"{name}/mask"], device=cuda_device) if f"{name}/mask" in GOLD.files else None
    val = ssim(X, Y, mask=mask, **kw)
    assert tuple(val.shape) == tuple(GOLD[f"{name}/value"].shape)
    np.testing.assert_allclose(val.detach().cpu().numpy(), GOLD[f"{name}/value"], rtol=2e-5, atol=2e-6)
    (val * torch.tensor(GOLD[f"{name}/cotangent"], device=cuda_device)).sum().backward()
    scale = float(np.abs(GOLD[f"{name}/grad_Y"]).max())
    np.testing.assert_allclose(Y.grad.cpu().numpy(), GOLD[f"{name}/grad_Y"], rtol=2e-3, atol=2e-5 * scale + 1e-12)
    np.testing.assert_allclose(X.grad.cpu().numpy(), GOLD[f"{name}/grad_X"], rtol=2e-3, atol=2e-5 * scale + 1e-12)


def test_module_call_as_mtgs_makes_it(cuda_device):
    """mtgs_scene_graph.py:822-840: 1 - self.ssim(gt[None], pred[None], mask=combined_mask), gradient to pred only."""
    from mtgs_b200.ssim import MaskedSSIM
    mod = MaskedSSIM(data_range=1.0, size_average=True, channel=3)
    X = torch.tensor(GOLD["module/X"], device=cuda_device)
    Y = torch.tensor(GOLD["module/Y"], device=cuda_device, requires_grad=True)
    val = mod(X, Y, mask=torch.tensor(GOLD["module/mask"], device=cuda_device))
    (1 - val).backward()
    np.testing.assert_allclose(val.item(), GOLD["module/value"], rtol=2e-5)
    scale = float(np.abs(GOLD["module/grad_Y"]).max())
    np.testing.assert_allclose(Y.grad.cpu().numpy(), GOLD["module/grad_Y"], rtol=2e-3, atol=2e-5 * scale)


@pytest.mark.parametrize("H,W", [(540, 960), (1080, 1920), (101, 77)])
def test_training_resolutions_against_oracle(cuda_device, H, W):
    """MTGS trains at 960x540 (mtgs/config/MTGS.py:43), the BASELINE metric is quoted at 1920x1080; ragged size too."""
    from mtgs_b200.ssim import ssim
    rng = np.random.default_rng(H)
    Xn = rng.random((1, 3, H, W), dtype=np.float32)
    Yn = np.clip(Xn * 0.8 + rng.normal(0, 0.1, Xn.shape), 0, 1).astype(np.float32)
    mn = rng.random((H, W, 1)) < 0.9
    X = torch.tensor(Xn, device=cuda_device)
    Y = torch.tensor(Yn, device=cuda_device, requires_grad=True)
    val = ssim(X, Y, data_range=1.0, mask=torch.tensor(mn, device=cuda_device))
    val.backward()
    want, _, gY = ssim_ref.ssim(Xn, Yn, data_range=1.0, mask=mn)
    np.testing.assert_allclose(val.item(), want, rtol=2e-5)
    scale = float(np.abs(gY).max())
    np.testing.assert_allclose(Y.grad.cpu().numpy(), gY, rtol=2e-3, atol=2e-5 * scale)
    assert X.grad is None


def test_edge_cases(cuda_device):
    from mtgs_b200.ssim import ssim
    X = torch.rand(1, 3, 32, 32, device=cuda_device)
    assert abs(ssim(X, X.clone(), data_range=1.0).item() - 1.0) < 1e-6                  # identical images
    empty = torch.zeros(32, 32, 1, dtype=torch.bool, device=cuda_device)
    assert torch.isnan(ssim(X, X.clone(), data_range=1.0, mask=empty))                   # mean of nothing, as the reference
    with pytest.raises(NotImplementedError):
        ssim(torch.rand(1, 3, 8, 40, device=cuda_device), torch.rand(1, 3, 8, 40, device=cuda_device))  # H < window
    with pytest.raises(ValueError):
        ssim(X, X, win_size=10)
