"""Masked-SSIM row on the CPU: the numpy oracle against the golden vectors that the reference's own
mtgs/utils/ssim.py produced (tests/golden/make_ssim_golden.py) -- this row's parity is PINNED to the reference."""
import ast
import os

import numpy as np
import pytest

from oracle import ssim_ref

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssim_reference_golden.npz"))
CASES = ["mtgs_call", "nomask_avg", "nomask_per_image", "nchw_mask_nonneg", "range255_win7"]


def _kw(name):
    return ast.literal_eval(str(GOLD[f"{name}/kwargs"]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    kw = _kw(name)
    mask = GOLD[f"{name}/mask"] if f"{name}/mask" in GOLD.files else None
    val, gX, gY = ssim_ref.ssim(GOLD[f"{name}/X"], GOLD[f"{name}/Y"], mask=mask, cotangent=GOLD[f"{name}/cotangent"], **kw)
    np.testing.assert_allclose(val, GOLD[f"{name}/value"], rtol=2e-5, atol=2e-6)
    # tolerance written here: fp32 reference vs float64 oracle; gradients are sums of ~121 fp32 products
    scale = float(np.abs(GOLD[f"{name}/grad_Y"]).max())
    np.testing.assert_allclose(gY, GOLD[f"{name}/grad_Y"], rtol=2e-3, atol=2e-5 * scale + 1e-12)
    np.testing.assert_allclose(gX, GOLD[f"{name}/grad_X"], rtol=2e-3, atol=2e-5 * scale + 1e-12)


def test_oracle_matches_reference_module_call():
    val, _, gY = ssim_ref.ssim(GOLD["module/X"], GOLD["module/Y"], data_range=1.0, mask=GOLD["module/mask"], cotangent=-1.0)
    np.testing.assert_allclose(val, GOLD["module/value"], rtol=2e-5)
    scale = float(np.abs(GOLD["module/grad_Y"]).max())
    np.testing.assert_allclose(gY, GOLD["module/grad_Y"], rtol=2e-3, atol=2e-5 * scale)


def test_window_matches_reference_constants():
    w = ssim_ref.gauss_window(11, 1.5)
    assert abs(float(w.sum()) - 1.0) < 1e-6 and np.allclose(w, w[::-1]) and w.argmax() == 5
    np.testing.assert_allclose(w[5], 0.26601171, rtol=1e-5)  # centre tap of the 11 / 1.5 window
    from mtgs_b200.ssim import _fspecial_gauss_1d
    for size, sigma in ((11, 1.5), (7, 1.0), (15, 2.5)):
        ours = _fspecial_gauss_1d(size, sigma)
        assert tuple(ours.shape) == (1, 1, size)
        np.testing.assert_allclose(ours.reshape(-1).numpy(), ssim_ref.gauss_window(size, sigma), rtol=1e-6)  # 1 ulp (exp)


def test_host_mirror_signature_and_errors():
    import inspect
    import torch
    from mtgs_b200 import ssim as ours
    sig = inspect.signature(ours.ssim)
    assert list(sig.parameters) == ["X", "Y", "data_range", "size_average", "win_size", "win_sigma", "win", "K",
                                    "nonnegative_ssim", "mask"]  # reference ssim.py:110-121
    assert sig.parameters["data_range"].default == 255 and sig.parameters["win_size"].default == 11
    m = ours.MaskedSSIM(data_range=1.0, size_average=True, channel=3)
    assert tuple(m.win.shape) == (3, 1, 1, 11)  # reference ssim.py:221
    with pytest.raises(ValueError):
        ours.ssim(torch.zeros(1, 3, 20, 20), torch.zeros(1, 3, 20, 21))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ours.ssim(torch.zeros(1, 3, 20, 20), torch.zeros(1, 3, 20, 20), data_range=1.0)
