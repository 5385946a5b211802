"""Comparison helpers shared by the parity tests."""
from __future__ import annotations

import numpy as np


def frac_mismatch(a, b, rtol, atol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    return float(bad.mean()), float(np.abs(a - b).max()) if a.size else 0.0


def assert_image_close(got, ref, name, rtol=1e-4, atol=2e-5, max_frac=3e-4, flip_bound=None):
    """Images agree to (rtol, atol) except for a small fraction of pixels where a hard threshold of the
    algorithm (alpha >= 1/255, T <= 1e-4) flipped under fp32 rounding / ex2.approx; those are
    bounded by the largest single-Gaussian contribution at the threshold."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    frac, mx = frac_mismatch(got, ref, rtol, atol)
    if flip_bound is None:
        flip_bound = 0.006 * max(1.0, float(np.abs(ref).max()))
    assert frac <= max_frac, f"{name}: {frac:.3e} of entries outside rtol={rtol}, atol={atol} (max abs {mx:.3e})"
    assert mx <= flip_bound, f"{name}: max abs diff {mx:.3e} > flip bound {flip_bound:.3e}"
    return frac, mx


def assert_grad_close(got, ref, name, rtol=5e-3, scale_atol=2e-4, max_frac=2e-3, outlier_bound=2e-2):
    """Gradients (sums over many pixel contributions; the CUDA side accumulates with fp32 atomics in
    arbitrary order and reconstructs T by division like upstream) agree to rtol plus an
    absolute slack proportional to the tensor's largest entry.  At most `max_frac` of the entries may fall
    outside that band (a hard threshold of the algorithm -- alpha >= 1/255, T <= 1e-4 -- flipped at some pixel
    under fp32 rounding), and even those are bounded: no entry may be off by more than `outlier_bound` x the
    tensor's largest entry."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    frac, mx = frac_mismatch(got, ref, rtol, scale_atol * scale + 1e-12)
    assert frac <= max_frac, f"{name}: {frac:.3e} of entries outside rtol={rtol} (+{scale_atol}*max); max abs {mx:.3e}, scale {scale:.3e}"
    assert mx <= outlier_bound * scale + 1e-12, \
        f"{name}: largest deviation {mx:.3e} exceeds {outlier_bound} x max|ref| = {outlier_bound * scale:.3e}"
    return frac, mx


def masked_psnr(pred, target, mask=None, data_range=1.0):
    """MaskedPSNR of the reference (mtgs/utils/pnsr.py:5-34: torchmetrics PeakSignalNoiseRatio(data_range) over the
    masked entries) = 10 log10(data_range^2 / mean squared error of the selected entries)."""
    pred = np.asarray(pred, np.float64)
    target = np.asarray(target, np.float64)
    if mask is not None:
        m = np.broadcast_to(np.asarray(mask, bool), pred.shape)
        pred, target = pred[m], target[m]
    mse = float(np.mean((pred - target) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(data_range ** 2 / mse)


def psnr(a, b, data_range=1.0):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(data_range ** 2 / mse)
