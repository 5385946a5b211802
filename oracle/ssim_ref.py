"""CPU oracle for the masked-SSIM row.  TEST INFRASTRUCTURE ONLY: imported by tests/, never by the
product path (mtgs_b200/).  Parity PINNED: checked against golden vectors produced by the reference's own
mtgs/utils/ssim.py (tests/golden/ssim_reference_golden.npz, tests/test_oracle_cpu.py).

Float64 numpy restatement of the reference:
  * window            mtgs/utils/ssim.py:11-25  (_fspecial_gauss_1d)
  * valid separable filter along H then W       ssim.py:28-53   (gaussian_filter; no padding)
  * SSIM / cs maps     ssim.py:78-99   (_ssim: C1, C2, mu, sigma, cs_map, ssim_map)
  * masked mean / per-channel means / relu      ssim.py:101-108, 150-153, 176-190
and the analytic gradient w.r.t. both images (adjoint of the filter + chain rule), which the reference obtains
through autograd.
"""
from __future__ import annotations

import numpy as np


def gauss_window(size: int, sigma: float) -> np.ndarray:
    c = np.arange(size, dtype=np.float32) - size // 2          # ssim.py:19-20 (float32 like the reference)
    g = np.exp(-(c ** 2) / np.float32(2 * sigma ** 2)).astype(np.float32)
    return (g / g.sum()).astype(np.float32)


def _filt(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    """valid separable filter over the last two axes (ssim.py:44-47)."""
    R = len(w)
    H, W = a.shape[-2:]
    v = sum(w[k] * a[..., k:H - R + 1 + k, :] for k in range(R))
    return sum(w[k] * v[..., :, k:W - R + 1 + k] for k in range(R))


def _filt_T(g: np.ndarray, w: np.ndarray, H: int, W: int) -> np.ndarray:
    """adjoint of _filt: scatters a valid-region map back onto the H x W image."""
    R = len(w)
    Ho, Wo = g.shape[-2:]
    v = np.zeros(g.shape[:-1] + (W,), g.dtype)
    for k in range(R):
        v[..., :, k:k + Wo] += w[k] * g
    out = np.zeros(g.shape[:-2] + (H, W), g.dtype)
    for k in range(R):
        out[..., k:k + Ho, :] += w[k] * v
    return out


def _expand_mask(mask, shape):
    if mask is None:
        return None
    m = np.asarray(mask).astype(bool)
    if m.ndim == 3:                       # (H, W, C) -> (1, C, H, W)   ssim.py:140-142
        m = np.transpose(m, (2, 0, 1))[None]
    return np.broadcast_to(m, shape)


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, K=(0.01, 0.03), nonnegative_ssim=False,
         mask=None, cotangent=None):
    """Returns (value, grad_X, grad_Y); gradients of sum(value * cotangent) (cotangent defaults to ones)."""
    X = np.asarray(X, np.float64)
    Y = np.asarray(Y, np.float64)
    N, C, H, W = X.shape
    w = gauss_window(win_size, win_sigma).astype(np.float64)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filt(X, w), _filt(Y, w)
    e11, e22, e12 = _filt(X * X, w), _filt(Y * Y, w), _filt(X * Y, w)
    s11, s22, s12 = e11 - mu1 ** 2, e22 - mu2 ** 2, e12 - mu1 * mu2
    A1, A2 = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    B1, B2 = mu1 ** 2 + mu2 ** 2 + C1, s11 + s22 + C2
    S = (A1 / B1) * (A2 / B2)                                   # ssim.py:98-99
    Ho, Wo = S.shape[-2:]
    m_full = _expand_mask(mask, X.shape)
    half = win_size // 2
    if m_full is not None:
        assert size_average is True                             # ssim.py:151
        m = m_full[..., half:H - half, half:W - half].astype(np.float64)   # ssim.py:152-153
        cnt = m.sum()
        val = (S * m).sum() / cnt if cnt > 0 else np.float64("nan")
        g_val = 1.0 if cotangent is None else float(np.asarray(cotangent))
        if nonnegative_ssim and val < 0:
            g_val, val = 0.0, 0.0
        gS = g_val * m / cnt if cnt > 0 else np.zeros_like(S)
    else:
        per = S.reshape(N, C, -1).mean(-1)                      # ssim.py:105
        relu_gate = np.ones_like(per)
        if nonnegative_ssim:
            relu_gate = (per > 0).astype(np.float64)
            per = np.maximum(per, 0)
        if size_average:
            val = per.mean()
            g_per = np.full((N, C), (1.0 if cotangent is None else float(np.asarray(cotangent))) / (N * C))
        else:
            val = per.mean(1)
            ct = np.ones(N) if cotangent is None else np.asarray(cotangent, np.float64)
            g_per = np.repeat(ct[:, None] / C, C, axis=1)
        gS = (g_per * relu_gate)[:, :, None, None] * np.ones_like(S) / (Ho * Wo)
    dmu2 = 2 * mu1 * (A2 - A1) / (B1 * B2) - S * 2 * mu2 * (1 / B1 - 1 / B2)
    dmu1 = 2 * mu2 * (A2 - A1) / (B1 * B2) - S * 2 * mu1 * (1 / B1 - 1 / B2)
    de_self = -S / B2
    de12 = 2 * A1 / (B1 * B2)
    gY = _filt_T(gS * dmu2, w, H, W) + 2 * Y * _filt_T(gS * de_self, w, H, W) + X * _filt_T(gS * de12, w, H, W)
    gX = _filt_T(gS * dmu1, w, H, W) + 2 * X * _filt_T(gS * de_self, w, H, W) + Y * _filt_T(gS * de12, w, H, W)
    return val, gX, gY
