"""The numpy oracle of the f3 image-space losses (oracle/losses_ref.py) against golden vectors produced by the
reference's own code (tests/golden/make_losses_golden.py): values and gradients."""
import os

import numpy as np

from oracle import losses_ref as L

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses_reference_golden.npz"))


def _close(a, b, rtol=2e-5, atol=1e-7):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def test_masked_l1_pinned():
    v, g = L.masked_l1(G["l1_pred"], G["l1_gt"], G["l1_mask"], grad_out=1.7)
    _close(v, G["l1_val"])
    _close(g, G["l1_grad"], rtol=1e-4, atol=1e-9)


def test_lidar_depth_losses_pinned():
    v, g = L.masked_l1(G["d_pred"], G["d_gt"], G["d_mask"], inverse=True, grad_out=0.5)
    _close(v, G["d_inv_val"], rtol=1e-4)
    _close(g, G["d_inv_grad"], rtol=2e-3, atol=1e-9)
    v, g = L.masked_l1(G["d_pred"], G["d_gt"], G["d_mask"], grad_out=0.5)
    _close(v, G["d_l1_val"])
    _close(g, G["d_l1_grad"], rtol=1e-4, atol=1e-9)


def test_tv_pinned():
    v, g = L.tv_loss(G["tv_in"], grad_out=0.3)
    _close(v, G["tv_val"])
    _close(g, G["tv_grad"], rtol=1e-4, atol=1e-9)


def test_ncc_pinned():
    for tag in ("ncc32", "ncc7", "ncc9"):
        patch, stride = (int(x) for x in G[tag + "_cfg"])
        v, g = L.depth_ncc_loss(G[tag + "_pred"], G[tag + "_gt"], patch, stride, G[tag + "_mask"], grad_out=0.1)
        _close(v, G[tag + "_val"], rtol=2e-4, atol=2e-6)
        ref = G[tag + "_grad"].reshape(g.shape)
        _close(g, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max())


def test_normal_from_depth_pinned():
    fx, fy, cx, cy = (float(x) for x in G["nd_k"])
    _close(L.normal_from_depth(G["nd_depth"], fx, fy, cx, cy, np.eye(4)), G["nd_eye"], rtol=1e-3, atol=2e-4)
    _close(L.normal_from_depth(G["nd_depth"], fx, fy, cx, cy, G["nd_c2w"]), G["nd_pose"], rtol=1e-3, atol=2e-4)
