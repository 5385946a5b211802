"""Golden vectors for the image-space losses of row f3, produced by the REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_losses_golden.py
The reference modules mtgs/utils/geometric_loss.py and mtgs/utils/camera_utils.py are loaded by file path.  Their
top-level imports of packages that are not installed here (cv2, torchmetrics, pyquaternion) are satisfied with empty
stand-in modules: none of the functions evaluated below touches them (they only need torch).  The masked L1 / inverse
L1 expressions are inline code of the reference's get_loss_dict (mtgs/scene_model/mtgs_scene_graph.py:825-828,
875-884, 929); they are evaluated here verbatim.  Inputs, outputs and autograd gradients go to losses_reference_golden.npz.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/mtgs/utils"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "losses_reference_golden.npz")


def _load():
    for name in ("cv2", "pyquaternion", "torchmetrics", "torchmetrics.image"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pyquaternion"].Quaternion = object
    sys.modules["torchmetrics.image"].MultiScaleStructuralSimilarityIndexMeasure = object
    sys.modules["torchmetrics.image"].StructuralSimilarityIndexMeasure = object
    for pkg in ("mtgs", "mtgs.utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    mods = {}
    for name in ("camera_utils", "geometric_loss"):
        spec = importlib.util.spec_from_file_location(f"mtgs.utils.{name}", os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"mtgs.utils.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["geometric_loss"]


def main():
    gl = _load()
    rng = np.random.default_rng(20261017)
    out = {}
    H, W = 70, 101

    # ---- masked L1 on RGB (mtgs_scene_graph.py:825) and on normals (:929)
    gt = torch.tensor(rng.random((H, W, 3), dtype=np.float32))
    pred = torch.tensor(np.clip(gt.numpy() + rng.normal(0, 0.1, (H, W, 3)), 0, 1).astype(np.float32), requires_grad=True)
    mask = torch.tensor(rng.random((H, W, 1)) < 0.7)
    l1 = torch.abs(gt - pred)[mask.squeeze(-1)].mean()
    (g,) = torch.autograd.grad(l1 * 1.7, pred)
    out.update(l1_gt=gt.numpy(), l1_pred=pred.detach().numpy(), l1_mask=mask.numpy(), l1_val=l1.item(), l1_grad=g.numpy())

    # ---- LiDAR depth losses (mtgs_scene_graph.py:875-884)
    gtd = torch.tensor((rng.random((H, W, 1)) * 90).astype(np.float32))
    gtd[torch.tensor(rng.random((H, W, 1)) < 0.6)] = 0.0  # sparse LiDAR
    prd = torch.tensor((rng.random((H, W, 1)) * 80 + 0.5).astype(np.float32), requires_grad=True)
    dmask = (gtd > 0.1) & (gtd < 80) & mask
    inv = torch.abs(1 / (gtd + 1e-5) - 1 / (prd + 1e-5))[dmask].mean()
    (gi,) = torch.autograd.grad(inv * 0.5, prd)
    pl = torch.abs(gtd - prd)[dmask].mean()
    (gp,) = torch.autograd.grad(pl * 0.5, prd)
    out.update(d_gt=gtd.numpy(), d_pred=prd.detach().numpy(), d_mask=dmask.numpy(), d_inv_val=inv.item(),
               d_inv_grad=gi.numpy(), d_l1_val=pl.item(), d_l1_grad=gp.numpy())

    # ---- TVLoss (geometric_loss.py:287-303)
    nrm = torch.tensor(rng.random((H, W, 3), dtype=np.float32), requires_grad=True)
    tv = gl.TVLoss()(nrm)
    (gt_,) = torch.autograd.grad(tv * 0.3, nrm)
    out.update(tv_in=nrm.detach().numpy(), tv_val=tv.item(), tv_grad=gt_.numpy())

    # ---- patch NCC (geometric_loss.py:322-348): the shipped config (patch 32 / stride 16, MTGS.py:110 +
    # mtgs_scene_graph.py:104-106), the function's own default (7 / 7) and an odd pair
    Hn, Wn = 96, 150
    yy, xx = np.meshgrid(np.arange(Hn), np.arange(Wn), indexing="ij")
    base = (5 + 0.05 * xx + 0.1 * yy + 2 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.float32)
    gtn = torch.tensor(base + rng.normal(0, 0.05, base.shape).astype(np.float32))[..., None]
    for tag, patch, stride in (("ncc32", 32, 16), ("ncc7", 7, 7), ("ncc9", 9, 4)):
        prn = torch.tensor(base * 1.1 + rng.normal(0, 0.3, base.shape).astype(np.float32))[..., None].requires_grad_(True)
        m = torch.tensor(rng.random((Hn, Wn, 1)) < 0.9995)
        m[:3] = True
        val = gl.calculate_depth_ncc_loss(prn, gtn, patch, stride, mask=m)
        (gn,) = torch.autograd.grad(val * 0.1, prn)
        out.update({f"{tag}_pred": prn.detach().numpy(), f"{tag}_gt": gtn.numpy(), f"{tag}_mask": m.numpy(),
                    f"{tag}_val": val.item(), f"{tag}_grad": gn.numpy(), f"{tag}_cfg": np.array([patch, stride])})

    # ---- normals from depth (geometric_loss.py:350-388), as get_loss_dict calls it (:912-922) and with a pose
    dep = torch.tensor((base[:64, :90] + 3).astype(np.float32))[..., None]
    fx, fy, cx, cy = 120.0, 118.0, 44.3, 31.6
    n0 = gl.normal_from_depth_image(dep, fx, fy, cx, cy, (90, 64), torch.eye(4), torch.device("cpu"), smooth=False)
    ang = 0.3
    c2w = torch.eye(4)
    c2w[:3, :3] = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float32)
    c2w[:3, 3] = torch.tensor([0.5, -1.0, 2.0])
    n1 = gl.normal_from_depth_image(dep, fx, fy, cx, cy, (90, 64), c2w, torch.device("cpu"), smooth=False)
    out.update(nd_depth=dep.numpy(), nd_k=np.array([fx, fy, cx, cy], np.float32), nd_eye=n0.numpy(), nd_c2w=c2w.numpy(),
               nd_pose=n1.numpy())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("_val")})


if __name__ == "__main__":
    main()
