"""Golden vectors for the masked-SSIM row, produced by the REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_ssim_golden.py
It loads /root/reference/mtgs/utils/ssim.py by file path (the module only needs torch; importing the ``mtgs`` package
would pull in nerfstudio, which is not installed), evaluates ``ssim`` / ``MaskedSSIM`` on seeded inputs on the CPU and
stores inputs, outputs and autograd gradients in ssim_reference_golden.npz.  /root/reference does not exist on the GPU
box; the tests only read the .npz.
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/mtgs/utils/ssim.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ssim_reference_golden.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_ssim", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(20260117)
    out = {}

    def case(name, N, C, H, W, mask_kind, **kw):
        X = torch.tensor(rng.random((N, C, H, W), dtype=np.float32))
        # Y = X + structured + random perturbation, clipped: SSIM values spread over (0, 1)
        Y = torch.tensor(np.clip(X.numpy() * rng.uniform(0.6, 1.0) + rng.normal(0, 0.15, (N, C, H, W)), 0, 1)
                         .astype(np.float32))
        X.requires_grad_(True)
        Y.requires_grad_(True)
        mask = None
        if mask_kind == "hwc":      # what MTGS passes: combined_mask [H, W, 1] (mtgs_scene_graph.py:820-838)
            mask = torch.tensor(rng.random((H, W, 1)) < 0.7)
        elif mask_kind == "nchw":
            mask = torch.tensor(rng.random((N, C, H, W)) < 0.5)
        val = ref.ssim(X, Y, mask=mask, **kw)
        w = torch.tensor(rng.normal(size=tuple(val.shape)).astype(np.float32)) if val.dim() else torch.tensor(1.0)
        (val * w).sum().backward()
        out[f"{name}/X"] = X.detach().numpy()
        out[f"{name}/Y"] = Y.detach().numpy()
        if mask is not None:
            out[f"{name}/mask"] = mask.numpy()
        out[f"{name}/value"] = val.detach().numpy()
        out[f"{name}/cotangent"] = w.numpy()
        out[f"{name}/grad_X"] = X.grad.numpy()
        out[f"{name}/grad_Y"] = Y.grad.numpy()
        out[f"{name}/kwargs"] = np.array(repr(kw))

    case("mtgs_call", 1, 3, 45, 61, "hwc", data_range=1.0, size_average=True)          # MaskedSSIM(data_range=1.0, channel=3)
    case("nomask_avg", 1, 3, 33, 40, None, data_range=1.0, size_average=True)
    case("nomask_per_image", 2, 3, 24, 37, None, data_range=1.0, size_average=False)
    case("nchw_mask_nonneg", 2, 2, 30, 30, "nchw", data_range=1.0, size_average=True, nonnegative_ssim=True)
    case("range255_win7", 1, 1, 19, 23, None, data_range=255, size_average=True, win_size=7, win_sigma=1.0)
    # the module form, exactly as constructed by MTGS
    mod = ref.MaskedSSIM(data_range=1.0, size_average=True, channel=3)
    X = torch.tensor(rng.random((1, 3, 40, 52), dtype=np.float32))
    Y = torch.tensor(np.clip(X.numpy() + rng.normal(0, 0.1, X.shape), 0, 1).astype(np.float32), requires_grad=True)
    mask = torch.tensor(rng.random((40, 52, 1)) < 0.8)
    val = mod(X, Y, mask=mask)
    (1 - val).backward()
    out["module/X"], out["module/Y"], out["module/mask"] = X.numpy(), Y.detach().numpy(), mask.numpy()
    out["module/value"], out["module/grad_Y"] = val.detach().numpy(), Y.grad.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.endswith("value")})


if __name__ == "__main__":
    main()
