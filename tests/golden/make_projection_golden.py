"""Golden vectors for the pinhole projection convention, produced by the REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_projection_golden.py
mtgs/utils/camera_utils.py:151-174 (`project_pix`) is how MTGS itself maps world points to pixel coordinates (OpenCV
camera-to-world pose, u = fx x / z + cx, no half-pixel offset) -- the same convention the rasterizer's projected means
(`info["means2d"]`) and depths must follow when it is handed viewmat = inverse(c2w) and K = [[fx,0,cx],[0,fy,cy],[0,0,1]]
(mtgs_scene_graph.py:548-640 builds them that way).  camera_utils.py is loaded by file path; its top-level imports of
packages that are not installed here (cv2, pyquaternion) are satisfied with empty stand-in modules (project_pix only
needs torch).  Inputs and outputs go to projection_reference_golden.npz.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/mtgs/utils/camera_utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "projection_reference_golden.npz")


def main():
    for name in ("cv2", "pyquaternion"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pyquaternion"].Quaternion = object
    spec = importlib.util.spec_from_file_location("ref_camera_utils", REF)
    cu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cu)

    rng = np.random.default_rng(20261018)
    # a camera pose (OpenCV axes: x right, y down, z forward): random rotation, translation of a few metres
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    c2w = np.eye(4)
    c2w[:3, :3] = R
    c2w[:3, 3] = rng.uniform(-5, 5, 3)
    fx, fy, cx, cy, W, H = 610.5, 598.25, 322.0, 178.0, 640, 360
    # points in front of the camera, inside and around the frustum
    pc = np.stack([rng.uniform(-1.2, 1.2, 400), rng.uniform(-0.8, 0.8, 400), np.ones(400)], 1) * rng.uniform(0.5, 80, (400, 1))
    pw = pc @ R.T + c2w[:3, 3]
    uvz = cu.project_pix(torch.tensor(pw, dtype=torch.float64), fx, fy, cx, cy, torch.tensor(c2w, dtype=torch.float64),
                         torch.device("cpu"), return_z_depths=True).numpy()
    np.savez_compressed(OUT, points=pw.astype(np.float32), c2w=c2w.astype(np.float64), fx=fx, fy=fy, cx=cx, cy=cy,
                        width=W, height=H, uvz=uvz.astype(np.float64))
    print("wrote", OUT, uvz[:2])


if __name__ == "__main__":
    main()
