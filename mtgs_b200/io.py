"""Readers for MTGS checkpoints and prepared nuPlan road blocks, without nerfstudio.

* ``load_checkpoint`` reads the ``step-XXXXXXXXX.ckpt`` files written by ``CustomTrainer.save_checkpoint``
  (``mtgs/scene_model/custom_trainer.py:137-170``: a ``torch.save``d dict with ``step`` and ``pipeline`` = the
  pipeline's ``state_dict``) and splits the scene graph's parameters per node the way
  ``MTGSSceneModel.load_state_dict`` does (``mtgs/scene_model/mtgs_scene_graph.py:1185-1216``: keys
  ``gaussian_models.<node>.<subkey>``; nerfstudio's pipeline adds the ``_model.`` prefix, DDP ``module.``).
  ``node_rasterizer_inputs`` applies the reference's activations (``gaussian_model/vanilla_gaussian_splatting.py:
  299-322``: exp scales, normalised quats, sigmoid opacities, SH = cat(features_dc, features_rest)) so that a loaded
  node can be handed to ``mtgs_b200.rendering.rasterization`` directly.
* ``load_video_scene_dict`` / ``cameras_from_video_scene`` read ``video_scene_dict.pkl`` (schema:
  ``docs/prepare_dataset.md:104-190``) and assemble per-image camera records as ``NuplanDataParser`` does
  (``mtgs/dataset/nuplan_dataparser.py:107-330``: pose = ego2global @ cam2ego, intrinsics, travel id from the video
  token, timestamps) -- in OpenCV convention, i.e. ready for the rasterizer's ``viewmats`` (the reference converts to
  nerfstudio's OpenGL axes here and back again at ``mtgs_scene_graph.py:601-613``).

Host-side parsing only; nothing here is on the hot path.
"""
from __future__ import annotations

import io as _io
import pickle
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

GAUSS_ATTRS = ("means", "scales", "quats", "opacities", "features_dc", "features_rest", "features_adapters")


@dataclass
class SceneCheckpoint:
    step: int
    nodes: Dict[str, Dict[str, Tensor]]          # node name -> {"gauss_params.means": ..., "instance_trans": ..., ...}
    other: Dict[str, Tensor] = field(default_factory=dict)  # camera optimizer, appearance model, ...

    def gauss_params(self, node: str) -> Dict[str, Tensor]:
        return {k[len("gauss_params."):]: v for k, v in self.nodes[node].items() if k.startswith("gauss_params.")}

    @property
    def num_gaussians(self) -> int:
        return sum(int(self.gauss_params(n)["means"].shape[0]) for n in self.nodes if "gauss_params.means" in self.nodes[n])


def _strip_prefix(key: str) -> str:
    for p in ("module.", "_model.", "module."):
        if key.startswith(p):
            key = key[len(p):]
    return key


def split_scene_state_dict(state: Dict[str, Tensor]) -> (Dict[str, Dict[str, Tensor]], Dict[str, Tensor]):
    """``gaussian_models.<node>.<subkey>`` -> per-node dicts (mtgs_scene_graph.py:1186-1195); the rest is returned apart."""
    nodes: Dict[str, Dict[str, Tensor]] = {}
    other: Dict[str, Tensor] = {}
    for raw, value in state.items():
        key = _strip_prefix(raw)
        if key.startswith("gaussian_models."):
            name, sub = key[len("gaussian_models."):].split(".", 1)
            nodes.setdefault(name, {})[sub] = value
        else:
            other[key] = value
    return nodes, other


def load_checkpoint(path: str, device: str = "cpu") -> SceneCheckpoint:
    try:
        ckpt = torch.load(path, map_location=device, weights_only=True)
    except Exception:
        # nerfstudio checkpoints may pickle small config objects next to the tensors
        ckpt = torch.load(path, map_location=device, weights_only=False)
    if "pipeline" not in ckpt:
        raise KeyError(f"{path}: not a trainer checkpoint (no 'pipeline' entry; keys: {sorted(ckpt)[:8]})")
    nodes, other = split_scene_state_dict(ckpt["pipeline"])
    return SceneCheckpoint(step=int(ckpt.get("step", 0)), nodes=nodes, other=other)


def save_checkpoint(path: str, step: int, nodes: Dict[str, Dict[str, Tensor]], other: Optional[Dict[str, Tensor]] = None
                    ) -> None:
    """Writes the same layout ``CustomTrainer.save_checkpoint`` does for a finished run (no optimizer state)."""
    state = {f"_model.gaussian_models.{n}.{k}": v.detach().cpu() for n, d in nodes.items() for k, v in d.items()}
    state.update({f"_model.{k}": v.detach().cpu() for k, v in (other or {}).items()})
    torch.save({"step": int(step), "pipeline": state}, path)


def node_rasterizer_inputs(params: Dict[str, Tensor], traversal: Optional[int] = None) -> Dict[str, Tensor]:
    """Reference activations of a (multi-colour) vanilla node: ``scales.exp()``, ``quats / ||quats||``,
    ``sigmoid(opacities).squeeze(-1)`` (vanilla_gaussian_splatting.py:299-307) and the SH coefficient block
    ``cat(features_dc[:, None], features_rest)`` (:312); for multi-colour nodes the per-traversal slices
    ``features_adapters[:, t]`` / ``features_rest[:, t]`` (multi_color_gaussian_splatting.py:77-87)."""
    out = {"means": params["means"], "scales": torch.exp(params["scales"]),
           "quats": params["quats"] / params["quats"].norm(dim=-1, keepdim=True),
           "opacities": torch.sigmoid(params["opacities"]).squeeze(-1)}
    dc, rest = params["features_dc"], params.get("features_rest")
    if rest is not None and rest.dim() == 4:  # [N, T, K-1, 3]
        if traversal is None:
            raise ValueError("multi-colour node: pass the traversal index")
        rest = rest[:, traversal]
        if "features_adapters" in params:
            dc = dc + params["features_adapters"][:, traversal]
    out["sh_coeffs"] = dc[:, None, :] if rest is None else torch.cat([dc[:, None, :], rest], dim=1)
    return out


# ------------------------------------------------------------------------------------------------
class _SceneUnpickler(pickle.Unpickler):
    """video_scene_dict.pkl holds builtins, numpy arrays / scalars and datetime.date objects only."""
    _ALLOWED = ("builtins", "numpy", "datetime", "collections", "_codecs", "copyreg", "pathlib")

    def find_class(self, module, name):
        if module.split(".")[0] in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a scene pickle")


def load_video_scene_dict(path: str) -> Dict[str, Dict[str, Any]]:
    with open(path, "rb") as f:
        d = _SceneUnpickler(_io.BytesIO(f.read())).load()
    if not isinstance(d, dict):
        raise TypeError(f"{path}: expected a dict keyed by video token")
    for tok, v in d.items():
        if "frame_infos" not in v:
            raise KeyError(f"{path}: video {tok!r} has no 'frame_infos' (docs/prepare_dataset.md)")
    return d


def quat_wxyz_to_rotmat(q: Sequence[float]) -> np.ndarray:
    w, x, y, z = (float(v) for v in q)
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def matrix_from_translation_and_quaternion(translation, quaternion, opencv2nf: bool = False) -> np.ndarray:
    """Same contract as the reference helper (mtgs/utils/camera_utils.py:276-283)."""
    m = np.eye(4)
    r = quat_wxyz_to_rotmat(quaternion)
    if opencv2nf:
        r = r @ np.diag([1.0, -1.0, -1.0])
    m[:3, :3] = r
    m[:3, 3] = np.asarray(translation, np.float64)
    return m


def cameras_from_video_scene(video_scene_dict: Dict[str, Dict[str, Any]], cameras: Iterable[str] = ("CAM_F0",),
                             travels: Optional[Iterable[int]] = None, skip_flagged: bool = True,
                             use_colmap: bool = True, origin: Optional[Sequence[float]] = None) -> List[Dict[str, Any]]:
    """One record per (frame, camera): ``viewmat`` (OpenCV world -> camera, float64 4x4), ``K`` (3x3), ``distortion``,
    ``travel_id``, ``frame_idx``, ``timestamp``, ``image_path``, ``frame_token``, ``camera``.  Mirrors the loop of
    nuplan_dataparser.py:229-330: frames flagged ``skipped`` are dropped for training (``filter_skipped_frames``),
    COLMAP-refined poses / intrinsics win when present, ``travel_id`` is the integer suffix of the video token.
    ``origin`` (e.g. the road block centre) is subtracted from the camera positions."""
    cams = list(cameras)
    keep = None if travels is None else {int(t) for t in travels}
    out: List[Dict[str, Any]] = []
    for token, video in video_scene_dict.items():
        travel_id = int(str(token).split("-")[-1])
        if keep is not None and travel_id not in keep:
            continue
        for frame_idx, info in enumerate(video["frame_infos"]):
            if skip_flagged and info.get("skipped", False):
                continue
            for cam in cams:
                ci = info["cams"][cam]
                cp = ci.get("colmap_param") if use_colmap else None
                if use_colmap and not ci.get("valid", True):
                    continue
                K = np.asarray((cp or ci)["cam_intrinsic"], np.float64).reshape(3, 3)
                dist = np.asarray((cp or ci).get("distortion", np.zeros(5)), np.float64)
                if cp is not None and "sensor2global_translation" in cp:
                    c2w = matrix_from_translation_and_quaternion(cp["sensor2global_translation"], cp["sensor2global_rotation"])
                else:
                    c2w = np.asarray(info["ego2global"], np.float64) @ matrix_from_translation_and_quaternion(
                        ci["sensor2ego_translation"], ci["sensor2ego_rotation"])
                if origin is not None:
                    c2w = c2w.copy()
                    c2w[:3, 3] -= np.asarray(origin, np.float64)
                out.append(dict(viewmat=np.linalg.inv(c2w), K=K, distortion=dist, travel_id=travel_id, frame_idx=frame_idx,
                                timestamp=int(ci.get("timestamp", info["timestamp"])), image_path=ci["data_path"],
                                frame_token=info["token"], camera=cam))
    return out
