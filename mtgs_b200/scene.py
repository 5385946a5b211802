"""Gaussian arena: all nodes of a scene graph in one structure-of-arrays buffer.

Reference: every node model (background, sky, one rigid node per vehicle ...) owns its parameters and turns them into
rasterizer inputs with its own chain of torch ops (``gaussian_model/vanilla_gaussian_splatting.py:299-322``,
``rigid_node.py:206-215, 243-252``); ``MTGSSceneModel.get_gaussians`` then concatenates the per-node results
attribute by attribute on every step (``mtgs_scene_graph.py:408-461``, the ``torch.cat`` at ``:454-455``).

Here the scene keeps ONE tensor per attribute (``means``, ``scales``, ``quats``, ``opacities``, ``sh``); a node is a row
range plus a per-frame pose.  ``GaussianArena.activated`` is one kernel launch (``csrc/arena.cu``) that reads the raw
rows and writes the rasterizer inputs of the whole scene -- activations, rigid-node transform and the view-dependent SH
colour included -- and one launch in the backward.  One ``FusedAdam`` (``mtgs_b200.optim``) over the five arena
tensors replaces the per-node, per-attribute optimizers.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from .rendering import _need_cuda, _ptr, _stream, rasterization


class _ArenaActivate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, scales, quats, opacities, sh, node_of, poses, campos, degree):
        lib = _lib.load()
        N, K = sh.shape[0], sh.shape[1]
        dev = means.device
        means_w = torch.empty(N, 3, dtype=torch.float32, device=dev)
        quats_w = torch.empty(N, 4, dtype=torch.float32, device=dev)
        scales_a = torch.empty(N, 3, dtype=torch.float32, device=dev)
        opac_a = torch.empty(N, dtype=torch.float32, device=dev)
        colors = torch.empty(N, 3, dtype=torch.float32, device=dev)
        mask = torch.empty(N, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.b2s_arena_fwd(_ptr(means), _ptr(scales), _ptr(quats), _ptr(opacities), _ptr(sh), _ptr(node_of),
                                         _ptr(poses), _ptr(campos), N, K, degree, _ptr(means_w), _ptr(quats_w),
                                         _ptr(scales_a), _ptr(opac_a), _ptr(colors), _ptr(mask), _stream()),
                       "b2s_arena_fwd")
        ctx.save_for_backward(means, scales, quats, opacities, sh, node_of, poses, campos, mask)
        ctx.degree = degree
        ctx.set_materialize_grads(False)
        return means_w, quats_w, scales_a, opac_a, colors

    @staticmethod
    def backward(ctx, v_means_w, v_quats_w, v_scales, v_opac, v_colors):
        lib = _lib.load()
        means, scales, quats, opacities, sh, node_of, poses, campos, mask = ctx.saved_tensors
        N, K = sh.shape[0], sh.shape[1]
        dev = means.device

        def z(v, *shape):
            return torch.zeros(*shape, dtype=torch.float32, device=dev) if v is None else v.contiguous()
        v_means_w, v_quats_w, v_scales = z(v_means_w, N, 3), z(v_quats_w, N, 4), z(v_scales, N, 3)
        v_opac, v_colors = z(v_opac, N), z(v_colors, N, 3)
        g_means, g_scales = torch.empty_like(means), torch.empty_like(scales)
        g_quats, g_opac, g_sh = torch.empty_like(quats), torch.empty_like(opacities), torch.empty_like(sh)
        with torch.cuda.device(dev):
            _lib.check(lib.b2s_arena_bwd(_ptr(means), _ptr(scales), _ptr(quats), _ptr(opacities), _ptr(sh), _ptr(node_of),
                                         _ptr(poses), _ptr(campos), N, K, ctx.degree, _ptr(mask), _ptr(v_means_w),
                                         _ptr(v_quats_w), _ptr(v_scales), _ptr(v_opac), _ptr(v_colors), _ptr(g_means),
                                         _ptr(g_quats), _ptr(g_scales), _ptr(g_opac), _ptr(g_sh), _stream()),
                       "b2s_arena_bwd")
        return g_means, g_scales, g_quats, g_opac, g_sh, None, None, None, None


def _quat_to_rotmat(q: Tensor) -> Tensor:
    w, x, y, z = (q / q.norm()).unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]).reshape(3, 3)


class GaussianArena(torch.nn.Module):
    """All Gaussians of a scene graph as five leaf tensors; nodes are row ranges.

    ``nodes``: mapping name -> dict with raw parameters ``means [n,3]``, ``scales [n,3]`` (log), ``quats [n,4]`` (wxyz),
    ``opacities [n]`` or ``[n,1]`` (logit), ``features_dc [n,3]``, ``features_rest [n,K-1,3]`` -- the reference's
    ``gauss_params`` of a vanilla / rigid node (``mtgs_b200.io.SceneCheckpoint.gauss_params`` gives exactly this)."""

    def __init__(self, nodes: Dict[str, Dict[str, Tensor]], device: Optional[torch.device] = None):
        super().__init__()
        names = list(nodes)
        if not names:
            raise ValueError("no nodes")
        dev = torch.device(device) if device is not None else nodes[names[0]]["means"].device
        cat = lambda key, f=lambda t: t: torch.cat([f(nodes[n][key]).to(dev, torch.float32) for n in names], dim=0)
        K = {1 + nodes[n]["features_rest"].shape[1] for n in names}
        if len(K) != 1:
            raise ValueError(f"nodes disagree on the number of SH bases: {sorted(K)}")
        self.means = torch.nn.Parameter(cat("means").contiguous())
        self.scales = torch.nn.Parameter(cat("scales").contiguous())
        self.quats = torch.nn.Parameter(cat("quats").contiguous())
        self.opacities = torch.nn.Parameter(cat("opacities", lambda t: t.reshape(-1)).contiguous())
        self.sh = torch.nn.Parameter(torch.cat([torch.cat([nodes[n]["features_dc"][:, None, :], nodes[n]["features_rest"]],
                                                          dim=1).to(dev, torch.float32) for n in names], dim=0).contiguous())
        self.slices: Dict[str, slice] = {}
        start = 0
        ids = []
        for i, n in enumerate(names):
            cnt = nodes[n]["means"].shape[0]
            self.slices[n] = slice(start, start + cnt)
            ids.append(torch.full((cnt,), i, dtype=torch.int32))
            start += cnt
        self.register_buffer("node_of", torch.cat(ids).to(dev), persistent=False)
        poses = torch.zeros(len(names), 16, dtype=torch.float32)
        poses[:, [0, 4, 8, 12]] = 1.0  # identity rotation, unit quaternion
        self.register_buffer("poses", poses.to(dev), persistent=False)
        self._index = {n: i for i, n in enumerate(names)}

    @property
    def num_points(self) -> int:
        return int(self.means.shape[0])

    def node(self, name: str) -> Dict[str, Tensor]:
        """Views of one node's raw rows (features_dc = sh[:, 0], features_rest = sh[:, 1:])."""
        s = self.slices[name]
        return dict(means=self.means[s], scales=self.scales[s], quats=self.quats[s], opacities=self.opacities[s],
                    features_dc=self.sh[s, 0], features_rest=self.sh[s, 1:])

    @torch.no_grad()
    def set_pose(self, name: str, quat: Tensor, trans: Tensor) -> None:
        """Pose of a rigid node for this frame (reference ``get_object_pose``: quaternion wxyz + translation)."""
        q = torch.as_tensor(quat, dtype=torch.float32, device=self.poses.device)
        row = torch.cat([_quat_to_rotmat(q).reshape(-1), torch.as_tensor(trans, dtype=torch.float32, device=q.device),
                         q / q.norm()])
        self.poses[self._index[name]] = row

    def activated(self, camera_to_world: Tensor, sh_degree: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
        """(means, quats, scales, opacities, colors) of the whole scene, ready for ``rasterization``."""
        _need_cuda(self.means)
        campos = camera_to_world[..., :3, 3].reshape(3).to(self.means.device, torch.float32).contiguous()
        return _ArenaActivate.apply(self.means, self.scales, self.quats, self.opacities, self.sh, self.node_of,
                                    self.poses, campos, int(sh_degree))

    def render(self, viewmat: Tensor, K: Tensor, width: int, height: int, sh_degree: int, **kwargs):
        """Activate + rasterize one camera.  ``viewmat``: world -> camera (OpenCV), [4,4]; kwargs go to ``rasterization``."""
        c2w = torch.linalg.inv(viewmat.detach())
        means, quats, scales, opac, colors = self.activated(c2w, sh_degree)
        kwargs.setdefault("packed", False)
        return rasterization(means, quats, scales, opac, colors, viewmat[None], K[None], width, height, **kwargs)
