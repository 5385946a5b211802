"""Row f4: checkpoint / scene readers.  The reference's own loaders need nerfstudio (not installed) and no released
checkpoint or prepared road block exists in the containers, so the fixtures are built here following the documented
schemas (custom_trainer.py:137-170, mtgs_scene_graph.py:1185-1216, docs/prepare_dataset.md:104-190); the activation
helper is checked against the reference's importable utilities where they exist."""
import datetime
import os
import pickle

import numpy as np
import pytest
import torch

from mtgs_b200 import io as mio

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fake_pipeline_state(n_bg=50, n_car=7, T=3):
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    st = {}
    for name, n, multi in (("background", n_bg, True), ("object_car_a", n_car, False)):
        p = f"_model.gaussian_models.{name}."
        st[p + "gauss_params.means"] = r(n, 3)
        st[p + "gauss_params.scales"] = r(n, 3)
        st[p + "gauss_params.quats"] = r(n, 4)
        st[p + "gauss_params.opacities"] = r(n, 1)
        st[p + "gauss_params.features_dc"] = r(n, 3)
        st[p + "gauss_params.features_rest"] = r(n, T, 15, 3) if multi else r(n, 15, 3)
        if multi:
            st[p + "gauss_params.features_adapters"] = r(n, T, 3)
    st["_model.gaussian_models.object_car_a.instance_trans"] = r(20, 3)
    st["_model.gaussian_models.object_car_a.instance_quats"] = r(20, 4)
    st["_model.camera_optimizer.pose_adjustment"] = r(40, 6)
    return st


def test_checkpoint_round_trip_and_split(tmp_path):
    st = _fake_pipeline_state()
    path = tmp_path / "step-000030000.ckpt"
    torch.save({"step": 30000, "pipeline": st, "optimizers": {}}, path)
    ck = mio.load_checkpoint(str(path))
    assert ck.step == 30000 and set(ck.nodes) == {"background", "object_car_a"}
    assert ck.num_gaussians == 57
    assert torch.equal(ck.nodes["object_car_a"]["instance_trans"], st["_model.gaussian_models.object_car_a.instance_trans"])
    assert "camera_optimizer.pose_adjustment" in ck.other
    gp = ck.gauss_params("background")
    assert set(gp) == {"means", "scales", "quats", "opacities", "features_dc", "features_rest", "features_adapters"}
    # DDP-wrapped pipelines prefix everything with "module."
    nodes, _ = mio.split_scene_state_dict({"module." + k: v for k, v in st.items()})
    assert set(nodes) == {"background", "object_car_a"}
    # write in the trainer's layout and read back
    p2 = tmp_path / "again.ckpt"
    mio.save_checkpoint(str(p2), 7, ck.nodes, ck.other)
    ck2 = mio.load_checkpoint(str(p2))
    assert ck2.step == 7 and all(torch.equal(ck2.nodes[n][k], ck.nodes[n][k]) for n in ck.nodes for k in ck.nodes[n])
    with pytest.raises(KeyError):
        torch.save({"foo": 1}, tmp_path / "bad.ckpt")
        mio.load_checkpoint(str(tmp_path / "bad.ckpt"))


def test_node_activations_follow_the_reference():
    ck_nodes, _ = mio.split_scene_state_dict(_fake_pipeline_state())
    gp = {k[len("gauss_params."):]: v for k, v in ck_nodes["background"].items() if k.startswith("gauss_params.")}
    out = mio.node_rasterizer_inputs(gp, traversal=1)
    assert out["sh_coeffs"].shape == (50, 16, 3) and out["opacities"].shape == (50,)
    assert torch.allclose(out["quats"].norm(dim=-1), torch.ones(50), atol=1e-6)
    assert torch.equal(out["scales"], torch.exp(gp["scales"]))
    assert torch.equal(out["sh_coeffs"][:, 0], gp["features_dc"] + gp["features_adapters"][:, 1])
    assert torch.equal(out["sh_coeffs"][:, 1:], gp["features_rest"][:, 1])
    with pytest.raises(ValueError):
        mio.node_rasterizer_inputs(gp)
    # the quaternion convention (w, x, y, z) against vectors produced by the reference's own quat_to_rotmat
    g = np.load(os.path.join(GOLD, "ref_utils_golden.npz"))
    key_q = [k for k in g.files if "quat" in k and "rot" not in k][0]
    key_r = [k for k in g.files if "rot" in k][0]
    for q, R in zip(g[key_q][:16], g[key_r][:16]):
        np.testing.assert_allclose(mio.quat_wxyz_to_rotmat(q), R, rtol=1e-5, atol=1e-6)


def _fake_video_scene():
    frames = []
    for i in range(4):
        e2g = np.eye(4)
        e2g[:3, 3] = [10.0 + i, 5.0, 0.3]
        cams = {"CAM_F0": dict(data_path=f"log/CAM_F0/{i:04d}.jpg", sensor2ego_rotation=[0.5, -0.5, 0.5, -0.5],
                               sensor2ego_translation=[1.6, 0.0, 1.5], cam_intrinsic=np.array([[1545.0, 0, 960], [0, 1545.0, 560], [0, 0, 1]]),
                               distortion=np.zeros(5), token=f"c{i}", timestamp=1000 + i),
                "CAM_L0": dict(data_path=f"log/CAM_L0/{i:04d}.jpg", sensor2ego_rotation=[1.0, 0, 0, 0],
                               sensor2ego_translation=[0, 0.5, 1.5], cam_intrinsic=np.eye(3), distortion=np.zeros(5),
                               token=f"l{i}", timestamp=1000 + i)}
        frames.append(dict(token=f"f{i}", frame_idx=i, skipped="low_velocity" if i == 2 else False, timestamp=1000 + i,
                           ego2global=e2g, cams=cams, gt_boxes=np.zeros((0, 7)), track_tokens=[]))
    return {"road_block-x-0": dict(video_token="road_block-x-0", date=datetime.date(2021, 5, 12), trajectory=np.zeros((4, 3)),
                                   frame_infos=frames),
            "road_block-x-7": dict(video_token="road_block-x-7", date=datetime.date(2021, 6, 1), trajectory=np.zeros((4, 3)),
                                   frame_infos=frames[:2])}


def test_video_scene_reader(tmp_path):
    p = tmp_path / "video_scene_dict.pkl"
    with open(p, "wb") as f:
        pickle.dump(_fake_video_scene(), f)
    vs = mio.load_video_scene_dict(str(p))
    recs = mio.cameras_from_video_scene(vs, cameras=("CAM_F0",))
    assert len(recs) == 3 + 2 and {r["travel_id"] for r in recs} == {0, 7}     # the skipped frame is dropped
    assert len(mio.cameras_from_video_scene(vs, cameras=("CAM_F0", "CAM_L0"), travels=[7])) == 4
    r0 = recs[0]
    # camera centre = ego position + sensor offset; the camera looks along the ego's +x (nuPlan front camera)
    c2w = np.linalg.inv(r0["viewmat"])
    np.testing.assert_allclose(c2w[:3, 3], [11.6, 5.0, 1.8], atol=1e-9)
    np.testing.assert_allclose(c2w[:3, 2], [1.0, 0.0, 0.0], atol=1e-9)       # OpenCV +z (forward) = world +x
    np.testing.assert_allclose(r0["K"][0, 0], 1545.0)
    rel = mio.cameras_from_video_scene(vs, origin=[10.0, 5.0, 0.0])[0]
    np.testing.assert_allclose(np.linalg.inv(rel["viewmat"])[:3, 3], [1.6, 0.0, 1.8], atol=1e-9)

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    with open(tmp_path / "evil.pkl", "wb") as f:
        pickle.dump({"a": Evil()}, f)
    with pytest.raises(pickle.UnpicklingError):
        mio.load_video_scene_dict(str(tmp_path / "evil.pkl"))
