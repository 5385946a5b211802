"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
the ctypes signature table covers the header, and the host-side mirror of the reference interface behaves
(argument validation, error conventions, module aliasing).  No compute kernels are launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200splat.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from mtgs_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libb200splat.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200splat.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), set(_lib.SIGNATURES) ^ set(syms)
    loaded = _lib.load()
    assert loaded.b2s_version() >= 100
    assert b"unsupported" in loaded.b2s_error_string(-2)
    assert isinstance(_lib.launch_count(), int)


def test_ctypes_tables_mirror_the_header_argument_by_argument():
    """Every prototype of include/b200splat.h has as many parameters as its ctypes argtypes entry in mtgs_b200/_lib.py
    (a missing or extra argument would shift every later pointer without any error from ctypes)."""
    from mtgs_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "b200splat.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    protos = dict(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S))
    assert set(protos) == set(_lib.SIGNATURES)
    def kind_h(a):
        a = " ".join(a.split())
        if "*" in a or "[" in a or "b2s_stream_t" in a:
            return "ptr"
        if a.startswith("float") or a.startswith("const float"):
            return "f32"
        if "long long" in a or "int64_t" in a or "size_t" in a:
            return "i64"
        if "uint32_t" in a or "unsigned" in a:
            return "u32"
        return "i32" if "int" in a else "?" + a

    def kind_c(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "ptr"
        return {ctypes.c_float: "f32", ctypes.c_longlong: "i64", ctypes.c_size_t: "i64", ctypes.c_ulonglong: "i64",
                ctypes.c_uint32: "u32", ctypes.c_int: "i32"}.get(t, "?" + str(t))

    for name, args in protos.items():
        args = " ".join(args.split())
        alist = [] if args in ("", "void") else args.split(",")
        table = _lib.SIGNATURES[name][1]
        assert len(alist) == len(table), (name, len(alist), len(table))
        for i, (a, t) in enumerate(zip(alist, table)):  # ... and the same kind of argument in every position
            assert kind_h(a) == kind_c(t), (name, i, a.strip(), t)


def test_header_cites_reference_call_sites():
    txt = open(os.path.join(ROOT, "include", "b200splat.h")).read()
    assert "mtgs_scene_graph.py:21, 641-662" in txt and "vanilla_gaussian_splatting.py:16, 317" in txt


def test_sm100a_code_is_embedded():
    """The shared library carries sm_100a SASS (not PTX-JIT for another arch)."""
    import subprocess
    from mtgs_b200 import _lib
    try:
        out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True, timeout=60)
    except FileNotFoundError:
        pytest.skip("cuobjdump not on PATH")
    assert "sm_100a" in out.stdout, out.stdout[:400]


def test_signature_mirror_and_error_conventions():
    import inspect
    from mtgs_b200.rendering import rasterization, padded_channels
    from mtgs_b200.cuda._wrapper import spherical_harmonics
    p = inspect.signature(rasterization).parameters
    # upstream gsplat v1.4.0 names, order and defaults
    names = ["means", "quats", "scales", "opacities", "colors", "viewmats", "Ks", "width", "height", "near_plane",
             "far_plane", "radius_clip", "eps2d", "sh_degree", "packed", "tile_size", "backgrounds", "render_mode",
             "sparse_grad", "absgrad", "rasterize_mode", "channel_chunk", "distributed", "camera_model", "covars"]
    assert list(p) == names
    assert p["near_plane"].default == 0.01 and p["far_plane"].default == 1e10 and p["eps2d"].default == 0.3
    assert p["packed"].default is True and p["tile_size"].default == 16 and p["render_mode"].default == "RGB"
    assert list(inspect.signature(spherical_harmonics).parameters) == ["degrees_to_use", "dirs", "coeffs", "masks"]
    assert [padded_channels(c) for c in (1, 3, 4, 5, 7, 8)] == [4, 4, 4, 8, 8, 8]
    with pytest.raises(NotImplementedError):
        padded_channels(9)
    # CPU tensors: loud failure, no fallback
    z = torch.zeros
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rasterization(z(4, 3), z(4, 4), z(4, 3), z(4), z(4, 3), torch.eye(4)[None], torch.eye(3)[None], 32, 32,
                      packed=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        spherical_harmonics(1, z(4, 3), z(4, 4, 3))
    with pytest.raises(AssertionError):
        spherical_harmonics(3, z(4, 3), z(4, 4, 3))  # (3+1)^2 > K, same assert as upstream


def test_gsplat_alias_resolves_mtgs_imports():
    import sys
    import mtgs_b200
    mtgs_b200.install_as_gsplat()
    from gsplat.rendering import rasterization  # mtgs_scene_graph.py:21
    from gsplat.cuda._wrapper import spherical_harmonics  # vanilla_gaussian_splatting.py:16
    assert rasterization.__module__ == "mtgs_b200.rendering"
    assert spherical_harmonics.__module__ == "mtgs_b200.cuda._wrapper"
    assert sys.modules["gsplat"].__b200__


def test_product_path_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under mtgs_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "mtgs_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("test infrastructure", ""), f"{f} mentions the oracle"
                assert "splat_oracle" not in src and "cpu_ref" not in src and "torch_ref" not in src


def test_scene_generators_are_deterministic():
    from mtgs_b200 import scenes
    a, b = scenes.street(n=1000, seed=1), scenes.street(n=1000, seed=1)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    c = scenes.config1()
    assert c["means"].shape == (10_000, 3) and (c["means"][:, 2] < 0.01).mean() > 0.05
    q = a["quats"]
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-5)
    # different traversal cameras share the Gaussians but not the pose
    d = scenes.street(n=1000, seed=1, camera=1)
    np.testing.assert_array_equal(a["means"], d["means"])
    assert np.abs(a["viewmat"] - d["viewmat"]).max() > 1e-3


def test_exchange_shard_layout():
    """Row ownership of the multi-GPU gradient exchange (pure host arithmetic of the C ABI; no GPU needed):
    shards are multiples of 256 rows (one projection-backward CTA never straddles two owners) and cover n_shared."""
    from mtgs_b200 import _lib
    lib = _lib.load()
    for n, w in [(0, 1), (1, 1), (255, 2), (256, 2), (257, 2), (2_000_000, 8), (3_000_001, 7), (500_000, 4)]:
        s = lib.b2s_exchange_shard_rows(n, w)
        assert s > 0 and s % 256 == 0
        assert s * w >= n
        assert (s - 256) * w < max(n, 1) + 256 * w  # not more than one CTA of slack per rank
    assert lib.b2s_exchange_shard_rows(10, 0) < 0 and lib.b2s_exchange_shard_rows(10, 9) < 0  # world in 1..8
    assert lib.b2s_exchange_shard_rows(-1, 2) < 0


def test_grad_exchange_needs_cuda_and_never_falls_back():
    import torch
    from mtgs_b200.parallel import GradExchange, current_exchange
    assert current_exchange() is None
    if not torch.cuda.is_available():
        import pytest
        with pytest.raises(Exception):
            GradExchange(n_shared=1024, d_in=3)


def test_tile_list_workspace_sizing_is_host_only_and_monotone():
    """b2s_bin_tiles_workspace_bytes is pure host arithmetic over the five list-size totals (M, S, E1, E3, n_vis):
    callable without a GPU, monotone in every total, 0 for unsupported tile grids (more than 512 tiles per axis
    would need more than 32 children per list)."""
    import ctypes as C
    from mtgs_b200 import _lib
    lib = _lib.load()

    def ws(tot, tw, th):
        return int(lib.b2s_bin_tiles_workspace_bytes((C.c_longlong * 5)(*tot), tw, th))

    base = [19_184_915, 3_727_319, 1_500_000, 4_700_000, 1_206_176]  # the bench workload's totals (1080p)
    w0 = ws(base, 120, 68)
    assert 0 < w0 < 1 << 31
    for k in range(5):
        bigger = list(base)
        bigger[k] *= 2
        assert ws(bigger, 120, 68) >= w0
    assert ws(base, 240, 135) > 0          # 4K: 16-row groups, 15 column groups
    assert ws(base, 275, 19) > 0 and ws(base, 19, 275) > 0   # > 256 tiles along one axis
    assert ws([0, 0, 0, 0, 0], 4, 3) > 0   # empty scene still has offset tables
    assert ws(base, 2000, 68) == 0 and ws(base, 120, 2000) == 0 and ws(base, 0, 68) == 0
    assert ws([-1, 0, 0, 0, 0], 120, 68) == 0
    assert int(lib.b2s_bin_depth_workspace_bytes(0)) > 0
    assert int(lib.b2s_bin_depth_workspace_bytes(3_000_000)) > int(lib.b2s_bin_depth_workspace_bytes(1_000_000))


def test_padded_channels_and_lazy_meta():
    import pytest
    from mtgs_b200.rendering import Meta, padded_channels
    assert [padded_channels(c) for c in range(1, 9)] == [4, 4, 4, 4, 8, 8, 8, 8]  # MTGS uses 3, 4, 6, 7 (Appendix B)
    with pytest.raises(NotImplementedError):
        padded_channels(9)
    m = Meta(a=1)
    calls = []
    # upstream's lists are built together, once, on the first access to any of the three keys
    m._lazy = lambda key: calls.append(key) or dict(flatten_ids="flat", isect_offsets="offs", isect_ids="ids")
    assert "isect_ids" in m and "flatten_ids" in m and "nothing" not in m
    assert m.get("isect_ids") == "ids" and m["isect_ids"] == "ids" and m["flatten_ids"] == "flat"
    assert m["isect_offsets"] == "offs" and calls == ["isect_ids"]
    assert m.get("missing", 7) == 7
    with pytest.raises(KeyError):
        m["missing"]


def test_ssim_alias_resolves_the_mtgs_import():
    import importlib
    import sys
    import mtgs_b200
    mtgs_b200.install_as_mtgs_ssim()
    mod = importlib.import_module("mtgs.utils.ssim") if "mtgs" in sys.modules else sys.modules["mtgs.utils.ssim"]
    assert mod.MaskedSSIM.__module__ == "mtgs_b200.ssim" and mod.ssim.__module__ == "mtgs_b200.ssim"
    mtgs_b200.install_as_mtgs_ssim()  # idempotent
