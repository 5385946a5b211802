"""ctypes binding of ``libb200splat.so`` (the C ABI declared in ``include/b200splat.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing the import
fails loudly.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``mtgs_b200/csrc/build.sh`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200splat.so")

_vp, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); mirrors include/b200splat.h one to one
SIGNATURES = {
    "b2s_version": (_i, []),
    "b2s_error_string": (C.c_char_p, [_i]),
    "b2s_launch_count": (_ll, []),
    "b2s_project_fwd": (_i, [_vp] * 7 + [_i] * 6 + [_f] * 4 + [_i] * 4 + [_vp] * 12 + [_vp]),
    "b2s_project_bwd": (_i, [_vp] * 6 + [_i] * 3 + [_f] + [_i] * 4 + [_vp] * 4 + [_i] + [_vp] * 8 + [_vp]),
    "b2s_exchange_shard_rows": (_i, [_i, _i]),
    "b2s_peer_alloc": (_i, [_sz, C.POINTER(C.c_void_p)]),
    "b2s_peer_free": (_i, [_vp]),
    "b2s_ipc_export": (_i, [_vp, C.c_char_p]),
    "b2s_ipc_import": (_i, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "b2s_ipc_close": (_i, [_vp]),
    "b2s_project_bwd_exchange": (_i, [_vp] * 6 + [_i] * 3 + [_f] + [_i] * 4 + [_vp] * 4 + [_i] + [_vp] * 3 +
                                 [_i, _i, _i, _i, _ll, _f, C.c_uint, _i, _f] + [_vp] * 4 + [_vp]),
    "b2s_bin_rect_totals": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "b2s_bin_depth_workspace_bytes": (_sz, [_i]),
    "b2s_bin_sort_depth": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "b2s_debug_sort_depth_phases": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _vp, _vp]),
    "b2s_bin_tiles_workspace_bytes": (_sz, [C.POINTER(_ll), _i, _i]),
    "b2s_bin_tiles": (_i, [_vp] * 3 + [C.POINTER(_ll)] + [_i] * 6 + [_vp, _vp, _i, _vp, _i] + [_vp] * 3 + [_sz, _vp]),
    "b2s_bin_tiles_l3_view": (_i, [C.POINTER(_ll), _i, _i, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i), C.POINTER(_i),
                                   C.POINTER(_i)]),
    "b2s_bin_isect_ids": (_i, [_vp, _i, _vp, _vp, _ll, _vp, _vp]),
    "b2s_blend_record_bytes": (_sz, [_ll, _i, _i]),
    "b2s_blend_record_blocks": (C.c_uint32, [_ll, _i]),
    "b2s_blend_fwd": (_i, [_vp] * 5 + [_i] * 9 + [_vp] * 4 + [C.c_uint32] + [_vp] * 3 + [_vp]),
    "b2s_blend_bwd": (_i, [_vp] * 2 + [_i] * 7 + [_vp] * 8 + [_i, _vp, _vp]),
    "b2s_ssim_fwd": (_i, [_vp] * 3 + [_ll, _ll] + [_i] * 4 + [_vp, _i, _f, _f] + [_vp] * 5 + [_vp]),
    "b2s_ssim_bwd": (_i, [_vp] * 6 + [_i] * 4 + [_vp, _i, _vp] + [_vp]),
    "b2s_masked_l1_fwd": (_i, [_vp] * 3 + [_ll, _i, _i, _f, _vp, _vp]),
    "b2s_masked_l1_bwd": (_i, [_vp] * 3 + [_ll, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "b2s_tv_fwd": (_i, [_vp] + [_i] * 4 + [_vp, _vp]),
    "b2s_tv_bwd": (_i, [_vp] + [_i] * 4 + [_vp, _vp, _vp]),
    "b2s_ncc_patch_grid": (_i, [_i] * 4 + [C.POINTER(_i), C.POINTER(_i)]),
    "b2s_ncc_fwd": (_i, [_vp] * 3 + [_i] * 4 + [_vp, _vp, _vp]),
    "b2s_ncc_bwd": (_i, [_vp] * 2 + [_i] * 4 + [_vp] * 4 + [_vp]),
    "b2s_normal_from_depth": (_i, [_vp, _i, _i] + [_f] * 4 + [_vp, _vp, _vp]),
    "b2s_arena_fwd": (_i, [_vp] * 8 + [_i] * 3 + [_vp] * 6 + [_vp]),
    "b2s_arena_bwd": (_i, [_vp] * 8 + [_i] * 3 + [_vp] * 11 + [_vp]),
    "b2s_adam_chunk": (_i, []),
    "b2s_adam_multi": (_i, [_vp, _vp, _vp, _i, _vp]),
    "b2s_densify_stats": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "b2s_mask_workspace_bytes": (_sz, [_ll]),
    "b2s_mask_scan": (_i, [_vp, _ll, _vp, _sz, _vp, _vp]),
    "b2s_mask_gather_rows": (_i, [_vp, _ll, _vp, _vp, _vp, _i, _vp]),
    "b2s_sh_fwd": (_i, [_i] + [_vp] * 3 + [_i, _i] + [_vp, _vp]),
    "b2s_sh_bwd": (_i, [_i] + [_vp] * 3 + [_i, _i] + [_vp] * 3 + [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library once; raise ImportError (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the B200 CUDA extension is not built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "mtgs_b200 has no CPU or PyTorch fallback path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the .so disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b2s_error_string(rc)
        text = f"{what} failed: {msg.decode() if msg else rc} (code {rc})"
        if rc == -2:  # B2S_ERR_UNSUPPORTED: a configuration this build does not instantiate
            raise NotImplementedError(text)
        raise RuntimeError(text)


def launch_count() -> int:
    return int(load().b2s_launch_count())
