"""ctypes binding of libpranetv2_b200.so (the C ABI declared in include/pv2.h).

There is NO fallback: if the shared library is missing (and cannot be built) or a call fails, the
product raises.  Nothing here touches `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_lib = None

c_void_pp = C.POINTER(C.c_void_p)
_i, _f, _p, _sz, _ll = C.c_int, C.c_float, C.c_void_p, C.c_size_t, C.c_longlong
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
# the 24 leading arguments shared by pv2_act_apply / pv2_bn_act_bwd (the forward description)
_APPLY = [_p, _i, _i, _i, _ll, _p, _p, _p, _i, _i, _i, _ll, _p, _p, _i, _p, _ll, _i, _i, _i, _i, _ll, _i, _i]

# name -> (restype, argtypes); mirrors include/pv2.h one to one
_SIGS = {
    "pv2_version": (_i, []),
    "pv2_last_error": (C.c_char_p, []),
    "pv2_launch_count": (C.c_ulonglong, []),
    "pv2_structure_loss_workspace_bytes": (_sz, [_i] * 4),
    "pv2_structure_loss_fwd": (_i, [c_void_pp, c_void_pp, _p, _p, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pv2_structure_loss_prepare": (_i, [_p, _i, _i, _i, _p, _sz, _p]),
    "pv2_structure_loss_fwd_prepared": (_i, [c_void_pp, c_void_pp, _p, _p, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pv2_structure_loss_bwd": (_i, [c_void_pp, c_void_pp, _p, _p, _p, c_void_pp, c_void_pp, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "pv2_structure_loss_lowres_workspace_bytes": (_sz, [_i] * 4),
    "pv2_structure_loss_lowres_fwd": (_i, [c_void_pp, c_void_pp, _ip, _ip, _fp, _fp, _p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pv2_structure_loss_lowres_bwd": (_i, [c_void_pp, c_void_pp, _ip, _ip, _fp, _fp, _p, _p, _p, c_void_pp, c_void_pp, _i, _i, _i, _i, _p, _sz, _p]),
    "pv2_bilinear_fwd": (_i, [_p, _p] + [_i] * 5 + [_f, _f, _i, _i, _p]),
    "pv2_bilinear_bwd": (_i, [_p, _p] + [_i] * 5 + [_f, _f, _i, _i, _p]),
    "pv2_bilinear_multi_fwd": (_i, [c_void_pp, c_void_pp, _ip, _ip, C.POINTER(C.c_float), C.POINTER(C.c_float)] + [_i] * 6 + [_p]),
    "pv2_bilinear_multi_bwd": (_i, [c_void_pp, c_void_pp, _ip, _ip, C.POINTER(C.c_float), C.POINTER(C.c_float)] + [_i] * 6 + [_p]),
    "pv2_dsra_fuse_fwd": (_i, [_p] * 4 + [_i] * 6 + [_f, _f, _i, _p]),
    "pv2_dsra_fuse_bwd": (_i, [_p] * 6 + [_i] * 6 + [_f, _f, _i, _p]),
    "pv2_ra_v1_scale_fwd": (_i, [_p] * 3 + [_i] * 4 + [_p]),
    "pv2_ra_v1_scale_bwd": (_i, [_p] * 5 + [_i] * 4 + [_p]),
    "pv2_mc_dual_loss_workspace_bytes": (_sz, [_i] * 4),
    "pv2_mc_dual_loss_fwd": (_i, [c_void_pp, c_void_pp, _p, _i, _i, _i, _i, _i, _i, _f, _f, _f, _p, _p, _sz, _p]),
    "pv2_mc_dual_loss_bwd": (_i, [c_void_pp, c_void_pp, _p, _p, c_void_pp, c_void_pp, _i, _i, _i, _i, _i, _i, _f, _f, _f, _p, _sz, _p]),
    # conv engine
    "pv2_conv_splits_hint": (_i, [_i] * 9),
    "pv2_conv_fwd": (_i, [_p, _ll, _p, _ll, _i, _i] + [_i] * 9 + [_i, _p, _i, _i, _p, _p, _p, _p]),
    "pv2_conv_sums_splits": (_i, []),
    "pv2_conv_set_cta_budget": (_i, [_i]),
    "pv2_bn_set_fused_grid": (_i, [_i]),
    "pv2_bn_fuse_workspace_floats": (_sz, [_ll, _i]),
    "pv2_conv_fuses_bn_stats": (_i, [_i, _i]),
    "pv2_bn_stats_group": (_i, [_p, _ll, _i, _ll, _i, _i, _p, _p]),
    "pv2_conv_wgrad_splits_hint": (_i, [_i] * 8),
    "pv2_conv_wgrad": (_i, [_p, _ll, _p, _ll, _i, _i] + [_i] * 10 + [_p, _i, _p]),
    "pv2_weight_pack": (_i, [_p, _p, _ll, _i, _i] + [_i] * 8 + [_p]),
    "pv2_wgrad_unpack": (_i, [_p, _ll, _i, _p] + [_i] * 6 + [_p]),
    "pv2_weight_pack_multi": (_i, [_p, _i, _i, _i, _p]),
    "pv2_wgrad_unpack_multi": (_i, [_p, _i, _p]),
    "pv2_pack_nchw": (_i, [_p, _i, _p, _ll, _i, _i] + [_i] * 5 + [_p]),
    "pv2_unpack_to_nchw": (_i, [c_void_pp, _ip, _ip, _i, _p, _i, _i, _i, _i, _i, _p]),
    "pv2_bn_workspace_floats": (_sz, [_ll, _i]),
    "pv2_bn_stats": (_i, [_p, _ll, _i, _ll, _i, _i, _p, _p, _f, _f] + [_p] * 8 + [_p]),
    "pv2_bn_eval_affine": (_i, [_i, _p, _p, _p, _p, _f, _p, _p, _p]),
    "pv2_act_apply": (_i, _APPLY + [_p, _ll, _i, _i, _i, _i, _p, _p, _i, _p]),
    "pv2_bn_act_bwd": (_i, _APPLY + [c_void_pp, _ip, _ip, _i, _p] + [_p] * 4 + [_i, _p, _i] + [_p, _ll, _i, _i] * 2 + [_p] * 6 + [_i, _p]),
    "pv2_up2_nhwc_fwd": (_i, [_p, _ll, _i, _i, _i, _p, _ll, _i, _i, _i] + [_i] * 5 + [_p]),
    "pv2_up2_nhwc_bwd": (_i, [c_void_pp, _ip, _ip, _i, _p, _i] + [_i] * 4 + [_p]),
    # inference tails
    "pv2_infer_tail_workspace_bytes": (_sz, [_i]),
    "pv2_infer_tail_binary": (_i, [c_void_pp, _ip, _ip, _fp, _fp] + [_i] * 6 + [_f, _f, _p, _p, _sz, _p]),
    "pv2_infer_tail_argmax": (_i, [c_void_pp, c_void_pp, _ip, _ip, _fp, _fp] + [_i] * 5 + [_p, _p]),
    # optimizer tail
    "pv2_adam_clamp_flat": (_i, [_p] * 4 + [_ll, _p, _p] + [_f] * 5 + [_i, _f, _f, _p]),
}


def declared_symbols():
    return list(_SIGS)


def library_path() -> str:
    return _build.LIB


def _have_nvcc() -> bool:
    try:
        _build._nvcc()
        return True
    except RuntimeError:
        return False


def load():
    """Load the shared library, (re)building it first when the sources are newer and nvcc is here.
    Raises if the library cannot be had -- there is no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.stale():
        if _have_nvcc():
            _build.build_library()
        elif not os.path.exists(path):
            raise RuntimeError(f"{path} is missing and nvcc is not available to build it")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export what pv2.h declares
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().pv2_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def launch_count() -> int:
    return int(load().pv2_launch_count())


def int_array(vals):
    arr = (C.c_int * len(vals))(*[int(v) for v in vals])
    return C.cast(arr, _ip), arr


def float_array(vals):
    arr = (C.c_float * len(vals))(*[float(v) for v in vals])
    return C.cast(arr, _fp), arr


def ptr_array(tensors):
    """Host array of device pointers; keep the second return value alive across the call."""
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return C.cast(arr, c_void_pp), arr
