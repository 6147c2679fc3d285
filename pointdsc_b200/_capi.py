"""ctypes binding of the C ABI in include/pointdsc_b200.h.

This is the whole "FFI" a maintainer of the reference would add (the reference is pure Python and
has none of its own): load the shared library, mirror the two structs, declare the entry points.
There is deliberately NO fallback: if the library is missing or no B200 is present the import /
engine creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POINTDSC_B200_LIB") or os.path.join(_HERE, "libpointdsc_b200.so")   # env: developer A/B builds

PRECISIONS = {"fp32": 0, "bf16x3": 1, "bf16": 2, "fp16x3": 3}
SPANS = ["sc", "linear", "attention", "head", "seeds", "knn", "nsm", "hypotheses", "refine", "total"]


class PdscError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("in_dim", C.c_int32), ("num_layers", C.c_int32), ("num_channels", C.c_int32),
        ("num_iterations", C.c_int32), ("ratio", C.c_float), ("inlier_threshold", C.c_float),
        ("sigma_d", C.c_float), ("k", C.c_int32), ("nms_radius", C.c_float),
        ("precision", C.c_int32), ("device", C.c_int32),
    ]


class StageIO(C.Structure):
    _fields_ = [
        ("in_features", C.c_void_p), ("in_confidence", C.c_void_p), ("in_seeds", C.c_void_p),
        ("in_knn_idx", C.c_void_p), ("in_seed_trans", C.c_void_p),
        ("out_sc", C.c_void_p), ("out_features", C.c_void_p), ("out_normed", C.c_void_p),
        ("out_confidence", C.c_void_p), ("out_seeds", C.c_void_p), ("out_knn_idx", C.c_void_p),
        ("out_compat", C.c_void_p), ("out_eig", C.c_void_p), ("out_power_iters", C.c_void_p),
        ("out_seed_trans", C.c_void_p), ("out_inlier_counts", C.c_void_p), ("out_best", C.c_void_p),
        ("out_init_trans", C.c_void_p), ("out_refine_solves", C.c_void_p),
        ("layer_tap", C.c_int32), ("out_layer_features", C.c_void_p), ("out_layer_debug", C.c_void_p), ("out_timeline", C.c_void_p),
    ]


# every symbol include/pointdsc_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pdsc_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "pdsc_destroy": (C.c_int, [C.c_void_p]),
    "pdsc_last_error": (C.c_char_p, []),
    "pdsc_version": (C.c_char_p, []),
    "pdsc_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "pdsc_commit_params": (C.c_int, [C.c_void_p]),
    "pdsc_set_precision": (C.c_int, [C.c_void_p, C.c_int32]),
    "pdsc_num_seeds": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pdsc_num_neighbours": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pdsc_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    "pdsc_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.POINTER(StageIO), C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdsc_forward_graph": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdsc_forward_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "pdsc_forward_host_submit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "pdsc_forward_host_wait": (C.c_int, [C.c_void_p, C.c_int32]),
    "pdsc_forward_eval": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(StageIO), C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdsc_eval_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "pdsc_leading_eigenvector_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "pdsc_leading_eigenvector": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdsc_match_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "pdsc_match": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                             C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_size_t, C.c_void_p]),
    "pdsc_voxel_down_sample_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "pdsc_voxel_down_sample": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdsc_fpfh_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "pdsc_estimate_normals": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdsc_compute_fpfh": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdsc_read_ply": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "pdsc_launches_per_forward": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32]),
    "pdsc_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "pdsc_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
}

_lib = None


def load():
    """Load libpointdsc_b200.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PdscError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(pointdsc_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


_util_engines = {}


def utility_engine(device_index: int):
    """An engine handle for the stateless entry points (pdsc_match, pdsc_eval_stats): they only need the device."""
    if device_index not in _util_engines:
        lib = load()
        cfg = Config(6, 12, 128, 10, 0.1, 0.1, 0.1, 40, 0.1, PRECISIONS["fp16x3"], device_index)
        handle = C.c_void_p()
        check(lib.pdsc_create(C.byref(cfg), C.byref(handle)))
        _util_engines[device_index] = handle
    return _util_engines[device_index]


def check(rc: int):
    if rc != 0:
        raise PdscError(f"pointdsc_b200 error {rc}: {load().pdsc_last_error().decode()}")
