"""pointdsc_b200 — B200-native engine for PointDSC's testing-mode forward path.

`PointDSC` mirrors the reference's `models.PointDSC.PointDSC` (constructor, forward(dict)->dict, state_dict
keys); the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in include/pointdsc_b200.h.
"""
from ._capi import LIB_PATH, PdscError  # noqa: F401
from .model import PointDSC  # noqa: F401

__all__ = ["PointDSC", "PdscError", "LIB_PATH"]
