"""Device-side correspondence front end.

Mirrors the lines every caller of `PointDSC.forward` runs first (reference datasets/ThreeDMatch.py:283-291 and :299-308,
datasets/KITTI.py:80-114, demo_registration.py:101-108): nearest neighbour in descriptor space, optional mutual check,
and the centred `corr_pos`.  One call = one pair; the outputs are device tensors in the layout the module consumes, so
the correspondences never visit the host (the one 4-byte read is the correspondence count, which fixes the shapes).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _capi


@torch.no_grad()
def match(src_desc: torch.Tensor, tgt_desc: torch.Tensor, src_keypts: torch.Tensor, tgt_keypts: torch.Tensor,
          use_mutual: bool = False) -> Dict[str, torch.Tensor]:
    """-> {'corr' [M,2] int64 (source, target), 'corr_pos' [1,M,6], 'src_keypts' [1,M,3], 'tgt_keypts' [1,M,3]} on the device.

    Descriptors are L2-normalised rows, float32 (FCGF) or float64 (FPFH); the arithmetic runs in their dtype, as numpy's
    does in the reference.  Ready for `model({'corr_pos': ..., 'src_keypts': ..., 'tgt_keypts': ..., 'testing': True})`."""
    if src_desc.device.type != "cuda":
        raise _capi.PdscError("pointdsc_b200.frontend.match runs on a B200 only: pass CUDA tensors (there is no CPU fallback)")
    if src_desc.dtype != tgt_desc.dtype or src_desc.dtype not in (torch.float32, torch.float64):
        raise ValueError("descriptors must both be float32 or both float64")
    if src_desc.dim() != 2 or tgt_desc.dim() != 2 or src_desc.shape[1] != tgt_desc.shape[1]:
        raise ValueError(f"expected descriptors [Ns,D] and [Nt,D], got {tuple(src_desc.shape)}, {tuple(tgt_desc.shape)}")
    dev = src_desc.device
    lib = _capi.load()
    engine = _capi.utility_engine(dev.index if dev.index is not None else torch.cuda.current_device())
    ns, nt, d = int(src_desc.shape[0]), int(tgt_desc.shape[0]), int(src_desc.shape[1])
    sd, td = src_desc.contiguous(), tgt_desc.contiguous()
    sk = src_keypts.to(device=dev, dtype=torch.float32).contiguous()
    tk = tgt_keypts.to(device=dev, dtype=torch.float32).contiguous()
    if sk.shape != (ns, 3) or tk.shape != (nt, 3):
        raise ValueError("key points must be [Ns,3] and [Nt,3]")
    corr = torch.empty(ns, 2, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    corr_pos = torch.empty(ns, 6, dtype=torch.float32, device=dev)
    out_src = torch.empty(ns, 3, dtype=torch.float32, device=dev)
    out_tgt = torch.empty(ns, 3, dtype=torch.float32, device=dev)
    scratch = torch.empty(int(lib.pdsc_match_scratch_bytes(ns, nt)) + 8, dtype=torch.uint8, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _capi.check(lib.pdsc_match(engine, ns, nt, d, C.c_void_p(sd.data_ptr()), C.c_void_p(td.data_ptr()),
                                   1 if sd.dtype == torch.float64 else 0, C.c_void_p(sk.data_ptr()), C.c_void_p(tk.data_ptr()),
                                   1 if use_mutual else 0, C.c_void_p(corr.data_ptr()), C.c_void_p(count.data_ptr()),
                                   C.c_void_p(corr_pos.data_ptr()), C.c_void_p(out_src.data_ptr()), C.c_void_p(out_tgt.data_ptr()),
                                   C.c_void_p((scratch.data_ptr() + 7) // 8 * 8), scratch.numel() - 8, stream))
    m = int(count.item())      # the only host read: it fixes the output shapes
    return {"corr": corr[:m].long(), "corr_pos": corr_pos[:m][None], "src_keypts": out_src[:m][None], "tgt_keypts": out_tgt[:m][None]}
