"""Device-side per-pair evaluation statistics.

Replaces, for a whole batch in one launch and without a host synchronisation, what the evaluation drivers compute per pair
with libs/loss.py:34-63 (TransformationLoss: RE / TE / success / RMSE) and libs/loss.py:94-100 (ClassificationLoss:
precision / recall / F1 through scikit-learn on the host) — evaluation/test_3DMatch.py:83-101.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi

COLUMNS = ("success", "re_deg", "te_cm", "gt_inliers", "gt_inlier_ratio", "kept_gt_inliers", "precision", "recall", "f1", "rmse")


@torch.no_grad()
def eval_stats(pred_trans: torch.Tensor, gt_trans: torch.Tensor, src_keypts: torch.Tensor, tgt_keypts: torch.Tensor,
               pred_labels: torch.Tensor, gt_labels: torch.Tensor, re_thre: float = 15.0, te_thre: float = 30.0) -> torch.Tensor:
    """[B,4,4] x2, [B,N,3] x2, [B,N] x2 (device)  ->  [B,10] device tensor, columns = COLUMNS.
    Thresholds as the drivers pass them: 3DMatch 15 deg / 30 cm (test_3DMatch.py), KITTI 5 deg / 60 cm (test_KITTI.py)."""
    if pred_trans.device.type != "cuda":
        raise _capi.PdscError("pointdsc_b200.metrics.eval_stats runs on a B200 only: pass CUDA tensors (there is no CPU fallback)")
    dev = pred_trans.device
    b, n = int(src_keypts.shape[0]), int(src_keypts.shape[1])
    f = lambda x: x.to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    pt, gt, s, t, pl, gl = f(pred_trans), f(gt_trans), f(src_keypts), f(tgt_keypts), f(pred_labels), f(gt_labels)
    if pt.shape != (b, 4, 4) or gt.shape != (b, 4, 4) or t.shape != (b, n, 3) or pl.shape != (b, n) or gl.shape != (b, n):
        raise ValueError("expected trans [B,4,4], key points [B,N,3], labels [B,N]")
    out = torch.empty(b, 10, dtype=torch.float32, device=dev)
    lib = _capi.load()
    engine = _capi.utility_engine(dev.index if dev.index is not None else torch.cuda.current_device())
    with torch.cuda.device(dev):
        _capi.check(lib.pdsc_eval_stats(engine, b, n, C.c_void_p(pt.data_ptr()), C.c_void_p(gt.data_ptr()), C.c_void_p(s.data_ptr()),
                                        C.c_void_p(t.data_ptr()), C.c_void_p(pl.data_ptr()), C.c_void_p(gl.data_ptr()),
                                        float(re_thre), float(te_thre), C.c_void_p(out.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out
