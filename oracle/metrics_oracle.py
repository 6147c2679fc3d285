"""CPU restatement of the per-pair evaluation statistics.

TEST INFRASTRUCTURE ONLY.  Reference: libs/loss.py:34-63 (`TransformationLoss.forward`: rotation error in degrees,
translation error in cm, success flag, RMSE under the predicted transform), libs/loss.py:96-102 (`ClassificationLoss`:
precision / recall / F1 of the predicted inlier labels, computed there with scikit-learn on the host) and the columns the
evaluation driver records per pair, evaluation/test_3DMatch.py:85-96.  Pinned by tests/golden/metrics_cases.npz, produced by
tests/golden/make_metrics_golden.py from the reference's own module.
"""
import math

import torch


def registration_errors(pred_trans: torch.Tensor, gt_trans: torch.Tensor):
    """(RE in degrees, TE in cm) of one 4x4 estimate against the ground truth (libs/loss.py:45-49): the angle of the relative
    rotation Rᵀ·R_gt through its trace (clamped into acos' domain) and the distance of the translations, in metres x 100."""
    R, t = pred_trans[:3, :3], pred_trans[:3, 3]
    Rg, tg = gt_trans[:3, :3], gt_trans[:3, 3]
    cos = torch.clamp((torch.trace(R.T @ Rg) - 1) / 2.0, min=-1, max=1)
    re = torch.acos(cos) * 180 / math.pi
    te = torch.sqrt(((t - tg) ** 2).sum()) * 100
    return re, te


def label_scores(pred_labels: torch.Tensor, gt_labels: torch.Tensor):
    """(precision, recall, f1) of `pred_labels > 0` against binary `gt_labels` with scikit-learn's binary conventions
    (a zero denominator scores 0), libs/loss.py:96-102."""
    p = pred_labels > 0
    g = gt_labels > 0
    tp = float((p & g).sum())
    fp = float((p & ~g).sum())
    fn = float((~p & g).sum())
    precision = tp / (tp + fp) if tp + fp > 0 else 0.0
    recall = tp / (tp + fn) if tp + fn > 0 else 0.0
    f1 = 2 * tp / (2 * tp + fp + fn) if 2 * tp + fp + fn > 0 else 0.0
    return precision, recall, f1


def stats_row(pred_trans, gt_trans, src_keypts, tgt_keypts, pred_labels, gt_labels, re_thre=15.0, te_thre=30.0):
    """Columns 0-8 of the evaluation driver's per-pair statistics (test_3DMatch.py:85-96) followed by the RMSE of the
    correspondences under the predicted transform (libs/loss.py:47-48):
    [success, RE deg, TE cm, #gt inliers, gt inlier ratio, #gt inliers among the kept, precision, recall, f1, rmse]."""
    re, te = registration_errors(pred_trans, gt_trans)
    success = 1.0 if (float(te) < te_thre and float(re) < re_thre) else 0.0
    warped = src_keypts @ pred_trans[:3, :3].T + pred_trans[:3, 3]
    rmse = (warped - tgt_keypts).norm(dim=-1).mean()
    precision, recall, f1 = label_scores(pred_labels, gt_labels)
    kept_inliers = float(gt_labels[pred_labels > 0].sum())
    return [success, float(re), float(te), float(gt_labels.sum()), float(gt_labels.float().mean()), kept_inliers, precision,
            recall, f1, float(rmse)]
