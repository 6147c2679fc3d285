#!/usr/bin/env python
"""Fixtures for SURVEY.md §8 row f3 (per-pair evaluation statistics of evaluation/test_3DMatch.py:85-99).

Imports the reference's own `libs/loss.py` (TransformationLoss, ClassificationLoss: importable here, torch + sklearn) and
evaluates seeded cases with bs = 1 exactly as the evaluation driver does; stores inputs and the statistics row
(columns 0-8 of `stats`) as tests/golden/metrics_cases.npz.  Build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_metrics_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from libs.loss import ClassificationLoss, TransformationLoss  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rand_rot(g, max_deg):
    axis = torch.randn(3, generator=g); axis = axis / axis.norm()
    ang = torch.rand(1, generator=g).item() * max_deg * np.pi / 180
    K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return torch.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def main():
    cases = []
    tl = {"3dmatch": TransformationLoss(re_thre=15, te_thre=30), "kitti": TransformationLoss(re_thre=5, te_thre=60)}
    cl = ClassificationLoss()
    for seed in range(24):
        g = torch.Generator().manual_seed(seed)
        ds = "3dmatch" if seed % 3 else "kitti"
        n = [40, 257, 1000][seed % 3]
        gt = torch.eye(4); gt[:3, :3] = rand_rot(g, 180); gt[:3, 3] = torch.rand(3, generator=g) * 2
        # predicted transform: from exact to clearly wrong, around both thresholds
        err_deg = [0.0, 2.0, 4.9, 5.1, 14.9, 15.1, 40.0, 179.0][seed % 8]
        err_t = [0.0, 0.05, 0.29, 0.31, 0.59, 0.61, 1.5, 0.0][(seed // 3) % 8]
        pred = torch.eye(4); pred[:3, :3] = rand_rot(g, 1e-9 + err_deg) @ gt[:3, :3] if err_deg else gt[:3, :3]
        dirv = torch.randn(3, generator=g); pred[:3, 3] = gt[:3, 3] + dirv / dirv.norm() * err_t
        src = torch.rand(1, n, 3, generator=g) * 3
        ratio = [0.0, 0.05, 0.5, 1.0][seed % 4]
        gt_labels = (torch.rand(1, n, generator=g) < ratio).float()
        tgt = src @ gt[:3, :3].T + gt[:3, 3] + 0.01 * torch.randn(1, n, 3, generator=g)
        flip = torch.rand(1, n, generator=g) < [0.0, 0.1, 0.5, 1.0][(seed // 4) % 4]
        pred_labels = torch.where(flip, 1 - gt_labels, gt_labels)
        stats = cl(pred_labels, gt_labels)
        loss, recall, re, te, rmse = tl[ds](pred[None], gt[None], src, tgt, pred_labels)
        row = np.array([float(recall / 100.0), float(re), float(te), int(gt_labels.sum()), float(gt_labels.mean()),
                        int(gt_labels[pred_labels > 0].sum()), stats["precision"], stats["recall"], stats["f1"], float(rmse)], dtype=np.float64)
        cases.append(dict(dataset=ds, pred=pred.numpy(), gt=gt.numpy(), src=src[0].numpy(), tgt=tgt[0].numpy(),
                          pred_labels=pred_labels[0].numpy(), gt_labels=gt_labels[0].numpy(), row=row))
    out = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            out[f"c{i}_{k}"] = np.array(v)
    out["num_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "metrics_cases.npz"), **out)
    print("wrote", len(cases), "cases; success flags:", [int(c["row"][0]) for c in cases])


if __name__ == "__main__":
    main()
