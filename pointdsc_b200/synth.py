"""Seeded synthetic correspondence sets.

The reference ships no data in-tree; its datasets build `corr_pos` as the
concatenated (src, tgt) coordinates minus their mean over the N correspondences
(reference datasets/ThreeDMatch.py:305-308, demo_registration.py:105-108).  This
module reproduces that input contract on synthetic geometry: points uniform in a
cube, one random rigid motion per pair, Gaussian noise on the inliers and the
trailing (1-inlier_ratio)*N target points replaced by uniform outliers.
"""
from __future__ import annotations

import torch

# name -> (cube side s [m], inlier noise sigma_n [m])
PRESETS = {
    "3dmatch": (3.0, 0.01),
    "kitti": (50.0, 0.1),
}


def random_rigid(gen: torch.Generator, scale: float):
    """Random rotation (QR of a Gaussian matrix, det forced to +1) and translation in [0, scale/3)^3."""
    a = torch.randn(3, 3, generator=gen)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r))[None, :]
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    t = torch.rand(3, generator=gen) * (scale / 3.0)
    return q.contiguous(), t


def make_pair(seed: int, n: int, preset: str = "3dmatch", inlier_ratio: float = 0.3):
    """One correspondence set.  Returns dict of float32 CPU tensors:
    corr_pos [n,6], src_keypts [n,3], tgt_keypts [n,3], gt_trans [4,4], gt_labels [n]."""
    scale, sigma_n = PRESETS[preset]
    gen = torch.Generator().manual_seed(int(seed))
    src = torch.rand(n, 3, generator=gen) * scale
    rot, t = random_rigid(gen, scale)
    tgt = src @ rot.T + t + sigma_n * torch.randn(n, 3, generator=gen)
    n_in = int(round(inlier_ratio * n))
    if n_in < n:
        tgt[n_in:] = torch.rand(n - n_in, 3, generator=gen) * scale
    corr = torch.cat([src, tgt], dim=-1)
    corr = corr - corr.mean(dim=0, keepdim=True)
    gt = torch.eye(4)
    gt[:3, :3] = rot
    gt[:3, 3] = t
    labels = torch.zeros(n)
    labels[:n_in] = 1.0
    return {
        "corr_pos": corr.float().contiguous(),
        "src_keypts": src.float().contiguous(),
        "tgt_keypts": tgt.float().contiguous(),
        "gt_trans": gt,
        "gt_labels": labels,
    }


def make_batch(seeds, n: int, preset: str = "3dmatch", inlier_ratio: float = 0.3):
    """Stack `make_pair` over a list of seeds -> tensors with a leading batch axis."""
    pairs = [make_pair(s, n, preset, inlier_ratio) for s in seeds]
    return {k: torch.stack([p[k] for p in pairs], dim=0) for k in pairs[0]}
