"""CPU restatement of the correspondence FRONT END that feeds `PointDSC.forward`.

TEST INFRASTRUCTURE ONLY, like oracle/pointdsc_oracle.py: imported by tests/ (and, once the device kernel exists, by its
parity tests); the product never calls it.  Pinned against the reference's own source lines, executed by
tests/golden/make_frontend_golden.py (fixtures tests/golden/frontend_*.npz).

Reference: datasets/ThreeDMatch.py:282-291 (matching, optional mutual check) and :299-308 (network input), the same
lines in datasets/KITTI.py:80-114 and demo_registration.py:101-108 (no mutual check there).  numpy semantics are part of
the contract: the arithmetic runs in the descriptors' dtype (float32 for FCGF, float64 for FPFH), `argmin` returns the
FIRST minimum, and the square root is taken BEFORE the argmin (it merges distances that differ by less than an ulp of
sqrt, so argmax of the dot product is not a substitute).
"""
import numpy as np


def feature_distance(src_desc: np.ndarray, tgt_desc: np.ndarray) -> np.ndarray:
    """[Ns, Nt] descriptor distance of L2-normalised rows: sqrt(2 - 2 <a, b> + 1e-6)  (ThreeDMatch.py:283)."""
    gram = src_desc @ tgt_desc.T
    return np.sqrt(2 - 2 * gram + 1e-6)


def match(src_desc: np.ndarray, tgt_desc: np.ndarray, use_mutual: bool = False) -> np.ndarray:
    """Putative correspondences [M, 2] (row = (source index, target index)), ThreeDMatch.py:283-291.

    Every source point takes its nearest target in feature space; with `use_mutual` only the pairs whose target also
    takes that source as ITS nearest neighbour survive (their order stays ascending in the source index)."""
    dist = feature_distance(src_desc, tgt_desc)
    nearest_tgt = dist.argmin(axis=1)
    src_ids = np.arange(nearest_tgt.shape[0])
    if use_mutual:
        nearest_src = dist.argmin(axis=0)
        keep = nearest_src[nearest_tgt] == src_ids
        return np.stack([src_ids[keep], nearest_tgt[keep]], axis=1)
    return np.stack([src_ids, nearest_tgt], axis=1)


def network_input(src_keypts: np.ndarray, tgt_keypts: np.ndarray, corr: np.ndarray):
    """(corr_pos [M,6], src_keypts [M,3], tgt_keypts [M,3]) for in_dim = 6: the matched key points side by side, centred
    by their mean over the M correspondences (ThreeDMatch.py:300-308, demo_registration.py:104-108)."""
    a = src_keypts[corr[:, 0]]
    b = tgt_keypts[corr[:, 1]]
    both = np.concatenate([a, b], axis=-1)
    return both - both.mean(0), a, b
