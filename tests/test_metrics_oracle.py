"""SURVEY.md §8 row f3 (next row): the evaluation-statistics oracle against fixtures computed by the reference's own
libs/loss.py (tests/golden/make_metrics_golden.py).  CPU only; the device kernel of this row does not exist yet."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_cases.npz"))
N = int(Z["num_cases"])


@pytest.mark.parametrize("i", range(N))
def test_stats_row_matches_reference(i):
    g = lambda k: torch.from_numpy(Z[f"c{i}_{k}"])  # noqa: E731
    thr = dict(re_thre=15.0, te_thre=30.0) if str(Z[f"c{i}_dataset"]) == "3dmatch" else dict(re_thre=5.0, te_thre=60.0)
    row = M.stats_row(g("pred"), g("gt"), g("src"), g("tgt"), g("pred_labels"), g("gt_labels"), **thr)
    ref = Z[f"c{i}_row"]
    assert row[0] == ref[0]                                   # success flag: exact
    assert row[3] == ref[3] and row[5] == ref[5]              # counts: exact
    # RE deg / TE cm / ratios / RMSE: fp32 arithmetic on both sides, 1e-4 relative + 1e-4 absolute (acos is ill-conditioned
    # near 0 deg: the reference's own value there is round-off of the trace)
    np.testing.assert_allclose(row[1:3] + row[4:5] + row[6:], np.concatenate([ref[1:3], ref[4:5], ref[6:]]), rtol=1e-4, atol=2e-2 if ref[1] < 0.1 else 1e-4)


def test_both_outcomes_and_degenerate_labels_are_covered():
    flags = [float(Z[f"c{i}_row"][0]) for i in range(N)]
    assert 0.0 in flags and 1.0 in flags
    assert any(float(Z[f"c{i}_gt_labels"].sum()) == 0 for i in range(N))      # no inlier at all: precision = recall = f1 = 0
    assert any(float(Z[f"c{i}_gt_labels"].mean()) == 1 for i in range(N))
