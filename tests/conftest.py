import glob
import json
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_snapshot(dataset):
    """Released state dict (key for key) from tests/golden/snapshot_<dataset>.npz."""
    z = np.load(os.path.join(GOLDEN, f"snapshot_{dataset}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def golden_cases(dataset=None, detail=None, n_max=None):
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz"))):
        meta = json.loads(str(np.load(path)["meta"]))
        if dataset and meta["dataset"] != dataset:
            continue
        if detail and meta["detail"] not in detail:
            continue
        if n_max and meta["n"] > n_max:
            continue
        out.append(path)
    return out


def load_case(path):
    z = np.load(path)
    case = {k: z[k] for k in z.files if k != "meta"}
    case["meta"] = json.loads(str(z["meta"]))
    return case


def registration_ok(case):
    """Oracle/reference registration succeeded (SURVEY.md §7 trap 8: failures are chaotic)."""
    scale = 1.0 if case["meta"]["dataset"] == "3dmatch" else 10.0
    return float(np.abs(case["final_trans"] - case["gt_trans"]).max()) < 0.05 * scale


@pytest.fixture(scope="session")
def snapshots():
    return {d: load_snapshot(d) for d in ("3dmatch", "kitti")}
