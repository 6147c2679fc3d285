"""SURVEY.md §8 row f1 (next row): the front-end oracle against fixtures produced by executing the reference's own source
lines (tests/golden/make_frontend_golden.py).  CPU only; the device kernel of this row does not exist yet."""
import glob
import os

import numpy as np
import pytest

from oracle import frontend_oracle as F

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "frontend_*.npz")))


def test_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[9:-4])
def test_front_end_oracle_matches_reference_lines(path):
    z = np.load(path)
    corr = F.match(z["src_desc"], z["tgt_desc"], bool(z["use_mutual"]))
    assert corr.dtype == z["corr"].dtype and np.array_equal(corr, z["corr"])            # indices: bit-exact
    corr_pos, a, b = F.network_input(z["src_keypts"], z["tgt_keypts"], corr)
    assert np.array_equal(a, z["input_src_keypts"]) and np.array_equal(b, z["input_tgt_keypts"])
    assert corr_pos.dtype == z["corr_pos"].dtype and np.array_equal(corr_pos, z["corr_pos"])   # same numpy ops: exact


def test_first_minimum_wins_on_duplicated_targets():
    z = np.load([p for p in CASES if "ties" in p][0])
    corr = F.match(z["src_desc"], z["tgt_desc"], False)
    assert int(corr[:, 1].max()) < 32          # targets 32.. duplicate 0..31: the first copy must be chosen


def test_mutual_check_is_a_subset_in_source_order():
    z = np.load([p for p in CASES if p.endswith("fcgf32_n300_m1.npz")][0])
    full = F.match(z["src_desc"], z["tgt_desc"], False)
    mutual = F.match(z["src_desc"], z["tgt_desc"], True)
    assert np.all(np.diff(mutual[:, 0]) > 0)
    assert all(tuple(r) in set(map(tuple, full)) for r in mutual)


def test_corr_pos_is_centred():
    z = np.load(CASES[0])
    corr_pos, _, _ = F.network_input(z["src_keypts"], z["tgt_keypts"], z["corr"])
    assert float(np.abs(corr_pos.mean(0)).max()) < 1e-5
