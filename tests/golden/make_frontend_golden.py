#!/usr/bin/env python
"""Fixtures for SURVEY.md §8 row f1 (the correspondence front end that feeds `PointDSC.forward`).

The reference has no function for it: the matching lives inline in `datasets/ThreeDMatch.py:282-291` (+ `:299-308` for the
network input) and `demo_registration.py:101-108`, and those modules do not import here (open3d).  This script therefore
EXECUTES THE REFERENCE'S OWN SOURCE LINES: it reads the two blocks out of /root/reference/datasets/ThreeDMatch.py at run
time (nothing is copied into this repository), runs them on seeded descriptors / key points with a stub `self`, and stores
inputs and outputs as tests/golden/frontend_*.npz.  Run in the build container only (needs /root/reference):

    python tests/golden/make_frontend_golden.py
"""
import os
import textwrap
import types

import numpy as np

REF = "/root/reference/datasets/ThreeDMatch.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_blocks():
    lines = open(REF).read().split("\n")
    # second occurrence = the test-split dataset class (ThreeDMatchTest uses the same lines; any occurrence is identical)
    start = [i for i, ln in enumerate(lines) if "construct the correspondence set by mutual nn in feature space" in ln][0]
    stop = [i for i, ln in enumerate(lines) if "build the ground truth label" in ln and i > start][0]
    match_src = textwrap.dedent("\n".join(lines[start:stop]))
    p0 = [i for i, ln in enumerate(lines) if "prepare input to the network" in ln and i > stop][0]
    p1 = [i for i, ln in enumerate(lines) if "elif self.in_dim == 9" in ln and i > p0][0]
    input_src = textwrap.dedent("\n".join(lines[p0:p1]))
    return match_src, input_src, (start + 1, stop, p0 + 1, p1)


def run_reference(src_desc, tgt_desc, src_keypts, tgt_keypts, use_mutual, in_dim):
    match_src, input_src, span = reference_blocks()
    ns = {"np": np, "self": types.SimpleNamespace(use_mutual=use_mutual, in_dim=in_dim), "src_desc": src_desc, "tgt_desc": tgt_desc,
          "src_keypts": src_keypts, "tgt_keypts": tgt_keypts}
    exec(match_src, ns)          # -> distance, source_idx, corr
    exec(input_src, ns)          # -> input_src_keypts, input_tgt_keypts, corr_pos
    return ns["corr"], ns["input_src_keypts"], ns["input_tgt_keypts"], ns["corr_pos"], span


def unit_rows(rng, n, d, dtype):
    x = rng.standard_normal((n, d)).astype(dtype)
    return (x / (np.linalg.norm(x, axis=1, keepdims=True) + 1e-6)).astype(dtype)


def main():
    cases = [  # name, n_src, n_tgt, descriptor dim, dtype, mutual, seed
        ("fcgf32_n300_m0", 300, 280, 32, np.float32, False, 0),
        ("fcgf32_n300_m1", 300, 280, 32, np.float32, True, 0),
        ("fpfh33_n500_m0", 500, 520, 33, np.float64, False, 1),
        ("fpfh33_n500_m1", 500, 520, 33, np.float64, True, 1),
        ("ties_n64_m0", 64, 64, 8, np.float32, False, 2),
    ]
    for name, ns_, nt, d, dtype, mutual, seed in cases:
        rng = np.random.default_rng(seed)
        src_desc, tgt_desc = unit_rows(rng, ns_, d, dtype), unit_rows(rng, nt, d, dtype)
        if name.startswith("ties"):          # duplicated target descriptors: argmin must take the first minimum
            tgt_desc[32:] = tgt_desc[:32]
        src_keypts = rng.uniform(0, 3, (ns_, 3)).astype(np.float32)
        tgt_keypts = rng.uniform(0, 3, (nt, 3)).astype(np.float32)
        corr, in_s, in_t, corr_pos, span = run_reference(src_desc, tgt_desc, src_keypts, tgt_keypts, mutual, 6)
        np.savez_compressed(os.path.join(HERE, f"frontend_{name}.npz"), src_desc=src_desc, tgt_desc=tgt_desc, src_keypts=src_keypts,
                            tgt_keypts=tgt_keypts, use_mutual=np.array(mutual), corr=corr, input_src_keypts=in_s,
                            input_tgt_keypts=in_t, corr_pos=corr_pos, reference_lines=np.array(span))
        print(name, corr.shape, corr_pos.dtype, "reference lines", span)


if __name__ == "__main__":
    main()
