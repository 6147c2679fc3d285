#!/usr/bin/env python
"""Developer tool (build container only: needs /root/reference): where does the CPU checker (oracle/pointdsc_oracle.py) leave the
UNMODIFIED reference on one synthetic set?  Runs the reference with the stage hooks of tests/golden/make_golden.py and the checker on
the same inputs and prints the deviation at every stage boundary.   python tools/oracle_vs_reference.py SEED INLIER_RATIO N"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)
import make_golden as G  # noqa: E402  (imports the reference from /root/reference)
from oracle import pointdsc_oracle as O  # noqa: E402

torch.set_num_threads(min(8, os.cpu_count() or 1))
seed, rho, n = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
model, _ = G.build_model("3dmatch")
ref = G.run_case(model, "3dmatch", n, seed, rho, "full")
z = np.load(os.path.join(ROOT, "tests", "golden", "snapshot_3dmatch.npz"))
o = O.forward_testing({k: torch.from_numpy(z[k]) for k in z.files}, O.default_config("3dmatch"), torch.from_numpy(ref["corr_pos"]),
                      torch.from_numpy(ref["src_keypts"]), torch.from_numpy(ref["tgt_keypts"]))
print(f"seed {seed}, inlier ratio {rho}, N {n}: reference vs checker")
for k in ["sc", "features", "normed", "confidence", "seeds", "knn_idx", "compat", "eig", "power_iters", "seed_weights", "seed_trans",
          "fitness", "best", "init_trans", "final_labels", "refine_solves", "final_trans"]:
    a, b = np.asarray(ref[k]), (o[k].numpy() if torch.is_tensor(o[k]) else np.asarray(o[k]))
    if a.dtype.kind in "iu" or b.dtype.kind in "iu":
        print(f"  {k:14s} {int((a.astype(np.int64) != b.astype(np.int64)).sum())} of {a.size} entries differ")
    else:
        print(f"  {k:14s} max abs diff {np.abs(a.astype(np.float64) - b.astype(np.float64)).max():.3e}")
f = np.asarray(ref["features"])
kr, ko = np.asarray(ref["knn_idx"]), o["knn_idx"].numpy()
best = int(ref["best"])
print(f"  feature scale: max |f| {np.abs(f).max():.1f}, rms {np.sqrt((f.astype(np.float64) ** 2).mean()):.1f}")
print(f"  seed rows with an identical neighbourhood SET: {sum(set(kr[i].tolist()) == set(ko[i].tolist()) for i in range(kr.shape[0]))} of {kr.shape[0]}")
print(f"  winning seed {best}: neighbourhood set identical: {set(kr[best].tolist()) == set(ko[best].tolist())}, "
      f"its transformation differs by {np.abs(np.asarray(ref['seed_trans'])[best] - o['seed_trans'].numpy()[best]).max():.2e}")
print(f"  the reference's five best inlier counts: {(np.sort(np.asarray(ref['fitness']))[::-1][:5] * n).round().astype(int).tolist()}")
