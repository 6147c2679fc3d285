#!/usr/bin/env python
"""Reference-generated fixture on REAL data: the reference's demo pair (BASELINE.json configs[0]).

Run in the build container only (needs /root/reference):   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_demo_golden.py

demo_data/cloud_bin_{0,1}.ply -> key points + FPFH with the CPU descriptor restatement (oracle/fpfh_oracle.py; open3d, which the
reference calls here, is not installed) -> the matching lines of demo_registration.py:101-108 in numpy -> the UNMODIFIED reference
module's testing-mode forward on CPU (through tests/golden/make_golden.py's recording wrappers).  Output:
demo_pair_3dmatch.npz = the N = 5 333 real correspondences (about 20 % inliers) + the reference's stage outputs for them.  The
descriptors only choose WHICH correspondences enter the network: the parity the fixture pins is that of the hot path."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (imports the reference)
from oracle import fpfh_oracle as F  # noqa: E402


def main():
    voxel = 0.05                                    # snapshot/PointDSC_3DMatch_release/config.json: downsample
    clouds = [F.read_ply(os.path.join(G.REF, "demo_data", f"cloud_bin_{i}.ply")) for i in (0, 1)]
    (src_pts, src_feat), (tgt_pts, tgt_feat) = (F.fpfh_descriptors(c, voxel) for c in clouds)
    # demo_registration.py:101-108
    distance = np.sqrt(2 - 2 * (src_feat @ tgt_feat.T) + 1e-6)
    source_idx = np.argmin(distance, axis=1)
    src_keypts = src_pts[np.arange(len(src_pts))].astype(np.float32)
    tgt_keypts = tgt_pts[source_idx].astype(np.float32)
    corr_pos = np.concatenate([src_keypts, tgt_keypts], axis=-1)
    corr_pos = corr_pos - corr_pos.mean(0)
    n = len(src_keypts)
    pair = {"corr_pos": torch.from_numpy(corr_pos).float(), "src_keypts": torch.from_numpy(src_keypts),
            "tgt_keypts": torch.from_numpy(tgt_keypts), "gt_trans": torch.eye(4), "gt_labels": torch.zeros(n)}
    G.make_pair = lambda *a, **k: pair
    torch.manual_seed(0)
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    model, _ = G.build_model("3dmatch")
    arrays = G.run_case(model, "3dmatch", n, 0, -1.0, "io")
    arrays.pop("gt_trans"); arrays.pop("gt_labels")           # no ground truth ships with the clouds
    path = os.path.join(HERE, "demo_pair_3dmatch.npz")
    np.savez_compressed(path, **arrays)
    print(f"{os.path.basename(path)}: N={n} inliers={int(arrays['final_labels'].sum())} iters={int(arrays['power_iters'])} "
          f"solves={int(arrays['refine_solves'])} {os.path.getsize(path) / 1e3:.0f} KB\n{arrays['final_trans']}")


if __name__ == "__main__":
    main()
