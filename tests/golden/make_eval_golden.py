#!/usr/bin/env python
"""Fixture for the forward WITHOUT the 'testing' key (SURVEY.md section 8 rows a6' / f4), FROM THE REFERENCE ITSELF.

Imports the unmodified reference, loads the released 3DMatch snapshot, puts the module in eval mode and calls it on a batch
of three seeded synthetic sets exactly as libs/trainer.py:186 does during validation (no 'testing' key, bs > 1): stores
the inputs and the returned final_trans / final_labels (= confidence logits) / M.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_eval_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from make_golden import build_model  # noqa: E402
from pointdsc_b200.synth import make_pair  # noqa: E402


def main():
    torch.manual_seed(0)
    model, _ = build_model("3dmatch")
    model.eval()
    pairs = [make_pair(40 + i, 256, "3dmatch", r) for i, r in enumerate((0.5, 0.3, 0.6))]
    data = {k: torch.stack([p[k] for p in pairs], 0) for k in ("corr_pos", "src_keypts", "tgt_keypts")}
    seen = {}
    orig = model.cal_seed_trans

    def spy(seeds, feats, s, t):
        seen["seeds"] = seeds.clone()
        return orig(seeds, feats, s, t)
    model.cal_seed_trans = spy
    with torch.no_grad():
        out = model(data)             # no 'testing' key
    model.cal_seed_trans = orig
    with torch.no_grad():      # the N x N use of the power iteration (PointDSC.py:170, disabled in the released forward)
        m_eig = model.cal_leading_eigenvector(out["M"], method="power")
    np.savez_compressed(os.path.join(HERE, "eval_3dmatch_n256_b3.npz"), corr_pos=data["corr_pos"].numpy(),
                        src_keypts=data["src_keypts"].numpy(), tgt_keypts=data["tgt_keypts"].numpy(),
                        gt_trans=torch.stack([p["gt_trans"] for p in pairs], 0).numpy(),
                        final_trans=out["final_trans"].numpy(), final_labels=out["final_labels"].numpy(), M=out["M"].numpy(),
                        seeds=seen["seeds"].numpy().astype(np.int32), M_eig=m_eig.numpy())
    err = (out["final_trans"] - torch.stack([p["gt_trans"] for p in pairs], 0)).abs().amax(dim=(1, 2))
    print("eval fixture: |T - gt| per set", err.tolist(), "M nonzero fraction", float((out["M"] > 0).float().mean()))


if __name__ == "__main__":
    main()
