#!/usr/bin/env python
"""Developer tool (GPU box): where does the engine leave the reference on the real-data fixture tests/golden/demo_pair_3dmatch.npz?
Prints, per precision and batch regime, the deviation at every stage boundary the fixture holds."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointdsc_b200 import PointDSC
z = np.load(os.path.join(ROOT, "tests/golden/demo_pair_3dmatch.npz"))
s = np.load(os.path.join(ROOT, "tests/golden/snapshot_3dmatch.npz")); sd = {k: torch.from_numpy(s[k]) for k in s.files}
taps = ["confidence", "seeds", "knn_idx", "eig", "power_iters", "seed_trans", "inlier_counts", "best", "init_trans", "refine_solves"]
for prec in ("fp32", "fp16x3", "bf16x3"):
    for B in (1, 3):
        m = PointDSC(num_layers=12, precision=prec, inlier_threshold=0.10, sigma_d=0.10, nms_radius=0.10).cuda().eval()
        m.load_state_dict(sd, strict=False)
        d = [torch.from_numpy(z[k])[None].repeat(B, 1, 1).cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts")]
        o = m.run(*d, taps=taps)
        g = {k: v[0].cpu().numpy() for k, v in o.items()}
        seeds_eq = g["seeds"] == z["seeds"]
        lead = int(np.argmax(~seeds_eq)) if (~seeds_eq).any() else len(seeds_eq)
        conf = np.abs(g["confidence"] - z["confidence"])
        print(f"{prec:7s} B={B}: dT={np.abs(g['final_trans'] - z['final_trans']).max():.2e} dInit={np.abs(g['init_trans'] - z['init_trans']).max():.2e} "
              f"flips={int((g['final_labels'] != z['final_labels']).sum())} best={int(g['best'])}/{int(z['best'])} "
              f"count[best]={int(g['inlier_counts'][int(g['best'])])} ref fitness max={float(z['fitness'].max()) * len(z['final_labels']):.0f} "
              f"iters={int(g['power_iters'])}/{int(z['power_iters'])} solves={int(g['refine_solves'])}/{int(z['refine_solves'])} "
              f"conf max/mean diff={conf.max():.2e}/{conf.mean():.2e} seeds common prefix={lead} "
              f"dSeedTrans[ref best]={np.abs(g['seed_trans'][int(z['best'])] - z['seed_trans'][int(z['best'])]).max():.2e}", flush=True)
        # with the reference's seeds and neighbourhoods injected: is the rest of the path exact?
        o2 = m.run(*d, taps=["best", "init_trans"], inject={"seeds": torch.from_numpy(z["seeds"])[None].repeat(B, 1).cuda(),
                                                              "confidence": torch.from_numpy(z["confidence"])[None].repeat(B, 1).cuda(),
                                                              "features": o["features"] if "features" in o else m.run(*d, taps=["features"])["features"]})
        print(f"        reference confidence + seeds injected: dT={np.abs(o2['final_trans'][0].cpu().numpy() - z['final_trans']).max():.2e} best={int(o2['best'][0])}", flush=True)
