#!/usr/bin/env python
"""Developer tool (GPU box): how often does the arithmetic mode change the selected hypothesis on FPFH-like correspondences?
For a few synthetic scene pairs (pointdsc_b200.synth_scene; descriptors + matching on the device) the engine in fp16x3 and fp32 is
compared with the CPU checker (which reproduces the unmodified reference to 2.7e-7 on the real demo pair), next to the checker's own
margin between its best and second-best hypothesis.   python tools/neartie_survey.py [pairs]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pointdsc_oracle as O
from pointdsc_b200 import PointDSC
from pointdsc_b200.descriptors import fpfh_descriptors
from pointdsc_b200.frontend import match
from pointdsc_b200.synth_scene import rigid, scene

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
s = np.load(os.path.join(ROOT, "tests/golden/snapshot_3dmatch.npz")); sd = {k: torch.from_numpy(s[k]) for k in s.files}
models = {}
for prec in ("fp16x3", "fp32"):
    m = PointDSC(num_layers=12, precision=prec, inlier_threshold=0.10, sigma_d=0.10, nms_radius=0.10).cuda().eval()
    m.load_state_dict(sd, strict=False)
    models[prec] = m
cfg = O.default_config("3dmatch")
print("pair | N | checker: inliers, best count, runner-up count | fp16x3: dT, flips, same best | fp32: dT, flips, same best")
for p in range(pairs):
    R, t = rigid(100 + p)
    src = torch.from_numpy(scene(60000, seed=2 * p, layout_seed=20 + p)).cuda()
    tgt = torch.from_numpy((scene(60000, seed=2 * p + 1, layout_seed=20 + p).astype(np.float64) @ R.T + t).astype(np.float32)).cuda()
    (skp, sf), (tkp, tf) = fpfh_descriptors(src, 0.08), fpfh_descriptors(tgt, 0.08)
    data = match(sf, tf, skp, tkp, use_mutual=False)
    d = [data[k] for k in ("corr_pos", "src_keypts", "tgt_keypts")]
    want = O.forward_testing(sd, cfg, *[x[0].float().cpu() for x in d])
    n = d[0].shape[1]
    counts = np.sort(np.round(want["fitness"].numpy() * n))[::-1]
    row = f"{p} | {n} | {int(want['final_labels'].sum())}, {int(counts[0])}, {int(counts[1])}"
    for prec, m in models.items():
        o = m.run(*d, taps=["best"])
        dT = np.abs(o["final_trans"][0].cpu().numpy() - want["final_trans"].numpy()).max()
        flips = int((o["final_labels"][0].cpu() != want["final_labels"]).sum())
        row += f" | {dT:.1e}, {flips}, {int(o['best'][0]) == int(want['best'])}"
    print(row, flush=True)
