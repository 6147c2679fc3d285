#!/usr/bin/env python
"""Developer tool: lifetime of the persistent attention CTAs (cycles, ns -> SM clock) per encoder layer, cold (single forward)
and sustained (after 30 back-to-back forwards)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdsc_b200 import PointDSC
from pointdsc_b200.synth import make_batch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(root, "tests/golden/snapshot_3dmatch.npz")); sd = {k: torch.from_numpy(z[k]) for k in z.files}
m = PointDSC(num_layers=12, precision="fp16x3"); m.load_state_dict(sd, strict=False); m = m.cuda().eval()
base = make_batch(range(16), 1000, "3dmatch", 0.3)
cp, s, t = (base[k].repeat(16, 1, 1)[:256].cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts"))
m.run(cp, s, t)
def probe(layer):
    tl = m.run(cp, s, t, taps=["timeline"], layer_tap=layer)["timeline"].cpu().numpy()[1]
    ns, cyc = int(tl[15, 1, 0]), int(tl[15, 1, 1])
    return dict(layer=layer, cta_max=int(tl[14, 0, 0]), cta_mean=int(tl[14, 0, 1]) // 148, cta0_cycles=cyc, cta0_ns=ns, ghz=round(cyc / max(ns, 1), 3))
for layer in (0, 1, 3, 6, 11):
    print("cold     ", probe(layer))
for _ in range(30): m.run(cp, s, t)
for layer in (0, 1, 3, 6, 11):
    for _ in range(5): m.run(cp, s, t)
    print("sustained", probe(layer))
