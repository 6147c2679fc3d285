#!/usr/bin/env python
"""Print the in-kernel clock64 timeline of CTA 0 (chain<PCQ> and attention) — developer tool."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdsc_b200 import PointDSC
from pointdsc_b200.synth import make_batch
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=1000); ap.add_argument("--b", type=int, default=256)
ap.add_argument("--precision", default="fp16x3"); a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(root, "tests/golden/snapshot_3dmatch.npz")); sd = {k: torch.from_numpy(z[k]) for k in z.files}
m = PointDSC(num_layers=12, precision=a.precision); m.load_state_dict(sd, strict=False); m = m.cuda().eval()
base = make_batch(range(16), a.n, "3dmatch", 0.3); rep = (a.b + 15) // 16
cp, s, t = (base[k].repeat(rep, 1, 1)[:a.b].cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts"))
m.run(cp, s, t)
out = m.run(cp, s, t, taps=["timeline"], layer_tap=3)
tl = out["timeline"].cpu().numpy()
for k, name, roles in ((0, "chain<PCQ>", ["mma", "loader", "epilogue", "epi-detail"]), (1, "attention", ["mma", "softmax-g0", "softmax-g1", "item2-tiles(ev0 = MMA saw P_j, ev5 = QK_{j+3} issued, ev7 = PV_j issued; the softmax loop carries no stamps)"])):
    d = tl[k]; t0 = d[:14][d[:14] > 0].min()
    print(f"== {name}: cycles since first stamp; rows = tile/iteration, per role events")
    for it in range(14 if k == 1 else 16):
        line = f"it{it:2d}"
        for ri, rn in enumerate(roles):
            ev = d[it, ri]; line += f" | {rn}:" + " ".join(f"{int(x - t0):7d}" if x > 0 else "      -" for x in ev)
        print(line)

d = tl[1]
ns, cyc = int(d[15, 1, 0]), int(d[15, 1, 1])
print("attention CTA lifetimes (cycles): max", int(d[14, 0, 0]), "mean", int(d[14, 0, 1]) // 148, "first 8 CTAs", [int(x) for x in d[15, 0]],
      "| CTA 0:", cyc, "cycles in", ns, "ns ->", round(cyc / max(ns, 1), 3), "GHz during the launch")
