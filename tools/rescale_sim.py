#!/usr/bin/env python
"""Developer tool (CPU): how often does the attention kernel's lazy-rescale path fire, per 32-row warp and 64-key tile, for
different key-tile visiting orders?  Uses the oracle's encoder arithmetic (test infrastructure) on a few synthetic sets.
    python tools/rescale_sim.py [N] [sets]"""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pointdsc_oracle as O
from pointdsc_b200.synth import make_batch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
SETS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
z = np.load("tests/golden/snapshot_3dmatch.npz"); sd = {k: torch.from_numpy(z[k]) for k in z.files}
cfg = O.default_config("3dmatch")
b = make_batch(range(SETS), N, "3dmatch", 0.3)
KT, QT = (N + 63) // 64, (N + 127) // 128
orders = {"natural": lambda qt: list(range(KT)),
          "diagonal first": lambda qt: [(2 * qt + j) % KT for j in range(KT)],
          "reverse": lambda qt: list(range(KT - 1, -1, -1))}
stats = {k: [0, 0, 0] for k in orders}          # warp-tiles with an advance, all warp-tiles (j > 0), group-tiles (any of 4 warps)
for s in range(SETS):
    sc = O.sc_matrix(b["src_keypts"][s], b["tgt_keypts"][s], cfg["sigma_d"])
    sc = sc[0] if isinstance(sc, tuple) else sc
    feat = O._lin(b["corr_pos"][s], sd["encoder.layer0.weight"], sd["encoder.layer0.bias"])
    for i in range(12):
        pre = f"encoder.blocks.PointCN_layer_{i}"
        feat = torch.relu(O._bn(O._lin(feat, sd[pre + ".0.weight"], sd[pre + ".0.bias"]), sd, pre + ".1"))
        pn = f"encoder.blocks.NonLocal_layer_{i}"
        q = O._lin(feat, sd[pn + ".projection_q.weight"], sd[pn + ".projection_q.bias"])
        k = O._lin(feat, sd[pn + ".projection_k.weight"], sd[pn + ".projection_k.bias"])
        logit2 = (sc * ((q @ k.t()) / math.sqrt(128.0)) * math.log2(math.e)).numpy()
        pad = np.full((QT * 128, KT * 64), -np.inf, np.float32); pad[:N, :N] = logit2
        tmax = pad.reshape(QT * 128, KT, 64).max(2)                      # [rows, KT] per-tile row maxima
        for name, fn in orders.items():
            for qt in range(QT):
                rows = tmax[qt * 128:(qt + 1) * 128]
                ref = np.full(128, -np.inf)
                for jj, t in enumerate(fn(qt)):
                    adv = (rows[:, t] > ref + 8.0) if jj else np.ones(128, bool)
                    ref = np.where(adv, rows[:, t], ref)
                    if jj:
                        w = adv.reshape(4, 32).any(1)
                        stats[name][0] += int(w.sum()); stats[name][1] += 4; stats[name][2] += int(w.any())
        feat = O.nonlocal_block(feat, sc, sd, pn)
for name, (a, t, g) in stats.items():
    print(f"{name:16s} warp-tiles with a rescale: {a / t:6.1%}   tiles where the group (any of 4 warps) rescales: {g / (t / 4):6.1%}")
