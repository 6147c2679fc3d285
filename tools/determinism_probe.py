#!/usr/bin/env python
"""Developer tool: run the same B=256, N=1000 batch repeatedly in one process and report every set whose result differs
bit-wise from the first run (a timing race or a read of stale memory would show up here), per stage tap."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdsc_b200 import PointDSC
from pointdsc_b200.synth import make_batch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(root, "tests/golden/snapshot_3dmatch.npz")); sd = {k: torch.from_numpy(z[k]) for k in z.files}
m = PointDSC(num_layers=12, precision=os.environ.get("PDSC_PRECISION", "fp16x3")); m.load_state_dict(sd, strict=False); m = m.cuda().eval()
B = int(os.environ.get("PDSC_B", "256"))
ratios = [0.05, 0.1, 0.2, 0.4]
from pointdsc_b200.synth import make_pair
pairs = [make_pair(g, 1000, "3dmatch", ratios[g % 4]) for g in range(B)]
cp, s, t = (torch.stack([p[k] for p in pairs]).cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts"))
taps = os.environ.get("PDSC_TAPS", "features,confidence,seeds,knn_idx,seed_trans,init_trans").split(",")
layer_tap = int(os.environ.get("PDSC_LAYER_TAP", "0"))
def run():
    try:
        return m.run(cp, s, t, taps=taps, layer_tap=layer_tap)
    except Exception as e:  # tap names differ between versions: fall back to the final outputs only
        print("taps unavailable:", e); return m.run(cp, s, t)
ref = run()
torch.cuda.synchronize()
bad_total = 0
for rep in range(int(os.environ.get("PDSC_REPS", "12"))):
    # perturb timing between repetitions: a competing memory stream on a second CUDA stream
    if rep % 2:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            junk = torch.empty(1 << 28, dtype=torch.uint8, device="cuda"); junk.fill_(rep)
    out = run()
    torch.cuda.synchronize()
    for k in ref:
        if ref[k].dtype.is_floating_point or ref[k].dtype in (torch.int32, torch.int64):
            diff = (out[k] != ref[k]).reshape(out[k].shape[0], -1).any(dim=1) if out[k].shape[0] == B else (out[k] != ref[k]).reshape(1, -1).any(dim=1)
            n = int(diff.sum())
            if n:
                bad_total += n
                idx = diff.nonzero().flatten()[:6].tolist()
                md = float((out[k].float() - ref[k].float()).abs().max())
                print(f"rep {rep}: tap {k}: {n} sets differ (first {idx}), max |delta| {md:.3e}")
print("DETERMINISTIC" if bad_total == 0 else f"NON-DETERMINISTIC: {bad_total} set-level differences")
