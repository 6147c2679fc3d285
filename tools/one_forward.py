#!/usr/bin/env python
"""Developer tool: a few testing-mode forwards of one shape (for ncu launch lists):  python tools/one_forward.py B N [reps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdsc_b200 import PointDSC
from pointdsc_b200.synth import make_batch
B, N = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(root, "tests/golden/snapshot_3dmatch.npz")); sd = {k: torch.from_numpy(z[k]) for k in z.files}
m = PointDSC(num_layers=12); m.load_state_dict(sd, strict=False); m = m.cuda().eval()
m.graph_rows = 0          # eager launches: every kernel is visible to the profiler
b = make_batch(range(B), N, "3dmatch", 0.3)
cp, s, t = (b[k].cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts"))
for _ in range(reps):
    out = m.run(cp, s, t)
torch.cuda.synchronize()
print("ok", float(out["final_trans"].abs().sum()))
