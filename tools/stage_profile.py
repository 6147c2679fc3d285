#!/usr/bin/env python
"""Developer tool: per-stage milliseconds (engine events) of one configuration:  python tools/stage_profile.py DATASET N B [k]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pointdsc_b200 import PointDSC
ds, N, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
k = int(sys.argv[4]) if len(sys.argv) > 4 else 40
m = PointDSC(num_layers=12, k=k, **bench.CTOR[ds]); m.load_state_dict(bench.load_snapshot(ds), strict=False); m = m.cuda().eval()
h = bench.make_inputs(N, B, ds, 0)
d = [h[x].cuda() for x in ("corr_pos", "src_keypts", "tgt_keypts")]
for _ in range(2): m.run(*d)
m.profile(True)
for _ in range(3): m.run(*d)
p = m.profile_read()
print(ds, N, B, k, {x: round(v[0] / 3, 3) for x, v in p.items()})
