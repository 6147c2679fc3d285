#!/usr/bin/env python
"""Developer tool: a small pass through every C-ABI entry for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pointdsc_b200 import PointDSC
from pointdsc_b200.frontend import match
from pointdsc_b200.metrics import eval_stats
from pointdsc_b200.spectral import leading_eigenvector
for k, n, b in ((40, 300, 3), (80, 257, 2)):
    m = PointDSC(num_layers=12, k=k, **bench.CTOR["3dmatch"]); m.load_state_dict(bench.load_snapshot("3dmatch"), strict=False); m = m.cuda().eval()
    h = bench.make_inputs(n, b, "3dmatch", 0)
    d = [h[x].cuda() for x in ("corr_pos", "src_keypts", "tgt_keypts")]
    out = m.run(*d, taps=["best"])                      # eager
    dbg = m.run(*d, taps=["layer_debug"], layer_tap=3)   # the debug tap un-blocks feat1
    out2 = m.run(*d)                                    # graph path (capture + replay)
    out2 = m.run(*d)
    ev = m({"corr_pos": d[0], "src_keypts": d[1], "tgt_keypts": d[2]})
    st = eval_stats(out["final_trans"], h["gt_trans"].cuda(), d[1], d[2], out["final_labels"], h["gt_labels"].cuda())
    v, it = leading_eigenvector(ev["M"], 10, True)
    host = m.run(h["corr_pos"], h["src_keypts"], h["tgt_keypts"])
    hd = {"corr_pos": h["corr_pos"].pin_memory(), "src_keypts": h["src_keypts"].pin_memory(), "tgt_keypts": h["tgt_keypts"].pin_memory(),
          "testing": True}
    streamed = list(m.forward_stream(hd for _ in range(3)))      # pdsc_forward_host_submit / _wait, two calls in flight
    assert all(torch.equal(o["final_trans"], host["final_trans"]) for o in streamed)
g = torch.Generator().manual_seed(0)
for dt in (torch.float32, torch.float64):
    sd = torch.nn.functional.normalize(torch.randn(301, 33, generator=g, dtype=dt), dim=1).cuda()
    td = torch.nn.functional.normalize(torch.randn(277, 33, generator=g, dtype=dt), dim=1).cuda()
    for mutual in (False, True):
        r = match(sd, td, torch.rand(301, 3).cuda(), torch.rand(277, 3).cuda(), use_mutual=mutual)
from pointdsc_b200 import descriptors as D
from pointdsc_b200.synth_scene import scene
kp, feat = D.fpfh_descriptors(torch.from_numpy(scene(4000, seed=0)).cuda(), 0.15)
torch.cuda.synchronize()
print("sanitize_smoke ok", float(st.sum()), int(it.sum()), r["corr"].shape, tuple(feat.shape))
