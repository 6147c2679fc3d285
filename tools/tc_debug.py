#!/usr/bin/env python
"""Compare the tensor-core encoder path against the fp32 SIMT path, stage by stage inside a layer."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdsc_b200 import PointDSC
from pointdsc_b200.synth import make_batch

QS = 1.4426950408889634 / 11.313708498984761

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256); ap.add_argument("--b", type=int, default=2)
    ap.add_argument("--precision", default="bf16x3"); ap.add_argument("--layers", default="0,1,11")
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests/golden/snapshot_3dmatch.npz"))
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    ms = {}
    for prec in ("fp32", a.precision):
        m = PointDSC(num_layers=12, precision=prec); m.load_state_dict(sd, strict=False); ms[prec] = m.cuda().eval()
    batch = make_batch(range(50, 50 + a.b), a.n, "3dmatch", 0.4)
    cp, s, t = (batch[k].cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts"))
    for layer in [int(x) for x in a.layers.split(",")]:
        outs = {p: m.run(cp, s, t, taps=["layer_debug", "layer_features"], layer_tap=layer) for p, m in ms.items()}
        torch.cuda.synchronize()
        ref, got = outs["fp32"], outs[a.precision]
        names = ["feat1", "q", "k", "v", "msg"]
        line = f"layer {layer}:"
        for i, nm in enumerate(names):
            r, g = ref["layer_debug"][i], got["layer_debug"][i]
            if nm == "q": g = g / QS
            d = (r - g).abs().max().item(); sc = r.abs().max().item()
            line += f" {nm} d={d:.2e}/{sc:.1e}"
            if not torch.isfinite(g).all(): line += "(NONFINITE)"
        r, g = ref["layer_features"], got["layer_features"]
        line += f" | feat d={(r - g).abs().max().item():.2e}/{r.abs().max().item():.1e}"
        print(line, flush=True)
    o1 = ms["fp32"].run(cp, s, t); o2 = ms[a.precision].run(cp, s, t)
    print("final dT", (o1["final_trans"] - o2["final_trans"]).abs().amax(dim=(1, 2)).cpu().numpy())
if __name__ == "__main__":
    main()
