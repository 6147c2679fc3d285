#!/usr/bin/env python
"""Per-stage device timing of one forward (CUDA events around cumulative prefixes via stage injection is
not possible from outside, so this times the whole forward plus encoder-less variants)."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdsc_b200 import PointDSC
from pointdsc_b200.synth import make_batch

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000); ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--precision", default="fp32"); ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/snapshot_3dmatch.npz"))
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    m = PointDSC(num_layers=12, precision=a.precision); m.load_state_dict(sd, strict=False); m = m.cuda().eval()
    base = make_batch(range(16), a.n, "3dmatch", 0.3)
    rep = (a.b + 15) // 16
    cp, s, t = (base[k].repeat(rep, 1, 1)[:a.b].cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts"))
    out = m.run(cp, s, t, taps=["features", "confidence"])
    feats, conf = out["features"], out["confidence"]
    def timeit(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters
    full = timeit(lambda: m.run(cp, s, t))
    tail = timeit(lambda: m.run(cp, s, t, inject={"features": feats, "confidence": conf}))
    print(f"N={a.n} B={a.b} precision={a.precision}: forward {full:.3f} ms ({a.b / full * 1e3:.1f} sets/s); "
          f"stages iii-v only {tail:.3f} ms; stages i-ii {full - tail:.3f} ms")
    err = (out["final_trans"].cpu() - base["gt_trans"].repeat(rep, 1, 1)[:a.b]).abs().amax(dim=(1, 2))
    print("registration ok fraction:", float((err < 0.05).float().mean()))
if __name__ == "__main__":
    main()
