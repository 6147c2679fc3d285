#!/usr/bin/env python
"""On the GPU box: distribution of |T_engine - T_oracle| over many synthetic sets, per precision."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pointdsc_oracle as O
from pointdsc_b200.synth import make_batch

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000); ap.add_argument("--pairs", type=int, default=48)
    ap.add_argument("--dataset", default="3dmatch"); ap.add_argument("--precisions", default="fp32,bf16x3,bf16")
    ap.add_argument("--refs-only", action="store_true")
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, f"tests/golden/snapshot_{a.dataset}.npz"))
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    cfg = O.default_config(a.dataset)
    torch.set_num_threads(os.cpu_count())
    ratios = [0.5, 0.3, 0.2, 0.1, 0.4, 0.15]
    parts = [make_batch(range(5000 + 8 * i, 5008 + 8 * i), a.n, a.dataset, ratios[i % 6]) for i in range(a.pairs // 8)]
    batch = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
    cache = os.path.join(root, "tools", "cache", f"eval_refs_{a.dataset}_n{a.n}_p{a.pairs}.npz")
    if os.path.exists(cache):
        zc = np.load(cache)
        ref_t, ref_l = torch.from_numpy(zc["ref_t"]), torch.from_numpy(zc["ref_l"])
    else:  # oracle on the CPU: do this in the build container, not on the GPU box's clock
        ref_t, ref_l = O.forward_batch(sd, cfg, batch["corr_pos"], batch["src_keypts"], batch["tgt_keypts"])
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.savez_compressed(cache, ref_t=ref_t.numpy(), ref_l=ref_l.numpy())
    if a.refs_only:
        return
    scale = 1.0 if a.dataset == "3dmatch" else 10.0
    ok = (ref_t - batch["gt_trans"]).abs().amax(dim=(1, 2)) < 0.05 * scale
    print(f"oracle registered {int(ok.sum())}/{len(ok)} sets")
    from pointdsc_b200 import PointDSC
    for prec in a.precisions.split(","):
        m = PointDSC(num_layers=12, inlier_threshold=cfg["inlier_threshold"], sigma_d=cfg["sigma_d"], nms_radius=cfg["nms_radius"], precision=prec)
        m.load_state_dict(sd, strict=False); m = m.cuda().eval()
        out = m.run(batch["corr_pos"].cuda(), batch["src_keypts"].cuda(), batch["tgt_keypts"].cuda())
        dT = (out["final_trans"].cpu() - ref_t).abs().amax(dim=(1, 2))[ok].numpy()
        flips = (out["final_labels"].cpu() != ref_l).sum(dim=1)[ok].numpy()
        print(f"{prec:7s} dT: median {np.median(dT):.2e} p90 {np.percentile(dT, 90):.2e} max {dT.max():.2e}  >1e-4: {(dT > 1e-4).sum()}/{len(dT)}  "
              f"label flips max {flips.max()}  worst sets {np.argsort(-dT)[:4].tolist()}")
if __name__ == "__main__":
    main()
