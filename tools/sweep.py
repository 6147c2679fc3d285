#!/usr/bin/env python
"""Throughput + parity sweep over the BASELINE.json configurations (run on the GPU box).

For every (dataset, N, k, B): device-resident sets/s (CUDA events, median of `--iters` forwards after warm-up) and
the engine-vs-oracle deviation on the first `--check` sets (oracle on the host cores; skipped with --check 0).
Prints one JSON line per configuration; `profiles/` keeps the output of the round."""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pointdsc_oracle as O
from pointdsc_b200 import PointDSC
from pointdsc_b200.synth import make_pair

CONFIGS = [  # name, dataset, N, k, B
    ("A 3DMatch", "3dmatch", 1000, 40, 64),
    ("B KITTI", "kitti", 5000, 40, 32),
    ("C k80", "3dmatch", 2000, 80, 256),
    ("D sweep", "3dmatch", 500, 40, 128),
    ("D sweep", "3dmatch", 1000, 40, 128),
    ("D sweep", "3dmatch", 2000, 40, 128),
    ("D sweep", "3dmatch", 5000, 40, 128),
    ("headline", "3dmatch", 1000, 40, 256),
]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16x3"); ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--check", type=int, default=2); ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    for name, ds, n, k, B in CONFIGS:
        if a.only and a.only not in name:
            continue
        z = np.load(os.path.join(ROOT, "tests", "golden", f"snapshot_{ds}.npz"))
        sd = {kk: torch.from_numpy(z[kk]) for kk in z.files}
        cfg = dict(O.default_config(ds)); cfg["k"] = k
        m = PointDSC(num_layers=12, inlier_threshold=cfg["inlier_threshold"], sigma_d=cfg["sigma_d"], k=k,
                     nms_radius=cfg["nms_radius"], precision=a.precision)
        m.load_state_dict(sd, strict=False); m = m.cuda().eval()
        ratios = [0.5, 0.3, 0.2, 0.4]
        base = [make_pair(9000 + g, n, ds, ratios[g % 4]) for g in range(min(B, 16))]
        stack = {kk: torch.stack([p[kk] for p in base], 0) for kk in base[0]}
        rep = (B + len(base) - 1) // len(base)
        cp, s, t = (stack[kk].repeat(rep, 1, 1)[:B].cuda() for kk in ("corr_pos", "src_keypts", "tgt_keypts"))
        out = m.run(cp, s, t); torch.cuda.synchronize()
        times = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); out = m.run(cp, s, t); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        gt = stack["gt_trans"].repeat(rep, 1, 1)[:B]
        scale = 0.05 if ds == "3dmatch" else 0.5
        reg = float(((out["final_trans"].cpu() - gt).abs().amax(dim=(1, 2)) < scale).float().mean())
        rec = {"config": name, "dataset": ds, "N": n, "k": k, "B": B, "precision": a.precision, "ms_per_forward": round(ms, 3),
               "sets_per_s": round(B / ms * 1e3, 1), "registered_fraction": round(reg, 3)}
        devs, cpu_s = [], []
        for i in range(min(a.check, len(base))):
            t0 = time.perf_counter()
            ref = O.forward_testing(sd, cfg, base[i]["corr_pos"], base[i]["src_keypts"], base[i]["tgt_keypts"])
            cpu_s.append(time.perf_counter() - t0)
            ok = float((ref["final_trans"] - base[i]["gt_trans"]).abs().max()) < scale
            d = float((out["final_trans"][i].cpu() - ref["final_trans"]).abs().max())
            flips = int((out["final_labels"][i].cpu() != ref["final_labels"]).sum())
            devs.append({"oracle_registered": ok, "max_abs_dT": d, "label_flips": flips})
        if devs:
            rec["vs_oracle"] = devs; rec["oracle_cpu_s_per_set"] = round(float(np.mean(cpu_s)), 3)
        print(json.dumps(rec), flush=True)
        del m; torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
