#!/usr/bin/env python
"""Design-time probe (CPU): how much does each candidate tensor-core operand format move the
final transform away from the fp32 oracle?  Emulates operand rounding of the encoder's
contractions (everything accumulates in fp32, like tcgen05 kind::f16/tf32 with f32 D).

    python tools/numerics_probe.py --n 1000 --pairs 12

Not part of the product or the tests; results are summarised in DESIGN.md.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pointdsc_oracle as O  # noqa: E402
from pointdsc_b200.synth import make_pair  # noqa: E402


def rnd(x, fmt):
    if fmt == "fp32":
        return x
    if fmt == "fp16":
        return x.half().float()
    if fmt == "bf16":
        return x.bfloat16().float()
    if fmt == "tf32":
        i = x.contiguous().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(fmt)


def mm(a, b, fmt):
    """a @ b with operands in `fmt`; '<f>x3' = hi/lo split, 3 products (hi*hi + hi*lo + lo*hi);
    '<f>x2a' = only the A operand is split (hi*hi + lo*hi)."""
    if fmt.endswith("x3"):
        f = fmt[:-2]
        ah, bh = rnd(a, f), rnd(b, f)
        al, bl = rnd(a - ah, f), rnd(b - bh, f)
        return ah @ bh + (ah @ bl + al @ bh)
    if fmt.endswith("x2a"):
        f = fmt[:-3]
        ah, bh = rnd(a, f), rnd(b, f)
        al = rnd(a - ah, f)
        return ah @ bh + al @ bh
    if fmt.endswith("x2b"):
        f = fmt[:-3]
        ah, bh = rnd(a, f), rnd(b, f)
        bl = rnd(b - bh, f)
        return ah @ bh + ah @ bl
    return rnd(a, fmt) @ rnd(b, fmt)


def encoder_emulated(corr_pos, sc, sd, num_layers, qk, pv, lin, scfmt):
    def L(x, wname, bname):
        return mm(x, sd[wname][:, :, 0].t(), lin) + sd[bname]

    def bnfold(x, prefix):
        return O._bn(x, sd, prefix)

    sc_r = rnd(sc, scfmt)
    feat = O._lin(corr_pos, sd["encoder.layer0.weight"], sd["encoder.layer0.bias"])
    for i in range(num_layers):
        pre = f"encoder.blocks.PointCN_layer_{i}"
        feat = torch.relu(bnfold(L(feat, pre + ".0.weight", pre + ".0.bias"), pre + ".1"))
        p = f"encoder.blocks.NonLocal_layer_{i}"
        q = L(feat, p + ".projection_q.weight", p + ".projection_q.bias")
        k = L(feat, p + ".projection_k.weight", p + ".projection_k.bias")
        v = L(feat, p + ".projection_v.weight", p + ".projection_v.bias")
        logits = sc_r * (mm(q, k.t(), qk) / (feat.shape[1] ** 0.5))
        # online-softmax form the kernel uses: unnormalised exp in the PV product, divide after
        mx = logits.max(dim=-1, keepdim=True)[0]
        e = torch.exp(logits - mx)
        msg = mm(e, v, pv) / e.sum(-1, keepdim=True)
        m = torch.relu(bnfold(L(msg, p + ".fc_message.0.weight", p + ".fc_message.0.bias"), p + ".fc_message.1"))
        m = torch.relu(bnfold(L(m, p + ".fc_message.3.weight", p + ".fc_message.3.bias"), p + ".fc_message.4"))
        m = L(m, p + ".fc_message.6.weight", p + ".fc_message.6.bias")
        feat = feat + m
    return feat


def forward_emulated(sd, cfg, pair, **fmt):
    with torch.no_grad():
        src, tgt, cp = pair["src_keypts"], pair["tgt_keypts"], pair["corr_pos"]
        n = cp.shape[0]
        src_dist, sc = O.sc_matrix(src, tgt, float(sd["sigma_spat"][0]))
        feat = encoder_emulated(cp, sc, sd, cfg["num_layers"], **fmt)
        normed = O.normalize_features(feat)
        conf = O.classify(feat, sd)
        seeds = O.pick_seeds(src_dist, conf, cfg["nms_radius"], int(n * cfg["ratio"]))
        k = min(cfg["k"], n - 1)
        knn_idx = O.knn_seed_rows(normed, seeds, k)
        compat = O.seed_compatibility(normed, src, tgt, knn_idx, float(sd["sigma"][0]), float(sd["sigma_spat"][0]))
        eig, _ = O.leading_eigenvector(compat, cfg["num_iterations"])
        _, st = O.seed_hypotheses(src, tgt, knn_idx, eig)
        _, best, init, labels = O.select_hypothesis(st, src, tgt, cfg["inlier_threshold"])
        final, _ = O.post_refinement(init, src, tgt, cfg["inlier_threshold"])
        return dict(final_trans=final, confidence=conf, seeds=seeds, features=feat, final_labels=labels)


CONFIGS = {
    "bf16": dict(qk="bf16", pv="bf16", lin="bf16", scfmt="fp32"),
    "bf16-attn": dict(qk="bf16", pv="bf16", lin="fp32", scfmt="fp32"),
    "fp16": dict(qk="fp16", pv="fp16", lin="fp16", scfmt="fp32"),
    "fp16-attn": dict(qk="fp16", pv="fp16", lin="fp32", scfmt="fp32"),
    "tf32": dict(qk="tf32", pv="tf32", lin="tf32", scfmt="fp32"),
    "fp16x3qk": dict(qk="fp16x3", pv="fp16", lin="fp16x3", scfmt="fp32"),
    "fp16x3": dict(qk="fp16x3", pv="fp16x3", lin="fp16x3", scfmt="fp32"),
    "fp16x3+scfp16": dict(qk="fp16x3", pv="fp16x3", lin="fp16x3", scfmt="fp16"),
    "bf16x3": dict(qk="bf16x3", pv="bf16x3", lin="bf16x3", scfmt="fp32"),
    "pv-x2a": dict(qk="fp16x3", pv="fp16x2a", lin="fp16x3", scfmt="fp32"),
    "pv-x2b": dict(qk="fp16x3", pv="fp16x2b", lin="fp16x3", scfmt="fp32"),
    "lin-x2a": dict(qk="fp16x3", pv="fp16x3", lin="fp16x2a", scfmt="fp32"),
    "lin-x2b": dict(qk="fp16x3", pv="fp16x3", lin="fp16x2b", scfmt="fp32"),
    "lin-fp16": dict(qk="fp16x3", pv="fp16x3", lin="fp16", scfmt="fp32"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--dataset", default="3dmatch")
    ap.add_argument("--configs", default=",".join(CONFIGS))
    args = ap.parse_args()
    gdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(gdir, f"snapshot_{args.dataset}.npz")).items()}
    cfg = O.default_config(args.dataset)
    ratios = [0.5, 0.3, 0.2, 0.1, 0.4, 0.15]
    stats = {c: [] for c in args.configs.split(",")}
    for s in range(args.pairs):
        pair = make_pair(1000 + s, args.n, args.dataset, ratios[s % len(ratios)])
        ref = O.forward_testing(sd, cfg, pair["corr_pos"], pair["src_keypts"], pair["tgt_keypts"])
        ok = float((ref["final_trans"] - pair["gt_trans"]).abs().max()) < 0.05 * (1 if args.dataset == "3dmatch" else 10)
        line = f"pair {s} ratio {ratios[s % len(ratios)]} oracle_ok={ok} |feat|max={float(ref['features'].abs().max()):.1f}"
        for c in stats:
            out = forward_emulated(sd, cfg, pair, **CONFIGS[c])
            dT = float((out["final_trans"] - ref["final_trans"]).abs().max())
            dC = float((out["confidence"] - ref["confidence"]).abs().max())
            dF = float((out["features"] - ref["features"]).abs().max())
            sm = float((out["seeds"] == ref["seeds"]).float().mean())
            stats[c].append((ok, dT, dC, dF, sm))
            line += f" | {c}: dT={dT:.1e} dC={dC:.1e} dF={dF:.1e} seeds={sm:.2f}"
        print(line, flush=True)
    print("\nsummary over oracle-successful pairs (max dT, median dT, max dConf):")
    for c, rows in stats.items():
        good = [r for r in rows if r[0]]
        if good:
            dts = np.array([r[1] for r in good])
            print(f"  {c:16s} max dT {dts.max():.2e}  median dT {np.median(dts):.2e}  max dConf {max(r[2] for r in good):.2e}"
                  f"  max dFeat {max(r[3] for r in good):.2e}  mean seed match {np.mean([r[4] for r in good]):.2f}")


if __name__ == "__main__":
    main()
