#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the unmodified reference (`models/PointDSC.py`, `models/common.py`, `utils/SE3.py`),
loads the released snapshots, runs the testing-mode forward on CPU (fp32) over seeded synthetic
correspondence sets (pointdsc_b200.synth) and records every stage-boundary tensor
by wrapping the reference's own methods (no reference code is copied or edited).  Outputs:

  snapshot_<dataset>.npz      the released state dict, key for key (incl. the stray `gamma`)
  case_<dataset>_n<N>_s<seed>.npz   inputs + stage intermediates + outputs of one forward

The reference publishes no golden vectors for this path, so these files are the
pin for `oracle/pointdsc_oracle.py` and, through it, for the CUDA engine.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("POINTDSC_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import models.PointDSC as ref_mod  # noqa: E402  (the reference)
from pointdsc_b200.synth import make_pair  # noqa: E402

SNAP = {"3dmatch": "PointDSC_3DMatch_release", "kitti": "PointDSC_KITTI_release"}
# ctor arguments exactly as the eval drivers pass them (test_3DMatch.py:215-224, test_KITTI.py:166-191)
CTOR = {
    "3dmatch": dict(inlier_threshold=0.10, sigma_d=0.10, nms_radius=0.10),
    "kitti": dict(inlier_threshold=0.6, sigma_d=1.2, nms_radius=0.6),
}
# (dataset, N, seed, inlier_ratio, detail[, k])   detail: "full" keeps N x N matrices and layer features, "full1k" the same
# without src_dist (bench-size case: SC tiling with 8 query tiles and per-layer features at N = 1000); k = ctor `k`
# (neighbourhood size, default: the snapshot's 40)
CASES = [
    ("3dmatch", 256, 0, 0.5, "full"),
    ("3dmatch", 256, 1, 0.2, "full"),
    ("3dmatch", 41, 2, 0.6, "full"),      # k clamps to N-1 = 40
    ("3dmatch", 30, 3, 0.6, "full"),      # k = 29 < 40, S = 3
    ("3dmatch", 1000, 0, 0.3, "feat"),
    ("3dmatch", 1000, 1, 0.1, "feat"),
    ("3dmatch", 1000, 2, 0.5, "io"),
    ("3dmatch", 1000, 3, 0.05, "io"),
    ("3dmatch", 1000, 4, 0.3, "io"),
    ("3dmatch", 1000, 5, 0.2, "io"),
    ("3dmatch", 500, 6, 0.3, "io"),
    ("3dmatch", 2000, 7, 0.2, "io"),
    ("kitti", 256, 0, 0.5, "full"),
    ("kitti", 1000, 1, 0.3, "feat"),
    ("kitti", 1000, 2, 0.1, "io"),
    ("kitti", 2000, 3, 0.3, "io"),
    # round 2: every BASELINE.json configuration (B: KITTI N = 5000; C: N = 2000 with k = 80; D: N = 5000 3DMatch),
    # the low inlier ratios of SURVEY.md §8(d), and one bench-size case with the N x N / per-layer tensors
    ("kitti", 5000, 4, 0.3, "io"),
    ("kitti", 5000, 5, 0.5, "io"),
    ("3dmatch", 5000, 8, 0.3, "io"),
    ("3dmatch", 2000, 9, 0.3, "io", 80),
    ("3dmatch", 2000, 10, 0.2, "io", 80),
    ("3dmatch", 2000, 11, 0.05, "io"),
    ("3dmatch", 2000, 12, 0.1, "io"),
    ("3dmatch", 1000, 13, 0.1, "io", 80),
    ("kitti", 1000, 6, 0.05, "io"),
    ("3dmatch", 1000, 14, 0.3, "full1k"),
    # sizes that are a multiple of no tile size (the same sets tests/test_gpu_parity.py::test_ragged_sizes_vs_oracle runs
    # against the oracle: here the REFERENCE is the authority)
    ("3dmatch", 257, 357, 0.6, "io"),
    ("3dmatch", 1003, 1103, 0.6, "io"),
]


def build_model(dataset, k=None):
    cfg = json.load(open(os.path.join(REF, "snapshot", SNAP[dataset], "config.json")))
    model = ref_mod.PointDSC(in_dim=cfg["in_dim"], num_layers=cfg["num_layers"], num_channels=cfg["num_channels"],
                             num_iterations=cfg["num_iterations"], ratio=cfg["ratio"], k=k or cfg["k"], **CTOR[dataset])
    sd = torch.load(os.path.join(REF, "snapshot", SNAP[dataset], "models", "model_best.pkl"), map_location="cpu")
    res = model.load_state_dict(sd, strict=False)
    assert res.missing_keys == [] and res.unexpected_keys == ["gamma"], res
    model.eval()
    return model, sd


def run_case(model, dataset, n, seed, ratio, detail):
    rec = {}
    pair = make_pair(seed, n, dataset, ratio)
    data = {"corr_pos": pair["corr_pos"][None], "src_keypts": pair["src_keypts"][None],
            "tgt_keypts": pair["tgt_keypts"][None], "testing": True}

    orig = dict(pick=model.pick_seeds, eig=model.cal_leading_eigenvector, seed=model.cal_seed_trans,
                refine=model.post_refinement, knn=ref_mod.knn, rigid=ref_mod.rigid_transform_3d)
    counters = {"eig_iters": 0, "rigid_calls": 0}

    def pick(dists, scores, R, max_num):
        rec["src_dist"] = dists[0].clone()
        rec["confidence"] = scores[0].clone()
        out = orig["pick"](dists, scores, R, max_num)
        rec["seeds"] = out[0].clone()
        return out

    def knn(x, k, ignore_self=False, normalized=True):
        rec["normed"] = x[0].clone()
        out = orig["knn"](x, k, ignore_self=ignore_self, normalized=normalized)
        rec["knn_all"] = out[0].clone()
        return out

    def eig(M, method="power"):
        rec["compat"] = M.clone()
        real_bmm = torch.bmm

        def counting_bmm(a, b):
            counters["eig_iters"] += 1
            return real_bmm(a, b)
        torch.bmm = counting_bmm
        try:
            out = orig["eig"](M, method)
        finally:
            torch.bmm = real_bmm
        rec["eig"] = out.clone()
        return out

    def rigid(A, B, weights=None, weight_threshold=0):
        counters["rigid_calls"] += 1
        if counters["rigid_calls"] == 1:
            rec["seed_weights"] = weights.clone()
        return orig["rigid"](A, B, weights, weight_threshold)

    def seed_trans_wrap(seeds, feats, s, t):
        out = orig["seed"](seeds, feats, s, t)
        rec["seed_trans"], rec["fitness"] = out[0][0].clone(), out[1][0].clone()
        rec["init_trans"], rec["final_labels"] = out[2][0].clone(), out[3][0].clone()
        return out

    layer_feats = []
    hooks = []
    for i in range(model.encoder.num_layers):
        hooks.append(model.encoder.blocks[f"NonLocal_layer_{i}"].register_forward_hook(
            lambda m, a, o: layer_feats.append(o[0].t().clone())))
    hooks.append(model.encoder.register_forward_hook(lambda m, a, o: rec.__setitem__("sc", a[1][0].clone())))

    model.pick_seeds, model.cal_leading_eigenvector, model.cal_seed_trans = pick, eig, seed_trans_wrap
    ref_mod.knn, ref_mod.rigid_transform_3d = knn, rigid
    try:
        with torch.no_grad():
            out = model(data)
    finally:
        model.pick_seeds, model.cal_leading_eigenvector, model.cal_seed_trans = orig["pick"], orig["eig"], orig["seed"]
        ref_mod.knn, ref_mod.rigid_transform_3d = orig["knn"], orig["rigid"]
        for h in hooks:
            h.remove()

    rec["final_trans"] = out["final_trans"][0]
    assert torch.equal(out["final_labels"][0], rec["final_labels"])
    rec["features"] = layer_feats[-1]
    rec["best"] = rec["fitness"].argmax()
    rec["knn_idx"] = rec["knn_all"][rec["seeds"]]
    rec["power_iters"] = torch.tensor(counters["eig_iters"])
    rec["refine_solves"] = torch.tensor(counters["rigid_calls"] - 1)
    rec["layer_features"] = torch.stack([layer_feats[i] for i in (0, 5, 11)], 0)

    keep_io = ["confidence", "seeds", "knn_idx", "eig", "seed_weights", "seed_trans", "fitness", "best",
               "init_trans", "final_labels", "final_trans", "power_iters", "refine_solves"]
    keep = {"io": keep_io, "feat": keep_io + ["features", "normed"],
            "full": keep_io + ["features", "normed", "sc", "src_dist", "layer_features", "compat"],
            "full1k": keep_io + ["features", "normed", "sc", "layer_features", "compat"]}[detail]
    arrays = {k: rec[k].detach().cpu().numpy() for k in keep}
    for k in ("seeds", "knn_idx"):
        arrays[k] = arrays[k].astype(np.int32)
    arrays.update(corr_pos=pair["corr_pos"].numpy(), src_keypts=pair["src_keypts"].numpy(),
                  tgt_keypts=pair["tgt_keypts"].numpy(), gt_trans=pair["gt_trans"].numpy(),
                  gt_labels=pair["gt_labels"].numpy().astype(np.uint8))
    meta = dict(dataset=dataset, n=n, seed=seed, inlier_ratio=ratio, detail=detail, torch=torch.__version__,
                layer_features_layers=[0, 5, 11], k=int(model.k), **CTOR[dataset])
    arrays["meta"] = np.array(json.dumps(meta))
    return arrays


def main():
    """Existing fixture files are kept (pass --force to regenerate everything)."""
    force = "--force" in sys.argv
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for dataset in SNAP:
        model, sd = build_model(dataset)
        snap_path = os.path.join(HERE, f"snapshot_{dataset}.npz")
        if force or not os.path.exists(snap_path):
            np.savez(snap_path, **{k: v.numpy() for k, v in sd.items()})
        for case in CASES:
            ds, n, seed, ratio, detail = case[:5]
            k = case[5] if len(case) > 5 else None
            if ds != dataset:
                continue
            path = os.path.join(HERE, f"case_{ds}_n{n}_s{seed}" + (f"_k{k}" if k else "") + ".npz")
            if os.path.exists(path) and not force:
                continue
            m = model if k is None else build_model(ds, k)[0]
            arrays = run_case(m, ds, n, seed, ratio, detail)
            np.savez_compressed(path, **arrays)
            gt, ft = arrays["gt_trans"], arrays["final_trans"]
            print(f"{os.path.basename(path)}: |T-gt|max={np.abs(gt - ft).max():.4f} inl={int(arrays['final_labels'].sum())} "
                  f"iters={int(arrays['power_iters'])} solves={int(arrays['refine_solves'])} "
                  f"{os.path.getsize(path) / 1e3:.0f} KB", flush=True)


if __name__ == "__main__":
    main()
