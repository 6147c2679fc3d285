"""Host-side logic of evaluate.py (the modernised evaluation driver): gt.log parsing, fragment loading,
ground-truth labels and the summary, on a tiny data set written in the reference's layout (datasets/ThreeDMatch.py:226-258, :338-351)."""
import os

import numpy as np
import torch

import evaluate


def _write_dataset(root, scene, poses):
    os.makedirs(os.path.join(root, "fragments", scene))
    os.makedirs(os.path.join(root, "gt_result", f"{scene}-evaluation"))
    rng = np.random.default_rng(0)
    for i in range(3):
        np.savez(os.path.join(root, "fragments", scene, f"cloud_bin_{i}_fpfh.npz"), xyz=rng.normal(size=(20, 3)),
                 feature=rng.uniform(0, 50, size=(20, 33)))
        np.savez(os.path.join(root, "fragments", scene, f"cloud_bin_{i}_fcgf.npz"), xyz=rng.normal(size=(20, 3)).astype(np.float32),
                 feature=rng.normal(size=(20, 32)).astype(np.float32))
    with open(os.path.join(root, "gt_result", f"{scene}-evaluation", "gt.log"), "w") as f:
        for (i, j), T in poses.items():
            f.write(f"{i}\t {j}\t 3\n")
            for r in range(4):
                f.write("\t ".join(f"{x:.8e}" for x in T[r]) + "\n")


def test_gt_log_and_fragments(tmp_path):
    T01 = np.eye(4); T01[:3, 3] = [0.1, -0.2, 0.3]
    a = 0.3
    T12 = np.array([[np.cos(a), -np.sin(a), 0, 1.0], [np.sin(a), np.cos(a), 0, 2.0], [0, 0, 1, 3.0], [0, 0, 0, 1]])
    root, scene = str(tmp_path), evaluate.SCENES_3DMATCH[0]
    _write_dataset(root, scene, {(0, 1): T01, (1, 2): T12})
    log = evaluate.read_gt_log(os.path.join(root, "gt_result", f"{scene}-evaluation", "gt.log"))
    assert sorted(log) == ["0_1", "1_2"] and np.allclose(log["1_2"], T12, atol=1e-7)
    pairs = evaluate.list_pairs(root, scene)
    assert [(p[0], p[1]) for p in pairs] == [(0, 1), (1, 2)]
    assert np.allclose(pairs[1][2] @ T12, np.eye(4), atol=1e-6)          # the logged pose is target -> source: inverted (ThreeDMatch.py:262)
    xyz, feat = evaluate.load_fragment(root, scene, 1, "fpfh", "cpu")
    assert xyz.dtype == torch.float32 and feat.dtype == torch.float64 and feat.shape == (20, 33)
    raw = np.load(os.path.join(root, "fragments", scene, "cloud_bin_1_fpfh.npz"))["feature"]
    assert np.allclose(feat.numpy(), raw / (np.linalg.norm(raw, axis=1, keepdims=True) + 1e-6), atol=1e-12)   # ThreeDMatch.py:257
    xyz, feat = evaluate.load_fragment(root, scene, 2, "fcgf", "cpu")
    assert feat.dtype == torch.float32 and feat.shape == (20, 32)


def test_gt_labels_and_summary():
    g = torch.Generator().manual_seed(1)
    src = torch.rand(1, 50, 3, generator=g)
    T = torch.eye(4); T[:3, 3] = torch.tensor([0.5, 0.0, -0.25])
    tgt = src + T[:3, 3]
    tgt[0, 10:20] += 1.0                                                  # ten outliers
    lab = evaluate.gt_labels({"src_keypts": src, "tgt_keypts": tgt}, T, 0.10)
    assert lab.shape == (1, 50) and int(lab.sum()) == 40 and float(lab[0, 10:20].sum()) == 0.0
    stats = np.zeros((4, len(evaluate.COLUMNS)))
    stats[:, 0] = [1, 1, 0, 1]; stats[:, 1] = [1.0, 2.0, 90.0, 3.0]; stats[:, 2] = [5.0, 6.0, 200.0, 7.0]
    stats[:, 6] = 0.8; stats[:, 7] = 0.9; stats[:, 8] = 0.85; stats[:, 11] = [0, 0, 1, 1]
    lines = []
    out = evaluate.summarise(stats, ["a", "b"], log=lines.append)
    assert out["pairs"] == 4 and abs(out["reg_recall"] - 0.75) < 1e-12
    assert abs(out["mean_re_deg"] - 2.0) < 1e-12 and abs(out["mean_te_cm"] - 6.0) < 1e-12   # successful pairs only (test_3DMatch.py:144-146)
    assert abs(out["scene_recall"] - 0.75) < 1e-12 and any("Scene a" in ln for ln in lines)
    assert evaluate.summarise(np.zeros((0, len(evaluate.COLUMNS))), log=lines.append) == {}
