"""CPU oracle for the PointDSC testing-mode forward path.  TEST INFRASTRUCTURE ONLY.

This file is a stage-by-stage restatement (torch CPU, fp32) of the algorithm in the
reference's `models/PointDSC.py:128-197` and its helpers.  It exists so the CUDA
engine can be checked stage by stage; only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  The product
package `pointdsc_b200` never imports it and has no CPU fallback.

Third-party arithmetic: the reference is pure PyTorch (pinned pytorch=1.6.0,
environment.yml:25); its path calls ATen matmul/softmax/topk/argsort and LAPACK
`torch.svd`.  The oracle calls the same library (torch 2.11 CPU, fp32) so its
summation orders are as close to the reference's as a restatement can be.

Pinning: the reference holds NO golden vectors or tests for this path.
The oracle is pinned instead against outputs of the reference itself, produced in
the build container by `tests/golden/make_golden.py` (imports /root/reference,
released snapshots) and committed under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` replays them.  Documented deviations from the
reference, both only in the order of exactly tied elements:
  * pick_seeds ranks with a STABLE descending sort (ties -> lowest index first);
    the reference's `torch.argsort` is unstable (PointDSC.py:217).
  * seed-row kNN breaks distance ties by lowest index; `torch.topk` tie order is
    unspecified (common.py:68).

Every function takes/returns plain tensors for ONE correspondence set (no batch
axis): the reference's testing mode asserts bs == 1 (PointDSC.py:210, :414), and a
batched engine call is defined as the loop of these per-pair calls.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping

import torch

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, reference PointDSC.py:14,59


# --------------------------------------------------------------------------------------
# stage i — spatial consistency matrix                              (PointDSC.py:150-153)
# --------------------------------------------------------------------------------------
def pairwise_length(p: torch.Tensor) -> torch.Tensor:
    """[N,3] -> [N,N] Euclidean lengths ||p_i - p_j|| (PointDSC.py:151)."""
    return torch.norm(p[:, None, :] - p[None, :, :], dim=-1)


def sc_matrix(src: torch.Tensor, tgt: torch.Tensor, sigma_d: float):
    """Returns (src_dist [N,N], SC [N,N]);  SC = max(0, 1 - (|xi-xj| - |yi-yj|)^2 / sigma_d^2)
    (PointDSC.py:151-153)."""
    src_dist = pairwise_length(src)
    diff = src_dist - pairwise_length(tgt)
    sig = torch.tensor(float(sigma_d), dtype=torch.float32)
    sc = torch.clamp(1.0 - diff ** 2 / sig ** 2, min=0)
    return src_dist, sc


# --------------------------------------------------------------------------------------
# stage ii — SCNonlocal encoder                                      (PointDSC.py:9-77)
# --------------------------------------------------------------------------------------
def _lin(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """1x1 Conv1d as a row-major linear map: x [N,Cin], w [Cout,Cin,1] -> [N,Cout]."""
    return x @ w[:, :, 0].t() + b


def _bn(x: torch.Tensor, sd: Mapping[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """Eval-mode BatchNorm1d over the channel axis of x [N,C]."""
    mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    return (x - mean) / torch.sqrt(var + BN_EPS) * sd[prefix + ".weight"] + sd[prefix + ".bias"]


def nonlocal_block(feat: torch.Tensor, sc: torch.Tensor, sd: Mapping[str, torch.Tensor], prefix: str):
    """One SCNonlocal block on feat [N,C] (PointDSC.py:27-45, single head).
    logits = SC * (Q K^T / sqrt(C)); P = softmax over keys; msg = P V; feat + fc_message(msg)."""
    c = feat.shape[1]
    q = _lin(feat, sd[prefix + ".projection_q.weight"], sd[prefix + ".projection_q.bias"])
    k = _lin(feat, sd[prefix + ".projection_k.weight"], sd[prefix + ".projection_k.bias"])
    v = _lin(feat, sd[prefix + ".projection_v.weight"], sd[prefix + ".projection_v.bias"])
    logits = sc * ((q @ k.t()) / (c ** 0.5))
    p = torch.softmax(logits, dim=-1)
    msg = p @ v
    m = _lin(msg, sd[prefix + ".fc_message.0.weight"], sd[prefix + ".fc_message.0.bias"])
    m = torch.relu(_bn(m, sd, prefix + ".fc_message.1"))
    m = _lin(m, sd[prefix + ".fc_message.3.weight"], sd[prefix + ".fc_message.3.bias"])
    m = torch.relu(_bn(m, sd, prefix + ".fc_message.4"))
    m = _lin(m, sd[prefix + ".fc_message.6.weight"], sd[prefix + ".fc_message.6.bias"])
    return feat + m


def encoder(corr_pos: torch.Tensor, sc: torch.Tensor, sd: Mapping[str, torch.Tensor], num_layers: int,
            keep_layers: bool = False):
    """NonLocalNet.forward (PointDSC.py:65-77): corr_pos [N,6] -> features [N,C]."""
    feat = _lin(corr_pos, sd["encoder.layer0.weight"], sd["encoder.layer0.bias"])
    per_layer = []
    for i in range(num_layers):
        pre = f"encoder.blocks.PointCN_layer_{i}"
        feat = _lin(feat, sd[pre + ".0.weight"], sd[pre + ".0.bias"])
        feat = torch.relu(_bn(feat, sd, pre + ".1"))
        feat = nonlocal_block(feat, sc, sd, f"encoder.blocks.NonLocal_layer_{i}")
        if keep_layers:
            per_layer.append(feat)
    return (feat, per_layer) if keep_layers else feat


def normalize_features(feat: torch.Tensor) -> torch.Tensor:
    """F.normalize(p=2, dim=-1), eps 1e-12 (PointDSC.py:156)."""
    return feat / feat.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def classify(feat: torch.Tensor, sd: Mapping[str, torch.Tensor]) -> torch.Tensor:
    """Confidence logits from the UN-normalised features (PointDSC.py:107-113, :171)."""
    h = torch.relu(_lin(feat, sd["classification.0.weight"], sd["classification.0.bias"]))
    h = torch.relu(_lin(h, sd["classification.2.weight"], sd["classification.2.bias"]))
    return _lin(h, sd["classification.4.weight"], sd["classification.4.bias"])[:, 0]


# --------------------------------------------------------------------------------------
# stage iii-a — NMS seed selection                                  (PointDSC.py:199-217)
# --------------------------------------------------------------------------------------
def local_max_mask(src_dist: torch.Tensor, scores: torch.Tensor, radius: float) -> torch.Tensor:
    """is_local_max_i = all_j (s_i >= s_j  or  dist_ij >= R)   (PointDSC.py:213-216)."""
    rel = (scores[:, None] >= scores[None, :]) | (src_dist >= radius)
    return rel.all(dim=-1)


def pick_seeds(src_dist: torch.Tensor, scores: torch.Tensor, radius: float, max_num: int) -> torch.Tensor:
    """Top-`max_num` of scores*is_local_max, descending; exact ties -> lowest index first.
    Returns int64 [max_num] (PointDSC.py:217)."""
    key = scores * local_max_mask(src_dist, scores, radius).float()
    order = torch.sort(key, descending=True, stable=True)[1]
    return order[:max_num]


def top_confidence_seeds(scores: torch.Tensor, max_num: int) -> torch.Tensor:
    """Non-testing seed rule (PointDSC.py:176), stable tie-break."""
    return torch.sort(scores, descending=True, stable=True)[1][:max_num]


# --------------------------------------------------------------------------------------
# stage iv — per-seed neighbourhoods and compatibility        (common.py:48-69, PointDSC.py:250-278)
# --------------------------------------------------------------------------------------
def knn_seed_rows(normed: torch.Tensor, seeds: torch.Tensor, k: int) -> torch.Tensor:
    """Feature-space kNN of the seed rows only.  The reference computes all N rows then gathers
    the seed rows (PointDSC.py:251-252) - identical result.  distance = 2 - 2 f_s.f_j; take the
    k+1 smallest (sorted ascending, ties -> lowest index) and drop the first (common.py:58-68)."""
    inner = 2 * (normed[seeds] @ normed.t())
    dist = 2 - inner
    order = torch.sort(dist, dim=-1, stable=True)[1]
    return order[:, 1:k + 1]


def seed_compatibility(normed, src, tgt, knn_idx, sigma: float, sigma_d: float):
    """[S,k,k] compatibility = feature term * spatial term, zero diagonal (PointDSC.py:257-278)."""
    sig = torch.tensor(float(sigma), dtype=torch.float32)
    sig_d = torch.tensor(float(sigma_d), dtype=torch.float32)
    f = normed[knn_idx]                                   # [S,k,C]
    fm = torch.clamp(1 - (1 - f @ f.transpose(1, 2)) / sig ** 2, min=0)
    a, b = src[knn_idx], tgt[knn_idx]                     # [S,k,3]
    la = ((a[:, :, None, :] - a[:, None, :, :]) ** 2).sum(-1) ** 0.5
    lb = ((b[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1) ** 0.5
    sm = torch.clamp(1 - (la - lb) ** 2 / sig_d ** 2, min=0)
    m = fm * sm
    idx = torch.arange(m.shape[1])
    m[:, idx, idx] = 0
    return m


# --------------------------------------------------------------------------------------
# stage iii-b — power iteration                                     (PointDSC.py:338-358)
# --------------------------------------------------------------------------------------
def leading_eigenvector(m: torch.Tensor, num_iterations: int):
    """Power iteration on [S,k,k] from the all-ones vector; stops when ALL S problems of the pair
    satisfy torch.allclose(new, last) (rtol 1e-5, atol 1e-8).  Returns ([S,k], iterations run)."""
    v = torch.ones_like(m[:, :, 0:1])
    last = v
    iters = 0
    for _ in range(num_iterations):
        v = torch.bmm(m, v)
        v = v / (torch.norm(v, dim=1, keepdim=True) + 1e-6)
        iters += 1
        if torch.allclose(v, last):
            break
        last = v
    return v.squeeze(-1), iters


# --------------------------------------------------------------------------------------
# stage iv core — weighted Procrustes / Kabsch                      (common.py:7-45)
# --------------------------------------------------------------------------------------
def weighted_kabsch(a: torch.Tensor, b: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """a,b [P,m,3], w [P,m] -> [P,4,4] with  b ~= R a + t.
    centroids use sum(w)+1e-6 (common.py:24-25); H = Am^T diag(w) Bm (:32-33);
    R = V diag(1,1,det(V U^T)) U^T (:36-41); t = cb - R ca (:42)."""
    w = torch.where(w < 0, torch.zeros_like(w), w)
    wsum = w.sum(dim=1, keepdim=True)[:, :, None] + 1e-6
    ca = (a * w[:, :, None]).sum(dim=1, keepdim=True) / wsum
    cb = (b * w[:, :, None]).sum(dim=1, keepdim=True) / wsum
    am, bm = a - ca, b - cb
    h = am.transpose(1, 2) @ (w[:, :, None] * bm)
    u, _, vh = torch.linalg.svd(h)
    v = vh.transpose(1, 2)
    d = torch.det(v @ u.transpose(1, 2))
    e = torch.eye(3)[None].repeat(a.shape[0], 1, 1)
    e[:, 2, 2] = d
    r = v @ e @ u.transpose(1, 2)
    t = cb.transpose(1, 2) - r @ ca.transpose(1, 2)
    out = torch.eye(4)[None].repeat(a.shape[0], 1, 1)
    out[:, :3, :3] = r
    out[:, :3, 3:4] = t
    return out


def seed_hypotheses(src, tgt, knn_idx, eig):
    """Normalise the eigenvector into weights (PointDSC.py:282) and solve one weighted Kabsch per
    seed over its k neighbours (:289-316).  Returns (weights [S,k], trans [S,4,4])."""
    w = eig / (eig.sum(dim=-1, keepdim=True) + 1e-6)
    return w, weighted_kabsch(src[knn_idx], tgt[knn_idx], w)


# --------------------------------------------------------------------------------------
# stage iv — hypothesis scoring and selection                       (PointDSC.py:325-336)
# --------------------------------------------------------------------------------------
def residuals(trans: torch.Tensor, src: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """||R x_n + t - y_n|| for trans [S,4,4] -> [S,N]."""
    pred = torch.einsum("snm,km->skn", trans[:, :3, :3], src) + trans[:, :3, 3][:, None, :]
    return torch.norm(pred - tgt[None], dim=-1)


def select_hypothesis(trans, src, tgt, inlier_threshold: float):
    """Returns (fitness [S], best index, best trans [4,4], labels [N] in {0,1})."""
    d = residuals(trans, src, tgt)
    fitness = (d < inlier_threshold).float().mean(dim=-1)
    best = int(fitness.argmax())
    return fitness, best, trans[best], (d[best] < inlier_threshold).float()


# --------------------------------------------------------------------------------------
# stage v — post refinement                                          (PointDSC.py:403-438)
# --------------------------------------------------------------------------------------
def refinement_threshold(ctor_inlier_threshold: float) -> float:
    """0.10 iff the constructor threshold is exactly 0.10, else 1.2 (PointDSC.py:415-418)."""
    return 0.10 if ctor_inlier_threshold == 0.10 else 1.2


def post_refinement(trans: torch.Tensor, src, tgt, ctor_inlier_threshold: float, max_iters: int = 20):
    """<= 20 reweighted Kabsch iterations over the current inliers, w = 1/(1+(d/tau)^2); stops when the
    inlier count repeats.  Returns ([4,4], number of Kabsch solves)."""
    tau = refinement_threshold(ctor_inlier_threshold)
    prev = 0
    solves = 0
    for _ in range(max_iters):
        warped = src @ trans[:3, :3].t() + trans[:3, 3][None, :]
        d = torch.norm(warped - tgt, dim=-1)
        inl = d < tau
        cnt = int(inl.sum())
        if cnt == prev:
            break
        prev = cnt
        w = (1 / (1 + (d / tau) ** 2))[inl]
        trans = weighted_kabsch(src[inl][None], tgt[inl][None], w[None])[0]
        solves += 1
    return trans, solves


# --------------------------------------------------------------------------------------
# whole path                                                         (PointDSC.py:128-197)
# --------------------------------------------------------------------------------------
def forward_testing(sd: Mapping[str, torch.Tensor], cfg: Mapping[str, float], corr_pos, src, tgt,
                    keep_layers: bool = False) -> Dict[str, torch.Tensor]:
    """Testing-mode forward for ONE correspondence set; returns every stage-boundary tensor.
    cfg keys: num_layers, num_iterations, ratio, inlier_threshold, k, nms_radius.  sigma_d and sigma
    are read from the state dict (`sigma_spat`, `sigma`) as the reference does."""
    with torch.no_grad():
        n = corr_pos.shape[0]
        sigma_d = float(sd["sigma_spat"][0])
        sigma = float(sd["sigma"][0])
        out: Dict[str, torch.Tensor] = {}
        src_dist, sc = sc_matrix(src, tgt, sigma_d)
        enc = encoder(corr_pos, sc, sd, int(cfg["num_layers"]), keep_layers=keep_layers)
        feat, per_layer = enc if keep_layers else (enc, None)
        normed = normalize_features(feat)
        conf = classify(feat, sd)
        num_seeds = int(n * float(cfg["ratio"]))
        seeds = pick_seeds(src_dist, conf, float(cfg["nms_radius"]), num_seeds)
        k = min(int(cfg["k"]), n - 1)
        knn_idx = knn_seed_rows(normed, seeds, k)
        compat = seed_compatibility(normed, src, tgt, knn_idx, sigma, sigma_d)
        eig, iters = leading_eigenvector(compat, int(cfg["num_iterations"]))
        weights, seed_trans = seed_hypotheses(src, tgt, knn_idx, eig)
        fitness, best, init_trans, labels = select_hypothesis(seed_trans, src, tgt, float(cfg["inlier_threshold"]))
        final_trans, solves = post_refinement(init_trans, src, tgt, float(cfg["inlier_threshold"]))
        out.update(src_dist=src_dist, sc=sc, features=feat, normed=normed, confidence=conf, seeds=seeds,
                   knn_idx=knn_idx, compat=compat, eig=eig, power_iters=torch.tensor(iters),
                   seed_weights=weights, seed_trans=seed_trans, fitness=fitness, best=torch.tensor(best),
                   init_trans=init_trans, final_labels=labels, final_trans=final_trans,
                   refine_solves=torch.tensor(solves))
        if keep_layers:
            out["layer_features"] = torch.stack(per_layer, 0)
        return out


def forward_batch(sd, cfg, corr_pos, src, tgt):
    """Loop of per-pair testing forwards: the definition of a batched engine call.
    Returns (final_trans [B,4,4], final_labels [B,N])."""
    tr, lb = [], []
    for b in range(corr_pos.shape[0]):
        o = forward_testing(sd, cfg, corr_pos[b], src[b], tgt[b])
        tr.append(o["final_trans"])
        lb.append(o["final_labels"])
    return torch.stack(tr, 0), torch.stack(lb, 0)


def feature_similarity(normed: torch.Tensor, sigma: float) -> torch.Tensor:
    """M = clamp(1 - (1 - F F^T) / sigma^2, 0, 1) with a zero diagonal  (PointDSC.py:160-165)."""
    m = normed @ normed.t()
    m = torch.clamp(1 - (1 - m) / sigma ** 2, min=0, max=1)
    m[torch.arange(m.shape[0]), torch.arange(m.shape[0])] = 0
    return m


def forward_validation(sd: Mapping[str, torch.Tensor], cfg: Mapping[str, float], corr_pos, src, tgt):
    """The forward WITHOUT the 'testing' key, eval-mode BatchNorm (PointDSC.py:158-165, :176, :182, :190-191), for a batch
    [B,N,...]: seeds = top-S by confidence (no suppression), ONE power-iteration exit for the whole batch (the reference's
    allclose spans [bs * S, k, 1]), no refinement, final_labels = confidence logits.
    Returns final_trans [B,4,4], confidence [B,N], M [B,N,N], seeds [B,S], iterations run."""
    with torch.no_grad():
        bsz, n = corr_pos.shape[0], corr_pos.shape[1]
        sigma_d, sigma = float(sd["sigma_spat"][0]), float(sd["sigma"][0])
        num_seeds = int(n * float(cfg["ratio"]))
        k = min(int(cfg["k"]), n - 1)
        confs, normeds, seeds, knns, compats = [], [], [], [], []
        for b in range(bsz):
            _, sc = sc_matrix(src[b], tgt[b], sigma_d)
            feat = encoder(corr_pos[b], sc, sd, int(cfg["num_layers"]))
            normed = normalize_features(feat)
            conf = classify(feat, sd)
            sd_b = top_confidence_seeds(conf, num_seeds)
            knn_idx = knn_seed_rows(normed, sd_b, k)
            confs.append(conf); normeds.append(normed); seeds.append(sd_b); knns.append(knn_idx)
            compats.append(seed_compatibility(normed, src[b], tgt[b], knn_idx, sigma, sigma_d))
        eig, iters = leading_eigenvector(torch.cat(compats, 0), int(cfg["num_iterations"]))     # batch-wide early exit
        eig = eig.view(bsz, num_seeds, k)
        trans = []
        for b in range(bsz):
            _, seed_trans = seed_hypotheses(src[b], tgt[b], knns[b], eig[b])
            _, _, init_trans, _ = select_hypothesis(seed_trans, src[b], tgt[b], float(cfg["inlier_threshold"]))
            trans.append(init_trans)
        m = torch.stack([feature_similarity(x, sigma) for x in normeds], 0)
        return dict(final_trans=torch.stack(trans, 0), final_labels=torch.stack(confs, 0), M=m, seeds=torch.stack(seeds, 0),
                    power_iters=iters)


def default_config(dataset: str = "3dmatch") -> Dict[str, float]:
    """Constructor arguments the reference's eval drivers use.
    3DMatch: evaluation/test_3DMatch.py:215-224 (inlier_threshold left at the ctor default 0.10,
    nms_radius = config.inlier_threshold = 0.1).  KITTI: evaluation/test_KITTI.py:166-191
    (inlier_threshold 0.6, sigma_d 1.2, nms_radius 0.6)."""
    if dataset == "3dmatch":
        return dict(num_layers=12, num_iterations=10, ratio=0.1, inlier_threshold=0.10, sigma_d=0.10, k=40,
                    nms_radius=0.10)
    if dataset == "kitti":
        return dict(num_layers=12, num_iterations=10, ratio=0.1, inlier_threshold=0.6, sigma_d=1.2, k=40,
                    nms_radius=0.6)
    raise ValueError(dataset)
