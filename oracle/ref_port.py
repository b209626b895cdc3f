"""oracle/ref_port.py -- CPU restatement of PTRanking's scoring-and-loss hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
module; the shipped package ``ptranking_b200`` never does (it fails loudly when
the CUDA library is missing).

The reference (wildltr/ptranking @ f1d366c) is pure PyTorch; its arithmetic
lives in ATen.  The faithful CPU restatement is therefore written with the same
ATen operators on CPU tensors (fp32), so ATen-owned semantics -- the BCE
``log >= -100`` / ``max(p(1-p), 1e-12)`` clamps, ``torch.sort`` tie order,
exact-erf GELU, ``BatchNorm1d(track_running_stats=False)`` -- are inherited
rather than re-guessed.  Every function cites the reference file:line it
follows (paths relative to the reference checkout).

Pinning: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference (imported
from /root/reference in the authoring container) on seeded inputs and commits
its outputs under ``tests/golden/``; ``tests/test_oracle_vs_golden.py`` checks
this restatement against those fixtures and against the reference's own metric
known-answer vectors (testing/metric/testing_metric.py:20-60).  Loss / gradient
values have no reference-side test, so for them the pin is
"outputs of the reference itself run here".
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

EPS_LAMBDALOSS = 1e-8  # ptranking/ltr_global.py:8


# --------------------------------------------------------------------------- #
# metric pieces (ptranking/metric/adhoc/adhoc_metric.py)
# --------------------------------------------------------------------------- #
def gains(labels: torch.Tensor) -> torch.Tensor:
    """MultiLabel gain 2^l - 1 (adhoc_metric.py:208-209)."""
    return torch.pow(2.0, labels) - 1.0


def dcg_at_k(rankings: torch.Tensor, cutoff: Optional[int] = None) -> torch.Tensor:
    """[B,n] labels in rank order -> [B,1] DCG@cutoff (adhoc_metric.py:197-217)."""
    k = rankings.size(1) if cutoff is None else cutoff
    num = gains(rankings[:, :k])
    disc = torch.log2(torch.arange(k, dtype=torch.float, device=num.device).expand_as(num) + 2.0)
    return torch.sum(num / disc, dim=1, keepdim=True)


def dcg_at_ks(rankings: torch.Tensor, max_cutoff: int) -> torch.Tensor:
    """[B,n] -> [B,max_cutoff] running DCG (adhoc_metric.py:219-235)."""
    num = gains(rankings[:, :max_cutoff])
    disc = torch.log2(torch.arange(max_cutoff, dtype=torch.float, device=num.device).expand_as(num) + 2.0)
    return torch.cumsum(num / disc, dim=1)


def ndcg_at_k(sys_rankings: torch.Tensor, ideal_rankings: torch.Tensor, k: int) -> torch.Tensor:
    """adhoc_metric.py:237-241 -- no guard against a zero ideal DCG."""
    return dcg_at_k(sys_rankings, k) / dcg_at_k(ideal_rankings, k)


def ndcg_at_ks(sys_rankings: torch.Tensor, ideal_rankings: torch.Tensor, ks: Sequence[int]) -> torch.Tensor:
    """adhoc_metric.py:243-260 -- cutoffs beyond the list length are zero padded."""
    n = sys_rankings.size(1)
    used = [k for k in ks if k <= n] if n < max(ks) else list(ks)
    idx = torch.tensor(used, dtype=torch.long) - 1
    sys_dcg = dcg_at_ks(sys_rankings, max(used))[:, idx]
    ideal_dcg = dcg_at_ks(ideal_rankings, max(used))[:, idx]
    out = sys_dcg / ideal_dcg
    if n < max(ks):
        padded = torch.zeros(sys_rankings.size(0), len(ks))
        padded[:, : len(used)] = out
        return padded
    return out


def rank_labels_by_scores(scores: torch.Tensor, labels: torch.Tensor):
    """Evaluator.ndcg_at_k's ranking step (base/ranker.py:50-56)."""
    _, order = torch.sort(scores, dim=1, descending=True)
    return torch.gather(labels, 1, order), order


def evaluator_ndcg_at_ks(scores, labels, ks, presort: bool):
    """Per-query nDCG@ks as Evaluator.ndcg_at_ks builds it (base/ranker.py:67-95)."""
    sys_rankings, _ = rank_labels_by_scores(scores, labels)
    ideal = labels if presort else torch.sort(labels, dim=1, descending=True)[0]
    return ndcg_at_ks(sys_rankings, ideal, ks)


# --------------------------------------------------------------------------- #
# losses.  Each returns the 0-dim batch loss (sum over queries), built from
# autograd-tracked ATen ops exactly as the reference does, so .backward()
# yields the reference gradient.
# --------------------------------------------------------------------------- #
def _pairwise_probs(preds, labels, sigma):
    """ltr_adhoc/util/lambda_utils.py:5-23."""
    s_ij = preds.unsqueeze(2) - preds.unsqueeze(1)
    p_ij = torch.sigmoid(sigma * s_ij)
    S_ij = torch.clamp(labels.unsqueeze(2) - labels.unsqueeze(1), min=-1.0, max=1.0)
    return p_ij, 0.5 * (1.0 + S_ij)


def _delta_ndcg(ideal_rankings, predict_rankings):
    """metric/metric_utils.py:19-45."""
    idcg = dcg_at_k(ideal_rankings)
    ng = gains(predict_rankings) / idcg
    ng_diff = ng.unsqueeze(2) - ng.unsqueeze(1)
    ranks = torch.arange(predict_rankings.size(1), dtype=torch.float, device=predict_rankings.device)
    disc = (1.0 / torch.log2(ranks + 2.0)).unsqueeze(0)
    disc_diff = disc.unsqueeze(2) - disc.unsqueeze(1)
    return torch.abs(ng_diff) * torch.abs(disc_diff)


def ranknet_loss(preds, labels, sigma=1.0):
    """ltr_adhoc/pairwise/ranknet.py:25-36."""
    p_ij, std_p_ij = _pairwise_probs(preds, labels, sigma)
    cell = F.binary_cross_entropy(input=torch.triu(p_ij, diagonal=1),
                                  target=torch.triu(std_p_ij, diagonal=1), reduction="none")
    return cell.sum(dim=(2, 1)).sum()


def lambdarank_loss(preds, labels, sigma=1.0):
    """ltr_adhoc/listwise/lambdarank.py:27-56 (labels must arrive presorted, :36)."""
    desc_preds, order = torch.sort(preds, dim=1, descending=True)
    pred_rankings = torch.gather(labels, 1, order)
    p_ij, std_p_ij = _pairwise_probs(desc_preds, pred_rankings, sigma)
    delta = _delta_ndcg(labels, pred_rankings)
    cell = F.binary_cross_entropy(input=torch.triu(p_ij, diagonal=1),
                                  target=torch.triu(std_p_ij, diagonal=1),
                                  weight=torch.triu(delta, diagonal=1), reduction="none")
    return cell.sum(dim=(2, 1)).sum()


def _rank_gap_discount(n, disc):
    """delta_ij of lambdaloss.py:36-42: index-wrapped lookup, diagonal zeroed."""
    ranks = torch.arange(n, device=disc.device).float() + 1.0
    gap = torch.abs(ranks[:, None] - ranks[None, :]).long()
    d = torch.abs(torch.pow(disc[gap - 1], -1.0) - torch.pow(disc[gap], -1.0))
    d.diagonal().zero_()
    return d


def lambdaloss_loss(preds, labels, k=5, sigma=1.0, loss_type="NDCG_Loss2++", mu=5.0, presort=True):
    """ltr_adhoc/listwise/lambdaloss.py:73-132."""
    if presort:
        target, ideal = preds, labels
    else:
        ideal, ideal_idx = torch.sort(labels, dim=1, descending=True)
        target = torch.gather(preds, 1, ideal_idx)
    desc_preds, order = torch.sort(target, dim=1, descending=True)
    pred_rankings = torch.gather(ideal, 1, order)
    n = target.size(1)
    disc = 1.0 / torch.log2(torch.arange(n, dtype=torch.float, device=preds.device) + 2.0)
    n_gains = gains(pred_rankings) / dcg_at_k(ideal)
    if loss_type == "NDCG_Loss1":                       # :33-34 (valid for B == 1 only)
        w = n_gains / disc
    elif loss_type == "NDCG_Loss2":                     # :36-45
        w = _rank_gap_discount(n, disc)[None] * torch.abs(n_gains[:, :, None] - n_gains[:, None, :])
    elif loss_type == "NDCG_Loss2++":                   # :47-58
        rho = torch.abs(torch.pow(disc[:, None], -1.0) - torch.pow(disc[None, :], -1.0))
        w = (rho + mu * _rank_gap_discount(n, disc)) * torch.abs(n_gains[:, :, None] - n_gains[:, None, :])
    else:
        raise NotImplementedError(loss_type)
    diffs = (desc_preds.unsqueeze(2) - desc_preds.unsqueeze(1)).clamp(min=-1e8, max=1e8)
    diffs = torch.where(torch.isnan(diffs), torch.zeros_like(diffs), diffs)     # :116
    wp = (torch.sigmoid(sigma * diffs).clamp(min=EPS_LAMBDALOSS) ** w).clamp(min=EPS_LAMBDALOSS)
    log_wp = torch.log2(wp)
    trunc = torch.zeros((n, n), dtype=torch.bool, device=preds.device)
    trunc[:k, :k] = True
    if loss_type in ("NDCG_Loss2", "NDCG_Loss2++"):
        pair_mask = (pred_rankings.unsqueeze(2) - pred_rankings.unsqueeze(1)) > 0
        picked = log_wp[pair_mask & trunc]
    else:
        picked = log_wp[trunc[None, :, :].expand_as(log_wp)]
    return -picked.sum()


def listnet_loss(preds, labels):
    """ltr_adhoc/listwise/listnet.py:39."""
    return torch.sum(-torch.sum(F.softmax(labels, dim=1) * F.log_softmax(preds, dim=1), dim=1))


def shuffle_ties_perm(labels: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """ltr_adhoc/util/sampling_utils.py:13-28: indices ordering labels descending,
    ties broken by a fresh random permutation per row."""
    B, n = labels.shape
    perms = torch.stack([torch.randperm(n, generator=generator) for _ in range(B)], dim=0).to(labels.device)   # (sampling_utils.py:16-22: B host-side randperms, then moved to the device)
    shuffled = torch.gather(labels, 1, perms)
    desc = torch.argsort(shuffled, descending=True)
    return torch.gather(perms, 1, desc)


def listmle_loss(preds, labels=None, perm: Optional[torch.Tensor] = None):
    """ltr_adhoc/listwise/listmle.py:81-97.  ``perm`` injects the tie-shuffled
    ordering (the reference draws it from the global RNG every call, :81)."""
    if perm is None:
        perm = shuffle_ties_perm(labels)
    z = torch.gather(preds, 1, perm.long())
    m, _ = torch.max(z, dim=1, keepdim=True)
    y = torch.exp(z - m)
    tail_sums = torch.flip(torch.cumsum(torch.flip(y, dims=[1]), dim=1), dims=[1])
    return torch.sum(torch.sum(torch.log(tail_sums) + m - z, dim=1))


class _RobustSigmoid(torch.autograd.Function):
    """base/utils.py:57-92: branch on the sign of the *unscaled* input, custom backward."""

    @staticmethod
    def forward(ctx, inp, sigma):
        x = inp if 1.0 == sigma else sigma * inp
        half = torch.tensor([0.5], dtype=torch.float, device=inp.device)
        pos = torch.where(inp > 0, 1.0 / (1.0 + torch.exp(-x)), half)
        ex = torch.exp(x)
        out = torch.where(inp < 0, ex / (1.0 + ex), pos)
        g = out * (1.0 - out) if 1.0 == sigma else sigma * out * (1.0 - out)
        ctx.save_for_backward(g)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return grad_out * ctx.saved_tensors[0], None


def approx_ranks(preds, alpha):
    """ltr_adhoc/listwise/approxNDCG.py:19-28."""
    diffs = preds.unsqueeze(2) - preds.unsqueeze(1)
    ind = _RobustSigmoid.apply(torch.transpose(diffs, 1, 2), alpha)
    return torch.sum(ind, dim=2) + 0.5


def approxndcg_loss(preds, labels, alpha=10.0, presort=True):
    """ltr_adhoc/listwise/approxNDCG.py:45-62 and :93-101.  The [B]/[B,1] division
    broadcasts to [B,B] (:58-59), coupling the queries of a batch -- kept."""
    if presort:
        target, ideal = preds, labels
    else:
        ideal, idx = torch.sort(labels, dim=1, descending=True)
        target = torch.gather(preds, 1, idx)
    hat_pi = approx_ranks(target, alpha)
    idcg = dcg_at_k(ideal)
    dcg = torch.sum(gains(ideal) / torch.log2(hat_pi + 1), dim=1)
    return -torch.sum(dcg / idcg)


# --------------------------------------------------------------------------- #
# sibling losses
# --------------------------------------------------------------------------- #
def rankmse_loss(preds, labels):
    """ltr_adhoc/pointwise/rank_mse.py:13-22: MEAN over the batch of the per-query summed squared error."""
    return torch.mean(torch.sum(F.mse_loss(preds, labels, reduction='none'), dim=1))


def rankcosine_loss(preds, labels):
    """ltr_adhoc/listwise/rank_cosine.py:15,33."""
    return torch.sum((1.0 - F.cosine_similarity(preds, labels, dim=1)) / 0.5)


def stlistnet_loss(preds, labels, temperature=1.0, unif: Optional[torch.Tensor] = None):
    """ltr_adhoc/listwise/st_listnet.py:41-49.  ``unif`` injects the uniform draw (the reference calls torch.rand, :41)."""
    if unif is None:
        unif = torch.rand(preds.size())
    gumbel = -torch.log(-torch.log(unif + 1e-20) + 1e-20)
    z = (preds + gumbel) / temperature
    return torch.sum(-torch.sum(F.softmax(labels, dim=1) * F.log_softmax(z, dim=1), dim=1))


def softrank_loss(preds, labels, delta=2.0, top_k=None):
    """ltr_adhoc/listwise/softrank.py:46-72 (metric nDCG, labels presorted descending)."""
    delta_t = torch.tensor([delta])
    pairsub = torch.unsqueeze(preds, dim=2) - torch.unsqueeze(preds, dim=1)
    pairsub_vars = 2 * delta_t ** 2
    phi0 = 0.5 * torch.erfc(pairsub / torch.sqrt(2 * pairsub_vars))
    phi0_offdiag = torch.triu(phi0, diagonal=1) + torch.tril(phi0, diagonal=-1)
    expt_ranks = torch.sum(phi0_offdiag, dim=2) + 1.0
    g = torch.pow(2.0, labels) - 1.0
    dists = 1.0 / torch.log2(expt_ranks + 1.0)
    idcgs = dcg_at_k(labels)
    if top_k is None:
        dcgs = dists * g
    else:
        k = min(top_k, labels.size(1))
        dcgs = dists[:, 0:k] * g[:, 0:k]
    return -torch.sum(torch.sum(dcgs / idcgs, dim=1))


def sinkstep(dist, log_nu, log_u, lam: float):
    """ltr_adhoc/listwise/wassrank/pytorch_wasserstein.py:277-291 (the CPU form of the CUDA kernel at :132-224)."""
    log_v = log_nu.clone()
    for b in range(log_u.size(0)):
        log_v[b] -= torch.logsumexp(-dist / lam + log_u[b, :, None], 0)
    return log_v


def sinkhorn_ot(mu, nu, dist, lam=1e-3, N=100):
    """SinkhornOT.forward / backward, pytorch_wasserstein.py:294-324 -> (distances[B], d/dmu per unit grad, d/dnu)."""
    d1, d2 = dist.size()
    log_mu, log_nu = mu.log(), nu.log()
    log_u = torch.full_like(mu, -math.log(d1))
    log_v = torch.full_like(nu, -math.log(d2))
    for _ in range(N):
        log_v = sinkstep(dist, log_nu, log_u, lam)
        log_u = sinkstep(dist.t(), log_mu, log_v, lam)
    distances = (-sinkstep(-dist.log() + dist / lam, -log_v, log_u, 1.0)).logsumexp(1).exp()
    return distances, log_u * lam, log_v * lam


def per_query_standard_scale(feature_mat, clip_max=None):
    """ptranking/data/data_utils.py:482-487: sklearn StandardScaler().fit_transform on ONE query's [n,F] feature matrix
    (float64, as np.vstack of the parsed rows gives), ISTELLA clip first when asked (:484-485)."""
    import numpy as np
    from sklearn.preprocessing import StandardScaler
    x = np.asarray(feature_mat, dtype=np.float64)
    if clip_max is not None:
        x = np.clip(x, a_min=None, a_max=clip_max)
    return StandardScaler().fit_transform(x)


LOSSES = {
    "RankMSE": rankmse_loss,
    "RankCosine": rankcosine_loss,
    "STListNet": stlistnet_loss,
    "SoftRank": softrank_loss,
    "RankNet": ranknet_loss,
    "LambdaRank": lambdarank_loss,
    "LambdaLoss": lambdaloss_loss,
    "ListNet": listnet_loss,
    "ListMLE": listmle_loss,
    "ApproxNDCG": approxndcg_loss,
}


def loss_and_grad(name: str, scores: torch.Tensor, labels: torch.Tensor, **params):
    """(loss, dloss/dscores) on a leaf copy of ``scores`` -- what the parity tests compare."""
    s = scores.detach().clone().float().requires_grad_(True)
    if name == "ListMLE":
        loss = listmle_loss(s, labels, perm=params.get("perm"))
    elif name in ("ListNet", "RankMSE", "RankCosine"):
        loss = LOSSES[name](s, labels)
    else:
        loss = LOSSES[name](s, labels, **params)
    loss.backward()
    g = s.grad if s.grad is not None else torch.zeros_like(s)
    return loss.detach(), g.detach()


# --------------------------------------------------------------------------- #
# scorers (ptranking/base/utils.py, point_ranker.py, list_ranker.py)
# --------------------------------------------------------------------------- #
def make_activation(code: str) -> nn.Module:
    """base/utils.py:101-143 (working branches only)."""
    table = {"R": nn.ReLU, "LR": nn.LeakyReLU, "E": nn.ELU, "SE": nn.SELU, "CE": nn.CELU,
             "GE": nn.GELU, "S": nn.Sigmoid, "T": nn.Tanh}
    if code not in table:
        raise NotImplementedError(code)
    return table[code]()


class BatchNormAcrossQueries(nn.Module):
    """LTRBatchNorm, base/utils.py:201-223: statistics over all B*n rows, train and eval alike."""

    def __init__(self, width, affine):
        super().__init__()
        self.bn = nn.BatchNorm1d(width, momentum=0.1, affine=affine, track_running_stats=False)

    def forward(self, x):
        return self.bn(x.permute(0, 2, 1)).permute(0, 2, 1) if x.dim() == 3 else self.bn(x)


class BatchNormPerQuery(nn.Module):
    """LTRBatchNorm2 + ltr_batch_norm, base/utils.py:227-282 (grad-enabled branch: the
    reference never evaluates under no_grad, SURVEY B4, so moving stats are never read)."""

    def __init__(self, width, affine):
        super().__init__()
        shape = (1, 1, width)
        self.gamma = nn.Parameter(torch.ones(shape))
        self.beta = nn.Parameter(torch.zeros(shape))
        self.affine = affine
        if affine:
            self.weight = nn.Parameter(torch.ones(shape))
            self.bias = nn.Parameter(torch.zeros(shape))

    def forward(self, x):
        mean = x.mean(dim=1, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=1, keepdim=True)
        y = self.gamma * ((x - mean) / torch.sqrt(var + 1e-5)) + self.beta
        return y * self.weight + self.bias if self.affine else y


def stacked_ffnet(ff_dims, AF, TL_AF, apply_tl_af, dropout=0.1, BN=True, bn_type=None, bn_affine=False):
    """get_stacked_FFNet, base/utils.py:288-356 (same module names => same state_dict keys)."""
    def norm(width):
        if bn_type == "BN":
            return BatchNormAcrossQueries(width, bn_affine)
        if bn_type == "BN2":
            return BatchNormPerQuery(width, bn_affine)
        raise NotImplementedError(bn_type)

    net = nn.Sequential()
    L = len(ff_dims)
    for i in range(1, L - 1):
        net.add_module(f"dr_{i}", nn.Dropout(dropout))
        lin = nn.Linear(ff_dims[i - 1], ff_dims[i])
        nn.init.xavier_normal_(lin.weight)
        net.add_module(f"ff_{i + 1}", lin)
        if BN:
            net.add_module(f"bn_{i + 1}", norm(ff_dims[i]))
        net.add_module(f"act_{i + 1}", make_activation(AF))
    last = nn.Linear(ff_dims[-2], ff_dims[-1])
    nn.init.xavier_normal_(last.weight)
    net.add_module(f"ff_{L}", last)
    if apply_tl_af:
        if BN:
            net.add_module(f"bn_{L}", norm(ff_dims[-1]))
        net.add_module(f"act_{L}", make_activation(TL_AF))
    return net


def point_scorer(num_features, h_dim=100, out_dim=1, num_layers=3, AF="R", TL_AF="S", apply_tl_af=False,
                 BN=True, bn_type=None, bn_affine=False, dropout=0.1):
    """PointNeuralRanker.ini_pointsf, base/point_ranker.py:30-42."""
    dims = [num_features] + [h_dim] * num_layers + [out_dim]
    return stacked_ffnet(dims, AF, TL_AF, apply_tl_af, dropout, BN, bn_type, bn_affine)


def point_forward(net, X):
    """base/point_ranker.py:45-55."""
    return net(X).view(-1, X.size(1))


class RefLayerNorm(nn.Module):
    """base/list_ranker.py:152-174: unbiased std, eps added to the std."""

    def __init__(self, width, eps=1e-6):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(width))
        self.b_2 = nn.Parameter(torch.zeros(width))
        self.eps = eps

    def forward(self, x):
        mu = x.mean(-1, keepdim=True)
        sd = x.std(-1, keepdim=True)
        return self.a_2 * (x - mu) / (sd + self.eps) + self.b_2


class RefMHSA(nn.Module):
    """base/list_ranker.py:176-254."""

    def __init__(self, width, n_heads, dropout=0.1):
        super().__init__()
        assert width % n_heads == 0
        self.width, self.n_heads = width, n_heads
        self.w_q, self.w_k, self.w_v = nn.Linear(width, width), nn.Linear(width, width), nn.Linear(width, width)
        self.fc = nn.Linear(width, width)
        self.do_dropout = nn.Dropout(dropout)
        self.scale = math.sqrt(width // n_heads)

    def forward(self, x):
        B = x.shape[0]
        d = self.width // self.n_heads
        split = lambda t: t.view(B, -1, self.n_heads, d).permute(0, 2, 1, 3)
        Q, K, V = split(self.w_q(x)), split(self.w_k(x)), split(self.w_v(x))
        att = torch.matmul(Q, K.permute(0, 1, 3, 2)) / torch.sqrt(torch.tensor([float(d)], device=Q.device))
        att = self.do_dropout(torch.softmax(att, dim=-1))
        out = torch.matmul(att, V).permute(0, 2, 1, 3).contiguous().view(B, -1, self.width)
        return self.fc(out)


class RefEncoderLayer(nn.Module):
    """EncoderLayer + SublayerConnection, base/list_ranker.py:87-149."""

    def __init__(self, width, n_heads, encoder_type, dropout):
        super().__init__()
        self.encoder_type = encoder_type
        self.mhsa = RefMHSA(width, n_heads, dropout)
        if encoder_type == "AllRank":
            self.norm0, self.norm1 = RefLayerNorm(width), RefLayerNorm(width)
            self.w1, self.w2 = nn.Linear(width, width), nn.Linear(width, width)
            self.drop0, self.drop1, self.drop_ff = nn.Dropout(dropout), nn.Dropout(dropout), nn.Dropout(dropout)
        else:
            self.norm = RefLayerNorm(width)

    def forward(self, x):
        if self.encoder_type == "AllRank":
            x = x + self.drop0(self.mhsa(self.norm0(x)))
            return x + self.drop1(self.w2(self.drop_ff(F.relu(self.w1(self.norm1(x))))))
        if self.encoder_type == "DASALC":
            return self.norm(self.mhsa(x))
        if self.encoder_type == "AttnDIN":
            return self.norm(x + self.mhsa(x))
        raise NotImplementedError(self.encoder_type)


class RefListScorer(nn.Module):
    """ListNeuralRanker.ini_listsf + forward, base/list_ranker.py:303-378."""

    def __init__(self, num_features, ff_dims=(128, 256, 512), out_dim=1, AF="R", TL_AF="GE", apply_tl_af=False,
                 BN=True, bn_type=None, bn_affine=False, n_heads=2, encoder_layers=3, dropout=0.1,
                 encoder_type="DASALC"):
        super().__init__()
        F_ = num_features
        self.encoder_type = encoder_type
        self.head = stacked_ffnet([F_, *ff_dims, F_], AF, AF, True, dropout, BN, bn_type, bn_affine)   # :309-314
        self.layers = nn.ModuleList([RefEncoderLayer(F_, n_heads, encoder_type, dropout)
                                     for _ in range(encoder_layers)])
        self.final_norm = RefLayerNorm(F_) if encoder_type == "AllRank" else None                      # :66-67
        self.tail = stacked_ffnet([F_, *ff_dims, out_dim], AF, TL_AF, apply_tl_af, 0.1, BN, bn_type, bn_affine)  # :337-341

    def encode(self, x):
        for layer in self.layers:
            x = layer(x)
        return self.final_norm(x) if self.final_norm is not None else x

    def forward(self, X):
        head = self.head(X)
        if self.encoder_type == "AllRank":
            z = self.encode(head)
        elif self.encoder_type == "DASALC":
            z = (self.encode(X) + 1.0) * head
        elif self.encoder_type == "AttnDIN":
            z = self.encode(head) + X
        else:
            raise NotImplementedError(self.encoder_type)
        return torch.squeeze(self.tail(z), dim=2)


def make_optimizer(params, opt="Adam", lr=1e-4, weight_decay=1e-3):
    """NeuralRanker.config_optimizer, base/ranker.py:512-525."""
    params = list(params)
    if opt == "Adam":
        o = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)
    elif opt == "RMS":
        o = torch.optim.RMSprop(params, lr=lr, weight_decay=weight_decay)
    elif opt == "Adagrad":
        o = torch.optim.Adagrad(params, lr=lr, weight_decay=weight_decay)
    else:
        raise NotImplementedError(opt)
    return o, torch.optim.lr_scheduler.StepLR(o, step_size=20, gamma=0.5)


def train_op(net, optimizer, loss_name, X, labels, point=True, **loss_params):
    """NeuralRanker.train_op + the loss class's zero_grad/backward/step tail
    (base/ranker.py:589-603, e.g. lambdarank.py:58-60)."""
    preds = point_forward(net, X) if point else net(X)
    if loss_name == "ListMLE":
        loss = listmle_loss(preds, labels, perm=loss_params.get("perm"))
    elif loss_name == "ListNet":
        loss = listnet_loss(preds, labels)
    else:
        loss = LOSSES[loss_name](preds, labels, **loss_params)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.detach()


# --------------------------------------------------------------------------- #
# P / AP / nERR (ptranking/metric/adhoc/adhoc_metric.py:18-193) -- SURVEY 8f "next" row 1
# --------------------------------------------------------------------------- #
def _used_ks(ks, n):
    return ([k for k in ks if k <= n], True) if n < max(ks) else (list(ks), False)


def _pad(vals, B, ks, used, padded):
    if not padded:
        return vals
    out = torch.zeros(B, len(ks))
    out[:, : len(used)] = vals
    return out


def precision_at_ks(sys_rankings, ks):
    """adhoc_metric.py:36-64."""
    used, padded = _used_ks(ks, sys_rankings.size(1))
    m = max(used)
    idx = torch.tensor(used, dtype=torch.long) - 1
    bi = torch.clamp(sys_rankings[:, :m], min=0, max=1)
    prec = torch.cumsum(bi, dim=1) / (torch.arange(m, dtype=torch.float) + 1.0)
    return _pad(prec[:, idx], sys_rankings.size(0), ks, used, padded)


def ap_at_ks(sys_rankings, ideal_rankings, ks):
    """adhoc_metric.py:95-128 -- the denominator sums the NON-binarised ideal labels."""
    used, padded = _used_ks(ks, sys_rankings.size(1))
    m = max(used)
    idx = torch.tensor(used, dtype=torch.long) - 1
    bi = torch.clamp(sys_rankings[:, :m], min=0, max=1)
    prec = torch.cumsum(bi, dim=1) / (torch.arange(m, dtype=torch.float) + 1.0)
    cum_prec = torch.cumsum(prec * bi, dim=1)
    ap = cum_prec / torch.cumsum(ideal_rankings, dim=1)[:, :m]
    return _pad(ap[:, idx], sys_rankings.size(0), ks, used, padded)


def _rankwise_err(rankings, max_label, k):
    """adhoc_metric.py:132-156 (point=False)."""
    labels = rankings[:, :k]
    satis = (torch.pow(torch.tensor([2.0]), labels) - 1.0) / torch.pow(torch.tensor([2.0]), max_label)
    cum_unsatis = torch.cumprod(1.0 - satis, dim=1)
    cascade = torch.ones_like(satis)
    cascade[:, 1:k] = cum_unsatis[:, : k - 1]
    expt_ranks = 1.0 / (torch.arange(k, dtype=torch.float) + 1.0)
    return torch.cumsum(expt_ranks * satis * cascade, dim=1)


def nerr_at_ks(sys_rankings, ideal_rankings, ks, max_label=None):
    """adhoc_metric.py:171-193; max_label defaults to the maximum over the whole batch."""
    used, padded = _used_ks(ks, sys_rankings.size(1))
    if max_label is None:
        max_label = torch.max(ideal_rankings)
    m = max(used)
    idx = torch.tensor(used, dtype=torch.long) - 1
    out = (_rankwise_err(sys_rankings, max_label, m) / _rankwise_err(ideal_rankings, max_label, m))[:, idx]
    return _pad(out, sys_rankings.size(0), ks, used, padded)


def evaluator_metrics_at_ks(scores, labels, ks, presort, max_label=None):
    """(nDCG, nERR, AP, P) per query as adhoc_performance_at_ks builds them (base/ranker.py:202-263)."""
    sys_r, _ = rank_labels_by_scores(scores, labels)
    ideal = labels if presort else torch.sort(labels, dim=1, descending=True)[0]
    return (ndcg_at_ks(sys_r, ideal, ks), nerr_at_ks(sys_r, ideal, ks, max_label), ap_at_ks(sys_r, ideal, ks),
            precision_at_ks(sys_r, ks))
