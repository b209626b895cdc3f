"""Golden fixtures for the sibling losses, produced by the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_siblings.py

RankMSE / RankCosine / STListNet / SoftRank: the reference classes' own custom_loss_function on a leaf score tensor
(a throw-away SGD(lr=0) stands in for the optimizer they step).  STListNet's torch.rand draw is captured by wrapping
torch.rand while the reference runs.  sinkstep / SinkhornOT: pytorch_wasserstein.py's CPU path.  Output: siblings.npz.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
REF = os.environ.get("PTRANKING_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from make_golden import synth_labels, point_sf_dict, MSLR_P, MQ_P, ML  # noqa: E402
from ptranking.ltr_adhoc.pointwise.rank_mse import RankMSE  # noqa: E402
from ptranking.ltr_adhoc.listwise.rank_cosine import RankCosine  # noqa: E402
from ptranking.ltr_adhoc.listwise.st_listnet import STListNet  # noqa: E402
from ptranking.ltr_adhoc.listwise.softrank import SoftRank  # noqa: E402
from ptranking.ltr_adhoc.listwise.wassrank.pytorch_wasserstein import sinkstep, SinkhornOT  # noqa: E402


def run(cls, model_para, s, y):
    sf = point_sf_dict(4)
    r = cls(sf_para_dict=sf, gpu=False, device="cpu") if model_para is None else \
        cls(sf_para_dict=sf, model_para_dict=model_para, gpu=False, device="cpu")
    leaf = torch.from_numpy(s).clone().requires_grad_(True)
    r.optimizer = torch.optim.SGD([leaf], lr=0.0)
    loss = r.custom_loss_function(leaf, torch.from_numpy(y), presort=True, label_type=ML)
    return float(loss.detach()), leaf.grad.detach().numpy().copy()


def main():
    rng = np.random.default_rng(4137)
    torch.manual_seed(4137)
    out = {}

    def put(name, case, s, y, loss, grad, **extra):
        key = f"{name}__{case}"
        out[key + "__scores"], out[key + "__labels"] = s, y
        out[key + "__loss"], out[key + "__grad"] = np.float64(loss), grad
        for k, v in extra.items():
            out[key + "__" + k] = v

    cases = []
    for (B, n) in [(1, 8), (3, 50), (2, 256), (1, 1024), (4, 37)]:
        y = synth_labels(rng, B, n, MQ_P if n == 50 else MSLR_P)
        for tag, s in (("sig", 1.0 / (1.0 + np.exp(-rng.standard_normal((B, n))))), ("wide", 2.5 * rng.standard_normal((B, n)))):
            cases.append((f"B{B}_n{n}_{tag}", s.astype(np.float32), y))
    for case, s, y in cases:
        put("RankMSE", case, s, y, *run(RankMSE, None, s, y))
        put("RankCosine", case, s, y, *run(RankCosine, None, s, y))
        for T in (1.0, 0.5):
            captured = {}
            real_rand = torch.rand

            def spy(*a, **k):
                u = real_rand(*a, **k)
                captured["u"] = u.detach().clone()
                return u
            torch.rand = spy
            try:
                l, g = run(STListNet, dict(model_id="STListNet", temperature=T), s, y)
            finally:
                torch.rand = real_rand
            put(f"STListNet_T{T}", case, s, y, l, g, unif=captured["u"].numpy())
        for delta, top_k in ((2.0, None), (0.5, None), (2.0, 5)):
            l, g = run(SoftRank, dict(model_id="SoftRank", delta=delta, metric="nDCG", top_k=top_k), s, y)
            put(f"SoftRank_delta{delta}_k{top_k}", case, s, y, l, g)
    # a zero-score row: RankCosine's eps branch
    s0 = np.zeros((2, 6), dtype=np.float32); s0[1] = rng.standard_normal(6)
    y0 = synth_labels(rng, 2, 6, MSLR_P)
    put("RankCosine", "B2_n6_zero", s0, y0, *run(RankCosine, None, s0, y0))

    # Sinkhorn half-step and the full SinkhornOT (forward distances + the gradients its backward returns)
    for (B, d1, d2, lam) in [(1, 8, 8, 0.1), (3, 50, 50, 0.1), (2, 33, 70, 1.0), (2, 256, 256, 0.05)]:
        dist = torch.from_numpy(rng.random((d1, d2)).astype(np.float32) + 0.01)
        log_nu = torch.log_softmax(torch.from_numpy(rng.standard_normal((B, d2)).astype(np.float32)), dim=1)
        log_u = torch.from_numpy(rng.standard_normal((B, d1)).astype(np.float32))
        if d1 == 33:
            log_nu[0, 3] = -float("inf")
            log_u[1, :] = -float("inf")
        log_v = sinkstep(dist, log_nu, log_u, lam)
        key = f"sinkstep__B{B}_{d1}x{d2}_lam{lam}"
        out[key + "__dist"], out[key + "__log_nu"], out[key + "__log_u"] = dist.numpy(), log_nu.numpy(), log_u.numpy()
        out[key + "__lam"], out[key + "__log_v"] = np.float64(lam), log_v.numpy()
    for (B, d, lam, N) in [(2, 16, 0.1, 20), (3, 50, 0.05, 20)]:
        dist = torch.from_numpy(np.abs(np.subtract.outer(np.arange(d), np.arange(d))).astype(np.float32) / d + 0.01)
        mu = torch.softmax(torch.from_numpy(rng.standard_normal((B, d)).astype(np.float32)), dim=1).requires_grad_(True)
        nu = torch.softmax(torch.from_numpy(rng.standard_normal((B, d)).astype(np.float32)), dim=1).requires_grad_(True)
        dists = SinkhornOT.apply(mu, nu, dist, lam, N)
        dists.sum().backward()
        key = f"sinkhorn__B{B}_d{d}_lam{lam}_N{N}"
        out[key + "__dist"], out[key + "__mu"], out[key + "__nu"] = dist.numpy(), mu.detach().numpy(), nu.detach().numpy()
        out[key + "__lam"], out[key + "__N"] = np.float64(lam), np.int64(N)
        out[key + "__distances"], out[key + "__dmu"], out[key + "__dnu"] = dists.detach().numpy(), mu.grad.numpy(), nu.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "siblings.npz"), **out)
    print(f"siblings.npz: {len(out)} arrays")


if __name__ == "__main__":
    main()
