"""GPU parity of the six fused loss kernels, called through the C ABI (ops -> libptranking_b200.so):
 - against outputs of the unmodified reference (tests/golden/losses.npz),
 - against the oracle restatement / float64 closed forms on fresh seeded inputs,
 - size-independent properties at BASELINE.json's full list lengths.
Tolerance (north_star): loss and gradient within 1e-5 relative, fp32."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_port as rp
from tests.helpers import loss_cases, parse_loss_key, rel_err

pytestmark = pytest.mark.gpu
CASES = loss_cases()
TOL = 1e-5


def _ops():
    from ptranking_b200 import ops
    return ops


def _run(name, s, y, **kw):
    ops = _ops()
    dev = torch.device("cuda", 0)
    st = torch.from_numpy(np.ascontiguousarray(s)).to(dev)
    yt = torch.from_numpy(np.ascontiguousarray(y)).to(dev)
    if "perm" in kw and kw["perm"] is not None:
        kw["perm"] = torch.from_numpy(np.ascontiguousarray(kw["perm"]).astype(np.int32)).to(dev)
    loss, loss_q, grad = ops.rank_loss_and_grad(name, st, yt, **kw)
    torch.cuda.synchronize()
    return float(loss.cpu()), loss_q.cpu().numpy(), grad.cpu().numpy()


@pytest.mark.parametrize("head,case,d", CASES, ids=[f"{h}-{c}" for h, c, _ in CASES])
def test_kernel_matches_reference_fixture(head, case, d):
    name, params, presort = parse_loss_key(head)
    kw = dict(params)
    if name in ("LambdaLoss", "ApproxNDCG"):
        kw["presort"] = presort
    if name == "ListMLE":
        kw["perm"] = d["perm"]
    loss, _, grad = _run(name, d["scores"], d["labels"], **kw)
    ref_loss, ref_grad = float(d["loss"]), d["grad"]
    # referee: where the fp32 reference itself is further than TOL from the float64 closed form,
    # the kernel only has to be as close to float64 as the reference is (its own rounding noise)
    tol_l, tol_g = TOL, TOL
    if "saturated" not in head:
        fl, fg = _closed(name, d, params, presort)
        tol_l = max(TOL, 2.0 * abs(ref_loss - fl) / max(abs(fl), 1.0))
        tol_g = max(TOL, 2.0 * rel_err(ref_grad, fg))
    else:
        tol_l, tol_g = 1e-4, 1e-4      # |sigma*ds| > 17: one fp32 ulp of exp() moves p across the clamp
    assert abs(loss - ref_loss) <= tol_l * max(abs(ref_loss), 1.0), (loss, ref_loss)
    assert rel_err(grad, ref_grad) <= tol_g


def _closed(name, d, params, presort):
    s, y = d["scores"], d["labels"]
    if name == "RankNet":
        return cf.ranknet(s, y, **params)
    if name == "LambdaRank":
        return cf.lambdarank(s, y, **params)
    if name == "LambdaLoss":
        return cf.lambdaloss(s, y, presort=presort, **params)
    if name == "ListNet":
        return cf.listnet(s, y)
    if name == "ListMLE":
        return cf.listmle(s, d["perm"])
    return cf.approxndcg(s, y, presort=presort, **params)


def _synth(B, n, seed, scale=1.0, sigmoid=False):
    rng = np.random.default_rng(seed)
    p = np.array([1940952, 1225770, 504958, 69010, 30435], dtype=np.float64); p /= p.sum()
    y = rng.choice(5, size=(B, n), p=p).astype(np.float32)
    y[:, 0] = np.maximum(y[:, 0], 1.0)
    y = -np.sort(-y, axis=1)
    s = rng.standard_normal((B, n)) * scale
    if sigmoid:
        s = 1.0 / (1.0 + np.exp(-s))
    return s.astype(np.float32), y


FRESH = [("RankNet", dict(sigma=1.0)), ("LambdaRank", dict(sigma=1.0)),
         ("LambdaLoss", dict(k=5, sigma=1.0, mu=5.0, loss_type="NDCG_Loss2++", presort=True)),
         ("LambdaLoss", dict(k=64, sigma=1.0, mu=5.0, loss_type="NDCG_Loss2", presort=True)),
         ("ListNet", {}), ("ListMLE", {}), ("ApproxNDCG", dict(alpha=10.0, presort=True))]


@pytest.mark.parametrize("name,params", FRESH, ids=[f"{n}-{i}" for i, (n, _) in enumerate(FRESH)])
@pytest.mark.parametrize("shape", [(8, 256), (4, 50), (2, 1024), (3, 33), (5, 1), (2, 2)])
def test_kernel_matches_oracle_fresh_inputs(name, params, shape):
    B, n = shape
    s, y = _synth(B, n, seed=1000 + B * 7 + n, sigmoid=(n % 2 == 0))
    kw = dict(params)
    if name == "ListMLE":
        kw["perm"] = rp.shuffle_ties_perm(torch.from_numpy(y), generator=torch.Generator().manual_seed(n)).numpy()
    loss, loss_q, grad = _run(name, s, y, **kw)
    okw = {k: (torch.from_numpy(v.astype(np.int64)) if k == "perm" else v) for k, v in kw.items()}
    o_loss, o_grad = rp.loss_and_grad(name, torch.from_numpy(s), torch.from_numpy(y), **okw)
    o_loss, o_grad = float(o_loss), o_grad.numpy()
    d = dict(scores=s, labels=y, perm=kw.get("perm"))
    cparams = {k: v for k, v in params.items() if k != "presort"}
    fl, fg = _closed(name, d, cparams, True)
    tol_l = max(TOL, 2.0 * abs(o_loss - fl) / max(abs(fl), 1.0))
    tol_g = max(TOL, 2.0 * rel_err(o_grad, fg))
    assert abs(loss - o_loss) <= tol_l * max(abs(o_loss), 1.0), (loss, o_loss, fl)
    assert rel_err(grad, o_grad) <= tol_g, (rel_err(grad, o_grad), rel_err(grad, fg), rel_err(o_grad, fg))
    # and the kernel itself is within TOL of float64 truth
    assert rel_err(grad, fg) <= 5 * TOL and abs(loss - fl) <= 5 * TOL * max(abs(fl), 1.0)
    assert abs(float(loss_q.sum()) - loss) <= 1e-5 * max(abs(loss), 1.0)


def test_approxndcg_per_query_mode_decouples_batch():
    s, y = _synth(4, 64, seed=5)
    _, lq_c, g_c = _run("ApproxNDCG", s, y, alpha=10.0, presort=True, batch_coupled=True)
    _, lq_u, g_u = _run("ApproxNDCG", s, y, alpha=10.0, presort=True, batch_coupled=False)
    fl, fg = cf.approxndcg(s, y, alpha=10.0, batch_coupled=False)
    assert abs(lq_u.sum() - fl) <= 2e-5 * abs(fl) and rel_err(g_u, fg) <= 5e-5
    for b in range(4):   # per-query mode == running the query alone
        _, lq1, g1 = _run("ApproxNDCG", s[b:b + 1], y[b:b + 1], alpha=10.0, presort=True, batch_coupled=True)
        assert abs(lq1[0] - lq_u[b]) <= 1e-6 * abs(lq1[0]) and rel_err(g_u[b], g1[0]) <= 1e-6
    assert abs(lq_c.sum() / lq_u.sum()) > 1.5      # the coupled loss is scaled by sum_a 1/iDCG_a


def test_shuffle_ties_perm_is_valid_and_random():
    ops = _ops()
    s, y = _synth(16, 200, seed=9)
    yt = torch.from_numpy(y).cuda()
    p1 = ops.shuffle_ties_perm(yt, seed=1, offset=1).cpu().numpy()
    p2 = ops.shuffle_ties_perm(yt, seed=1, offset=2).cpu().numpy()
    p1b = ops.shuffle_ties_perm(yt, seed=1, offset=1).cpu().numpy()
    assert np.array_equal(p1, p1b) and not np.array_equal(p1, p2)
    for p in (p1, p2):
        assert np.array_equal(np.sort(p, axis=1), np.tile(np.arange(200), (16, 1)))
        ordered = np.take_along_axis(y, p.astype(np.int64), 1)
        assert np.all(np.diff(ordered, axis=1) <= 0)       # labels descending
    # unsorted labels too
    yu = np.random.default_rng(3).permuted(y, axis=1)
    pu = ops.shuffle_ties_perm(torch.from_numpy(yu).cuda(), seed=2, offset=1).cpu().numpy()
    assert np.all(np.diff(np.take_along_axis(yu, pu.astype(np.int64), 1), axis=1) <= 0)


@pytest.mark.parametrize("name,params", [("RankNet", dict(sigma=1.0)), ("LambdaRank", dict(sigma=1.0)),
                                         ("LambdaLoss", dict(k=5, loss_type="NDCG_Loss2++")),
                                         ("ListNet", {}), ("ApproxNDCG", dict(alpha=10.0))])
def test_full_size_properties(name, params):
    """BASELINE.json sizes (B=1024 lists of 256 / 16 of 1024): properties that need no oracle."""
    for (B, n) in [(1024, 256), (16, 1024)]:
        s, y = _synth(B, n, seed=B + n, sigmoid=True)
        loss, lq, g = _run(name, s, y, **params)
        loss2, lq2, g2 = _run(name, s, y, **params)
        assert np.array_equal(g, g2) and np.array_equal(lq, lq2)           # deterministic, bit for bit
        assert np.isfinite(g).all() and np.isfinite(lq).all()
        # translation invariance of every loss => gradients of a query sum to zero
        assert np.abs(g.sum(1)).max() <= 2e-4 * max(np.abs(g).max(), 1e-12) * np.sqrt(n)
        # permutation equivariance: shuffling the documents of a query permutes its gradient
        if name in ("RankNet", "ListNet"):      # the others require presorted labels
            rng = np.random.default_rng(1)
            perm = rng.permutation(n)
            _, lqp, gp = _run(name, s[:, perm], y[:, perm], **params)
            if name == "ListNet":
                assert np.allclose(lqp, lq, rtol=1e-5) and rel_err(gp, g[:, perm]) <= 1e-5
        # queries are independent: a sub-batch reproduces its rows exactly
        if name != "ApproxNDCG":
            _, lqs, gs = _run(name, s[:3], y[:3], **params)
            assert np.array_equal(gs, g[:3]) and np.array_equal(lqs, lq[:3])


def test_listmle_full_size_properties():
    B, n = 512, 1024
    s, y = _synth(B, n, seed=77)
    ops = _ops()
    perm = ops.shuffle_ties_perm(torch.from_numpy(y).cuda(), seed=5, offset=3).cpu().numpy()
    loss, lq, g = _run("ListMLE", s, y, perm=perm)
    assert np.isfinite(g).all() and np.abs(g.sum(1)).max() <= 1e-3
    fl, fg = cf.listmle(s[:4], perm[:4])
    assert rel_err(g[:4], fg) <= 5e-5


def test_error_codes():
    from ptranking_b200 import _lib
    ops = _ops()
    s = torch.zeros(1, 5000, device="cuda")
    with pytest.raises(_lib.B200LibraryError, match="PTRB200_MAX_LIST_LEN"):
        ops.rank_loss_and_grad("LambdaRank", s, s)
    with pytest.raises(ValueError):
        ops.rank_loss_and_grad("LambdaRank", torch.zeros(2, 4, device="cuda"), torch.zeros(2, 5, device="cuda"))
