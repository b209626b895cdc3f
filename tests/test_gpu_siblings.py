"""GPU parity of the sibling-loss kernels (RankMSE, RankCosine, STListNet, SoftRank) and the Sinkhorn
half-step, through the C ABI: against outputs of the unmodified reference (tests/golden/siblings.npz), against the
oracle on fresh inputs, size-independent properties at full list lengths, and the drop-in classes' train step."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_port as rp
from tests.helpers import parse_sibling_key, rel_err, sibling_cases, sinkhorn_cases
from tests.test_gpu_losses import _run, _synth

pytestmark = pytest.mark.gpu
CASES = sibling_cases()
TOL = 1e-5
DEV = "cuda:0"


def _closed(name, s, y, params, unif=None):
    if name == "RankMSE":
        return cf.rankmse(s, y)
    if name == "RankCosine":
        return cf.rankcosine(s, y)
    if name == "STListNet":
        return cf.stlistnet(s, y, unif, **params)
    return cf.softrank(s, y, **params)


@pytest.mark.parametrize("head,case,d", CASES, ids=[f"{h}-{c}" for h, c, _ in CASES])
def test_kernel_matches_reference_fixture(head, case, d):
    name, params = parse_sibling_key(head)
    kw = dict(params)
    if name == "STListNet":
        kw["unif"] = torch.from_numpy(d["unif"]).to(DEV)
    loss, lq, grad = _run(name, d["scores"], d["labels"], **kw)
    ref_loss, ref_grad = float(d["loss"]), d["grad"]
    fl, fg = _closed(name, d["scores"], d["labels"], params, d.get("unif"))
    # where the fp32 reference is itself further than TOL from float64, the kernel must be as close to float64 as it is
    tol_l = max(TOL, 2.0 * abs(ref_loss - fl) / max(abs(fl), 1.0))
    tol_g = max(TOL, 2.0 * rel_err(ref_grad, fg))
    assert abs(loss - ref_loss) <= tol_l * max(abs(ref_loss), 1.0), (loss, ref_loss, fl)
    assert rel_err(grad, ref_grad) <= tol_g, (rel_err(grad, ref_grad), rel_err(grad, fg))
    assert abs(float(lq.sum()) - loss) <= 1e-5 * max(abs(loss), 1.0)


FRESH = [("RankMSE", {}), ("RankCosine", {}), ("STListNet", dict(temperature=1.0)), ("STListNet", dict(temperature=0.3)),
         ("SoftRank", dict(delta=2.0, top_k=None)), ("SoftRank", dict(delta=0.7, top_k=10))]


@pytest.mark.parametrize("name,params", FRESH, ids=[f"{n}-{i}" for i, (n, _) in enumerate(FRESH)])
@pytest.mark.parametrize("shape", [(8, 256), (4, 50), (2, 1024), (3, 33), (5, 1), (2, 2), (1, 2000)])
def test_kernel_matches_oracle_fresh_inputs(name, params, shape):
    B, n = shape
    s, y = _synth(B, n, seed=2000 + B * 7 + n, sigmoid=(n % 2 == 0))
    kw = dict(params)
    unif = None
    if name == "STListNet":
        unif = np.random.default_rng(n).random((B, n), dtype=np.float32)
        kw["unif"] = torch.from_numpy(unif).to(DEV)
    loss, lq, grad = _run(name, s, y, **kw)
    okw = dict(params)
    if unif is not None:
        okw["unif"] = torch.from_numpy(unif)
    o_loss, o_grad = rp.loss_and_grad(name, torch.from_numpy(s), torch.from_numpy(y), **okw)
    o_loss, o_grad = float(o_loss), o_grad.numpy()
    fl, fg = _closed(name, s, y, params, unif)
    tol_l = max(TOL, 2.0 * abs(o_loss - fl) / max(abs(fl), 1.0))
    tol_g = max(TOL, 2.0 * rel_err(o_grad, fg))
    assert abs(loss - o_loss) <= tol_l * max(abs(o_loss), 1.0), (loss, o_loss, fl)
    # (a gradient that is identically zero in exact arithmetic -- RankCosine on one-document lists -- is compared absolutely)
    floor = 1e-6 if np.abs(o_grad).max() < 1e-6 else 0.0
    assert rel_err(grad, o_grad) <= tol_g or np.abs(grad - o_grad).max() <= floor, (rel_err(grad, o_grad), rel_err(grad, fg), rel_err(o_grad, fg))
    assert (rel_err(grad, fg) <= 5 * TOL or np.abs(grad - fg).max() <= floor) and abs(loss - fl) <= 5 * TOL * max(abs(fl), 1.0)


def test_stlistnet_device_noise_is_keyed_and_gumbel():
    """Without an injected draw the kernel generates its own uniforms: reproducible per (seed, offset), different across
    offsets, and the implied noise has the Gumbel mean (Euler-Mascheroni) -- recovered from the gradient identity
    grad * T + softmax(y) = softmax((s + g) / T) on a constant-score list."""
    B, n = 64, 1024
    s = np.zeros((B, n), dtype=np.float32)
    y = np.zeros((B, n), dtype=np.float32)
    _, _, g1 = _run("STListNet", s, y, temperature=1.0, seed=7, offset=1)
    _, _, g1b = _run("STListNet", s, y, temperature=1.0, seed=7, offset=1)
    _, _, g2 = _run("STListNet", s, y, temperature=1.0, seed=7, offset=2)
    assert np.array_equal(g1, g1b) and not np.array_equal(g1, g2)
    p = g1.astype(np.float64) + 1.0 / n                      # softmax(gumbel) per row
    gum = np.log(p) - np.log(p).mean(1, keepdims=True)       # gumbel noise up to a per-row constant
    assert abs(np.var(gum) - np.pi ** 2 / 6) < 0.05          # Var[Gumbel(0,1)] = pi^2/6
    assert np.abs(g1.sum(1)).max() < 1e-5


@pytest.mark.parametrize("name,params", [("RankMSE", {}), ("RankCosine", {}), ("SoftRank", dict(delta=2.0, top_k=None)),
                                         ("STListNet", dict(temperature=1.0, seed=3, offset=9))])
def test_full_size_properties(name, params):
    for (B, n) in [(1024, 256), (16, 1024)]:
        s, y = _synth(B, n, seed=B + n, sigmoid=True)
        loss, lq, g = _run(name, s, y, **params)
        _, lq2, g2 = _run(name, s, y, **params)
        assert np.array_equal(g, g2) and np.array_equal(lq, lq2)           # deterministic, bit for bit
        assert np.isfinite(g).all() and np.isfinite(lq).all()
        if name in ("SoftRank", "STListNet"):      # translation-invariant losses: per-query gradients sum to zero
            assert np.abs(g.sum(1)).max() <= 2e-4 * max(np.abs(g).max(), 1e-12) * np.sqrt(n)
        if name == "RankCosine":                   # scale invariance: the gradient is orthogonal to the scores
            assert np.abs((g * s).sum(1)).max() <= 1e-5
        if name != "RankMSE":                      # RankMSE carries 1/B; the others are independent per query
            _, lqs, gs = _run(name, s[:3], y[:3], **params) if name != "STListNet" else (None, None, None)
            if gs is not None:
                assert np.array_equal(gs, g[:3]) and np.array_equal(lqs, lq[:3])
        else:
            assert rel_err(g, 2.0 * (s - y) / B) <= 1e-6


@pytest.mark.parametrize("case,d", sinkhorn_cases("sinkstep"), ids=[c for c, _ in sinkhorn_cases("sinkstep")])
def test_sinkstep_matches_reference(case, d):
    from ptranking_b200 import ops
    got = ops.sinkstep(torch.from_numpy(d["dist"]).to(DEV), torch.from_numpy(d["log_nu"]).to(DEV),
                       torch.from_numpy(d["log_u"]).to(DEV), float(d["lam"])).cpu().numpy()
    want = d["log_v"]
    inf = np.isinf(want)
    assert np.array_equal(np.isinf(got), inf) and np.array_equal(got[inf], want[inf])
    assert np.abs(got[~inf] - want[~inf]).max() <= 1e-5 * max(np.abs(want[~inf]).max(), 1.0)


@pytest.mark.parametrize("case,d", sinkhorn_cases("sinkhorn"), ids=[c for c, _ in sinkhorn_cases("sinkhorn")])
def test_sinkhorn_ot_matches_reference(case, d):
    from ptranking_b200 import ops
    mu = torch.from_numpy(d["mu"]).to(DEV).requires_grad_(True)
    nu = torch.from_numpy(d["nu"]).to(DEV).requires_grad_(True)
    dist = ops.SinkhornOT.apply(mu, nu, torch.from_numpy(d["dist"]).to(DEV), float(d["lam"]), int(d["N"]))
    dist.sum().backward()
    assert rel_err(dist.detach().cpu().numpy(), d["distances"]) <= 1e-4
    assert rel_err(mu.grad.cpu().numpy(), d["dmu"]) <= 1e-4 and rel_err(nu.grad.cpu().numpy(), d["dnu"]) <= 1e-4


def test_sinkstep_large():
    """d = 1024 (one histogram bin per document of a 1024-doc list), B = 64: against float64."""
    from ptranking_b200 import ops
    rng = np.random.default_rng(0)
    B, d1, d2 = 64, 1024, 1000
    dist = rng.random((d1, d2)).astype(np.float32)
    log_nu = np.log(rng.dirichlet(np.ones(d2), B)).astype(np.float32)
    log_u = rng.standard_normal((B, d1)).astype(np.float32)
    got = ops.sinkstep(torch.from_numpy(dist).to(DEV), torch.from_numpy(log_nu).to(DEV), torch.from_numpy(log_u).to(DEV), 0.1).cpu().numpy()
    M = -dist.astype(np.float64)[None] / 0.1 + log_u.astype(np.float64)[:, :, None]
    mx = M.max(1, keepdims=True)
    want = log_nu - (np.log(np.exp(M - mx).sum(1)) + mx[:, 0])
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("model", ["RankMSE", "RankCosine", "STListNet", "SoftRank"])
def test_drop_in_class_train_step(model):
    """The drop-in class runs a training step (scorer forward, fused loss kernel, scorer backward, optimizer) and the
    loss it returns equals the oracle's loss on the same scores."""
    import ptranking_b200
    from ptranking_b200 import LABEL_TYPE
    from tests.test_oracle_vs_golden import point_cfg
    F = 46
    sf = dict(sf_id="pointsf", opt="Adam", lr=1e-3, pointsf=point_cfg(F, num_layers=2))
    paras = {"RankMSE": None, "RankCosine": None, "STListNet": dict(model_id="STListNet", temperature=1.0),
             "SoftRank": dict(model_id="SoftRank", delta=2.0, metric="nDCG", top_k=None)}[model]
    cls = getattr(ptranking_b200, model)
    torch.manual_seed(3)
    r = cls(sf_para_dict=sf, gpu=True, device=DEV) if paras is None else cls(sf_para_dict=sf, model_para_dict=paras, gpu=True, device=DEV)
    r.init()
    s, y = _synth(6, 40, seed=8)
    X = torch.randn(6, 40, F, generator=torch.Generator().manual_seed(1)).to(DEV)
    yt = torch.from_numpy(y).to(DEV)
    before = r.grad_bucket.flat_param.clone()
    with torch.no_grad():
        scores = r.predict(X).detach().cpu()
    kw = dict(presort=True, label_type=LABEL_TYPE.MultiLabel)
    okw = {}
    if model == "STListNet":
        unif = torch.rand(6, 40, generator=torch.Generator().manual_seed(2))
        kw["unif"] = unif.to(DEV)
        okw = dict(temperature=1.0, unif=unif)
    if model == "SoftRank":
        okw = dict(delta=2.0, top_k=None)
    loss, stop = r.train_op(X, yt, **kw)
    want, _ = rp.loss_and_grad(model, scores, torch.from_numpy(y), **okw)
    assert not stop and abs(float(loss) - float(want)) <= 2e-5 * max(abs(float(want)), 1.0)
    assert not torch.equal(before, r.grad_bucket.flat_param)          # the optimizer moved the weights
