"""Pin the oracle's sibling-loss restatements (RankMSE, RankCosine, STListNet, SoftRank, the Sinkhorn step)
against outputs of the unmodified reference (tests/golden/siblings.npz, made by tests/golden/make_golden_siblings.py)."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_port as rp
from tests.helpers import parse_sibling_key, rel_err, sibling_cases, sinkhorn_cases

CASES = sibling_cases()
IDS = [f"{h}-{c}" for h, c, _ in CASES]


@pytest.mark.parametrize("head,case,d", CASES, ids=IDS)
def test_port_matches_reference(head, case, d):
    name, params = parse_sibling_key(head)
    s, y = torch.from_numpy(d["scores"]), torch.from_numpy(d["labels"])
    kw = dict(params)
    if name == "STListNet":
        kw["unif"] = torch.from_numpy(d["unif"])
    loss, grad = rp.loss_and_grad(name, s, y, **kw)
    assert abs(float(loss) - float(d["loss"])) <= 2e-6 * max(1.0, abs(float(d["loss"])))
    assert rel_err(grad.numpy(), d["grad"]) <= 2e-6


@pytest.mark.parametrize("head,case,d", CASES, ids=IDS)
def test_closed_form_matches_reference(head, case, d):
    name, params = parse_sibling_key(head)
    s, y = d["scores"], d["labels"]
    if name == "RankMSE":
        loss, grad = cf.rankmse(s, y)
    elif name == "RankCosine":
        loss, grad = cf.rankcosine(s, y)
    elif name == "STListNet":
        loss, grad = cf.stlistnet(s, y, d["unif"], **params)
    else:
        loss, grad = cf.softrank(s, y, **params)
    assert abs(loss - float(d["loss"])) <= 2e-5 * max(1.0, abs(float(d["loss"])))
    assert rel_err(grad, d["grad"]) <= 5e-5


@pytest.mark.parametrize("case,d", sinkhorn_cases("sinkstep"), ids=[c for c, _ in sinkhorn_cases("sinkstep")])
def test_sinkstep_port(case, d):
    got = rp.sinkstep(torch.from_numpy(d["dist"]), torch.from_numpy(d["log_nu"]), torch.from_numpy(d["log_u"]), float(d["lam"]))
    assert np.array_equal(got.numpy(), d["log_v"], equal_nan=True)


@pytest.mark.parametrize("case,d", sinkhorn_cases("sinkhorn"), ids=[c for c, _ in sinkhorn_cases("sinkhorn")])
def test_sinkhorn_ot_port(case, d):
    dist, gmu, gnu = rp.sinkhorn_ot(torch.from_numpy(d["mu"]), torch.from_numpy(d["nu"]), torch.from_numpy(d["dist"]),
                                    float(d["lam"]), int(d["N"]))
    assert np.array_equal(dist.numpy(), d["distances"])
    assert np.array_equal(gmu.numpy(), d["dmu"]) and np.array_equal(gnu.numpy(), d["dnu"])
