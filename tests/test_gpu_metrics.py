"""GPU parity of the nDCG@ks kernel: integer ranks bit-exact, values vs the reference fixtures."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from tests.helpers import load

pytestmark = pytest.mark.gpu


def test_known_answer_vector():
    from ptranking_b200 import ops
    z = load("metrics.npz")
    sys_l, std_l = z["kat_sys"], z["kat_std"]
    # scores that rank the documents in the KAT's system order: descending by position
    n = sys_l.shape[1]
    scores = torch.arange(n, 0, -1, dtype=torch.float32).view(1, n).cuda()
    got = ops.ndcg_at_ks(scores, torch.from_numpy(sys_l).cuda(), list(z["kat_ks"]), presort=False).cpu().numpy()
    assert np.allclose(got, z["kat_ndcg_at_ks"], rtol=0, atol=1e-6)
    assert np.allclose(got[0], z["kat_expected_4dp"], atol=5e-5)


@pytest.mark.parametrize("key", ["B5_n50", "B3_n256", "B2_n7", "B2_n1024"])
def test_ndcg_fixture(key):
    from ptranking_b200 import ops
    z = load("metrics.npz")
    s, y, ks = z[key + "__scores"], z[key + "__labels"], [int(k) for k in z[key + "__ks"]]
    out, order = ops.ndcg_at_ks(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), ks, presort=True, return_order=True)
    assert np.array_equal(order.cpu().numpy(), z[key + "__order"])          # integer ranks: bit exact
    assert np.allclose(out.cpu().numpy(), z[key + "__ndcg_at_ks"], rtol=0, atol=1e-6)
    out_ns = ops.ndcg_at_ks(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), ks, presort=False)
    assert np.allclose(out_ns.cpu().numpy(), z[key + "__ndcg_at_ks"], rtol=0, atol=1e-6)   # labels are sorted already
    if key + "__ndcg_at_10" in z.files:
        o10 = ops.ndcg_at_ks(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), [10], presort=True)
        assert np.allclose(o10.cpu().numpy(), z[key + "__ndcg_at_10"], rtol=0, atol=1e-6)


def test_ties_unsorted_and_cutoff_order():
    from ptranking_b200 import ops
    rng = np.random.default_rng(0)
    B, n = 64, 300
    s = np.round(rng.standard_normal((B, n)), 1).astype(np.float32)        # many exact ties
    s[0, :5] = [0.0, -0.0, 0.0, -0.0, 0.0]
    y = rng.integers(0, 5, (B, n)).astype(np.float32)
    ks = [10, 1, 500, 5]                                                  # unsorted cutoffs, one beyond n
    out, order = ops.ndcg_at_ks(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), ks, presort=False, return_order=True)
    want, worder = cf.ndcg_at_ks(s, y, ks, presort=False)
    assert np.array_equal(order.cpu().numpy(), worder.astype(np.int32))    # stable order among equal scores
    assert np.allclose(out.cpu().numpy(), want, rtol=0, atol=2e-6)
    assert np.all(out.cpu().numpy()[:, 2] == 0.0)
