"""Host logic of the fused Q|K|V projection (CPU): `ops.adjacent_rows` stacks parameter tensors that sit side by side in the
flat parameter buffer WITHOUT copying them, falls back to a copy otherwise, and hands every tensor its rows of the
gradient; `dist.GradBucket` keeps the order it is given, which is how the list ranker puts an attention block's three
projection matrices next to each other (base/list_ranker.py: get_parameters)."""
import pytest
import torch

from ptranking_b200 import dist as b200dist


def _ops():
    from ptranking_b200 import ops
    return ops


def test_adjacent_rows_is_a_view_of_the_flat_buffer_and_splits_the_gradient():
    ops = _ops()
    F = 12
    ps = [torch.nn.Parameter(torch.randn(F, F)) for _ in range(3)] + [torch.nn.Parameter(torch.randn(F)) for _ in range(3)]
    bucket = b200dist.GradBucket(ps, align=4)
    flat = bucket.flatten_params()
    wq, wk, wv, bq, bk, bv = ps
    W = ops.adjacent_rows(wq, wk, wv)
    b = ops.adjacent_rows(bq, bk, bv)
    assert W.shape == (3 * F, F) and b.shape == (3 * F,)
    assert W.data_ptr() == wq.data_ptr() == flat.data_ptr()            # no copy: a strided view of the flat buffer
    assert b.data_ptr() == bq.data_ptr()
    assert torch.equal(W, torch.cat([wq, wk, wv]).detach()) and torch.equal(b, torch.cat([bq, bk, bv]).detach())
    x = torch.randn(5, F)
    (x @ W.t() + b).square().sum().backward()
    # the same computation with three separate projections
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    sum((x @ ref[i].t() + ref[3 + i]).square().sum() for i in range(3)).backward()
    for p, r in zip(ps, ref):
        assert torch.allclose(p.grad, r.grad, rtol=1e-6, atol=1e-6)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket._views()))   # accumulated in place


def test_adjacent_rows_copies_when_the_tensors_are_not_adjacent():
    ops = _ops()
    a, c = torch.nn.Parameter(torch.randn(4, 6)), torch.nn.Parameter(torch.randn(4, 6))
    W = ops.adjacent_rows(a, c)
    assert W.data_ptr() != a.data_ptr() and torch.equal(W, torch.cat([a, c]).detach())
    W.sum().backward()
    assert torch.equal(a.grad, torch.ones_like(a)) and torch.equal(c.grad, torch.ones_like(c))
    # adjacent in memory but in the wrong order is not adjacent either
    flat = torch.randn(48)
    lo, hi = flat[:24].view(4, 6).requires_grad_(), flat[24:].view(4, 6).requires_grad_()
    assert ops.adjacent_rows(hi, lo).data_ptr() not in (lo.data_ptr(), hi.data_ptr())
    assert ops.adjacent_rows(lo, hi).data_ptr() == lo.data_ptr()


def test_list_ranker_lists_projection_weights_side_by_side():
    """Pure ordering logic of ListNeuralRanker.get_parameters, exercised without a GPU through the module classes."""
    from ptranking_b200.base.list_ranker import MultiheadAttention
    m = MultiheadAttention(hid_dim=8, n_heads=2)
    order = m.projection_parameters()
    assert [id(p) for p in order[:3]] == [id(m.w_q.weight), id(m.w_k.weight), id(m.w_v.weight)]
    assert [id(p) for p in order[3:6]] == [id(m.w_q.bias), id(m.w_k.bias), id(m.w_v.bias)]
    assert {id(p) for p in order} == {id(p) for p in m.parameters()}
    bucket = b200dist.GradBucket(order, align=4)
    bucket.flatten_params()
    ops = _ops()
    assert ops.adjacent_rows(m.w_q.weight, m.w_k.weight, m.w_v.weight).data_ptr() == m.w_q.weight.data_ptr()
