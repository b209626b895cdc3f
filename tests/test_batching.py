"""Host-side length-bucketed batching: uniform n per batch, presort contract, document-count targeting,
data-parallel sharding.  CPU only."""
import numpy as np
import pytest
import torch

from ptranking_b200.data import LengthBucketedBatches, presort_query


def _queries(rng, lengths, F=7):
    out = []
    for i, n in enumerate(lengths):
        out.append((f"q{i}", rng.standard_normal((n, F)).astype(np.float32), rng.integers(0, 5, size=n).astype(np.float32)))
    return out


def test_presort_is_stable_descending():
    X = np.arange(12, dtype=np.float32).reshape(6, 2)
    y = np.array([1, 3, 1, 0, 3, 2], dtype=np.float32)
    Xs, ys = presort_query(X, y)
    assert ys.tolist() == [3, 3, 2, 1, 1, 0]
    assert Xs[:, 0].tolist() == [2, 8, 10, 0, 4, 6]            # ties keep dataset order (stable), like the reference's sort


def test_batches_are_uniform_cover_every_query_once_and_hit_the_document_target():
    rng = np.random.default_rng(0)
    lengths = [5] * 37 + [12] * 20 + [40] * 3 + [1] * 4 + [0]
    qs = _queries(rng, lengths)
    it = LengthBucketedBatches(qs, docs_per_batch=120, pin_memory=False)
    seen = []
    for ids, X, y in it:
        B, n, F = X.shape
        assert y.shape == (B, n) and len(ids) == B and F == 7
        assert B <= max(1, 120 // n)
        assert torch.all(y[:, :-1] >= y[:, 1:])                  # presorted descending per query
        for b, q in enumerate(ids):
            src = qs[int(q[1:])]
            assert src[1].shape[0] == n
            assert sorted(map(tuple, X[b].numpy().tolist())) == sorted(map(tuple, src[1].tolist()))
        seen += ids
    assert sorted(seen) == sorted(f"q{i}" for i, n in enumerate(lengths) if n > 0)     # empty query skipped, nothing duplicated
    st = it.stats()
    assert st["queries"] == 64 and st["docs"] == 37 * 5 + 20 * 12 + 3 * 40 + 4 and st["lengths"] == 4
    assert it.batch_size(5) == 24 and it.batch_size(40) == 3 and it.batch_size(1000) == 1


def test_shuffle_changes_order_between_epochs_but_not_membership():
    rng = np.random.default_rng(1)
    qs = _queries(rng, [8] * 50)
    it = LengthBucketedBatches(qs, docs_per_batch=64, shuffle_seed=5, pin_memory=False)
    e1 = [tuple(ids) for ids, _, _ in it]
    e2 = [tuple(ids) for ids, _, _ in it]
    assert e1 != e2 and sorted(sum(map(list, e1), [])) == sorted(sum(map(list, e2), []))
    fixed = LengthBucketedBatches(qs, docs_per_batch=64, pin_memory=False)
    assert [tuple(i) for i, _, _ in fixed] == [tuple(i) for i, _, _ in fixed]


def test_ranks_get_disjoint_equal_length_shards():
    rng = np.random.default_rng(2)
    qs = _queries(rng, [6] * 45 + [9] * 31)
    shards = [LengthBucketedBatches(qs, docs_per_batch=36, shuffle_seed=3, rank=r, world=3, pin_memory=False) for r in range(3)]
    per_rank = [[tuple(ids) for ids, _, _ in s] for s in shards]
    assert len({len(p) for p in per_rank}) == 1                  # same number of steps on every rank (one all-reduce per step)
    flat = [q for p in per_rank for b in p for q in b]
    assert len(flat) == len(set(flat))                           # no query trained twice in an epoch


def test_bad_input_is_rejected():
    with pytest.raises(ValueError):
        LengthBucketedBatches([("a", np.zeros((3, 2), np.float32), np.zeros(4, np.float32))], pin_memory=False)
    with pytest.raises(ValueError):
        LengthBucketedBatches([("a", np.zeros((3, 2), np.float32), np.zeros(3, np.float32)),
                               ("b", np.zeros((3, 5), np.float32), np.zeros(3, np.float32))], pin_memory=False)
    with pytest.raises(ValueError):
        LengthBucketedBatches([], rank=2, world=2)
