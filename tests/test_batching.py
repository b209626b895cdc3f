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


def test_ragged_batches_pack_by_documents_and_bucket_by_length():
    """RaggedBatches: queries of any length packed up to the document target; inside a batch longest first, with length
    buckets that cover the batch in order and bound every list they hold."""
    from ptranking_b200.data import RaggedBatches, length_buckets
    rng = np.random.default_rng(0)
    lens = np.clip(rng.lognormal(4.45, 0.85, 600), 1, 1251).astype(int)
    queries = [(f"q{i}", rng.standard_normal((n, 5)).astype(np.float32), np.sort(rng.integers(0, 5, n))[::-1].astype(np.float32))
               for i, n in enumerate(lens)]
    queries.append(("empty", np.zeros((0, 5), dtype=np.float32), np.zeros(0, dtype=np.float32)))     # skipped like the reference does
    loader = RaggedBatches(queries, docs_per_batch=20000, presort=False, pin_memory=False)
    seen, docs = [], 0
    for ids, X, y, offsets, max_len, buckets in loader:
        B = len(ids)
        ln = (offsets[1:] - offsets[:-1]).numpy()
        assert offsets[0] == 0 and offsets[-1] == X.shape[0] == y.shape[0] and X.shape[1] == 5
        assert (np.diff(ln) <= 0).all() and ln.max() == max_len and ln.min() >= 1
        assert buckets[0][0] == 0 and buckets[-1][1] == B and all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))
        for b0, b1, ml in buckets:
            assert ln[b0:b1].max() == ml
        assert X.shape[0] <= 20000 or B == 1
        for i, q in enumerate(ids):                       # every query's rows are its own
            src = queries[int(q[1:])]
            assert np.array_equal(X[offsets[i]: offsets[i + 1]].numpy(), src[1]) and np.array_equal(y[offsets[i]: offsets[i + 1]].numpy(), src[2])
        seen += ids
        docs += X.shape[0]
    assert sorted(seen) == sorted(f"q{i}" for i in range(600)) and docs == int(lens.sum())
    st = loader.stats()
    assert st["queries"] == 600 and st["max_len"] == int(lens.max()) and st["batches"] == len(loader)
    # data-parallel shards: disjoint, equal batch counts
    a = RaggedBatches(queries, docs_per_batch=20000, presort=False, pin_memory=False, rank=0, world=2)
    b = RaggedBatches(queries, docs_per_batch=20000, presort=False, pin_memory=False, rank=1, world=2)
    ia, ib = [q for batch in a for q in batch[0]], [q for batch in b for q in batch[0]]
    assert len(a) == len(b) and not set(ia) & set(ib)
    assert length_buckets([]) == [] and length_buckets([7]) == [(0, 1, 7)]


def test_ragged_batches_take_custom_length_classes():
    """Finer classes for the list scorer (attention pads every class to its longest list): the classes still cover the batch
    in order, each holds lists no longer than its first one, and bad edges are rejected."""
    from ptranking_b200.data import RaggedBatches
    rng = np.random.default_rng(0)
    lens = np.clip(rng.lognormal(4.0, 0.8, 300), 1, 400).astype(int)
    queries = [(f"q{i}", rng.standard_normal((n, 5)).astype(np.float32), rng.integers(0, 3, n).astype(np.float32)) for i, n in enumerate(lens)]
    rb = RaggedBatches(queries, docs_per_batch=1 << 20, pin_memory=False, bucket_edges=(16, 32, 64, 128, 256))
    (ids, X, y, offsets, max_len, buckets), = list(rb)
    assert len(buckets) >= 4 and buckets[0][0] == 0 and buckets[-1][1] == len(ids)
    blens = (offsets[1:] - offsets[:-1]).numpy()
    for (a0, a1, ml), nxt in zip(buckets, buckets[1:] + [None]):
        assert ml == blens[a0] == blens[a0:a1].max()
        if nxt is not None:
            assert nxt[0] == a1
    with pytest.raises(ValueError):
        RaggedBatches(queries, bucket_edges=(64, 32))


def test_length_buckets_invariants_property():
    """For any descending length vector and any increasing edges: the classes tile [0, N) in order and every class records
    the length of its first (= longest) list -- what the kernels size their CTAs / padding by."""
    from hypothesis import given, settings, strategies as st
    from ptranking_b200.data import length_buckets

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(1, 2000), min_size=0, max_size=300),
           st.lists(st.integers(1, 1500), min_size=1, max_size=5, unique=True))
    def check(lens, edges):
        lens = sorted(lens, reverse=True)
        edges = sorted(edges)
        out = length_buckets(lens, edges=edges)
        if not lens:
            assert out == []
            return
        assert out[0][0] == 0 and out[-1][1] == len(lens)
        for (a0, a1, ml), nxt in zip(out, out[1:] + [None]):
            assert a0 < a1 and ml == lens[a0] == max(lens[a0:a1])
            if nxt is not None:
                assert nxt[0] == a1

    check()
