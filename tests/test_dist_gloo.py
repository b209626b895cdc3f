"""world_size-2 gloo test of the data-parallel host logic (query sharding + flat-bucket
all-reduce SUM): sharded gradients == single-process full-batch gradients.  The compute inside
each rank is the oracle (CPU); the product's kernels are exercised by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_port as rp
from ptranking_b200 import dist as b200dist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _make(seed=137):
    torch.manual_seed(seed)
    net = rp.point_scorer(num_features=12, num_layers=2, h_dim=16, AF="R", TL_AF="S", apply_tl_af=True,
                          BN=True, bn_type="BN2", bn_affine=False, dropout=0.0)
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(6, 10, 12, generator=g)
    y = torch.sort(torch.randint(0, 5, (6, 10), generator=g).float(), dim=1, descending=True)[0]
    y[:, 0] = torch.clamp(y[:, 0], min=1.0)
    return net, X, y


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    b200dist.init_from_env(backend="gloo")
    net, X, y = _make()
    params = list(net.parameters())
    bucket = b200dist.GradBucket(params)
    mine = b200dist.shard_queries(X.size(0), rank, world)
    idx = torch.tensor(list(mine))
    bucket.zero()
    loss = rp.lambdarank_loss(rp.point_forward(net, X[idx]), y[idx], sigma=1.0)
    loss.backward()
    bucket.all_reduce()
    total = b200dist.all_reduce_sum_(loss.detach().clone().reshape(1))
    if rank == 0:
        out["flat"] = bucket.flat.clone().numpy()
        out["loss"] = float(total)
    dist.destroy_process_group()


def test_sharded_gradients_sum_to_full_batch():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    net, X, y = _make()
    loss = rp.lambdarank_loss(rp.point_forward(net, X), y, sigma=1.0)
    loss.backward()
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy()
    assert abs(out["loss"] - float(loss)) <= 1e-5 * abs(float(loss))
    assert np.abs(out["flat"] - full).max() <= 1e-5 * max(np.abs(full).max(), 1e-6)


def test_shard_queries_partition():
    for nq in (1, 7, 8, 1024, 1025):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                seen += list(b200dist.shard_queries(nq, r, w))
            assert seen == list(range(nq))
            sizes = [len(b200dist.shard_queries(nq, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_grad_bucket_views_survive_zero():
    p = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    b = b200dist.GradBucket(p)
    (p[0].sum() * 2 + p[1].sum() * 3).backward()
    assert torch.allclose(b.flat, torch.cat([torch.full((12,), 2.0), torch.full((5,), 3.0)]))
    b.zero()
    assert float(b.flat.abs().sum()) == 0.0 and p[0].grad.data_ptr() == b.flat.data_ptr()


def test_aligned_bucket_and_flat_parameters_share_one_layout():
    """align=4 keeps every tensor 16-byte aligned; flatten_params re-homes the parameters into a buffer with the same
    offsets, keeps their values and identity, and gradients keep flowing into the flat gradient buffer."""
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)                       # weight 15 elements, bias 3: neither is a multiple of 4
    extra = torch.nn.Parameter(torch.randn(1))
    params = list(lin.parameters()) + [extra]
    before = [p.detach().clone() for p in params]
    b = b200dist.GradBucket(params, align=4)
    assert b.offsets == [0, 16, 20] and b.flat.numel() == 24
    assert not b.params_are_flat()
    flat = b.flatten_params()
    assert b.params_are_flat() and flat.numel() == b.flat.numel()
    assert all(torch.equal(p.detach(), q) for p, q in zip(params, before))
    assert list(lin.parameters())[0] is params[0]                      # same Parameter objects, new storage
    assert all(p.data_ptr() == flat.data_ptr() + 4 * o for p, o in zip(params, b.offsets))
    assert float(flat[15]) == 0.0 and float(flat[19:20].abs().sum()) == 0.0      # padding stays zero
    (lin(torch.ones(2, 5)).sum() + 3 * extra.sum()).backward()
    assert torch.allclose(b.flat[16:19], torch.full((3,), 2.0)) and float(b.flat[20]) == 3.0
    # an in-place update of the flat buffer is an update of every parameter (what the fused optimizer does)
    flat.add_(1.0)
    assert all(torch.allclose(p.detach(), q + 1.0) for p, q in zip(params, before))
    lin.load_state_dict({k: v.clone() for k, v in lin.state_dict().items()})
    assert b.params_are_flat()
    params[0].data = params[0].data.clone()                              # re-allocation is detected
    assert not b.params_are_flat()


def _bcast_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    b200dist.init_from_env(backend="gloo")
    for flat in (False, True):
        net, _, _ = _make(seed=100 + rank)              # every rank draws different initial weights
        bucket = b200dist.GradBucket(list(net.parameters()), align=4)
        if flat:
            bucket.flatten_params()
        b200dist.broadcast_parameters(bucket, src=0)
        out[(rank, flat)] = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy()
    dist.destroy_process_group()


def test_replicas_start_from_rank0_weights():
    """ranks seeded differently must hold identical parameters after the optimizer is configured
    (NeuralRanker.config_optimizer ends with dist.broadcast_parameters)."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bcast_worker, args=(2, port, out), nprocs=2, join=True)
    ref, _, _ = _make(seed=100)
    want = torch.cat([p.detach().reshape(-1) for p in ref.parameters()]).numpy()
    for flat in (False, True):
        assert np.array_equal(out[(0, flat)], want)
        assert np.array_equal(out[(1, flat)], want)
