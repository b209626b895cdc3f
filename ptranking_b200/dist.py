"""Data-parallel plumbing: one process per GPU, queries sharded across ranks, ONE all-reduce
(sum, fp32) of a flat gradient buffer per step.

The reference has no distributed code ("Parallelism strategies present: none");
this layer is new.  Every reference loss is a SUM over queries (lambdarank.py:56,
listnet.py:39 ...), so summing shard gradients reproduces the single-device large-batch
gradient up to fp32 re-association (exceptions: batch-level BN statistics and ApproxNDCG's
batch coupling).  The module is compute-agnostic so the host logic can be tested
with the gloo backend on CPU.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def init_from_env(backend: str = "nccl") -> tuple:
    """(rank, local_rank, world_size) from torchrun's environment; initialises the process group
    when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_queries(num_queries: int, rank: int, world: int) -> range:
    """Contiguous block of query indices owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(num_queries, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class GradBucket:
    """Flat fp32 gradient buffer: every parameter's ``.grad`` is a view into it, so a step needs
    one ``all_reduce(SUM)`` (220.8 KB for the default pointwise scorer) instead of one per tensor."""

    def __init__(self, params: Sequence[torch.nn.Parameter], align: int = 1):
        """``align``: every tensor starts at a multiple of ``align`` elements (4 keeps float4 access legal when the
        parameters themselves are re-homed into a buffer of the same layout, see :meth:`flatten_params`)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += -(-p.numel() // align) * align
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_param = None
        for p, v in zip(self.params, self._views()):
            p.grad = v

    def flatten_params(self) -> torch.Tensor:
        """Re-home every parameter into one flat buffer laid out exactly like the gradient buffer (values kept), so
        an optimizer can update all of them in a single elementwise pass.  Modules keep their Parameter objects."""
        if self.flat_param is None:
            self.flat_param = torch.zeros_like(self.flat)
            for p, o in zip(self.params, self.offsets):
                view = self.flat_param[o: o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        return self.flat_param

    def params_are_flat(self) -> bool:
        return self.flat_param is not None and all(
            p.data.data_ptr() == self.flat_param.data_ptr() + 4 * o for p, o in zip(self.params, self.offsets))

    def zero(self, skip_memset: bool = False) -> None:
        """optimizer.zero_grad() that keeps the views alive.  ``skip_memset``: every gradient is overwritten
        (not accumulated) by the backward kernels this step, so clearing the buffer is unnecessary."""
        if not skip_memset:
            self.flat.zero_()
        for p, v in zip(self.params, self._views()):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def _views(self):
        for p, off in zip(self.params, self.offsets):
            yield self.flat[off: off + p.numel()].view_as(p)

    def all_reduce(self) -> None:
        if is_distributed():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def broadcast_parameters(bucket: "GradBucket", src: int = 0) -> None:
    """Make every replica's parameters equal to rank ``src``'s (no-op when not distributed): one broadcast of the flat
    parameter buffer when the parameters were re-homed into it, else one per tensor."""
    if not is_distributed():
        return
    if bucket.flat_param is not None and bucket.params_are_flat():
        dist.broadcast(bucket.flat_param, src=src)
    else:
        for p in bucket.params:
            dist.broadcast(p.data, src=src)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
