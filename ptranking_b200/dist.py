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

import contextlib
import os
from typing import Iterable, List, Optional, Sequence

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


# --------------------------------------------------------------------------- #
# host hook of the C library: collectives in the middle of a forward / backward call
# --------------------------------------------------------------------------- #
_sync_bn = os.environ.get("PTRANKING_B200_SYNC_BN", "0") == "1"
_call = {"ws": None, "layer_targets": None}
_active_bucket: Optional["GradBucket"] = None
_hook_installed = False


def set_sync_bn(on: bool) -> None:
    """Batch-level ``BN`` statistics over the GLOBAL batch (sum over data-parallel ranks) instead of per rank: with it a
    sharded step computes exactly what one GPU would on the concatenated batch (LTRBatchNorm, base/utils.py:201-223,
    normalises over every document it is handed).  Costs one 2*C+1-double all-reduce per normalised layer in the forward
    pass and one in the backward pass; off by default (PTRANKING_B200_SYNC_BN=1 turns it on)."""
    global _sync_bn
    _sync_bn = bool(on)


def sync_bn_active() -> bool:
    if _sync_bn and is_distributed():
        _install_hook()
        return True
    return False


def _install_hook() -> None:
    global _hook_installed
    if not _hook_installed:
        from . import _lib
        _lib.set_hook(_hook)
        _hook_installed = True


def _hook(what, layer, ptr, count, stream) -> int:
    """Runs on the launching thread between kernel launches (include/ptranking_b200.h: ptrb200_set_hook)."""
    from . import _lib
    try:
        if what == _lib.HOOK_ALLREDUCE_F64:
            ws = _call["ws"]
            if ws is None:
                return 1
            off = ptr - ws.data_ptr()
            if off < 0 or off + 8 * count > ws.numel():
                return 1
            dist.all_reduce(ws[off: off + 8 * count].view(torch.float64), op=dist.ReduceOp.SUM)
        elif what == _lib.HOOK_LAYER_GRADS_READY:
            if _active_bucket is not None and _call["layer_targets"] is not None:
                _active_bucket.layer_ready(layer, _call["layer_targets"])
        return 0
    except Exception as e:      # never let an exception cross the C boundary
        print(f"ptranking_b200.dist hook failed: {e!r}", flush=True)
        return 1


@contextlib.contextmanager
def call_context(ws, layer_targets):
    """Tells the hook which workspace tensor / per-layer gradient tensors the running C call uses."""
    prev = dict(_call)
    _call["ws"], _call["layer_targets"] = ws, layer_targets
    try:
        yield
    finally:
        _call.update(prev)


def shard_queries(num_queries: int, rank: int, world: int) -> range:
    """Contiguous block of query indices owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(num_queries, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class _DeviceMemory:
    """A raw device allocation presented through __cuda_array_interface__ so that torch can view it without a copy."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerExchange:
    """Every rank's gradient buffers mapped into every process of the node (CUDA IPC over NVLink / NVSwitch peer
    memory): the data-parallel gradient SUM is then read by the optimizer kernel itself (csrc/optim.cu), one launch
    instead of ncclAllReduce + the step.  Layout of a rank's allocation: a 256-byte pad (``world`` arrival flags, an error
    word) and two gradient buffers used in alternation -- a rank that is already writing step e+1 gradients can never
    touch what a slower rank still reads for step e, with a single flag exchange per step.  Handles travel through the
    process group (``all_gather_object``); the library only exports / maps memory (ptrb200_peer_alloc / _open)."""

    PAD, ERR_OFF = 256, 128

    def __init__(self, count: int, device):
        from . import _lib
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f"peer exchange supports up to {_lib.MAX_PEERS} ranks")
        self.count, self.device = int(count), torch.device(device)
        self.stride = -(-self.count * 4 // 256) * 256
        # every rank takes part in every collective below whatever happens locally, and all ranks leave through the same
        # exit (all succeed or all raise) -- a rank that bailed out alone would leave the others inside a collective
        self.own, handle, problem = None, None, None
        with torch.cuda.device(self.device):
            try:
                self.own, handle = _lib.peer_alloc(self.PAD + 2 * self.stride)
            except Exception as e:
                problem = e
            handles = [None] * self.world
            dist.all_gather_object(handles, handle)
            self.ptrs = []
            if all(h is not None for h in handles):
                try:
                    self.ptrs = [self.own if r == self.rank else _lib.peer_open(handles[r]) for r in range(self.world)]
                    self.bufs = [torch.as_tensor(_DeviceMemory(self.own + self.PAD + k * self.stride, self.count), device=self.device)
                                 for k in (0, 1)]
                except Exception as e:
                    problem = e
            elif problem is None:
                problem = RuntimeError("another rank could not export its gradient buffer")
            agreed = torch.tensor([0 if problem is not None else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(agreed, op=dist.ReduceOp.MIN)      # also orders "every pad is mapped" before the first signal
        if int(agreed.item()) == 0:
            raise RuntimeError(f"peer exchange unavailable: {problem!r}" if problem is not None else
                               "peer exchange unavailable on another rank")
        self.epoch = 0

    def group(self, k: int):
        """The descriptor of the NEXT exchange over buffer ``k`` (advances the step counter)."""
        from . import _lib
        self.epoch += 1
        g = _lib.PeerGroup()
        g.world, g.rank, g.epoch = self.world, self.rank, self.epoch & 0xffffffff
        for r, p in enumerate(self.ptrs):
            g.grads[r] = p + self.PAD + k * self.stride
            g.flags[r] = p
        g.error = self.own + self.ERR_OFF
        return g

    def error(self) -> int:
        """0, or 1 + the rank that never arrived at some exchange (synchronises the device)."""
        with torch.cuda.device(self.device):
            word = torch.as_tensor(_DeviceMemory(self.own + self.ERR_OFF, 1), device=self.device)
            return int(word.view(torch.int32).item())


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
        self.distributed = True            # False: a purely local bucket even inside an initialised process group
        self._reduced_from = None          # overlapped all-reduce: flat[_reduced_from:] is already on the side stream
        self._side = None
        self.overlap_from_layer = 1        # start reducing when this layer's gradients are complete (layers above it too)
        self.peer: Optional[PeerExchange] = None   # gradient sum over NVLink peer memory inside the optimizer kernel
        self._peer_k = 0
        for p, v in zip(self.params, self._views()):
            p.grad = v

    def enable_peer(self) -> bool:
        """Data parallel on one node: move the gradient buffer into peer-mapped memory so that the optimizer kernel sums
        the ranks' gradients itself (:class:`PeerExchange`).  All ranks agree on the outcome; on any failure (no peer
        access, IPC unavailable) every rank keeps the NCCL all-reduce.  PTRANKING_B200_PEER=0 switches it off."""
        if self.peer is not None:
            return True
        if not (is_distributed() and self.distributed and self.flat.is_cuda and dist.get_backend() == "nccl"
                and os.environ.get("PTRANKING_B200_PEER", "1") == "1"):
            return False
        try:
            ex = PeerExchange(self.flat.numel(), self.flat.device)      # all ranks succeed or all raise (see its __init__)
        except Exception as e:
            print(f"ptranking_b200.dist: rank {dist.get_rank()}: {e}; gradients go through the NCCL all-reduce", flush=True)
            return False
        self.peer = ex
        self._peer_k = 0
        self.flat = ex.bufs[0]
        for p, v in zip(self.params, self._views()):
            p.grad = v
        return True

    def peer_group(self):
        """Descriptor for this step's fused exchange + optimizer launch (None: gradients are reduced by all_reduce())."""
        return self.peer.group(self._peer_k) if (self.peer is not None and self.distributed) else None

    def advance(self) -> None:
        """After the step of a peer-exchange bucket: the next step's gradients go to the other buffer."""
        if self.peer is None:
            return
        self._peer_k ^= 1
        self.flat = self.peer.bufs[self._peer_k]
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
        """Finish the step's gradient reduction: whatever :meth:`layer_ready` has not already put on the side stream is
        reduced now; then the compute stream waits for the side stream."""
        global _active_bucket
        _active_bucket = None
        if not (is_distributed() and self.distributed) or self.peer is not None:     # peer: the optimizer kernel sums
            return
        upto = self._reduced_from if self._reduced_from is not None else self.flat.numel()
        if upto > 0:
            dist.all_reduce(self.flat[:upto], op=dist.ReduceOp.SUM)
        if self._reduced_from is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
            self._reduced_from = None

    def begin_overlap(self) -> None:
        """Arm the overlapped reduction for the backward pass that follows (CUDA + NCCL only): the scorer's backward
        call reports each layer's gradients complete (last layer first) and the tail of the flat buffer goes onto a side
        stream while the layers below are still computing."""
        global _active_bucket
        self._reduced_from = None
        if self.peer is not None:
            return
        if is_distributed() and self.distributed and self.flat.is_cuda and os.environ.get("PTRANKING_B200_OVERLAP", "0") == "1":
            _install_hook()
            _active_bucket = self

    def layer_ready(self, layer: int, layer_targets) -> None:
        if layer != self.overlap_from_layer or self._reduced_from is not None or layer >= len(layer_targets):
            return
        tens = layer_targets[layer]
        if not tens:
            return
        start = min((t.data_ptr() - self.flat.data_ptr()) // 4 for t in tens)
        if not (0 < start < self.flat.numel()):
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.flat.device)
        cur = torch.cuda.current_stream(self.flat.device)
        self._side.wait_stream(cur)                       # everything enqueued so far (layers >= `layer`) is complete
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.flat[start:], op=dist.ReduceOp.SUM)
        self._reduced_from = int(start)


def broadcast_parameters(bucket: "GradBucket", src: int = 0) -> None:
    """Make every replica's parameters equal to rank ``src``'s (no-op when not distributed): one broadcast of the flat
    parameter buffer when the parameters were re-homed into it, else one per tensor."""
    if not (is_distributed() and bucket.distributed):
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
