"""Batch-sharded data parallelism: one process per GPU, replicated parameters, one flat all-reduce of the
gradients per optimiser step over NCCL/NVLink (gloo on CPU for tests).  Replaces the reference's
single-process nn.DataParallel (train_multi_stylegan.py:67-70), which re-broadcasts both models on every
forward."""
import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> int:
    """Initialise torch.distributed from torchrun's environment; returns the local rank (0 if single)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
    return local_rank


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


@torch.no_grad()
def broadcast_parameters(modules: Iterable[torch.nn.Module], src: int = 0, group=None) -> None:
    """Once at start-up (the reference's DataParallel does this on every forward call)."""
    if world_size(group) == 1:
        return
    for m in modules:
        tensors = [p.data for p in m.parameters()] + [b.data for b in m.buffers()]
        if not tensors:
            continue
        flat = torch._utils._flatten_dense_tensors(tensors)
        dist.broadcast(flat, src=src, group=group)
        for t, f in zip(tensors, torch._utils._unflatten_dense_tensors(flat, tensors)):
            t.copy_(f)


class PendingReduce(object):
    """An averaging all-reduce in flight: `wait()` orders the current stream after the collective and writes the
    averages back into the tensors (one multi-tensor copy)."""

    def __init__(self, tensors, flat, work, divide_by):
        self.tensors, self.flat, self.work, self.divide_by = tensors, flat, work, divide_by

    @torch.no_grad()
    def wait(self) -> int:
        if self.work is not None:
            self.work.wait()
        if self.divide_by != 1:
            self.flat.div_(self.divide_by)
        views = torch._utils._unflatten_dense_tensors(self.flat, self.tensors)
        if self.flat.is_cuda:
            torch._foreach_copy_(list(self.tensors), list(views))
        else:
            for g, f in zip(self.tensors, views):
                g.copy_(f)
        return self.flat.numel()


@torch.no_grad()
def all_reduce_tensors_begin(tensors: List[torch.Tensor], group=None) -> Optional[PendingReduce]:
    """Start averaging the given tensors across ranks with ONE collective on a flat fp32 buffer (NCCL: ReduceOp.AVG,
    asynchronous on NCCL's stream, so kernels issued before `wait()` overlap the transfer)."""
    ws = world_size(group)
    if ws == 1 or not tensors:
        return None
    flat = torch._utils._flatten_dense_tensors(tensors)
    nccl = dist.get_backend(group) == "nccl"
    work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM, group=group, async_op=True)
    return PendingReduce(tensors, flat, work, 1 if nccl else ws)


@torch.no_grad()
def all_reduce_tensors(tensors: List[torch.Tensor], group=None) -> int:
    """Average the given tensors across ranks in place with ONE collective on a flat fp32 buffer; returns the
    number of elements reduced."""
    pending = all_reduce_tensors_begin(tensors, group)
    return 0 if pending is None else pending.wait()


@torch.no_grad()
def all_reduce_gradients(parameters: Iterable[torch.nn.Parameter], group=None) -> int:
    """Average gradients across ranks with ONE collective.  Parameters that never receive a gradient (the
    generator's unobservable second branch) are skipped on every rank alike."""
    if world_size(group) == 1:
        return 0
    return all_reduce_tensors([p.grad for p in parameters if p.grad is not None], group)


@torch.no_grad()
def all_reduce_mean_(t: torch.Tensor, group=None) -> torch.Tensor:
    ws = world_size(group)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(ws)
    return t
