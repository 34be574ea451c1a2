"""Data-parallel plumbing of the head: one process per GPU, batch sharded, head parameters replicated.

The only exchange step of the path is the gradient all-reduce of the trainable head parameters
(`prototype_vectors`, `prototype_vectors_global`, `add_on_layers.*`; the last layers are frozen,
protopformer.py:130-131) -- what DDP does implicitly in the reference (main.py:370).  Here the gradients of those
parameters are views into ONE flat fp32 buffer, so a step issues a single NCCL all-reduce (3.2 MB at the CUB shape,
latency-bound over NVLink/NVSwitch) on the backward stream right behind the prototype-gradient kernel.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def head_parameters(module):
    """Trainable head parameters in a fixed order (names follow tools/create_optimizer.py:31-39)."""
    names = ["prototype_vectors", "prototype_vectors_global", "add_on_layers.0.weight", "add_on_layers.0.bias"]
    params = dict(module.named_parameters())
    return [(n, params[n]) for n in names if n in params and params[n].requires_grad]


class FlatGradReducer:
    """Keeps `.grad` of the given parameters as views of one flat buffer and averages it across ranks."""

    def __init__(self, named_params, process_group=None):
        self.named = list(named_params)
        self.group = process_group
        total = sum(p.numel() for _, p in self.named)
        ref = self.named[0][1]
        self.flat = torch.zeros(total, dtype=torch.float32, device=ref.device)
        off = 0
        self.views = []
        for _, p in self.named:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)
            off += p.numel()

    def zero(self):
        """Zero the flat buffer and re-attach the views (autograd then accumulates in place)."""
        self.flat.zero_()
        for (_, p), v in zip(self.named, self.views):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def check_attached(self):
        """Raises if some parameter's .grad no longer aliases the flat buffer (e.g. after an optimizer's
        zero_grad(set_to_none=True)): reducing the flat buffer would then average stale values while autograd
        accumulates somewhere else."""
        for (n, p), v in zip(self.named, self.views):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                raise RuntimeError(f"gradient of {n} is detached from the flat all-reduce buffer: call "
                                   "FlatGradReducer.zero() (not optimizer.zero_grad(set_to_none=True)) between steps")

    def segment(self, names):
        """[lo, hi) of the flat buffer covered by the (contiguous, in construction order) parameters `names`."""
        off, lo, hi = 0, None, None
        for n, p in self.named:
            if n in names:
                lo = off if lo is None else lo
                hi = off + p.numel()
            off += p.numel()
        return lo, hi

    def allreduce(self, async_op: bool = False, lo: int = 0, hi: int | None = None, check: bool = True):
        """Average flat[lo:hi] (default: all of it) over the process group (no-op for a single process)."""
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(self.group)
        if world == 1:
            return None
        if check:
            self.check_attached()
        buf = self.flat if (lo == 0 and hi is None) else self.flat[lo:hi]
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=False)
        buf.div_(world)
        return work


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous batch shard [lo, hi) of rank `rank` (images are independent: no data-path collective)."""
    per, rem = divmod(n_items, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)
