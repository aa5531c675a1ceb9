"""Data-parallel gradient exchange for the PoET hot path.

The image batch shards across ranks with no forward communication; the only collective is one
sum all-reduce of the parameter gradients per step (the reference's DDP, main.py:282).  Instead of
DDP's bucketing + autograd hooks we keep ONE flat fp32 gradient arena: every ``p.grad`` is a view
into it, so backward accumulates straight into the arena and a single NCCL all-reduce over
NVLink/NVSwitch (NVLS in-switch reduction when available) replaces all buckets.  Parameters that
receive no gradient (``transformer.reference_points.*`` in bbox mode — the reason the reference
needs ``find_unused_parameters=True``) simply stay zero, so every rank reduces the same length.

Works with any backend (NCCL on the GPUs; gloo in the CPU unit tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, average: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.offsets, total = [], 0
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("all parameters must share device and dtype")
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4               # keep every view 16-byte aligned
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        self.group = process_group
        self.average = average
        self.bind()

    def bind(self) -> None:
        """(Re-)point every p.grad at its slice of the arena and register the slices as direct-accumulation slots:
        the backward kernels of poet_b200.ops add parameter gradients straight into them (no AccumulateGrad node, hence
        no gradient hooks: this reducer replaces DistributedDataParallel, it does not combine with it)."""
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        if self.flat.is_cuda:
            from . import ops
            ops.register_direct_grad_slots([p.grad for p in self.params])

    def zero(self) -> None:
        self.flat.zero_()
        rebound = []
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                rebound.append(p.grad)
        if rebound and self.flat.is_cuda:
            from . import ops
            ops.register_direct_grad_slots(rebound)

    def __del__(self):
        try:
            if self.flat.is_cuda:
                from . import ops
                ops.unregister_direct_grad_slots([self.flat[off:off + p.numel()] for p, off in zip(self.params, self.offsets)])
        except Exception:
            pass

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self, async_op: bool = False):
        """Sum (and average) the arena across ranks.  Returns the work handle when async_op."""
        ws = self.world_size()
        if ws == 1:
            return None
        if self.average and dist.get_backend(self.group) == "nccl":
            # ncclAvg: the 1/world scaling happens inside the collective, no extra pass over the arena
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        if self.average:
            self.flat.mul_(1.0 / ws)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous image shard of rank `rank` (rank r gets [r*B/G, (r+1)*B/G))."""
    if n_items % world:
        raise ValueError(f"batch {n_items} does not divide over {world} ranks")
    per = n_items // world
    return rank * per, (rank + 1) * per
