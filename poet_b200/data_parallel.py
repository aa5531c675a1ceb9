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

    # ---- all-reduce overlapped with the backward pass ------------------------------------------------------------
    # The arena is cut into segments in the order their gradients complete (pose heads, decoder, encoder layers from
    # the last to the first, the rest); poet_b200.ops.grad_marker nodes in the model's forward report progress during
    # backward and each report launches the all-reduce of the segment finished ONE stage earlier (so the result does not
    # depend on the order autograd picks among ready nodes) on a communication stream that waits for the calling stream
    # and for every side stream that received work in this step.  finish() reduces what is left and joins.  Inside a
    # CUDA-graph capture all of this becomes graph nodes: the replay needs no host-side NCCL call.
    def plan_overlap(self, named_parameters) -> bool:
        """named_parameters: the (name, param) pairs of the model whose parameters this reducer was built from, in
        the same order.  Returns False (and leaves overlap off) if a group is not contiguous in the arena."""
        names = [n for n, p in named_parameters if p.requires_grad]
        if len(names) != len(self.params):
            return False
        ends = self.offsets[1:] + [self.flat.numel()]

        def group_of(n):
            parts = n.split(".")
            if parts[0] in ("translation_head", "rotation_head"):
                return ("heads",)
            if parts[0] == "transformer" and parts[1] == "decoder":
                return ("decoder",)
            if parts[0] == "transformer" and parts[1] == "encoder" and parts[2] == "layers":
                return ("encoder", int(parts[3]))
            return ("rest",)

        ranges = {}
        for n, a, b in zip(names, self.offsets, ends):
            ranges.setdefault(group_of(n), []).append((a, b))
        segs = {}
        for g, rs in ranges.items():
            if g == ("rest",):
                continue
            rs.sort()
            if any(rs[i][1] != rs[i + 1][0] for i in range(len(rs) - 1)):
                if g == ("heads",):
                    continue                                   # translation / rotation heads around other tensors: leave to finish()
                return False
            segs[g] = (rs[0][0], rs[-1][1])
        n_enc = 1 + max([g[1] for g in segs if g[0] == "encoder"], default=-1)
        # marker key -> segment whose gradients are certainly complete when that marker fires
        self._ready = {("decoder", 0): segs.get(("heads",))}
        for i in range(n_enc):
            self._ready[("encoder", i)] = segs.get(("decoder",)) if i == n_enc - 1 else segs.get(("encoder", i + 1))
        import os
        if os.environ.get("POET_OVERLAP_COARSE", "0") != "0":
            # one early collective only: the decoder's segment as soon as the last encoder layer's marker fires (its
            # all-reduce then runs next to the whole encoder backward); everything else in finish()
            self._ready = {("encoder", n_enc - 1): segs.get(("decoder",))}
        self._overlap = True
        self._comm_stream = None
        self._works, self._done = [], []
        return True

    def begin_step(self) -> None:
        self._works, self._done = [], []
        if self.flat.is_cuda:
            from . import ops
            ops.reset_touched_side_streams()
            ops.set_grad_marker_callback(self.on_marker)

    def _launch(self, a: int, b: int) -> None:
        self._done.append((a, b))
        if not self.flat.is_cuda:                              # host logic under test (gloo): same segments, blocking calls
            seg = self.flat[a:b]
            if self.average:
                seg.mul_(1.0 / self.world_size())
            dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group)
            return
        from . import ops
        dev = self.flat.device
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        self._comm_stream.wait_stream(cur)
        for s in ops.touched_side_streams():
            self._comm_stream.wait_stream(s)
        with torch.cuda.stream(self._comm_stream):
            op = dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM
            self._works.append(dist.all_reduce(self.flat[a:b], op=op, group=self.group, async_op=True))

    def on_marker(self, key) -> None:
        seg = self._ready.get(key)
        if seg is not None and seg not in self._done:
            self._launch(*seg)

    def finish(self) -> None:
        """All-reduce every part of the arena no marker has covered, then make the calling stream wait for all of it."""
        if self.flat.is_cuda:
            from . import ops
            ops.set_grad_marker_callback(None)
        pos = 0
        for a, b in sorted(self._done):
            if a > pos:
                self._launch(pos, a)
            pos = max(pos, b)
        if pos < self.flat.numel():
            self._launch(pos, self.flat.numel())
        for w in self._works:
            w.wait()
        if self.flat.is_cuda and self._comm_stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._comm_stream)
        self._works = []

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous image shard of rank `rank` (rank r gets [r*B/G, (r+1)*B/G))."""
    if n_items % world:
        raise ValueError(f"batch {n_items} does not divide over {world} ranks")
    per = n_items // world
    return rank * per, (rank + 1) * per
