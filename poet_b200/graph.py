"""Whole-step CUDA graph for the PoET hot path.

The path is ~500 kernel launches per forward+backward at fixed shapes; the GPU work of the
launch-bound decoder / head kernels is a few microseconds each, so launching them from Python one by
one leaves the B200 idle most of the step.  `GraphedStep` captures one forward + loss + backward
(all libpoet_b200 launches, the tiny ATen plumbing and the memsets of the gradient arena) into a
single CUDA graph over static input buffers and replays it per step: streams and graphs instead of
a tracing compiler.

    step = GraphedStep(model, loss_fn, srcs, masks, boxes, labels)     # warm-up + capture
    loss, out = step.run(srcs, masks, boxes, labels)                   # copy-in, replay

Inputs may be host (pinned) or device tensors; they are copied into the static buffers on the
current stream.  Shapes (batch, pyramid, queries) are fixed at capture time.

Input pipelining: `prefetch(...)` starts the host-to-device copy of the NEXT step's inputs on a
dedicated copy stream into one of two landing buffer sets, so it overlaps the current step's replay;
`run()` without arguments then consumes the oldest prefetched set (device-to-device into the static
buffers, a few microseconds) before replaying:

    step.prefetch(batch0)
    for batch in batches[1:]:
        step.prefetch(batch)              # H2D of step i+1 ...
        loss, out = step.run()            # ... overlaps the replay of step i
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch

from .data_parallel import FlatGradReducer


class GraphedStep:
    def __init__(self, model, loss_fn: Callable, srcs: Sequence[torch.Tensor], masks: Sequence[torch.Tensor],
                 boxes, labels, reducer: Optional[FlatGradReducer] = None, warmup: int = 3, backward: bool = True,
                 optimizer=None, entry: str = "pyramid", overlap_allreduce: bool = False):
        """optimizer (poet_b200.optim.FusedClipAdamW, optional): its step() writes the bf16 planes of the updated
        weights, so the captured step contains no split pass; call optimizer.step() after every run()."""
        dev = next(model.parameters()).device
        # entry "pyramid": srcs = post-input_proj maps, masks = their masks (the benchmarked path);
        # entry "features": srcs = backbone feature maps, masks = their masks + [padded-image mask] (input_proj included)
        self.entry = entry
        self.model, self.loss_fn, self.backward = model, loss_fn, backward
        # overlap_allreduce: the gradient all-reduce is issued segment by segment DURING the captured backward pass
        # (FlatGradReducer.plan_overlap); run() then returns with the arena already averaged over the ranks
        self.reduces = False
        self.external_planes = optimizer is not None and getattr(optimizer, "planes", None) is not None
        self.reducer = reducer if reducer is not None else (FlatGradReducer(model.parameters()) if backward else None)
        self.s_srcs = [torch.empty(s.shape, dtype=torch.float32, device=dev) for s in srcs]
        self.s_masks = [torch.empty(m.shape, dtype=torch.bool, device=dev) for m in masks]
        B, Q = srcs[0].shape[0], model.n_queries
        self.s_boxes = torch.empty((B, Q, 4), dtype=torch.float32, device=dev)
        self.s_classes = torch.empty((B, Q), dtype=torch.int64, device=dev)
        self.s_counts = torch.empty((B,), dtype=torch.int32, device=dev)
        self.counts_host = None
        self._landing, self._pending, self._copy_stream, self._n_prefetched = None, [], None, 0
        self._copy_in(srcs, masks, boxes, labels)
        if overlap_allreduce and backward and self.reducer is not None and self.reducer.world_size() > 1:
            self.reduces = self.reducer.plan_overlap(list(model.named_parameters()))

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.out = self._body()
        if self.reducer is not None:
            self.reducer.bind()

    # -- the captured region ----------------------------------------------------------------
    def _body(self):
        from . import ops
        if self.reducer is not None:
            self.reducer.zero()
        if self.reduces:
            self.reducer.begin_step()
        with ops.planes_scope(self.model, refresh=not self.external_planes):
            if self.entry == "features":
                out = self.model.forward_features_padded(self.s_srcs, self.s_masks[:-1], self.s_masks[-1], self.s_boxes,
                                                         self.s_classes, self.s_counts)
            else:
                out = self.model.forward_padded(self.s_srcs, self.s_masks, self.s_boxes, self.s_classes, self.s_counts)
        loss = self.loss_fn(out)
        if self.backward:
            loss.backward()
        if self.reduces:
            self.reducer.finish()
        return loss.detach(), out

    def _copy_in(self, srcs, masks, boxes, labels):
        for dst, src in zip(self.s_srcs, srcs):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(self.s_masks, masks):
            dst.copy_(src, non_blocking=True)
        pb, pc, counts, n_dev = self.model._pad_boxes(boxes, labels, self.s_boxes.device)
        self.s_boxes.copy_(pb, non_blocking=True)
        self.s_classes.copy_(pc, non_blocking=True)
        self.s_counts.copy_(n_dev, non_blocking=True)
        self.counts_host = counts

    # -- pipelined input copies ---------------------------------------------------------------
    def _static_set(self):
        return [*self.s_srcs, *self.s_masks, self.s_boxes, self.s_classes, self.s_counts]

    def prefetch(self, srcs, masks, boxes, labels) -> None:
        """Start copying the inputs of a FUTURE run() into a landing buffer set on the copy stream (at most two
        prefetches may be outstanding).  Host tensors should be pinned for the copy to be asynchronous."""
        dev = self.s_boxes.device
        if self._landing is None:
            self._landing = [[torch.empty_like(t) for t in self._static_set()] for _ in range(2)]
            self._free_ev = [None, None]                       # landing set consumed (recorded on the compute stream)
            self._copy_stream = torch.cuda.Stream(device=dev)
            # persistent pinned staging for the padded boxes / classes / counts (allocating pinned memory per call
            # costs more than the copies themselves)
            self._pinned = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in
                             (self.s_boxes, self.s_classes, self.s_counts)] for _ in range(2)]
            self._h2d_ev = [None, None]
        if len(self._pending) >= 2:
            raise RuntimeError("GraphedStep.prefetch: two prefetched input sets are already waiting for run()")
        slot = self._n_prefetched % 2
        self._n_prefetched += 1
        pb, pc, counts, n_dev = self.model._pad_boxes(boxes, labels, boxes[0].device, pin=False)
        if not pb.is_cuda:
            if self._h2d_ev[slot] is not None:
                self._h2d_ev[slot].synchronize()               # the previous copy out of this pinned set has finished
            for dst, src in zip(self._pinned[slot], (pb, pc, n_dev)):
                dst.copy_(src)
            pb, pc, n_dev = self._pinned[slot]
        src_set = [*srcs, *masks, pb, pc, n_dev]
        with torch.cuda.stream(self._copy_stream):
            if self._free_ev[slot] is not None:
                self._copy_stream.wait_event(self._free_ev[slot])
            for dst, src in zip(self._landing[slot], src_set):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._h2d_ev[slot] = ev
        self._pending.append((slot, ev, counts))

    def run(self, srcs=None, masks=None, boxes=None, labels=None):
        """Copy new inputs (if given; else the oldest prefetched set, if any) into the static buffers and replay.
        Returns (loss, out_dict); both alias static graph memory: read them before the next run()."""
        if srcs is not None:
            self._copy_in(srcs, masks, boxes, labels)
        elif self._pending:
            slot, ev, counts = self._pending.pop(0)
            main = torch.cuda.current_stream(self.s_boxes.device)
            main.wait_event(ev)
            for dst, src in zip(self._static_set(), self._landing[slot]):
                dst.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            self._free_ev[slot] = done
            self.counts_host = counts
        self.graph.replay()
        return self.loss, self.out

    def n_boxes_per_sample(self):
        return self.counts_host
