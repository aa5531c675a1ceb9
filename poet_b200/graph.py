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
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch

from .data_parallel import FlatGradReducer


class GraphedStep:
    def __init__(self, model, loss_fn: Callable, srcs: Sequence[torch.Tensor], masks: Sequence[torch.Tensor],
                 boxes, labels, reducer: Optional[FlatGradReducer] = None, warmup: int = 3, backward: bool = True):
        dev = next(model.parameters()).device
        self.model, self.loss_fn, self.backward = model, loss_fn, backward
        self.reducer = reducer if reducer is not None else (FlatGradReducer(model.parameters()) if backward else None)
        self.s_srcs = [torch.empty(s.shape, dtype=torch.float32, device=dev) for s in srcs]
        self.s_masks = [torch.empty(m.shape, dtype=torch.bool, device=dev) for m in masks]
        B, Q = srcs[0].shape[0], model.n_queries
        self.s_boxes = torch.empty((B, Q, 4), dtype=torch.float32, device=dev)
        self.s_classes = torch.empty((B, Q), dtype=torch.int64, device=dev)
        self.s_counts = torch.empty((B,), dtype=torch.int32, device=dev)
        self.counts_host = None
        self._copy_in(srcs, masks, boxes, labels)

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
        ops.clear_weight_split_cache()          # the bf16 weight planes must be re-derived inside the graph
        if self.reducer is not None:
            self.reducer.zero()
        out = self.model.forward_padded(self.s_srcs, self.s_masks, self.s_boxes, self.s_classes, self.s_counts)
        loss = self.loss_fn(out)
        if self.backward:
            loss.backward()
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

    def run(self, srcs=None, masks=None, boxes=None, labels=None):
        """Copy new inputs (if given) into the static buffers and replay.  Returns (loss, out_dict);
        both alias static graph memory: read them before the next run()."""
        if srcs is not None:
            self._copy_in(srcs, masks, boxes, labels)
        self.graph.replay()
        return self.loss, self.out

    def n_boxes_per_sample(self):
        return self.counts_host
