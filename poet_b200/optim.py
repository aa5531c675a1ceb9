"""Fused optimizer step for the PoET hot path.

Reference behaviour (engine.py:77-81, main.py:253-277): `clip_grad_norm_(model.parameters(), max_norm)` followed
by `torch.optim.AdamW(param_dicts, lr, weight_decay).step()` with three learning-rate groups selected by name
(`lr_backbone_names`, `lr_linear_proj_names` = reference_points / sampling_offsets, everything else).

Here the gradients already live in ONE flat arena (`FlatGradReducer`, which is also the all-reduce buffer), so the
step is two launches with no host round trip: `poet_sumsq` over the arena and `poet_adamw_clip_multi` over a
pointer table of all parameters.  The second kernel also writes the bf16 hi/lo planes of the updated weight
matrices into the model's `WeightPlanes` arena, so the next forward skips its split pass.

    reducer = FlatGradReducer(model.parameters())
    opt = FusedClipAdamW(model, reducer, lr=2e-4, weight_decay=1e-4, max_norm=0.1,
                         lr_backbone=2e-5, lr_linear_proj_mult=0.1)
    ... forward / backward / reducer.all_reduce() ...
    opt.step()
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Iterable, Optional, Sequence

import torch

from . import _lib, ops
from .data_parallel import FlatGradReducer


def _match(name: str, keywords: Sequence[str]) -> bool:          # main.py:241-248 match_name_keywords
    return any(k in name for k in keywords)


class FusedClipAdamW:
    def __init__(self, model: torch.nn.Module, reducer: FlatGradReducer, lr: float = 2e-4, weight_decay: float = 1e-4,
                 max_norm: float = 0.1, betas=(0.9, 0.999), eps: float = 1e-8, lr_backbone: float = 2e-5,
                 lr_linear_proj_mult: float = 0.1, lr_backbone_names: Iterable[str] = ("backbone.0",),
                 lr_linear_proj_names: Iterable[str] = ("reference_points", "sampling_offsets"),
                 skip: Iterable[str] = ("transformer.reference_points",), emit_weight_planes: bool = True):
        """`skip`: parameters that never receive a gradient on this path (the reference leaves their .grad at None, so
        torch's AdamW does not touch them; in the flat arena they are zeros and must not be weight-decayed)."""
        self.reducer = reducer
        self.lrs = [lr, lr_backbone, lr * lr_linear_proj_mult]                # same order as main.py's param_dicts
        self.weight_decay, self.max_norm, self.betas, self.eps = weight_decay, max_norm, betas, eps
        self.step_count = 0
        names = {id(p): n for n, p in model.named_parameters()}
        dev = reducer.flat.device
        self.m = torch.zeros_like(reducer.flat)
        self.v = torch.zeros_like(reducer.flat)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.planes: Optional[ops.WeightPlanes] = None
        if emit_weight_planes and ops.get_gemm_precision() != "fp32":
            self.planes = getattr(model, "_poet_weight_planes", None) or ops.WeightPlanes(model)
            object.__setattr__(model, "_poet_weight_planes", self.planes)
            self.planes.refresh()        # parameters the optimizer skips never change: their planes are written here, once
        plane_off = {}
        if self.planes is not None:
            for base, _nbytes, off in self.planes.ranges:
                plane_off[base] = off
        raw, chunk, self.n_tensors = bytearray(), 0, 0
        for p, off in zip(reducer.params, reducer.offsets):
            name = names.get(id(p), "")
            if _match(name, tuple(skip)):
                continue
            if not p.is_contiguous() or p.data_ptr() % 16 or off % 4:
                raise ValueError(f"parameter {name} is not a contiguous 16-byte aligned tensor")
            group = 1 if _match(name, tuple(lr_backbone_names)) else 2 if _match(name, tuple(lr_linear_proj_names)) else 0
            hi = lo = 0
            if self.planes is not None and p.data_ptr() in plane_off:
                po = plane_off[p.data_ptr()]
                hi = self.planes.hi.data_ptr() + 2 * po
                lo = (self.planes.lo.data_ptr() + 2 * po) if self.planes.with_lo else 0
            raw += struct.pack("<QqQQqqii", p.data_ptr(), off // 4, hi, lo, p.numel(), chunk, group, 0)
            chunk += ((p.numel() + 3) // 4 + 1023) // 1024
            self.n_tensors += 1
        self.table = torch.frombuffer(raw, dtype=torch.uint8).clone().to(dev)
        self.chunks = chunk
        self._lr_arr = (C.c_float * 3)(*self.lrs)

    def set_lr(self, lr: float, lr_backbone: float, lr_linear_proj: float) -> None:
        """StepLR etc.: the learning rates are plain kernel arguments."""
        self.lrs = [lr, lr_backbone, lr_linear_proj]
        self._lr_arr = (C.c_float * 3)(*self.lrs)

    @torch.no_grad()
    def step(self) -> None:
        """clip (global L2 norm over the whole arena) + AdamW, on the current stream; the gradient arena is read only."""
        self.step_count += 1
        flat = self.reducer.flat
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        if self.max_norm > 0:
            ops._call("poet_sumsq", flat.data_ptr(), flat.numel(), self.sumsq.data_ptr(), stream)
        ops._call("poet_adamw_clip_multi", self.table.data_ptr(), self.n_tensors, self.chunks, flat.data_ptr(),
                  self.m.data_ptr(), self.v.data_ptr(), self.sumsq.data_ptr(), float(self.max_norm), self._lr_arr, 3,
                  float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                  self.step_count, stream)
        if self.planes is not None:
            self.planes.mark_fresh()

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm of the last step() (device scalar; what engine.py logs as grad_norm)."""
        return self.sumsq.sqrt().float()
