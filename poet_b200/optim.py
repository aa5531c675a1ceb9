"""Fused optimizer step for the PoET hot path.

Reference behaviour (engine.py:77-81, main.py:253-277): `clip_grad_norm_(model.parameters(), max_norm)` followed
by `torch.optim.AdamW(param_dicts, lr, weight_decay).step()` with three learning-rate groups selected by name
(`lr_backbone_names`, `lr_linear_proj_names` = reference_points / sampling_offsets, everything else).

Here the gradients already live in ONE flat arena (`FlatGradReducer`, which is also the all-reduce buffer), so the
step is two launches with no host round trip: `poet_grad_sumsq_multi` (per-tensor squared norms) and
`poet_adamw_clip_multi` over a pointer table of all parameters.  The second kernel also writes the bf16 hi/lo planes
of the updated weight matrices into the model's `WeightPlanes` arena, so the next forward skips its split pass.

`FusedClipAdamW` is a `torch.optim.Optimizer`: `param_groups` are the reference's three groups in the reference's
order (so `torch.optim.lr_scheduler.StepLR` drives the learning rates, main.py:277), and `state_dict()` /
`load_state_dict()` use torch.optim.AdamW's format (per-parameter `step`, `exp_avg`, `exp_avg_sq`), so the
`optimizer` entry of a reference checkpoint (main.py:302,362) resumes here and vice versa.

    reducer = FlatGradReducer(model.parameters())
    opt = FusedClipAdamW(model, reducer, lr=2e-4, weight_decay=1e-4, max_norm=0.1,
                         lr_backbone=2e-5, lr_linear_proj_mult=0.1)
    sched = torch.optim.lr_scheduler.StepLR(opt, lr_drop)
    ... forward / opt.zero_grad() / backward / reducer.all_reduce() ...
    opt.step()

Parameters that receive no gradient in a step (their arena slice is exactly zero: `transformer.reference_points.*`
in bbox mode, the intermediate heads with aux_loss=False) are skipped by the kernel -- no weight decay, no moment
decay -- which is what torch.optim.AdamW does for parameters whose .grad is None.
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


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, model: torch.nn.Module, reducer: FlatGradReducer, lr: float = 2e-4, weight_decay: float = 1e-4,
                 max_norm: float = 0.1, betas=(0.9, 0.999), eps: float = 1e-8, lr_backbone: float = 2e-5,
                 lr_linear_proj_mult: float = 0.1, lr_backbone_names: Iterable[str] = ("backbone.0",),
                 lr_linear_proj_names: Iterable[str] = ("reference_points", "sampling_offsets"),
                 skip: Iterable[str] = (), emit_weight_planes: bool = True):
        """`skip`: name substrings of parameters to leave out of the step altogether (parameters that merely receive
        no gradient need not be listed: the kernel skips every tensor whose gradient is exactly zero)."""
        self.reducer = reducer
        self.max_norm = max_norm
        self.step_count = 0
        bb, lp, skip = tuple(lr_backbone_names), tuple(lr_linear_proj_names), tuple(skip)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        groups = [                                                            # main.py:253-271, same order
            {"params": [p for n, p in named if not _match(n, bb) and not _match(n, lp)], "lr": lr},
            {"params": [p for n, p in named if _match(n, bb)], "lr": lr_backbone},
            {"params": [p for n, p in named if _match(n, lp)], "lr": lr * lr_linear_proj_mult},
        ]
        super().__init__(groups, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        group_of = {id(p): gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
        names = {id(p): n for n, p in named}
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
        raw, chunk = bytearray(), 0
        self.entries = []                                                     # (param, arena offset) per table row
        for p, off in zip(reducer.params, reducer.offsets):
            name = names.get(id(p), "")
            if id(p) not in group_of or (skip and _match(name, skip)):
                continue
            if not p.is_contiguous() or p.data_ptr() % 16 or off % 4:
                raise ValueError(f"parameter {name} is not a contiguous 16-byte aligned tensor")
            hi = lo = 0
            if self.planes is not None and p.data_ptr() in plane_off:
                po = plane_off[p.data_ptr()]
                hi = self.planes.hi.data_ptr() + 2 * po
                lo = (self.planes.lo.data_ptr() + 2 * po) if self.planes.with_lo else 0
            raw += struct.pack("<QqQQqqii", p.data_ptr(), off // 4, hi, lo, p.numel(), chunk, group_of[id(p)], 0)
            chunk += ((p.numel() + 3) // 4 + 1023) // 1024
            self.entries.append((p, off))
        self.n_tensors = len(self.entries)
        self.table = torch.frombuffer(raw, dtype=torch.uint8).clone().to(dev)
        self.chunks = chunk
        self.tensor_sumsq = torch.zeros(self.n_tensors, device=dev, dtype=torch.float32)
        self.touched = torch.zeros(self.n_tensors, device=dev, dtype=torch.float32)   # > 0: the tensor owns Adam state

    # ---- torch.optim.Optimizer surface ------------------------------------------------------------
    @property
    def lrs(self):
        return [float(g["lr"]) for g in self.param_groups]

    def set_lr(self, lr: float, lr_backbone: float, lr_linear_proj: float) -> None:
        for g, v in zip(self.param_groups, (lr, lr_backbone, lr_linear_proj)):
            g["lr"] = v

    def zero_grad(self, set_to_none: bool = True) -> None:
        """engine.py:75: the gradients are views of the reducer's arena; 'none' would unbind them, so the arena is
        zero-filled instead (one memset)."""
        self.reducer.zero()

    @torch.no_grad()
    def step(self, closure=None):
        """clip (global L2 norm over the whole arena) + AdamW, on the current stream; the gradient arena is read only."""
        loss = closure() if closure is not None else None
        self.step_count += 1
        flat = self.reducer.flat
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        g0 = self.param_groups[0]
        lr_arr = (C.c_float * 3)(*self.lrs)
        ops._call("poet_grad_sumsq_multi", self.table.data_ptr(), self.n_tensors, self.chunks, flat.data_ptr(),
                  self.tensor_sumsq.data_ptr(), stream)
        ops._call("poet_adamw_clip_multi", self.table.data_ptr(), self.n_tensors, self.chunks, flat.data_ptr(),
                  self.m.data_ptr(), self.v.data_ptr(), self.sumsq.data_ptr(), self.tensor_sumsq.data_ptr(),
                  self.touched.data_ptr(), float(self.max_norm), lr_arr, 3, float(g0["betas"][0]), float(g0["betas"][1]),
                  float(g0["eps"]), float(g0["weight_decay"]), self.step_count, stream)
        if self.planes is not None:
            self.planes.mark_fresh()
        return loss

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm of the last step() (device scalar; what engine.py logs as grad_norm)."""
        return self.sumsq.sqrt().float()

    # ---- checkpoint format of torch.optim.AdamW (reference main.py:302, 357-369) -----------------
    def _index_of(self):
        """torch numbers parameters consecutively over the groups."""
        idx, k = {}, 0
        for g in self.param_groups:
            for p in g["params"]:
                idx[id(p)] = k
                k += 1
        return idx

    def state_dict(self):
        idx = self._index_of()
        touched = self.touched.cpu()
        state = {}
        for t, (p, off) in enumerate(self.entries):
            if float(touched[t]) == 0.0:
                continue                                                   # never received a gradient: no state, like torch
            n = p.numel()
            state[idx[id(p)]] = {"step": torch.tensor(float(self.step_count)),
                                 "exp_avg": self.m[off:off + n].view_as(p).clone(),
                                 "exp_avg_sq": self.v[off:off + n].view_as(p).clone()}
        groups, k = [], 0
        for g in self.param_groups:
            d = {key: val for key, val in g.items() if key != "params"}
            d["params"] = list(range(k, k + len(g["params"])))
            k += len(g["params"])
            groups.append(d)
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, sd) -> None:
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"])
                                                         for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has different parameter groups (reference layout: main.py:253-271)")
        for dst, src in zip(self.param_groups, groups):
            for key, val in src.items():
                if key != "params":
                    dst[key] = val
        by_idx = {i: (t, p, off) for t, (p, off) in enumerate(self.entries) for i in [self._index_of()[id(p)]]}
        self.m.zero_()
        self.v.zero_()
        touched = torch.zeros(self.n_tensors, dtype=torch.float32)
        steps = []
        for i, st in sd["state"].items():
            hit = by_idx.get(int(i))
            if hit is None:
                continue
            t, p, off = hit
            n = p.numel()
            self.m[off:off + n].copy_(st["exp_avg"].reshape(-1).to(self.m.device, torch.float32))
            self.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1).to(self.v.device, torch.float32))
            touched[t] = 1.0
            steps.append(int(float(st["step"])))
        self.touched.copy_(touched)
        # one step counter for the whole model: every parameter that has state was updated at every step
        self.step_count = max(steps) if steps else 0
