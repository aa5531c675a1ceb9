"""Drop-in for the reference's external dependency ``from deformable_attention import MSDeformAttn``
(reference models/deformable_transformer.py:24; constructed :177,:248; called :201,:283-285;
re-initialised :58-59).  Seam B-py1 of SURVEY.md §8b: registering this module as
``sys.modules['deformable_attention']`` makes the *unmodified* reference model run on the
poet_b200 CUDA kernels.

Sub-module names (sampling_offsets, attention_weights, value_proj, output_proj) are load-bearing:
checkpoint keys and the lr-group substring match of reference main.py:41,267-269.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops


def host_shapes(spatial_shapes) -> Tuple[Tuple[int, int], ...]:
    """(H_l, W_l) per level as host ints.  A tensor costs one D2H sync unless it carries the
    `_poet_host` attribute our own DeformableTransformer attaches."""
    if isinstance(spatial_shapes, torch.Tensor):
        cached = getattr(spatial_shapes, "_poet_host", None)
        if cached is not None:
            return cached
        return tuple((int(h), int(w)) for h, w in spatial_shapes.tolist())
    return tuple((int(h), int(w)) for h, w in spatial_shapes)


class MSDeformAttn(nn.Module):
    def __init__(self, d_model: int = 256, n_levels: int = 4, n_heads: int = 8, n_points: int = 4):
        super().__init__()
        if d_model % n_heads:
            raise ValueError(f"d_model ({d_model}) must be divisible by n_heads ({n_heads})")
        self.im2col_step = 64                 # kept for attribute compatibility; the kernels do not chunk
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    @torch.no_grad()
    def _reset_parameters(self) -> None:
        M, L, P = self.n_heads, self.n_levels, self.n_points
        self.sampling_offsets.weight.zero_()
        angle = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
        ray = torch.stack((angle.cos(), angle.sin()), -1)
        ray = ray / ray.abs().max(-1, keepdim=True)[0]                    # onto the unit square
        step = torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, P, 1)
        self.sampling_offsets.bias.copy_((ray.view(M, 1, 1, 2).expand(M, L, P, 2) * step).reshape(-1))
        self.attention_weights.weight.zero_()
        self.attention_weights.bias.zero_()
        nn.init.xavier_uniform_(self.value_proj.weight)
        self.value_proj.bias.zero_()
        nn.init.xavier_uniform_(self.output_proj.weight)
        self.output_proj.bias.zero_()

    @staticmethod
    def _mask_u8(input_padding_mask):
        if input_padding_mask is None:
            return None
        m = input_padding_mask.reshape(-1)
        return m if m.dtype == torch.uint8 else m.to(torch.uint8)

    def project_value(self, input_flatten: torch.Tensor, input_padding_mask: Optional[torch.Tensor] = None):
        """value_proj(input_flatten) with padded rows zeroed.  Exposed so that a caller whose keys do not
        depend on the queries (the decoder: `memory` is fixed) can issue it early / on another stream."""
        # value feeds only the gather kernel, whose backward hands us a private grad buffer
        return ops.linear(input_flatten, self.value_proj.weight, self.value_proj.bias,
                          row_mask=self._mask_u8(input_padding_mask), mask_grad_inplace=True)

    def forward(self, query: torch.Tensor, reference_points: torch.Tensor, input_flatten: torch.Tensor,
                input_spatial_shapes, input_level_start_index=None,
                input_padding_mask: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None,
                output_bias_grad_elsewhere: bool = False, value_grad_buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        """query [N,Lq,C]; reference_points [N,Lq,L,2] in [0,1]; input_flatten [N,S,C];
        input_spatial_shapes [(H_l,W_l)]; input_padding_mask [N,S] True = padded  ->  [N,Lq,C].
        `value` (ours, optional): the result of project_value() computed beforehand.
        `output_bias_grad_elsewhere` (ours): the caller feeds the result to ops.add_layernorm(..., r_bias=output_proj.bias),
        whose backward kernel produces that bias gradient (no separate pass over the gradient rows).
        `value_grad_buf` (ours): zero-filled buffer for the gradient of `value` (ops.grad_value_buffer), filled by the caller
        next to project_value() so that the backward does not start with a 26 MB zero-fill in its dependent chain."""
        shapes = host_shapes(input_spatial_shapes)
        N, S, _ = input_flatten.shape
        if sum(h * w for h, w in shapes) != S:
            raise ValueError("spatial shapes do not add up to the flattened input length")
        if reference_points.shape[-1] != 2:
            raise NotImplementedError("only 2-d reference points (PoET 'bbox' mode) are implemented; "
                                      "4-d box references are reachable in the reference but in no PoET config")
        forked = None
        if value is None:
            if ops.parallel_streams_enabled() and input_flatten.is_cuda:
                # value projection and the [offsets | logits] projection are independent GEMMs
                forked = ops.fork(4, input_flatten.device)
                forked.uses(input_flatten, input_padding_mask)
                with forked:
                    value = self.project_value(input_flatten, input_padding_mask)
                    value_grad_buf = ops.grad_value_buffer(value)
                    mark = forked.checkpoint()
            else:
                value = self.project_value(input_flatten, input_padding_mask)
                value_grad_buf = ops.grad_value_buffer(value)
        # one projection for [offsets | logits]: the gather kernel reads both out of the same row
        oa = ops.proj_cat(query, self.sampling_offsets.weight, self.sampling_offsets.bias,
                          self.attention_weights.weight, self.attention_weights.bias)
        if forked is not None:
            forked.wait(mark, value, value_grad_buf)
        out = ops.msda_block(value, oa, reference_points, shapes, self.n_heads, self.n_levels, self.n_points,
                             grad_value_buf=value_grad_buf)
        return ops.linear(out, self.output_proj.weight, self.output_proj.bias, bias_grad_elsewhere=output_bias_grad_elsewhere)


def ms_deform_attn_core(value: torch.Tensor, spatial_shapes, sampling_locations: torch.Tensor,
                        attention_weights: torch.Tensor) -> torch.Tensor:
    """Functional core with upstream MSDeformAttnFunction semantics: value [N,S,M,D],
    sampling_locations [N,Lq,M,L,P,2], attention_weights [N,Lq,M,L,P] -> [N,Lq,M*D]."""
    return ops.msda_core(value, host_shapes(spatial_shapes), sampling_locations, attention_weights)
