"""B200-native DeformableTransformer behind the reference's nn.Module surface.

Mirror of reference models/deformable_transformer.py (class names, constructor signatures,
attribute / state_dict names, forward signatures and return tuple) so it can replace
``models.deformable_transformer`` (seam B-py2); every numeric step is a
libpoet_b200 kernel via ``poet_b200.ops``:

  flatten + level_embed        :124-140  -> ops.flatten_levels           (poet_nchw_to_tokens)
  encoder reference points     :217-230  -> ops.enc_reference_points
  encoder layer                :199-208  -> MSDeformAttn + ops.add_layernorm + ops.mlp
  decoder layer                :275-292  -> in_proj GEMMs + ops.mha_smallq + MSDeformAttn + LN + FFN
  next-layer query (x + pos)             -> second output of the LayerNorm kernel (fused)

Dropout (reference :178,184,186,249,253-254,260,262 and the attention-probability dropout inside
nn.MultiheadAttention; default 0.1, main.py:94) is counter-based and fused into the kernels that own the
dropped tensors (the LayerNorm kernels for the residual branches, the FFN1 GEMM epilogue for the hidden
activation, the attention kernel for the probabilities): active in train() with dropout > 0, the identity in
eval() or with dropout = 0 -- the parity path (PyTorch's Philox stream cannot be reproduced, so
train-mode checks are statistical).
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import ops
from .deformable_attention import MSDeformAttn


def _clones(module: nn.Module, n: int) -> nn.ModuleList:
    return nn.ModuleList(copy.deepcopy(module) for _ in range(n))


_shape_cache = {}


def _shape_tensors(shapes, device):
    """spatial_shapes / level_start_index device tensors of the reference API, built once per
    (pyramid, device): a host->device copy per forward would also break CUDA-graph capture.  The
    host copy rides along as `_poet_host`, so no kernel launch ever needs a device->host sync."""
    key = (shapes, str(device))
    if key not in _shape_cache:
        ss = torch.as_tensor(shapes, dtype=torch.long, device=device)
        ss._poet_host = shapes
        ls = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
        _shape_cache[key] = (ss, ls)
    return _shape_cache[key]


def _p_drop(mod: nn.Module) -> float:
    """Dropout probability in effect: the layer's `dropout` in train(), 0 in eval()."""
    return float(mod.p_drop) if mod.training else 0.0


# dropout site ids: (layer base) + (site within the layer); bases are assigned by DeformableTransformer.__init__
_SITE_ATTN_PROB, _SITE_D1, _SITE_D2, _SITE_HIDDEN, _SITE_D4 = 0, 1, 2, 3, 5


class _SelfAttentionParams(nn.Module):
    """Parameter container with nn.MultiheadAttention's state_dict layout
    (in_proj_weight, in_proj_bias, out_proj.{weight,bias})."""

    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    def forward(self, qk_in: torch.Tensor, v_in: torch.Tensor, drop_site: int = 0, out_bias_grad_elsewhere: bool = False) -> torch.Tensor:
        """qk_in = tgt + query_pos, v_in = tgt, both [B,Q,C] (batch-first) -> attention output [B,Q,C]."""
        C = self.embed_dim
        qk, v = ops.in_proj_qk_v(qk_in, v_in, self.in_proj_weight, self.in_proj_bias)
        o = ops.mha_smallq(qk, v, self.num_heads, drop_p=float(self.dropout) if self.training else 0.0, drop_site=drop_site)
        return ops.linear(o, self.out_proj.weight, self.out_proj.bias, bias_grad_elsewhere=out_bias_grad_elsewhere)


class DeformableTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("poet_b200 implements the 'relu' FFN used by every PoET config")
        self.p_drop = dropout
        self.site_base = 0x100                     # re-assigned per layer by DeformableTransformer
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None,
                query=None, emit_next_query=False):
        """Reference signature plus two optional arguments used by our encoder stack: `query`
        (= src + pos, produced by the previous layer's LayerNorm kernel) and `emit_next_query`."""
        p, sb = _p_drop(self), self.site_base
        if query is None:
            query = src if pos is None else ops.add_tensors(src, pos)
        attn = self.self_attn(query, reference_points, src, spatial_shapes, level_start_index, padding_mask,
                              output_bias_grad_elsewhere=True)
        # norm1's result has two readers (linear1 and the residual of norm2): two autograd handles, their gradients arrive
        # at norm1's backward kernel as dy / dy2 (its two-pointer form) instead of through a 26 MB accumulation kernel
        src, src_mlp = ops.add_layernorm(src, attn, self.norm1.weight, self.norm1.bias, eps=self.norm1.eps,
                                         drop_p=p, drop_site=sb + _SITE_D1, r_bias=self.self_attn.output_proj.bias,
                                         n_alias=1)                                                   # dropout1
        # linear1 / relu / dropout2 / linear2 / dropout3 / norm2 as one autograd node (ops._FFNBlock)
        return ops.ffn_block(src, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                             self.norm2.weight, self.norm2.bias, pos=pos if (emit_next_query and pos is not None) else None,
                             eps=self.norm2.eps, drop_p=p, site_hidden=sb + _SITE_HIDDEN, site_res=sb + _SITE_D2,
                             x_mlp=src_mlp)


class DeformableTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.gemm_precision = None      # None: the global ops precision; 'bf16': single-pass MMAs (throughput mode, cfg4)

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device=None):
        from .deformable_attention import host_shapes
        return ops.enc_reference_points(valid_ratios.float().contiguous(), host_shapes(spatial_shapes))

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        ref = self.get_reference_points(spatial_shapes, valid_ratios, device=src.device)
        out, query = src, None
        with ops.precision_scope(self.gemm_precision):
            for i, layer in enumerate(self.layers):
                last = i == self.num_layers - 1
                out = ops.grad_marker(out, ("encoder", i))       # backward: layers > i have issued all their gradient kernels
                res = layer(out, pos, ref, spatial_shapes, level_start_index, padding_mask, query=query,
                            emit_next_query=not last)
                out, query = res if isinstance(res, tuple) else (res, None)
        return out


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("poet_b200 implements the 'relu' FFN used by every PoET config")
        self.p_drop = dropout
        self.site_base = 0x10100                   # re-assigned per layer by DeformableTransformer
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = _SelfAttentionParams(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index,
                src_padding_mask=None, value=None, handles=None, fan_out: int = 0, emit_next_query: bool = False,
                value_grad_buf=None):
        """Reference signature plus optional arguments used by our decoder stack:
        `handles` = (q, tgt_v, tgt_res): the layer input as separate autograd handles for its three readers -- the
        self-attention query/key input (ALREADY tgt + query_pos), the value input and the residual of norm2 -- so that their
        gradients reach the producing LayerNorm backward as separate pointers instead of through accumulation kernels;
        `fan_out` / `emit_next_query`: return (out, q_next, alias_1 .. alias_fan_out) with q_next = out + query_pos (or None)
        produced by the last LayerNorm kernel."""
        p, sb = _p_drop(self), self.site_base
        if tgt.shape[1] > 32:
            raise NotImplementedError("decoder self-attention kernel supports at most 32 object queries")
        if handles is None:
            q = tgt if query_pos is None else ops.add_tensors(tgt, query_pos)
            tgt_v = tgt_res = tgt
        else:
            q, tgt_v, tgt_res = handles
        sa = self.self_attn(q, tgt_v, drop_site=sb + _SITE_ATTN_PROB, out_bias_grad_elsewhere=True)   # MHA(dropout=p)
        ob = self.self_attn.out_proj.bias
        if query_pos is not None:
            tgt, q2 = ops.add_layernorm(tgt_res, sa, self.norm2.weight, self.norm2.bias, pos=query_pos, eps=self.norm2.eps,
                                        drop_p=p, drop_site=sb + _SITE_D2, r_bias=ob)                # dropout2
        else:
            tgt = ops.add_layernorm(tgt_res, sa, self.norm2.weight, self.norm2.bias, eps=self.norm2.eps,
                                    drop_p=p, drop_site=sb + _SITE_D2, r_bias=ob)
            q2 = tgt
        ca = self.cross_attn(q2, reference_points, src, src_spatial_shapes, level_start_index, src_padding_mask,
                             value=value, output_bias_grad_elsewhere=True, value_grad_buf=value_grad_buf)
        # norm1's result has two readers (linear1 and the residual of norm3): two handles, gradients summed in norm1's backward
        tgt, tgt_mlp = ops.add_layernorm(tgt, ca, self.norm1.weight, self.norm1.bias, eps=self.norm1.eps,
                                         drop_p=p, drop_site=sb + _SITE_D1, r_bias=self.cross_attn.output_proj.bias,
                                         n_alias=1)                                                   # dropout1
        # linear1 / relu / dropout3 / linear2 / dropout4 / norm3 as one autograd node
        want_q = emit_next_query and query_pos is not None
        res = ops.ffn_block(tgt, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                            self.norm3.weight, self.norm3.bias, eps=self.norm3.eps, drop_p=p,
                            site_hidden=sb + _SITE_HIDDEN, site_res=sb + _SITE_D4, x_mlp=tgt_mlp,
                            pos=query_pos if want_q else None, n_alias=fan_out)
        if handles is None and fan_out == 0 and not emit_next_query:
            return res
        res = res if isinstance(res, tuple) else (res,)
        out = res[0]
        q_next = res[1] if want_q else None
        return (out, q_next) + tuple(res[2 if want_q else 1:])


class DeformableTransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, return_intermediate=False):
        super().__init__()
        self.layers = _clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.return_intermediate = return_intermediate
        self.bbox_embed = None      # PoET never refines boxes (reference :302,:321)
        self.class_embed = None
        self.value_gemm_precision = None  # precision of the [B*S, d] x [d, d] value projections of `memory` (throughput mode)

    def forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index, src_valid_ratios,
                query_pos=None, src_padding_mask=None, layer_callback=None):
        """Reference signature; `layer_callback(i, out_i)` (ours, optional) is invoked as soon as decoder layer i
        is issued, so the caller can start that layer's pose heads without waiting for the whole stack."""
        if self.bbox_embed is not None:
            raise NotImplementedError("iterative box refinement is not part of PoET")
        if reference_points.shape[-1] != 2:
            raise NotImplementedError("only 2-d reference points (PoET 'bbox' mode) are implemented")
        ref_in = (reference_points[:, :, None] * src_valid_ratios[:, None]).contiguous()     # [B,Q,L,2]
        # `src` (the encoder memory) is the same for every layer: project it for all layers up front on a
        # side stream, so these large GEMMs (and their dgrad/wgrad in backward) overlap the launch-bound
        # self-attention / FFN chain of the queries instead of sitting on its critical path
        values = [None] * len(self.layers)
        gv_bufs = [None] * len(self.layers)
        forked, marks = None, []
        if ops.parallel_streams_enabled() and src.is_cuda:
            forked = ops.fork(0, src.device)
            forked.uses(src, src_padding_mask)
            with forked, ops.precision_scope(self.value_gemm_precision), ops.background_gemms():
                for i, layer in enumerate(self.layers):
                    values[i] = layer.cross_attn.project_value(src, src_padding_mask)
                    gv_bufs[i] = ops.grad_value_buffer(values[i])       # zero-filled here, off the backward's dependent chain
                    marks.append(forked.checkpoint())
        out, inter, inter_ref = tgt, [], []
        # the output of layer i is read by the pose heads / the intermediate stack (main handle), and by layer i+1 three
        # times (query = out + query_pos from the LayerNorm kernel itself, value input, residual): one autograd handle per
        # reader, so the backward has no accumulation kernels between the layers
        handles = None
        n_layers = len(self.layers)
        for i, layer in enumerate(self.layers):
            if forked is not None:
                forked.wait(marks[i], values[i], gv_bufs[i])
            last = i == n_layers - 1
            res = layer(out, query_pos, ref_in, src, src_spatial_shapes, src_level_start_index, src_padding_mask,
                        value=values[i], handles=handles, fan_out=0 if last else 2, emit_next_query=not last,
                        value_grad_buf=gv_bufs[i])
            if last:
                out, handles = (res[0] if isinstance(res, tuple) else res), None
            else:
                out, q_next, a_v, a_res = res
                handles = ((q_next if q_next is not None else a_v), a_v, a_res)
                if q_next is None:                            # no query_pos: the query input is one more reader of out
                    handles = (out, a_v, a_res)
            if layer_callback is not None:
                layer_callback(i, out)
            if self.return_intermediate:
                inter.append(out)
                inter_ref.append(reference_points)
        if self.return_intermediate:
            return torch.stack(inter), torch.stack(inter_ref)
        return out, reference_points


class DeformableTransformer(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=1024,
                 dropout=0.1, activation="relu", return_intermediate_dec=False, num_feature_levels=4,
                 dec_n_points=4, enc_n_points=4):
        super().__init__()
        self.d_model, self.nhead = d_model, nhead
        self.encoder = DeformableTransformerEncoder(
            DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                              nhead, enc_n_points), num_encoder_layers)
        self.decoder = DeformableTransformerDecoder(
            DeformableTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                              nhead, dec_n_points), num_decoder_layers, return_intermediate_dec)
        self.level_embed = nn.Parameter(torch.empty(num_feature_levels, d_model))
        # unused when reference points come from boxes (reference :157-158) but part of the checkpoint format
        self.reference_points = nn.Linear(d_model, 2)
        self.dropout = dropout
        for i, layer in enumerate(self.encoder.layers):            # distinct dropout sites per layer (ops: counter-based masks)
            layer.site_base = 0x100 * (i + 1)
        for i, layer in enumerate(self.decoder.layers):
            layer.site_base = 0x10000 + 0x100 * (i + 1)
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.xavier_uniform_(self.reference_points.weight, gain=1.0)
        nn.init.zeros_(self.reference_points.bias)
        nn.init.normal_(self.level_embed)

    def set_throughput_mode(self, on: bool = True) -> None:
        """BASELINE.json cfg4 (bf16 training mode; the reference itself has no mixed precision: SURVEY.md section 5): the
        token-row GEMMs -- every encoder layer and the decoder's value projections of `memory`, 97 % of the FLOPs --
        issue single-pass bf16 tcgen05 MMAs (fp32 accumulation, fp32 activations in HBM); the query-row GEMMs of the
        decoder and the pose heads stay split-bf16 (fp32-grade).  Tolerance of this mode: tests/test_gpu_model.py."""
        self.encoder.gemm_precision = "bf16" if on else None
        self.decoder.value_gemm_precision = "bf16" if on else None

    @staticmethod
    def get_valid_ratio(mask: torch.Tensor) -> torch.Tensor:
        _, H, W = mask.shape
        vh = (~mask[:, :, 0]).sum(1).float() / H
        vw = (~mask[:, 0, :]).sum(1).float() / W
        return torch.stack((vw, vh), -1)

    def forward(self, srcs, masks, pos_embeds, query_embed=None, reference_points=None, pos_tokens=None,
                layer_callback=None, src_tokens=None):
        # all weight matrices -> bf16 hi/lo planes in one launch (no-op if an enclosing module already did it)
        with ops.planes_scope(self):
            return self._forward_impl(srcs, masks, pos_embeds, query_embed, reference_points, pos_tokens, layer_callback,
                                      src_tokens)

    def _forward_impl(self, srcs, masks, pos_embeds, query_embed=None, reference_points=None, pos_tokens=None,
                      layer_callback=None, src_tokens=None):
        """Reference signature (deformable_transformer.py:120).  `pos_tokens` (optional, ours):
        lvl_pos_embed_flatten [B,S,C] already in token layout with level_embed added."""
        if query_embed is None:
            raise ValueError("query_embed is required")
        if self.training and self.dropout > 0.0 and masks[0].is_cuda:
            ops.begin_dropout_forward(masks[0].device)       # this forward's dropout seed (bumped in place: graph-safe)
        if reference_points is None:
            raise NotImplementedError("learned reference points are not used by PoET ('bbox' mode only)")
        if src_tokens is not None:               # ours: the pyramid already in token layout (ops.input_proj_tokens)
            shapes = tuple((int(m.shape[1]), int(m.shape[2])) for m in masks)
            src = src_tokens
        else:
            shapes = tuple((int(s.shape[2]), int(s.shape[3])) for s in srcs)
            src = ops.flatten_levels(list(srcs))
        pos = pos_tokens if pos_tokens is not None else ops.flatten_levels(list(pos_embeds), self.level_embed)
        spatial_shapes, level_start = _shape_tensors(shapes, src.device)
        if src.is_cuda and len(masks) <= 4:
            pad, valid_ratios = ops.mask_prep(masks)          # flattened padding mask + valid ratios: one launch
        else:
            mask = torch.cat([m.flatten(1) for m in masks], 1)
            valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
            pad = mask.to(torch.uint8)          # converted once; every MSDeformAttn layer reuses it

        memory = self.encoder(src, spatial_shapes, level_start, valid_ratios, pos, pad)

        memory = ops.grad_marker(memory, ("decoder", 0))         # backward: decoder + heads have issued all their gradient kernels
        C = memory.shape[2]
        if query_embed.dim() == 2:
            query_embed = query_embed.unsqueeze(0).expand(memory.shape[0], -1, -1)
        query_pos = query_embed[..., :C].contiguous()
        tgt = query_embed[..., C:].contiguous()
        hs, inter_refs = self.decoder(tgt, reference_points, memory, spatial_shapes, level_start, valid_ratios,
                                      query_pos, pad, layer_callback=layer_callback)
        return hs, reference_points, inter_refs, None, None


def build_deforamble_transformer(args):      # (sic) the reference spells it this way, :358
    return DeformableTransformer(
        d_model=args.hidden_dim, nhead=args.nheads, num_encoder_layers=args.enc_layers,
        num_decoder_layers=args.dec_layers, dim_feedforward=args.dim_feedforward, dropout=args.dropout,
        activation="relu", return_intermediate_dec=True, num_feature_levels=args.num_feature_levels,
        dec_n_points=args.dec_n_points, enc_n_points=args.enc_n_points)
