"""CPU oracle for the PoET deformable encoder/decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``poet_b200/`` may import this file; the only
callers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs.  The product path is the CUDA library and fails loudly
without it.

What it is: a functional, plain-PyTorch-on-CPU restatement (fp32 or fp64) of the
reference's algorithm for SURVEY.md §8 rows A0-A10, written over a flat ``dict`` of
tensors that uses the reference's ``state_dict`` key names.  It is *not* the reference
code: every function cites the reference lines whose arithmetic it follows.

The multi-scale deformable attention op itself is a third-party, un-vendored and
un-pinned dependency of the reference (``from deformable_attention import
MSDeformAttn``, reference ``models/deformable_transformer.py:24``; upstream
fundamentalvision/Deformable-DETR ``models/ops``, branch main).  Its published
algorithm (``ms_deform_attn_core_pytorch``: ``grid_sample(bilinear, zeros,
align_corners=False)`` on ``2*loc-1``; module = value_proj / sampling_offsets /
attention_weights+softmax / output_proj) is restated in ``msda_core`` and
``msda_module``; an identical restatement exists on disk in transformers 5.5
``modeling_deformable_detr.py:170-223,519-620`` and is used as a cross-check in
``tests/test_oracle.py``.

Pinning: the reference ships no tests or golden vectors for this path (SURVEY.md §4,
§8c) so the oracle is pinned against the reference itself: ``oracle/make_golden.py``
imports the unmodified reference ``models/`` package from ``/root/reference`` (with a
grid_sample ``deformable_attention`` shim for the missing third-party op), runs it on
seeded inputs/weights and commits the outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this file against those fixtures.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# A0  image position encoding                      reference models/position_encoding.py:40-60
# --------------------------------------------------------------------------------------
def sine_position_embedding(mask: Tensor, num_pos_feats: int = 128, temperature: float = 10000.0,
                            normalize: bool = True, scale: float = 2 * math.pi) -> Tensor:
    """mask [B,H,W] bool (True = padded) -> [B, 2*num_pos_feats, H, W] fp32.

    Channels [0,F) encode y, [F,2F) encode x; even channels sin, odd channels cos
    (position_encoding.py:56-59).  Always fp32, like the reference (cumsum dtype :45-46).
    """
    keep = ~mask
    ey = keep.cumsum(1, dtype=torch.float32)
    ex = keep.cumsum(2, dtype=torch.float32)
    if normalize:                                             # :47-50
        ey = (ey - 0.5) / (ey[:, -1:, :] + 1e-6) * scale
        ex = (ex - 0.5) / (ex[:, :, -1:] + 1e-6) * scale
    k = torch.arange(num_pos_feats, dtype=torch.float32)
    freq = temperature ** (2 * (k // 2) / num_pos_feats)       # :52-53

    def _interleave(arg: Tensor) -> Tensor:
        out = torch.empty_like(arg)
        out[..., 0::2] = arg[..., 0::2].sin()
        out[..., 1::2] = arg[..., 1::2].cos()
        return out

    py = _interleave(ey[..., None] / freq)
    px = _interleave(ex[..., None] / freq)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------------------
# A1  bounding-box embedding + dummy padding       position_encoding.py:71-84,
#                                                  pose_estimation_transformer.py:203-239
# --------------------------------------------------------------------------------------
def bbox_sine_embedding(boxes: Tensor, num_pos_feats: int = 32) -> Tensor:
    """boxes [n,4] (cx,cy,w,h) -> [n, 8*num_pos_feats]: per coord [sin(c*2^k) | cos(c*2^k)]."""
    pow2 = 2 ** torch.arange(num_pos_feats, dtype=torch.float32)
    parts = []
    for c in range(4):
        arg = boxes[:, c, None] * pow2
        parts.append(torch.cat((arg.sin(), arg.cos()), dim=-1))
    return torch.cat(parts, dim=-1)


def build_queries(boxes: Sequence[Tensor], labels: Sequence[Tensor], n_queries: int,
                  hidden_dim: int) -> Tuple[Tensor, Tensor, Tensor, List[int]]:
    """gt/jitter bbox mode of PoET.forward (pose_estimation_transformer.py:203-239, 307-309).

    Returns query_embeds [B,Q,2*hidden], pred_boxes [B,Q,4], pred_classes [B,Q] (int64),
    n_boxes_per_sample.  Dummies: embed -10, box -1, class -1.
    """
    q_all, b_all, c_all, n_all = [], [], [], []
    for bx, lb in zip(boxes, labels):
        n = bx.shape[0]
        n_all.append(n)
        emb = bbox_sine_embedding(bx.float(), hidden_dim // 8).repeat(1, 2)      # :217-219
        pad = n_queries - n
        if pad > 0:                                                              # :225-236
            bx = torch.cat((bx.float(), torch.full((pad, 4), -1.0)), 0)
            emb = torch.cat((emb, torch.full((pad, 2 * hidden_dim), -10.0)), 0)
            lb = torch.cat((lb.to(torch.int64), torch.full((pad,), -1, dtype=torch.int64)), 0)
        q_all.append(emb)
        b_all.append(bx.float())
        c_all.append(lb.to(torch.int64))
    return torch.stack(q_all), torch.stack(b_all), torch.stack(c_all), n_all


# --------------------------------------------------------------------------------------
# A2  MSDA core                  upstream ms_deform_attn_core_pytorch (third party, see header)
# --------------------------------------------------------------------------------------
def msda_core(value: Tensor, shapes: Sequence[Tuple[int, int]], loc: Tensor, attn: Tensor) -> Tensor:
    """value [B,S,M,D]; shapes [(H_l,W_l)]; loc [B,Lq,M,L,P,2] (x,y in [0,1]); attn [B,Lq,M,L,P]
    -> [B,Lq,M*D].  grid_sample formulation (bilinear, zero padding, align_corners=False)."""
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    chunks = value.split([h * w for h, w in shapes], dim=1)
    grid = 2 * loc - 1
    sampled = []
    for l, (h, w) in enumerate(shapes):
        v = chunks[l].flatten(2).transpose(1, 2).reshape(B * M, D, h, w)
        g = grid[:, :, :, l].transpose(1, 2).flatten(0, 1)              # [B*M, Lq, P, 2]
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(B * M, 1, Lq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * a).sum(-1)        # [B*M, D, Lq]
    return out.view(B, M * D, Lq).transpose(1, 2).contiguous()


def msda_core_direct(value: Tensor, shapes: Sequence[Tuple[int, int]], loc: Tensor, attn: Tensor) -> Tensor:
    """Same op written in the CUDA-kernel convention (used to pin the kernel's index math):
    pixel x = loc_x*W - 0.5, y = loc_y*H - 0.5; a sample contributes iff -1 < x < W and
    -1 < y < H; each of the four corners is bounds-checked; flat index start_l + y*W + x."""
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    out = value.new_zeros(B, Lq, M, D)
    start = 0
    bi = torch.arange(B)[:, None, None, None]
    mi = torch.arange(M)[None, None, :, None]
    for l, (h, w) in enumerate(shapes):
        x = loc[:, :, :, l, :, 0] * w - 0.5                              # [B,Lq,M,P]
        y = loc[:, :, :, l, :, 1] * h - 0.5
        inside = (x > -1) & (y > -1) & (x < w) & (y < h)
        x0 = torch.floor(x)
        y0 = torch.floor(y)
        fx = x - x0
        fy = y - y0
        acc = value.new_zeros(B, Lq, M, P, D)
        for dy, dx, wgt in ((0, 0, (1 - fy) * (1 - fx)), (0, 1, (1 - fy) * fx),
                            (1, 0, fy * (1 - fx)), (1, 1, fy * fx)):
            xi = (x0 + dx).long()
            yi = (y0 + dy).long()
            ok = inside & (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
            idx = start + yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)
            corner = value[bi, idx, mi]                                  # [B,Lq,M,P,D]
            acc = acc + corner * (wgt * ok)[..., None]
        out = out + (acc * attn[:, :, :, l, :, None]).sum(3)
        start += h * w
    return out.reshape(B, Lq, M * D)


# --------------------------------------------------------------------------------------
# A3  MSDeformAttn module        upstream models/ops/modules/ms_deform_attn.py (third party)
# --------------------------------------------------------------------------------------
def msda_module(P: Params, pre: str, query: Tensor, ref: Tensor, src: Tensor,
                shapes: Sequence[Tuple[int, int]], padding_mask: Optional[Tensor],
                n_heads: int, n_points: int) -> Tensor:
    """query [B,Lq,C]; ref [B,Lq,L,2]; src [B,S,C]; padding_mask [B,S] bool -> [B,Lq,C].
    Call sites in the reference: deformable_transformer.py:201 (encoder), :283-285 (decoder)."""
    B, Lq, C = query.shape
    S = src.shape[1]
    L = len(shapes)
    M, Pn = n_heads, n_points
    value = F.linear(src, P[pre + "value_proj.weight"], P[pre + "value_proj.bias"])
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)
    value = value.view(B, S, M, C // M)
    off = F.linear(query, P[pre + "sampling_offsets.weight"], P[pre + "sampling_offsets.bias"])
    off = off.view(B, Lq, M, L, Pn, 2)
    logits = F.linear(query, P[pre + "attention_weights.weight"], P[pre + "attention_weights.bias"])
    attn = F.softmax(logits.view(B, Lq, M, L * Pn), -1).view(B, Lq, M, L, Pn)
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype)    # normaliser is (W_l, H_l)
    loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    out = msda_core(value, shapes, loc, attn)
    return F.linear(out, P[pre + "output_proj.weight"], P[pre + "output_proj.bias"])


def msda_reset_parameters(P: Params, pre: str, n_heads: int, n_levels: int, n_points: int) -> None:
    """Upstream ``MSDeformAttn._reset_parameters`` (called at deformable_transformer.py:58-59):
    zero offset weights, directional offset bias, zero attention weights/bias, xavier
    value/output projections with zero bias.  In place on P."""
    M, L, Pn = n_heads, n_levels, n_points
    P[pre + "sampling_offsets.weight"].zero_()
    th = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
    g = torch.stack([th.cos(), th.sin()], -1)
    g = (g / g.abs().max(-1, keepdim=True)[0]).view(M, 1, 1, 2).repeat(1, L, Pn, 1)
    for i in range(Pn):
        g[:, :, i, :] *= i + 1
    P[pre + "sampling_offsets.bias"].copy_(g.view(-1))
    P[pre + "attention_weights.weight"].zero_()
    P[pre + "attention_weights.bias"].zero_()
    torch.nn.init.xavier_uniform_(P[pre + "value_proj.weight"])
    P[pre + "value_proj.bias"].zero_()
    torch.nn.init.xavier_uniform_(P[pre + "output_proj.weight"])
    P[pre + "output_proj.bias"].zero_()


# --------------------------------------------------------------------------------------
# A4/A5  encoder                 deformable_transformer.py:193-208 (layer), :217-238 (stack)
# --------------------------------------------------------------------------------------
def _ln(P: Params, pre: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), P[pre + "weight"], P[pre + "bias"], 1e-5)


def _ffn(P: Params, pre: str, x: Tensor) -> Tensor:
    return F.linear(F.relu(F.linear(x, P[pre + "linear1.weight"], P[pre + "linear1.bias"])),
                    P[pre + "linear2.weight"], P[pre + "linear2.bias"])


def encoder_reference_points(shapes: Sequence[Tuple[int, int]], valid_ratios: Tensor) -> Tensor:
    """deformable_transformer.py:217-230: pixel centres / (valid_ratio * size), then scaled by
    every level's valid ratio -> [B,S,L,2] (x,y)."""
    per_level = []
    for l, (h, w) in enumerate(shapes):
        ys = torch.linspace(0.5, h - 0.5, h, dtype=torch.float32)
        xs = torch.linspace(0.5, w - 0.5, w, dtype=torch.float32)
        gy = ys[:, None].expand(h, w).reshape(-1)[None] / (valid_ratios[:, None, l, 1] * h)
        gx = xs[None, :].expand(h, w).reshape(-1)[None] / (valid_ratios[:, None, l, 0] * w)
        per_level.append(torch.stack((gx, gy), -1))
    pts = torch.cat(per_level, 1)
    return pts[:, :, None] * valid_ratios[:, None]


def encoder_layer(P: Params, pre: str, src: Tensor, pos: Tensor, ref: Tensor, shapes, padding_mask,
                  n_heads: int, n_points: int) -> Tensor:
    a = msda_module(P, pre + "self_attn.", src + pos, ref, src, shapes, padding_mask, n_heads, n_points)
    src = _ln(P, pre + "norm1.", src + a)                                 # :202-203 (dropout = identity)
    return _ln(P, pre + "norm2.", src + _ffn(P, pre, src))                # :193-197


# --------------------------------------------------------------------------------------
# A7/A8  decoder                 deformable_transformer.py:275-292 (layer), :305-340 (stack)
# --------------------------------------------------------------------------------------
def mha_self_attention(P: Params, pre: str, qk_in: Tensor, v_in: Tensor, n_heads: int) -> Tensor:
    """nn.MultiheadAttention as used at deformable_transformer.py:277-278: q = k = tgt+pos,
    v = tgt, no masks (dummy queries attend and are attended).  Inputs [B,Q,C]."""
    B, Q, C = qk_in.shape
    D = C // n_heads
    W, b = P[pre + "in_proj_weight"], P[pre + "in_proj_bias"]
    q = F.linear(qk_in, W[:C], b[:C]).view(B, Q, n_heads, D).transpose(1, 2)
    k = F.linear(qk_in, W[C:2 * C], b[C:2 * C]).view(B, Q, n_heads, D).transpose(1, 2)
    v = F.linear(v_in, W[2 * C:], b[2 * C:]).view(B, Q, n_heads, D).transpose(1, 2)
    p = F.softmax((q * (1.0 / math.sqrt(D))) @ k.transpose(-1, -2), -1)
    o = (p @ v).transpose(1, 2).reshape(B, Q, C)
    return F.linear(o, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])


def decoder_layer(P: Params, pre: str, tgt: Tensor, qpos: Tensor, ref_in: Tensor, memory: Tensor,
                  shapes, padding_mask, n_heads: int, n_points: int) -> Tensor:
    a = mha_self_attention(P, pre + "self_attn.", tgt + qpos, tgt, n_heads)
    tgt = _ln(P, pre + "norm2.", tgt + a)                                 # :279-280
    c = msda_module(P, pre + "cross_attn.", tgt + qpos, ref_in, memory, shapes, padding_mask,
                    n_heads, n_points)
    tgt = _ln(P, pre + "norm1.", tgt + c)                                 # :286-287
    return _ln(P, pre + "norm3.", tgt + _ffn(P, pre, tgt))                # :267-271


# --------------------------------------------------------------------------------------
# A6  DeformableTransformer.forward                 deformable_transformer.py:111-166
# --------------------------------------------------------------------------------------
def valid_ratio(mask: Tensor) -> Tensor:
    _, H, W = mask.shape                                                   # :111-118
    vh = (~mask[:, :, 0]).sum(1).float() / H
    vw = (~mask[:, 0, :]).sum(1).float() / W
    return torch.stack((vw, vh), -1)


def transformer_forward(P: Params, cfg: dict, srcs: Sequence[Tensor], masks: Sequence[Tensor],
                        pos_embeds: Sequence[Tensor], query_embed: Tensor, reference_points: Tensor,
                        pre: str = "transformer.", capture: Optional[dict] = None):
    """srcs/pos L x [B,C,H,W]; masks L x [B,H,W]; query_embed [B,Q,2C]; reference_points [B,Q,2].
    Returns (hs [Dl,B,Q,C], init_reference [B,Q,2], inter_references [Dl,B,Q,2])."""
    dt = srcs[0].dtype
    M, Pn = cfg["nheads"], cfg["n_points"]
    shapes = [tuple(s.shape[-2:]) for s in srcs]
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    pmask = torch.cat([m.flatten(1) for m in masks], 1)
    pos = torch.cat([(p.flatten(2).transpose(1, 2) + P[pre + "level_embed"][l].view(1, 1, -1))
                     for l, p in enumerate(pos_embeds)], 1)
    vr = torch.stack([valid_ratio(m) for m in masks], 1).to(dt)           # [B,L,2]

    ref_enc = encoder_reference_points(shapes, vr.float()).to(dt)
    mem = src
    for i in range(cfg["enc_layers"]):
        mem = encoder_layer(P, f"{pre}encoder.layers.{i}.", mem, pos, ref_enc, shapes, pmask, M, Pn)
    if capture is not None:
        capture["memory"] = mem

    C = mem.shape[-1]
    qpos, tgt = query_embed[..., :C], query_embed[..., C:]                # :156
    ref_in = reference_points[:, :, None] * vr[:, None]                   # :316-317
    hs, refs = [], []
    for i in range(cfg["dec_layers"]):
        tgt = decoder_layer(P, f"{pre}decoder.layers.{i}.", tgt, qpos, ref_in, mem, shapes, pmask, M, Pn)
        hs.append(tgt)
        refs.append(reference_points)
    return torch.stack(hs), reference_points, torch.stack(refs)


# --------------------------------------------------------------------------------------
# A9  heads                      pose_estimation_transformer.py:357-393, 434-451, 677-689
# --------------------------------------------------------------------------------------
def mlp3(P: Params, pre: str, x: Tensor) -> Tensor:
    for k in range(3):
        x = F.linear(x, P[f"{pre}layers.{k}.weight"], P[f"{pre}layers.{k}.bias"])
        if k < 2:
            x = F.relu(x)
    return x


def rotation_6d_to_matrix(r6: Tensor) -> Tensor:
    """[...,6] -> [...,3,3]; columns (x,y,z): x=norm(a1), z=norm(x x a2), y = z x x (:434-451)."""
    a1, a2 = r6[..., 0:3], r6[..., 3:6]
    x = F.normalize(a1, p=2, dim=-1)
    z = F.normalize(torch.cross(x, a2, dim=-1), p=2, dim=-1)
    y = torch.cross(z, x, dim=-1)
    return torch.stack((x, y, z), dim=-1)


def pose_heads(P: Params, cfg: dict, hs: Tensor, pred_classes: Tensor):
    """hs [Dl,B,Q,C] -> (translation [Dl,B,Q,3], rotation [Dl,B,Q,3,3], rot6d [Dl,B,Q,6])."""
    Dl, B, Q, _ = hs.shape
    specific = cfg["class_mode"] == "specific"
    slot = torch.where(pred_classes > 0, pred_classes, 0).view(-1)        # :354
    rows = torch.arange(B * Q)
    t_all, r_all, r6_all = [], [], []
    for l in range(Dl):
        r = mlp3(P, f"rotation_head.{l}.", hs[l])
        t = mlp3(P, f"translation_head.{l}.", hs[l])
        if specific:                                                      # :365-374
            n_cls = cfg["n_classes"] + 1
            r = r.view(B * Q, n_cls, 6)[rows, slot].view(B, Q, 6)
            t = t.view(B * Q, n_cls, 3)[rows, slot].view(B, Q, 3)
        t_all.append(t)
        r6_all.append(r)
        r_all.append(rotation_6d_to_matrix(r))
    return torch.stack(t_all), torch.stack(r_all), torch.stack(r6_all)


# --------------------------------------------------------------------------------------
# N1  input_proj (row "next")     pose_estimation_transformer.py:100-135, 313-335
# --------------------------------------------------------------------------------------
def input_proj(P: Params, cfg: dict, feats: Sequence[Tensor], feat_masks: Sequence[Tensor],
               image_mask: Tensor):
    """feats: backbone maps [B,Cin,H,W]; returns (srcs, masks, pos) for all n_levels levels."""
    srcs, masks = [], []
    for l, f in enumerate(feats):
        y = F.conv2d(f, P[f"input_proj.{l}.0.weight"], P[f"input_proj.{l}.0.bias"])
        srcs.append(F.group_norm(y, 32, P[f"input_proj.{l}.1.weight"], P[f"input_proj.{l}.1.bias"], 1e-5))
        masks.append(feat_masks[l])
    for l in range(len(feats), cfg["n_levels"]):
        x = feats[-1] if l == len(feats) else srcs[-1]
        y = F.conv2d(x, P[f"input_proj.{l}.0.weight"], P[f"input_proj.{l}.0.bias"], stride=2, padding=1)
        y = F.group_norm(y, 32, P[f"input_proj.{l}.1.weight"], P[f"input_proj.{l}.1.bias"], 1e-5)
        m = F.interpolate(image_mask[None].float(), size=y.shape[-2:]).to(torch.bool)[0]
        srcs.append(y)
        masks.append(m)
    pos = [sine_position_embedding(m, cfg["d_model"] // 2).to(srcs[0].dtype) for m in masks]
    return srcs, masks, pos


# --------------------------------------------------------------------------------------
# A10  whole path: post-input_proj pyramid + boxes -> output dict
# --------------------------------------------------------------------------------------
def poet_path_forward(P: Params, cfg: dict, srcs: Sequence[Tensor], masks: Sequence[Tensor],
                      boxes: Sequence[Tensor], labels: Sequence[Tensor], capture: Optional[dict] = None):
    """The benchmarked path: pyramids + gt boxes -> PoET output dict
    (pose_estimation_transformer.py:398-414).  Returns (out, n_boxes_per_sample)."""
    dt = srcs[0].dtype
    pos = [sine_position_embedding(m, cfg["d_model"] // 2).to(dt) for m in masks]
    qe, pboxes, pcls, n_boxes = build_queries(boxes, labels, cfg["num_queries"], cfg["d_model"])
    qe, pboxes_dt = qe.to(dt), pboxes.to(dt)
    hs, _, _ = transformer_forward(P, cfg, srcs, masks, pos, qe, pboxes_dt[:, :, :2], capture=capture)
    t, R, r6 = pose_heads(P, cfg, hs, pcls)
    if capture is not None:
        capture.update(hs=hs, rot6d=r6, translation_all=t, rotation_all=R)
    out = {"pred_translation": t[-1], "pred_rotation": R[-1], "pred_boxes": pboxes, "pred_classes": pcls}
    if cfg.get("aux_loss", True):
        out["aux_outputs"] = [{"pred_translation": t[i], "pred_rotation": R[i], "pred_boxes": pboxes,
                               "pred_classes": pcls} for i in range(t.shape[0] - 1)]
    return out, n_boxes


def synthetic_loss(capture_or_tR, g_t: Tensor, g_R: Tensor) -> Tensor:
    """sum_l <translation_l, g_t[l]> + <rotation_l, g_R[l]> (fixed cotangents)."""
    t, R = capture_or_tR
    return (t * g_t).sum() + (R * g_R).sum()


# --------------------------------------------------------------------------------------
# N2: SetCriterion + PoseMatcher in 'gt' bbox mode
# --------------------------------------------------------------------------------------
def pose_criterion_gt(t_all, R_all, tgt_t, tgt_R, n_boxes, w_trans=1.0, w_rot=1.0):
    """Restatement of reference models/pose_estimation_transformer.py SetCriterion.forward (:635-662) with
    losses ['translation', 'rotation'] (:478-494, :519-537) and models/matcher.py PoseMatcher in bbox_mode 'gt'
    (:158-173, :192-195): the predicted boxes of the first n_boxes[i] queries ARE the target boxes, so the L1
    cost matrix has a zero diagonal and the Hungarian assignment is query j <-> target j.

    t_all [L,B,Q,3], R_all [L,B,Q,3,3] (decoder layers, last = final output); tgt_t / tgt_R: lists (len B) of
    [n_i,3] / [n_i,3,3]; n_boxes: list of ints.  Returns (dict with the reference's keys: 'loss_trans',
    'loss_rot' for the last layer and '<k>_<i>' for aux layer i, weighted total as in engine.py:60-61)."""
    L = t_all.shape[0]
    n_obj = sum(int(n) for n in n_boxes)
    losses = {}
    for l in range(L):
        src_t = torch.cat([t_all[l, b, :n] for b, n in enumerate(n_boxes)], 0)
        src_R = torch.cat([R_all[l, b, :n] for b, n in enumerate(n_boxes)], 0)
        gt_t = torch.cat([t[:n] for t, n in zip(tgt_t, n_boxes)], 0)
        gt_R = torch.cat([r[:n] for r, n in zip(tgt_R, n_boxes)], 0)
        lt = torch.sqrt(((src_t - gt_t) ** 2).sum(1)).sum() / n_obj                       # :486-490
        prod = torch.bmm(src_R, gt_R.transpose(1, 2))                                     # :531
        trace = prod.diagonal(dim1=1, dim2=2).sum(1)
        theta = torch.clamp(0.5 * (trace - 1), -1 + 1e-6, 1 - 1e-6)                       # :533
        lr = torch.acos(theta).sum() / n_obj
        suffix = "" if l == L - 1 else f"_{l}"
        losses["loss_trans" + suffix] = lt
        losses["loss_rot" + suffix] = lr
    total = sum(v * (w_trans if k.startswith("loss_trans") else w_rot) for k, v in losses.items())
    return losses, total
