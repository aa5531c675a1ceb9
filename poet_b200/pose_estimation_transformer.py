"""B200-native PoET behind the reference's nn.Module surface and output-dict contract.

Mirror of reference models/pose_estimation_transformer.py:32-451 (`PoET`), :677-689 (`MLP`),
:692-739 (`build`): same constructor arguments, attribute names (transformer, input_proj,
translation_head, rotation_head, bbox_embedding ...), state_dict keys and
``forward(samples, targets) -> (out_dict, n_boxes_per_sample)`` contract, so it
drops into the reference's engine.py unchanged.  What changed underneath:

  query construction   :203-239  per-image Python loop  -> one padded batch + poet_bbox_embed_pad
  transformer          :346      -> poet_b200.deformable_transformer (CUDA kernels)
  heads                :357-393  per-row Python list comprehension -> poet_gemm MLPs +
                                  poet_heads_select_rot6d (class select + Gram-Schmidt fused)

Out of the hot path and therefore plain PyTorch: the frozen detector backbone
and input_proj (row 8, "next" N1).  Unsupported-but-reachable flags raise NotImplementedError like
the reference does for its own unsupported modes (:82, :98, :154, :307).
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .deformable_transformer import build_deforamble_transformer
from .position_encoding import BoundingBoxEmbeddingSine


class MLP(nn.Module):
    """Linear/ReLU stack with the reference's parameter names (`layers.K.{weight,bias}`)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(i, o) for i, o in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        return ops.mlp(x, [(l.weight, l.bias) for l in self.layers])


class _Nested:
    """Minimal stand-in for the reference's util.misc.NestedTensor (tensors + mask)."""

    def __init__(self, tensors, mask):
        self.tensors, self.mask = tensors, mask

    def decompose(self):
        return self.tensors, self.mask


def _as_nested(samples):
    if hasattr(samples, "tensors") and hasattr(samples, "mask"):
        return samples
    imgs = list(samples)                                           # list of [3,H,W] images: pad to the largest
    H = max(i.shape[1] for i in imgs)
    W = max(i.shape[2] for i in imgs)
    batch = imgs[0].new_zeros((len(imgs), imgs[0].shape[0], H, W))
    mask = torch.ones((len(imgs), H, W), dtype=torch.bool, device=imgs[0].device)
    for k, im in enumerate(imgs):
        batch[k, :, : im.shape[1], : im.shape[2]] = im
        mask[k, : im.shape[1], : im.shape[2]] = False
    return _Nested(batch, mask)


class PoET(nn.Module):
    def __init__(self, backbone, transformer, num_queries, num_feature_levels, n_classes, bbox_mode="gt",
                 ref_points_mode="bbox", query_embedding_mode="bbox", rotation_mode="6d", class_mode="agnostic",
                 aleatoric=False, aux_loss=True, backbone_type="yolo"):
        super().__init__()
        self.transformer = transformer
        hidden_dim = transformer.d_model
        self.hidden_dim = hidden_dim
        self.backbone = backbone
        self.backbone_type = backbone_type
        self.aux_loss = aux_loss
        self.n_queries = num_queries
        self.n_classes = n_classes + 1                    # +1 dummy/background slot (reference :63)
        self.bbox_mode = bbox_mode
        self.ref_points_mode = ref_points_mode
        self.query_embedding_mode = query_embedding_mode
        self.rotation_mode = rotation_mode
        self.class_mode = class_mode
        self.aleatoric = aleatoric
        if aleatoric:
            raise NotImplementedError("aleatoric heads are outside the poet_b200 hot path (no BASELINE config uses them)")
        if rotation_mode != "6d":
            raise NotImplementedError("poet_b200 implements the '6d' rotation representation of the PoET configs")
        if ref_points_mode != "bbox" or query_embedding_mode != "bbox":
            raise NotImplementedError("poet_b200 implements bbox reference points / bbox query embeddings")
        if class_mode not in ("agnostic", "specific"):
            raise NotImplementedError("Class mode is not supported.")
        self.t_dim, self.rot_dim = 3, 6
        slots = self.n_classes if class_mode == "specific" else 1
        # registration (= optimizer parameter numbering, an on-disk format: main.py:302) and RNG order of the reference
        # (:86-96): the two prototype heads first, input_proj next, then the per-layer copies replace the prototypes
        self.translation_head = MLP(hidden_dim, hidden_dim, self.t_dim * slots, 3)
        self.rotation_head = MLP(hidden_dim, hidden_dim, self.rot_dim * slots, 3)

        self.num_feature_levels = num_feature_levels
        if backbone is not None:
            n_outs = len(backbone.strides)
            proj = []
            in_ch = hidden_dim
            for n in range(n_outs if num_feature_levels > 1 else 1):
                in_ch = backbone.num_channels[n]
                proj.append(nn.Sequential(nn.Conv2d(in_ch, hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim)))
            for _ in range(num_feature_levels - n_outs if num_feature_levels > 1 else 0):
                proj.append(nn.Sequential(nn.Conv2d(in_ch, hidden_dim, kernel_size=3, stride=2, padding=1),
                                          nn.GroupNorm(32, hidden_dim)))
                in_ch = hidden_dim
            self.input_proj = nn.ModuleList(proj)
            for p in self.input_proj:
                nn.init.xavier_uniform_(p[0].weight, gain=1)
                nn.init.constant_(p[0].bias, 0)
        else:
            self.input_proj = nn.ModuleList()             # pyramid-only use (forward_pyramid)

        n_pred = transformer.decoder.num_layers
        self.translation_head = nn.ModuleList(copy.deepcopy(self.translation_head) for _ in range(n_pred))
        self.rotation_head = nn.ModuleList(copy.deepcopy(self.rotation_head) for _ in range(n_pred))
        self.bbox_embedding = BoundingBoxEmbeddingSine(num_pos_feats=hidden_dim / 8)

    # ------------------------------------------------------------------ queries (A1)
    def _pad_boxes(self, boxes: Sequence[torch.Tensor], classes: Sequence[torch.Tensor], device, pin: bool = True):
        """Lists of per-image boxes [n_i,4] / classes [n_i] -> padded [B,Q,4] (-1), [B,Q] int64 (-1), counts."""
        B, Q = len(boxes), self.n_queries
        counts = [min(int(b.shape[0]), Q) for b in boxes]
        src_dev = boxes[0].device
        # pad where the boxes live (host lists from the data loader stay on the host: one H2D copy)
        pb = torch.full((B * Q, 4), -1.0, dtype=torch.float32, device=src_dev)
        pc = torch.full((B * Q,), -1, dtype=torch.int64, device=src_dev)
        rows = [i * Q + j for i, n in enumerate(counts) for j in range(n)]
        if rows:
            idx = torch.tensor(rows, dtype=torch.int64).to(src_dev)
            pb[idx] = torch.cat([b[:n].to(torch.float32) for b, n in zip(boxes, counts)], 0)
            pc[idx] = torch.cat([c[:n].to(torch.int64) for c, n in zip(classes, counts)], 0)
        n_dev = torch.tensor(counts, dtype=torch.int32)
        if pin and src_dev.type == "cpu" and torch.device(device).type == "cuda":
            pb, pc, n_dev = pb.pin_memory(), pc.pin_memory(), n_dev.pin_memory()
        pb = pb.view(B, Q, 4).to(device, non_blocking=True)
        pc = pc.view(B, Q).to(device, non_blocking=True)
        return pb, pc, counts, n_dev.to(device, non_blocking=True)

    def build_queries(self, boxes, classes, device):
        pb, pc, counts, n_dev = self._pad_boxes(boxes, classes, device)
        qe = ops.bbox_embed_pad(pb, n_dev, int(self.hidden_dim // 8))
        return qe, pb, pc, counts

    def _queries_from_targets(self, targets, device):
        key = "boxes" if self.bbox_mode == "gt" else "jitter_boxes"
        for t in targets:
            if t[key].shape[0] > self.n_queries:
                raise ValueError("more target boxes than object queries")
        return self.build_queries([t[key] for t in targets], [t["labels"] for t in targets], device)

    def _queries_from_backbone(self, pred_objects, image_hw, device):
        """reference :240-305: xyxy -> normalised cxcywh, top-Q by score, dummy padding."""
        boxes, classes = [], []
        for pred in pred_objects:
            if pred is None or pred.shape[0] == 0:
                boxes.append(torch.zeros((0, 4), device=device))
                classes.append(torch.zeros((0,), dtype=torch.int64, device=device))
                continue
            xyxy = pred[:, :4]
            cxcywh = torch.stack(((xyxy[:, 0] + xyxy[:, 2]) / 2, (xyxy[:, 1] + xyxy[:, 3]) / 2,
                                  xyxy[:, 2] - xyxy[:, 0], xyxy[:, 3] - xyxy[:, 1]), -1)
            h, w = image_hw
            cxcywh = cxcywh / torch.tensor([w, h, w, h], dtype=torch.float32, device=cxcywh.device)
            cls = pred[:, 5].to(torch.int64)
            if cxcywh.shape[0] > self.n_queries:
                order = torch.sort(pred[:, 4], dim=0, descending=True)[1][: self.n_queries]
                cxcywh, cls = cxcywh[order], cls[order]
            boxes.append(cxcywh)
            classes.append(cls)
        return self.build_queries(boxes, classes, device)

    # ------------------------------------------------------------------ heads (A9)
    def _head_layer(self, l: int, h: torch.Tensor, pred_classes: torch.Tensor):
        slots = self.n_classes if self.class_mode == "specific" else 1
        cls = pred_classes if slots > 1 else None
        rot = self.rotation_head[l](h)
        tr = self.translation_head[l](h)
        t, R, _ = ops.heads_select_rot6d(rot, tr, cls, slots)
        return t, R

    def _heads(self, hs: torch.Tensor, pred_classes: torch.Tensor):
        t_all, R_all = [], []
        for l in range(hs.shape[0]):
            t, R = self._head_layer(l, hs[l], pred_classes)
            t_all.append(t)
            R_all.append(R)
        return t_all, R_all

    def _run_with_heads(self, srcs, masks, pos, qe, ref, pred_classes, pos_tokens=None, src_tokens=None):
        """transformer + heads; with stream forking on, layer l's heads are issued on a side stream the moment
        decoder layer l is issued, so they (and their backward) overlap the rest of the decoder chain."""
        with ops.planes_scope(self):          # transformer + head weights -> bf16 planes, one launch
            return self._run_with_heads_impl(srcs, masks, pos, qe, ref, pred_classes, pos_tokens, src_tokens)

    def _run_with_heads_impl(self, srcs, masks, pos, qe, ref, pred_classes, pos_tokens=None, src_tokens=None):
        if not (ops.parallel_streams_enabled() and masks[0].is_cuda):
            hs = self.transformer(srcs, masks, pos, qe, ref, pos_tokens=pos_tokens, src_tokens=src_tokens)[0]
            return self._heads(hs, pred_classes)
        dev = masks[0].device
        pending = []

        def on_layer(l, out):
            f = ops.fork(1 + (l % 3), dev)
            f.uses(out, pred_classes)
            with f:
                t, R = self._head_layer(l, out, pred_classes)
                pending.append((f, f.checkpoint(), t, R))

        self.transformer(srcs, masks, pos, qe, ref, pos_tokens=pos_tokens, layer_callback=on_layer, src_tokens=src_tokens)
        t_all, R_all = [], []
        for f, ev, t, R in pending:
            f.wait(ev, t, R)
            t_all.append(t)
            R_all.append(R)
        return t_all, R_all

    def _pack(self, t_all, R_all, pred_boxes, pred_classes):
        out = {"pred_translation": t_all[-1], "pred_rotation": R_all[-1], "pred_boxes": pred_boxes,
               "pred_classes": pred_classes}
        if self.aux_loss:
            out["aux_outputs"] = [{"pred_translation": t, "pred_rotation": r, "pred_boxes": pred_boxes,
                                   "pred_classes": pred_classes} for t, r in zip(t_all[:-1], R_all[:-1])]
        return out

    # ------------------------------------------------------------------ entry points
    def forward_pyramid(self, srcs: Sequence[torch.Tensor], masks: Sequence[torch.Tensor], boxes, classes):
        """The benchmarked path: post-input_proj pyramid + boxes -> output dict.
        Position encodings are written straight into token layout (no NCHW pos tensors)."""
        pb, pc, counts, n_dev = self._pad_boxes(boxes, classes, srcs[0].device)
        return self.forward_padded(srcs, masks, pb, pc, n_dev), counts

    # Images are independent, so the batch can be cut into `micro_batches` slices issued on different streams:
    # the launch-latency-bound decoder / head chain of one slice (GPU mostly idle) then overlaps the
    # throughput-bound encoder kernels of another.  Parameter gradients of all slices accumulate into the
    # same .grad slots (atomic beta = 1 epilogues).  1 = off.
    micro_batches = 1

    def forward_padded(self, srcs, masks, pred_boxes, pred_classes, n_boxes_dev):
        """Device-only part of the path (no host work, no syncs: capturable in a CUDA graph):
        padded boxes [B,Q,4] / classes [B,Q] int64 / counts [B] int32, all on the device."""
        mb, B = int(self.micro_batches), srcs[0].shape[0]
        if mb <= 1 or B % mb or not srcs[0].is_cuda:
            return self._forward_padded_slice(srcs, masks, pred_boxes, pred_classes, n_boxes_dev)
        dev, per = srcs[0].device, B // mb
        main = torch.cuda.current_stream(dev)
        with ops.planes_scope(self):                         # weight planes: one refresh for all slices
            streams = [main] + [ops.side_stream(32 + i, dev) for i in range(1, mb)]
            for st in streams[1:]:
                st.wait_stream(main)                         # fork before any slice is enqueued
            outs = []
            for i, st in enumerate(streams):
                sl = slice(i * per, (i + 1) * per)
                with torch.cuda.stream(st), ops.stream_namespace(i):
                    o = self._forward_padded_slice([s[sl] for s in srcs], [m[sl] for m in masks], pred_boxes[sl],
                                                   pred_classes[sl], n_boxes_dev[sl], pack=False)
                    if st is not main:                       # this slice's backward tail must rejoin the caller's stream
                        o = ([ops.join_after_backward(t, st) for t in o[0]], [ops.join_after_backward(t, st) for t in o[1]])
                if st is not main:
                    for t in (*srcs, *masks, pred_boxes, pred_classes, n_boxes_dev):
                        t.record_stream(st)
                outs.append(o)
            for st, o in zip(streams[1:], outs[1:]):
                main.wait_stream(st)
                for t in (*o[0], *o[1]):
                    t.record_stream(main)
        n_layers = len(outs[0][0])
        t_all = [torch.cat([o[0][l] for o in outs], 0) for l in range(n_layers)]
        R_all = [torch.cat([o[1][l] for o in outs], 0) for l in range(n_layers)]
        return self._pack(t_all, R_all, pred_boxes, pred_classes)

    def _forward_padded_slice(self, srcs, masks, pred_boxes, pred_classes, n_boxes_dev, pack=True):
        qe = ops.bbox_embed_pad(pred_boxes, n_boxes_dev, int(self.hidden_dim // 8))
        C = srcs[0].shape[1]
        S = sum(int(s.shape[2] * s.shape[3]) for s in srcs)
        pos_tokens = _PosTokens.apply(self.transformer.level_embed, C, S, *masks)
        t_all, R_all = self._run_with_heads(srcs, masks, None, qe, pred_boxes[:, :, :2].contiguous(), pred_classes,
                                            pos_tokens=pos_tokens)
        return self._pack(t_all, R_all, pred_boxes, pred_classes) if pack else (t_all, R_all)

    def forward_features(self, feats: Sequence[torch.Tensor], feat_masks: Sequence[torch.Tensor], image_mask: torch.Tensor,
                         boxes, classes):
        """Backbone feature maps -> output dict: input_proj + the hot path, all on our kernels.
        feats[l] [B,Cin,H_l,W_l] with masks feat_masks[l] [B,H_l,W_l]; image_mask [B,H,W] is the padded-image mask the
        extra level's mask is interpolated from (reference :326-334).  The projections are written token-major."""
        if len(self.input_proj) == 0:
            raise RuntimeError("this PoET was built without a backbone: no input_proj parameters")
        pb, pc, counts, n_dev = self._pad_boxes(boxes, classes, feats[0].device)
        qe = ops.bbox_embed_pad(pb, n_dev, int(self.hidden_dim // 8))
        return self._features_to_outputs(feats, feat_masks, image_mask, qe, pb, pc), counts

    def forward_features_padded(self, feats, feat_masks, image_mask, pred_boxes, pred_classes, n_boxes_dev):
        """Device-only form of forward_features (padded boxes / classes / counts already on the device): capturable."""
        qe = ops.bbox_embed_pad(pred_boxes, n_boxes_dev, int(self.hidden_dim // 8))
        return self._features_to_outputs(feats, feat_masks, image_mask, qe, pred_boxes, pred_classes)

    def _features_to_outputs(self, feats, feat_masks, image_mask, qe, pb, pc):
        masks = list(feat_masks)
        h, w = int(feats[-1].shape[2]), int(feats[-1].shape[3])
        for _ in range(len(feats), self.num_feature_levels):
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            masks.append(F.interpolate(image_mask[None].float(), size=(h, w)).to(torch.bool)[0])
        with ops.planes_scope(self):
            levels = [(ip[0].weight, ip[0].bias, ip[1].weight, ip[1].bias) for ip in self.input_proj]
            src_tokens = ops.input_proj_tokens(list(feats), levels, groups=self.input_proj[0][1].num_groups,
                                               eps=self.input_proj[0][1].eps)
            C, S = src_tokens.shape[2], src_tokens.shape[1]
            pos_tokens = _PosTokens.apply(self.transformer.level_embed, C, S, *masks)
            t_all, R_all = self._run_with_heads(None, masks, None, qe, pb[:, :, :2].contiguous(), pc,
                                                pos_tokens=pos_tokens, src_tokens=src_tokens)
        return self._pack(t_all, R_all, pb, pc)

    def _default_sine_posenc(self) -> bool:
        """True when backbone[1] is the sine position encoding every PoET config builds (position_encoding.py:87-99:
        hidden_dim/2 features, temperature 1e4, normalize, scale 2*pi): only then may the token-layout kernel rebuild
        the encodings from the masks instead of using the `pos` tensors the backbone returned."""
        import math
        try:
            pe = self.backbone[1]
        except (TypeError, IndexError, KeyError):
            return False
        return (type(pe).__name__ == "PositionEmbeddingSine" and int(getattr(pe, "num_pos_feats", -1)) * 2 == self.hidden_dim
                and float(getattr(pe, "temperature", 0)) == 10000.0 and bool(getattr(pe, "normalize", False))
                and abs(float(getattr(pe, "scale", 0.0)) - 2 * math.pi) < 1e-12)

    def forward(self, samples, targets=None):
        samples = _as_nested(samples)
        image_hw = (int(samples.tensors.shape[-2]), int(samples.tensors.shape[-1]))
        features, pos, pred_objects = self.backbone(samples)
        dev = features[0].tensors.device
        if self.bbox_mode in ("gt", "jitter") and targets is not None:
            qe, pb, pc, counts = self._queries_from_targets(targets, dev)
        elif self.bbox_mode == "backbone":
            qe, pb, pc, counts = self._queries_from_backbone(pred_objects, image_hw, dev)
        else:
            raise NotImplementedError("PoET Bounding Box Mode not implemented!")

        if (dev.type == "cuda" and self.num_feature_levels > 1 and len(self.input_proj) == self.num_feature_levels
                and self.num_feature_levels - len(features) in (0, 1)
                and (all(pe is None for pe in pos) or self._default_sine_posenc())):
            # input_proj + path on our kernels; position encodings are rebuilt from the masks in token layout
            fm = [feat.decompose() for feat in features]
            if any(m is None for _, m in fm):
                raise ValueError("backbone features need padding masks")
            return self._features_to_outputs([f for f, _ in fm], [m for _, m in fm], samples.mask, qe, pb, pc), counts

        srcs, masks, pos = [], [], list(pos)
        for lvl, feat in enumerate(features):
            src, mask = feat.decompose()
            if mask is None:
                raise ValueError("backbone features need padding masks")
            srcs.append(self.input_proj[lvl](src))
            masks.append(mask)
        for lvl in range(len(srcs), self.num_feature_levels):
            src = self.input_proj[lvl](features[-1].tensors if lvl == len(features) else srcs[-1])
            mask = F.interpolate(samples.mask[None].float(), size=src.shape[-2:]).to(torch.bool)[0]
            pos.append(self.backbone[1](_Nested(src, mask)).to(src.dtype))
            srcs.append(src)
            masks.append(mask)

        t_all, R_all = self._run_with_heads(srcs, masks, pos, qe, pb[:, :, :2].contiguous(), pc)
        return self._pack(t_all, R_all, pb, pc), counts


class _PosTokens(torch.autograd.Function):
    """lvl_pos_embed_flatten [B,S,C] = sine(mask_l) + level_embed[l], written token-major by the
    posenc kernel; backward = per-level column sums into level_embed's gradient."""

    @staticmethod
    def forward(ctx, level_embed, C, S, *masks):
        B = masks[0].shape[0]
        out = torch.empty((B, S, C), device=masks[0].device, dtype=torch.float32)
        off, hws = 0, []
        for l, m in enumerate(masks):
            ops.posenc_sine_tokens_(out, m, level_embed[l], off, F=C // 2)
            hws.append(int(m.shape[1] * m.shape[2]))
            off += hws[-1]
        ctx.meta = (B, C, S, hws)
        ctx.le_param = level_embed
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, S, hws = ctx.meta
        g = g.contiguous()
        slot = ops._grad_slot(ctx.le_param)               # accumulate straight into level_embed.grad when it exists
        gle = slot if slot is not None else torch.zeros((len(hws), C), device=g.device, dtype=torch.float32)
        off = 0
        for l, hw in enumerate(hws):
            ops._call("poet_tokens_to_nchw", g.data_ptr(), None, gle[l].data_ptr(), B, C, hw, S, off, ops._stream(g))
            off += hw
        return (None if slot is not None else gle, None, None, *([None] * len(hws)))


def build_poet(args, backbone=None):
    """The model of reference build() (:692-712).  The detector backbone is outside the hot path: pass the
    reference's own (`models.backbone.build_backbone(args)`) or any module with the same contract as `backbone`
    / `args.backbone_module`."""
    backbone = backbone if backbone is not None else getattr(args, "backbone_module", None)
    transformer = build_deforamble_transformer(args)
    return PoET(backbone, transformer, num_queries=args.num_queries, num_feature_levels=args.num_feature_levels,
                n_classes=args.n_classes, bbox_mode=args.bbox_mode, ref_points_mode=args.reference_points,
                query_embedding_mode=args.query_embedding, rotation_mode=args.rotation_representation,
                class_mode=args.class_mode, aleatoric=args.aleatoric, aux_loss=args.aux_loss,
                backbone_type=getattr(args, "backbone", "maskrcnn"))


def build(args):
    """reference :692-739: `models.build_model(args)` -> (model, criterion, matcher) (models/__init__.py:10, called at
    main.py:205 and inference_tools/inference_engine.py:31).  The backbone comes from `args.backbone_module` when set,
    else from the reference's `models.backbone.build_backbone` if that package is importable."""
    from .criterion import build_criterion
    backbone = getattr(args, "backbone_module", None)
    if backbone is None:
        try:
            from models.backbone import build_backbone           # the reference's detector, unmodified
            backbone = build_backbone(args)
        except ImportError:
            backbone = None
    model = build_poet(args, backbone)
    criterion, matcher = build_criterion(args)
    dev = getattr(args, "device", None)
    if dev is not None:
        criterion.to(torch.device(dev))
    return model, criterion, matcher


def build_model(args):
    """models/__init__.py:10."""
    return build(args)
