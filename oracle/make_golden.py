"""Generate tests/golden/*.pt by running the UNMODIFIED reference (read-only, /root/reference).

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

The reference's MSDeformAttn op is an un-vendored third-party package
(``from deformable_attention import MSDeformAttn``, models/deformable_transformer.py:24), so a
``deformable_attention`` module implementing upstream's published pure-PyTorch formulation
(``oracle.poet_oracle.msda_core``: grid_sample bilinear / zeros / align_corners=False) is
injected; every other line executed is the reference's own code: ``models.position_encoding``,
``models.deformable_transformer.DeformableTransformer`` and
``models.pose_estimation_transformer.PoET`` (forward incl. query construction, input_proj,
class-specific head select, 6D->R).  The frozen Mask R-CNN backbone is replaced by a stub that
returns seeded feature maps (it is outside the hot path).

Fixtures hold only outputs + fingerprints; inputs and weights are re-generated from seeds by
``poet_b200.synthetic`` (CPU torch.Generator => bit-identical wherever this image runs).
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import types

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import poet_oracle as O            # noqa: E402
from poet_b200 import synthetic as S           # noqa: E402


# --------------------------------------------------------------------------------------
def install_shim():
    """deformable_attention.MSDeformAttn with upstream's module semantics on the grid_sample core."""
    mod = types.ModuleType("deformable_attention")

    class MSDeformAttn(nn.Module):
        def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
            super().__init__()
            self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
            self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
            self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
            self.value_proj = nn.Linear(d_model, d_model)
            self.output_proj = nn.Linear(d_model, d_model)
            self._reset_parameters()

        def _reset_parameters(self):
            P = {k: v.data for k, v in self.named_parameters()}
            O.msda_reset_parameters(P, "", self.n_heads, self.n_levels, self.n_points)

        def forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                    input_level_start_index, input_padding_mask=None):
            P = dict(self.named_parameters())
            shapes = [tuple(int(v) for v in hw) for hw in input_spatial_shapes.tolist()]
            return O.msda_module(P, "", query, reference_points, input_flatten, shapes,
                                 input_padding_mask, self.n_heads, self.n_points)

    mod.MSDeformAttn = MSDeformAttn
    sys.modules["deformable_attention"] = mod
    if REF not in sys.path:
        sys.path.insert(0, REF)


def ref_args(cfg):
    return types.SimpleNamespace(hidden_dim=cfg["d_model"], nheads=cfg["nheads"], enc_layers=cfg["enc_layers"],
                                 dec_layers=cfg["dec_layers"], dim_feedforward=cfg["dim_ff"], dropout=0.0,
                                 num_feature_levels=cfg["n_levels"], dec_n_points=cfg["n_points"],
                                 enc_n_points=cfg["n_points"])


def sample_indices(numel: int, k: int = 256, seed: int = 7):
    g = torch.Generator().manual_seed(seed + numel)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def grad_digest(named_grads):
    """name -> (sampled indices are re-derivable) sampled values + l2 norm."""
    out = {}
    for name, gr in named_grads.items():
        if gr is None:
            out[name] = None
            continue
        flat = gr.detach().flatten()
        out[name] = dict(samples=flat[sample_indices(flat.numel())].clone(), norm=float(flat.double().norm()))
    return out


# --------------------------------------------------------------------------------------
def golden_posenc():
    from models.position_encoding import PositionEmbeddingSine, BoundingBoxEmbeddingSine
    from util.misc import NestedTensor
    out = {}
    pe = PositionEmbeddingSine(128, normalize=True)
    for name, (B, H, W) in {"6x8": (3, 6, 8), "15x20": (2, 15, 20), "4x5": (2, 4, 5)}.items():
        mask = torch.zeros(B, H, W, dtype=torch.bool)
        mask[1, :, W - max(1, W // 4):] = True
        if B > 2:
            mask[2, H - 2:, :] = True
        out[f"pos_{name}"] = dict(mask=mask, pos=pe(NestedTensor(torch.zeros(B, 1, H, W), mask)))
    g = torch.Generator().manual_seed(99)
    boxes = torch.cat((torch.rand(7, 2, generator=g) * 0.5 + 0.25, torch.rand(7, 2, generator=g) * 0.25 + 0.05), 1)
    boxes = torch.cat((boxes, torch.tensor([[0.5, 0.5, 1.0, 1.0], [0.0, 1.0, 0.123456, 0.999]])), 0)
    out["bbox"] = dict(boxes=boxes, embed=BoundingBoxEmbeddingSine(num_pos_feats=256 / 8)(boxes))
    return out


def golden_transformer(cfg_name: str, pad: bool):
    from models.deformable_transformer import build_deforamble_transformer
    cfg = S.CONFIGS[cfg_name]
    torch.manual_seed(0)
    model = build_deforamble_transformer(ref_args(cfg)).eval()
    P = S.make_params(cfg)
    sd = {k[len("transformer."):]: v for k, v in P.items() if k.startswith("transformer.")}
    model.load_state_dict(sd, strict=True)
    inp = S.make_inputs(cfg, pad_columns=pad)
    pos = [O.sine_position_embedding(m, cfg["d_model"] // 2) for m in inp["masks"]]
    qe, pboxes, pcls, _ = O.build_queries(inp["boxes"], inp["labels"], cfg["num_queries"], cfg["d_model"])
    mem_box = {}
    h = model.encoder.register_forward_hook(lambda m, i, o: mem_box.__setitem__("memory", o.detach()))
    with torch.no_grad():
        hs, init_ref, inter_ref, _, _ = model(inp["srcs"], inp["masks"], pos, qe, pboxes[:, :, :2])
    h.remove()
    stride = 8 if S.n_tokens(cfg) > 200 else 1
    return dict(cfg=cfg_name, pad=pad, hs=hs, init_ref=init_ref, inter_ref=inter_ref,
                memory_rows=mem_box["memory"][:, ::stride].clone(), memory_stride=stride,
                fp_inputs=S.fingerprint(inp["srcs"]), fp_params=S.fingerprint([P["transformer.level_embed"],
                                                                              P["transformer.encoder.layers.0.linear1.weight"]]))


class _StubDetector(nn.Module):
    """Stands in for MaskRCNNBackbone: returns seeded feature maps as NestedTensors."""
    def __init__(self, feats, masks, predictions=None):
        super().__init__()
        self.feats, self.masks, self.predictions = feats, masks, predictions
        self.train_backbone = False
        self.strides = [8, 16, 32]
        self.num_channels = [feats[0].shape[1]] * 3

    def forward(self, samples):
        from util.misc import NestedTensor
        return self.predictions, {str(i): NestedTensor(f, m) for i, (f, m) in enumerate(zip(self.feats, self.masks))}


def golden_poet(cfg_name: str, pad: bool):
    """Full reference PoET.forward (+ backward through a fixed-cotangent loss) on stub backbone features."""
    from models.backbone import Joiner
    from models.position_encoding import PositionEmbeddingSine
    from models.deformable_transformer import build_deforamble_transformer
    from models.pose_estimation_transformer import PoET
    from util.misc import NestedTensor
    cfg = S.CONFIGS[cfg_name]
    inp = S.make_inputs(cfg, pad_columns=pad)
    feats = [f.clone().requires_grad_(True) for f in inp["srcs"][:3]]       # backbone maps = first 3 pyramid levels
    fmasks = inp["masks"][:3]
    H0, W0 = feats[0].shape[-2:]
    image_mask = torch.zeros(cfg["batch"], H0 * 16, W0 * 16, dtype=torch.bool)
    if pad:
        image_mask[1::2, :, (W0 - max(1, W0 // 8)) * 16:] = True
    torch.manual_seed(0)
    joiner = Joiner(_StubDetector(feats, fmasks), PositionEmbeddingSine(cfg["d_model"] // 2, normalize=True))
    model = PoET(joiner, build_deforamble_transformer(ref_args(cfg)), num_queries=cfg["num_queries"],
                 num_feature_levels=cfg["n_levels"], n_classes=cfg["n_classes"], bbox_mode="gt",
                 ref_points_mode="bbox", query_embedding_mode="bbox", rotation_mode="6d",
                 class_mode=cfg["class_mode"], aleatoric=False, aux_loss=True, backbone_type="maskrcnn").eval()
    P = S.make_params(cfg, with_input_proj=True)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected and all(k.startswith("backbone.") for k in missing), (missing, unexpected)
    targets = [dict(boxes=b, labels=l) for b, l in zip(inp["boxes"], inp["labels"])]
    samples = NestedTensor(torch.zeros(cfg["batch"], 3, image_mask.shape[1], image_mask.shape[2]), image_mask)
    out, n_boxes = model(samples, targets)
    t_all = torch.stack([a["pred_translation"] for a in out["aux_outputs"]] + [out["pred_translation"]])
    R_all = torch.stack([a["pred_rotation"] for a in out["aux_outputs"]] + [out["pred_rotation"]])
    g_t, g_R = S.make_cotangents(cfg)
    loss = O.synthetic_loss((t_all, R_all), g_t, g_R)
    loss.backward()
    named = {k: p.grad for k, p in model.named_parameters() if not k.startswith("backbone.")}
    for l, f in enumerate(feats):
        named[f"__feat{l}"] = f.grad
    return dict(cfg=cfg_name, pad=pad, translation=t_all.detach(), rotation=R_all.detach(),
                pred_boxes=out["pred_boxes"].detach(), pred_classes=out["pred_classes"].detach(),
                n_boxes=n_boxes, loss=float(loss), grads=grad_digest(named), image_mask_shape=tuple(image_mask.shape),
                fp_inputs=S.fingerprint(inp["srcs"][:3]))


def backbone_predictions(cfg, image_hw, seed=321):
    """Detector output per image for the 'backbone' bbox mode: [n, 6] = (x1, y1, x2, y2, score, class) in pixels, or None.
    Image 0: nothing detected; image 1: fewer than Q boxes; image 2: more than Q (top-Q by score is taken)."""
    g = torch.Generator().manual_seed(seed)
    Q, (H, W) = cfg["num_queries"], image_hw
    preds = [None]
    for n in (max(1, Q - 2), Q + 3):
        cxy = torch.rand(n, 2, generator=g) * 0.5 + 0.25
        wh = torch.rand(n, 2, generator=g) * 0.25 + 0.05
        x1y1, x2y2 = (cxy - wh / 2) * torch.tensor([W, H]), (cxy + wh / 2) * torch.tensor([W, H])
        score = torch.rand(n, 1, generator=g)
        cls = torch.randint(1, cfg["n_classes"] + 1, (n, 1), generator=g).float()
        preds.append(torch.cat((x1y1, x2y2, score, cls), 1))
    return preds


def golden_poet_backbone_mode(cfg_name: str = "tiny16"):
    """Reference PoET.forward in inference mode (bbox_mode='backbone', targets=None): query construction from detector
    output (pose_estimation_transformer.py:240-305: xyxy -> normalised cxcywh, top-Q by score, the `predictions is None`
    branch) + the path.  SURVEY.md section 8f N4."""
    from models.backbone import Joiner
    from models.position_encoding import PositionEmbeddingSine
    from models.deformable_transformer import build_deforamble_transformer
    from models.pose_estimation_transformer import PoET
    from util.misc import NestedTensor
    cfg = dict(S.CONFIGS[cfg_name], batch=3)
    inp = S.make_inputs(cfg, pad_columns=False)
    feats, fmasks = inp["srcs"][:3], inp["masks"][:3]
    H0, W0 = feats[0].shape[-2:]
    image_mask = torch.zeros(cfg["batch"], H0 * 16, W0 * 16, dtype=torch.bool)
    preds = backbone_predictions(cfg, (H0 * 16, W0 * 16))
    torch.manual_seed(0)
    joiner = Joiner(_StubDetector(feats, fmasks, preds), PositionEmbeddingSine(cfg["d_model"] // 2, normalize=True))
    model = PoET(joiner, build_deforamble_transformer(ref_args(cfg)), num_queries=cfg["num_queries"],
                 num_feature_levels=cfg["n_levels"], n_classes=cfg["n_classes"], bbox_mode="backbone",
                 ref_points_mode="bbox", query_embedding_mode="bbox", rotation_mode="6d",
                 class_mode=cfg["class_mode"], aleatoric=False, aux_loss=True, backbone_type="maskrcnn").eval()
    P = S.make_params(cfg, with_input_proj=True)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected and all(k.startswith("backbone.") for k in missing), (missing, unexpected)
    samples = NestedTensor(torch.zeros(cfg["batch"], 3, image_mask.shape[1], image_mask.shape[2]), image_mask)
    with torch.no_grad():
        out, n_boxes = model(samples, None)
    t_all = torch.stack([a["pred_translation"] for a in out["aux_outputs"]] + [out["pred_translation"]])
    R_all = torch.stack([a["pred_rotation"] for a in out["aux_outputs"]] + [out["pred_rotation"]])
    return dict(cfg=cfg_name, batch=3, translation=t_all, rotation=R_all, pred_boxes=out["pred_boxes"],
                pred_classes=out["pred_classes"], n_boxes=n_boxes)


def criterion_case(seed=99, L=3, B=4, Q=6, n_boxes=(6, 3, 1, 4)):
    """Seeded inputs of the criterion fixture (shared with the tests through poet_b200.synthetic-style seeding)."""
    g = torch.Generator().manual_seed(seed)
    t_all = torch.randn(L, B, Q, 3, generator=g)
    R_all = O.rotation_6d_to_matrix(torch.randn(L, B, Q, 6, generator=g))
    boxes, labels, tgt_t, tgt_R = [], [], [], []
    for n in n_boxes:
        boxes.append(torch.rand(n, 4, generator=g) * 0.5 + 0.2)
        labels.append(torch.randint(1, 22, (n,), generator=g))
        tgt_t.append(torch.randn(n, 3, generator=g))
        tgt_R.append(O.rotation_6d_to_matrix(torch.randn(n, 6, generator=g)))
    return t_all, R_all, boxes, labels, tgt_t, tgt_R, list(n_boxes)


def golden_criterion():
    """Reference SetCriterion + PoseMatcher(bbox_mode='gt') on seeded predictions / targets: loss dict and the
    gradients of the weighted total w.r.t. every layer's predictions."""
    from models.matcher import PoseMatcher
    from models.pose_estimation_transformer import SetCriterion
    t_all, R_all, boxes, labels, tgt_t, tgt_R, n_boxes = criterion_case()
    L, B, Q = t_all.shape[:3]
    t_all.requires_grad_(True)
    R_all.requires_grad_(True)
    pb = torch.full((B, Q, 4), -1.0)
    pc = torch.full((B, Q), -1, dtype=torch.int64)
    for b, n in enumerate(n_boxes):
        pb[b, :n], pc[b, :n] = boxes[b], labels[b]
    targets = [dict(boxes=boxes[b], labels=labels[b], relative_position=tgt_t[b], relative_rotation=tgt_R[b]) for b in range(B)]
    w = {"loss_trans": 2.0, "loss_rot": 0.5}
    weight_dict = dict(w)
    for i in range(L - 1):
        weight_dict.update({k + f"_{i}": v for k, v in w.items()})
    crit = SetCriterion(PoseMatcher(cost_bbox=1, cost_class=1, bbox_mode="gt", class_mode="specific"), weight_dict,
                        ["translation", "rotation"])
    mk = lambda l: {"pred_translation": t_all[l], "pred_rotation": R_all[l], "pred_boxes": pb, "pred_classes": pc}
    outputs = mk(L - 1)
    outputs["aux_outputs"] = [mk(l) for l in range(L - 1)]
    losses = crit(outputs, targets, n_boxes)
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)          # engine.py:60-61
    total.backward()
    return dict(losses={k: float(v) for k, v in losses.items()}, total=float(total), weights=w,
                grad_t=t_all.grad.clone(), grad_R=R_all.grad.clone())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default=None, help="regenerate only the fixtures whose key starts with this prefix")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    install_shim()
    torch.set_num_threads(os.cpu_count() or 1)
    gold = {}
    if args.only is None or "posenc".startswith(args.only):
        gold["posenc"] = golden_posenc()
    for name, pad in (("tiny", False), ("tiny", True), ("tiny16", True), ("cfg1", False)):
        if args.only is None or f"transformer/{name}".startswith(args.only):
            gold[f"transformer/{name}/pad{int(pad)}"] = golden_transformer(name, pad)
    for name, pad in (("tiny", True), ("tiny16", False), ("cfg1", False), ("cfg2_b2", True)):
        if args.only is None or f"poet/{name}".startswith(args.only):
            gold[f"poet/{name}/pad{int(pad)}"] = golden_poet(name, pad)
    if args.only is None or "poet_backbone_mode/tiny16".startswith(args.only):
        gold["poet_backbone_mode/tiny16"] = golden_poet_backbone_mode("tiny16")
    if args.only is None or "criterion/gt".startswith(args.only):
        gold["criterion/gt"] = golden_criterion()
    meta = dict(torch=torch.__version__, reference=REF, note="generated by oracle/make_golden.py")
    for key, val in gold.items():
        fn = os.path.join(args.out, key.replace("/", "__") + ".pt")
        torch.save(dict(meta=meta, data=val), fn)
        print(f"{fn}: {os.path.getsize(fn) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
