"""Seeded synthetic workloads for the PoET hot path.

Everything is drawn from CPU ``torch.Generator``s so that the CPU oracle, the golden
fixtures and the CUDA path see bit-identical inputs and weights.  Shapes follow what the
reference really feeds its transformer (measured in SURVEY.md fact 5): at 640x480 the
Mask R-CNN path yields the pyramid [(30,40),(15,20),(8,10),(4,5)] (S = 1600).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

PYRAMIDS = {
    "REF640": [(30, 40), (15, 20), (8, 10), (4, 5)],        # S = 1600
    "REF1280": [(60, 80), (30, 40), (15, 20), (8, 10)],     # S = 6380
    "S8_640": [(60, 80), (30, 40), (15, 20), (8, 10)],      # nominal stride-8 variant
    "TINY": [(6, 8), (3, 4), (2, 2), (1, 1)],               # unit tests
}


def _cfg(**kw):
    base = dict(d_model=256, dim_ff=1024, n_levels=4, n_points=4, class_mode="specific",
                rotation_mode="6d", aux_loss=True, pyramid="REF640")
    base.update(kw)
    return base


# BASELINE.json `configs`, in order (cfg1..cfg5); "tiny*" are unit-test sized.
CONFIGS: Dict[str, dict] = {
    "cfg1": _cfg(batch=1, enc_layers=2, dec_layers=2, nheads=8, num_queries=5, n_classes=21),
    "cfg2": _cfg(batch=16, enc_layers=5, dec_layers=5, nheads=16, num_queries=10, n_classes=21),
    "cfg3": _cfg(batch=16, enc_layers=5, dec_layers=5, nheads=16, num_queries=10, n_classes=8),
    "cfg4": _cfg(batch=16, enc_layers=5, dec_layers=5, nheads=16, num_queries=10, n_classes=21),
    "cfg5": _cfg(batch=4, enc_layers=6, dec_layers=6, nheads=8, num_queries=25, n_classes=21,
                 pyramid="REF1280"),
    "tiny": _cfg(batch=2, enc_layers=1, dec_layers=1, nheads=8, num_queries=4, n_classes=3,
                 pyramid="TINY"),
    "tiny16": _cfg(batch=2, enc_layers=2, dec_layers=2, nheads=16, num_queries=6, n_classes=5,
                   pyramid="TINY"),
    "cfg2_b2": _cfg(batch=2, enc_layers=5, dec_layers=5, nheads=16, num_queries=10, n_classes=21),
    "cfg3_b2": _cfg(batch=2, enc_layers=5, dec_layers=5, nheads=16, num_queries=10, n_classes=8),       # LM-O heads 27/54
    "cfg5_b1": _cfg(batch=1, enc_layers=6, dec_layers=6, nheads=8, num_queries=25, n_classes=21,
                    pyramid="REF1280"),                                                                  # 1280x960, S=6380
}


def pyramid_of(cfg: dict) -> List[Tuple[int, int]]:
    return list(PYRAMIDS[cfg["pyramid"]]) if isinstance(cfg["pyramid"], str) else list(cfg["pyramid"])


def n_tokens(cfg: dict) -> int:
    return sum(h * w for h, w in pyramid_of(cfg))


# --------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------
def make_inputs(cfg: dict, seed: int = 1234, pad_columns: bool = False, batch: int | None = None):
    """Returns dict(srcs, masks, boxes, labels).  srcs[l] ~ N(0,1) [B,C,H_l,W_l]; masks False
    (``pad_columns`` pads the right ~1/8 of each level's columns on odd images, to exercise
    valid_ratios); image i carries Q - (i mod 4) boxes, cx,cy~U(.25,.75), w,h~U(.05,.30)."""
    g = torch.Generator().manual_seed(seed)
    B = cfg["batch"] if batch is None else batch
    C, Q = cfg["d_model"], cfg["num_queries"]
    srcs, masks = [], []
    for (h, w) in pyramid_of(cfg):
        srcs.append(torch.randn(B, C, h, w, generator=g))
        m = torch.zeros(B, h, w, dtype=torch.bool)
        if pad_columns:
            keep = max(1, w - max(1, w // 8))
            m[1::2, :, keep:] = True
        masks.append(m)
    boxes, labels = [], []
    for i in range(B):
        n = max(1, Q - (i % 4))
        cxy = torch.rand(n, 2, generator=g) * 0.5 + 0.25
        wh = torch.rand(n, 2, generator=g) * 0.25 + 0.05
        boxes.append(torch.cat((cxy, wh), 1))
        labels.append(torch.randint(1, cfg["n_classes"] + 1, (n,), generator=g))
    return dict(srcs=srcs, masks=masks, boxes=boxes, labels=labels)


def make_cotangents(cfg: dict, seed: int = 4321, batch: int | None = None):
    g = torch.Generator().manual_seed(seed)
    B = cfg["batch"] if batch is None else batch
    Dl, Q = cfg["dec_layers"], cfg["num_queries"]
    return torch.randn(Dl, B, Q, 3, generator=g), torch.randn(Dl, B, Q, 3, 3, generator=g)


# --------------------------------------------------------------------------------------
# parameters (reference state_dict key names)
# --------------------------------------------------------------------------------------
def param_shapes(cfg: dict, with_input_proj: bool = False) -> "OrderedDict[str, Tuple[int, ...]]":
    C, F_, M, L, Pn = cfg["d_model"], cfg["dim_ff"], cfg["nheads"], cfg["n_levels"], cfg["n_points"]
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def lin(name, o, i):
        out[name + ".weight"] = (o, i)
        out[name + ".bias"] = (o,)

    def ln(name):
        out[name + ".weight"] = (C,)
        out[name + ".bias"] = (C,)

    def msda(name):
        lin(name + ".sampling_offsets", M * L * Pn * 2, C)
        lin(name + ".attention_weights", M * L * Pn, C)
        lin(name + ".value_proj", C, C)
        lin(name + ".output_proj", C, C)

    for i in range(cfg["enc_layers"]):
        p = f"transformer.encoder.layers.{i}"
        msda(p + ".self_attn"); ln(p + ".norm1"); lin(p + ".linear1", F_, C); lin(p + ".linear2", C, F_); ln(p + ".norm2")
    for i in range(cfg["dec_layers"]):
        p = f"transformer.decoder.layers.{i}"
        msda(p + ".cross_attn"); ln(p + ".norm1")
        out[p + ".self_attn.in_proj_weight"] = (3 * C, C)
        out[p + ".self_attn.in_proj_bias"] = (3 * C,)
        lin(p + ".self_attn.out_proj", C, C)
        ln(p + ".norm2"); lin(p + ".linear1", F_, C); lin(p + ".linear2", C, F_); ln(p + ".norm3")
    out["transformer.level_embed"] = (L, C)
    lin("transformer.reference_points", 2, C)
    n_slots = (cfg["n_classes"] + 1) if cfg["class_mode"] == "specific" else 1
    for head, dim in (("translation_head", 3), ("rotation_head", 6)):
        for i in range(cfg["dec_layers"]):
            lin(f"{head}.{i}.layers.0", C, C)
            lin(f"{head}.{i}.layers.1", C, C)
            lin(f"{head}.{i}.layers.2", dim * n_slots, C)
    if with_input_proj:
        for l in range(L):
            k = 1 if l < 3 else 3
            out[f"input_proj.{l}.0.weight"] = (C, C, k, k)
            out[f"input_proj.{l}.0.bias"] = (C,)
            out[f"input_proj.{l}.1.weight"] = (C,)
            out[f"input_proj.{l}.1.bias"] = (C,)
    return out


def msda_directional_bias(M: int, L: int, Pn: int) -> torch.Tensor:
    """Upstream MSDeformAttn._reset_parameters offset bias: head m looks along angle 2*pi*m/M,
    normalised to the unit square, point p at distance p+1."""
    th = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
    g = torch.stack([th.cos(), th.sin()], -1)
    g = (g / g.abs().max(-1, keepdim=True)[0]).view(M, 1, 1, 2).repeat(1, L, Pn, 1)
    g = g * torch.arange(1, Pn + 1, dtype=torch.float32).view(1, 1, Pn, 1)
    return g.reshape(-1)


def make_params(cfg: dict, seed: int = 42, with_input_proj: bool = False,
                dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic 'trained-like' weights: xavier-uniform matrices, N(0,.02) biases,
    LayerNorm/GroupNorm affine 1+N(0,.1) / N(0,.1), MSDA offsets = directional bias with
    N(0,.02)-perturbed weights and attention logits (SURVEY.md §4 trap: fresh-init zeros would
    not exercise those projections or the softmax)."""
    g = torch.Generator().manual_seed(seed)
    M, L, Pn = cfg["nheads"], cfg["n_levels"], cfg["n_points"]
    P: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in param_shapes(cfg, with_input_proj).items():
        leaf = name.rsplit(".", 1)[-1]
        is_norm = ".norm" in name or (name.startswith("input_proj") and ".1." in name)
        if is_norm:
            t = torch.randn(shape, generator=g) * 0.1 + (1.0 if leaf == "weight" else 0.0)
        elif "sampling_offsets" in name:
            t = torch.randn(shape, generator=g) * 0.02
            if leaf == "bias":
                t = t * 0 + msda_directional_bias(M, L, Pn)
        elif "attention_weights" in name:
            t = torch.randn(shape, generator=g) * 0.02
        elif name == "transformer.level_embed":
            t = torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_out = shape[0] * (shape[2] * shape[3] if len(shape) == 4 else 1)
            fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            t = torch.randn(shape, generator=g) * 0.02
        P[name] = t.to(dtype)
    return P


def fingerprint(tensors) -> List[float]:
    """Cheap cross-machine check that seeded generation reproduced the same bits."""
    out = []
    for t in tensors:
        f = t.detach().double().flatten()
        out.append(float(f.sum()))
        out.append(float((f * torch.arange(1, f.numel() + 1, dtype=torch.float64)).sum() / f.numel()))
    return out
