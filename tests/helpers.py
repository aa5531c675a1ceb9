"""Shared test helpers: golden loading, oracle drivers, sampled-gradient comparison."""
import os

import torch

from oracle import poet_oracle as O
from poet_b200 import synthetic as S

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(key: str):
    return torch.load(os.path.join(GOLDEN_DIR, key.replace("/", "__") + ".pt"), weights_only=False)["data"]


def sample_indices(numel: int, k: int = 256, seed: int = 7):
    g = torch.Generator().manual_seed(seed + numel)           # mirrors oracle/make_golden.py
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def image_mask_for(cfg, pad):
    H0, W0 = S.pyramid_of(cfg)[0]
    m = torch.zeros(cfg["batch"], H0 * 16, W0 * 16, dtype=torch.bool)
    if pad:
        m[1::2, :, (W0 - max(1, W0 // 8)) * 16:] = True
    return m


def oracle_poet_from_feats(cfg, pad, dtype=torch.float32, need_grad=True):
    """Mirror of make_golden.golden_poet on the oracle: stub-backbone feats -> input_proj -> path."""
    P = {k: v.to(dtype).requires_grad_(need_grad) for k, v in S.make_params(cfg, with_input_proj=True).items()}
    inp = S.make_inputs(cfg, pad_columns=pad)
    feats = [f.to(dtype).requires_grad_(need_grad) for f in inp["srcs"][:3]]
    srcs, masks, _ = O.input_proj(P, cfg, feats, inp["masks"][:3], image_mask_for(cfg, pad))
    cap = {}
    out, n_boxes = O.poet_path_forward(P, cfg, srcs, masks, inp["boxes"], inp["labels"], capture=cap)
    return P, feats, srcs, masks, inp, out, n_boxes, cap


def same_fingerprint(a, b, rel=1e-9):
    """Fingerprints are fp64 sums whose association order depends on the host's thread count."""
    import math
    return len(a) == len(b) and all(math.isclose(x, y, rel_tol=rel, abs_tol=1e-9) for x, y in zip(a, b))
