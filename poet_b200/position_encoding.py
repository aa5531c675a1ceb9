"""Host-side mirror of reference models/position_encoding.py on the poet_b200 kernels.

PositionEmbeddingSine  <- position_encoding.py:24-60   (kernel: poet_posenc_sine)
BoundingBoxEmbeddingSine <- position_encoding.py:63-84 (kernel: poet_bbox_embed_pad)
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import ops


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats: int = 64, temperature: float = 10000, normalize: bool = False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats = int(num_pos_feats)
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi if scale is None else scale

    def forward(self, tensor_list) -> torch.Tensor:
        """tensor_list: anything with a `.mask` [B,H,W] bool (the reference's NestedTensor) -> [B,2F,H,W]."""
        mask = tensor_list.mask
        if mask is None:
            raise ValueError("PositionEmbeddingSine needs a padding mask")
        return ops.posenc_sine_nchw(mask, self.num_pos_feats, self.temperature, self.normalize, self.scale)


class BoundingBoxEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=32):
        super().__init__()
        self.num_pos_feats = num_pos_feats

    def forward(self, bboxes: torch.Tensor) -> torch.Tensor:
        """bboxes [n,4] (cx,cy,w,h) of one image -> [n, 8F]."""
        n = bboxes.shape[0]
        F = int(self.num_pos_feats)
        if n == 0:
            return bboxes.new_zeros((0, 8 * F))
        cnt = torch.full((1,), n, dtype=torch.int32, device=bboxes.device)
        both = ops.bbox_embed_pad(bboxes.float().reshape(1, n, 4), cnt, F)        # [1,n,16F] = [emb|emb]
        return both[0, :, : 8 * F]


def build_position_encoding(args):
    n_steps = args.hidden_dim // 2
    if args.position_embedding in ("v2", "sine"):
        return PositionEmbeddingSine(n_steps, normalize=True)
    raise NotImplementedError(f"position embedding '{args.position_embedding}' is outside the poet_b200 hot path "
                              "(only 'sine' is used by the PoET configs)")
