"""poet_b200: B200-native (sm_100a) implementation of PoET's deformable encoder/decoder hot path.

Public surface (mirrors the reference's modules for this path):
  poet_b200.deformable_attention.MSDeformAttn
  poet_b200.deformable_transformer.DeformableTransformer / build_deforamble_transformer
  poet_b200.position_encoding.PositionEmbeddingSine / BoundingBoxEmbeddingSine
  poet_b200.pose_estimation_transformer.PoET / MLP / build
All numerics run in libpoet_b200.so (C ABI: include/poet_b200.h); there is no CPU fallback.
"""
import sys as _sys

__version__ = "0.1.0"


def install_into_reference() -> None:
    """Register the MSDeformAttn seam so the unmodified reference imports our CUDA op
    (`from deformable_attention import MSDeformAttn`, reference models/deformable_transformer.py:24)."""
    from . import deformable_attention
    _sys.modules["deformable_attention"] = deformable_attention


def build_model(args):
    """Seam B-py2: drop-in for the reference's `models.build_model(args)` -> (model, criterion, matcher)."""
    from .pose_estimation_transformer import build
    return build(args)
