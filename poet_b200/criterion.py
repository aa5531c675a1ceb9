"""Pose loss on the device: the reference's SetCriterion (translation L2 + rotation geodesic,
final + auxiliary decoder layers; models/pose_estimation_transformer.py:455-674) under the PoseMatcher assignment
(models/matcher.py:104-229), as ONE kernel launch with no host synchronisation.

In bbox_mode 'gt' / 'jitter' the first n_i queries of image i carry the target boxes in target order, so the
Hungarian assignment is the identity on them (matcher.py:169-183: zero-cost diagonal) and is built on the device
from the box counts.  For other modes pass an explicit `assign` tensor computed by any matcher.

    crit = PoseCriterion(weight_dict={'loss_trans': 1.0, 'loss_rot': 1.0})
    losses, total = crit(outputs, tgt_t, tgt_R, n_boxes_dev)        # same keys as the reference's loss dict
    total.backward()
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops


class _PoseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t_all, R_all, tgt_t, tgt_R, assign, n_obj, w_trans, w_rot):
        t_all, R_all = ops._chk(t_all), ops._chk(R_all)
        L, B, Q = t_all.shape[:3]
        T = tgt_t.shape[1]
        losses = torch.empty((L, 2), device=t_all.device, dtype=torch.float32)
        gt, gR = torch.empty_like(t_all), torch.empty_like(R_all)
        ops._call("poet_pose_loss", t_all.data_ptr(), R_all.data_ptr(), ops._chk(tgt_t).data_ptr(), ops._chk(tgt_R).data_ptr(),
                  ops._chk(assign, torch.int32).data_ptr(), ops._chk(n_obj, torch.int32).data_ptr(), losses.data_ptr(),
                  gt.data_ptr(), gR.data_ptr(), L, B, Q, T, float(w_trans), float(w_rot), ops._stream(t_all))
        ctx.save_for_backward(gt, gR)
        total = (losses[:, 0] * w_trans + losses[:, 1] * w_rot).sum()
        ctx.mark_non_differentiable(losses)
        return total, losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        gt, gR = ctx.saved_tensors
        return gt * g_total, gR * g_total, None, None, None, None, None, None


class PoseCriterion(torch.nn.Module):
    def __init__(self, weight_dict: Optional[Dict[str, float]] = None):
        super().__init__()
        wd = weight_dict or {"loss_trans": 1.0, "loss_rot": 1.0}
        self.w_trans, self.w_rot = float(wd["loss_trans"]), float(wd["loss_rot"])

    @staticmethod
    def identity_assignment(n_boxes_dev: torch.Tensor, Q: int) -> torch.Tensor:
        """[B] int32 box counts -> [B,Q] int32: j for j < n_i else -1 (PoseMatcher 'gt' mode on the device)."""
        j = torch.arange(Q, device=n_boxes_dev.device, dtype=torch.int32)[None, :]
        return torch.where(j < n_boxes_dev[:, None].to(torch.int32), j, torch.full_like(j, -1)).contiguous()

    def forward(self, outputs: dict, tgt_t: torch.Tensor, tgt_R: torch.Tensor, n_boxes_dev: torch.Tensor,
                assign: Optional[torch.Tensor] = None):
        """outputs: the model's dict (pred_translation [B,Q,3], pred_rotation [B,Q,3,3], aux_outputs);
        tgt_t [B,T,3] / tgt_R [B,T,3,3]: padded targets; n_boxes_dev [B] int32.  Returns (loss dict with the
        reference's keys, weighted total)."""
        layers = list(outputs.get("aux_outputs", [])) + [outputs]
        t_all = torch.stack([o["pred_translation"] for o in layers])
        R_all = torch.stack([o["pred_rotation"] for o in layers])
        L, B, Q = t_all.shape[:3]
        if assign is None:
            assign = self.identity_assignment(n_boxes_dev, Q)
        n_obj = (assign >= 0).sum().to(torch.int32).reshape(1)
        total, losses = _PoseLoss.apply(t_all, R_all.reshape(L, B, Q, 9), tgt_t, tgt_R.reshape(B, -1, 9), assign, n_obj,
                                        self.w_trans, self.w_rot)
        out = {}
        for l in range(L):
            suffix = "" if l == L - 1 else f"_{l}"
            out["loss_trans" + suffix] = losses[l, 0]
            out["loss_rot" + suffix] = losses[l, 1]
        return out, total
