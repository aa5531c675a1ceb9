"""Seam B-py2 end to end on the GPU: `poet_b200.build_model(args)` driven by the exact call sequence of the reference's
engine.train_one_epoch (engine.py:38,55-81) -- model.train(); model(samples, targets); criterion(outputs, targets,
n_boxes); weighted sum over criterion.weight_dict; optimizer.zero_grad(); backward; clip; optimizer.step() -- with the
reference default dropout 0.1, plus StepLR (main.py:278) and a checkpoint save / resume (main.py:287-317, 357-369).
The reference package itself is not on the GPU box, so the loop is restated here line for line; the structural half of
the seam (our classes inside the unmodified reference code) is tests/test_seams.py."""
import io
import math
import types

import pytest
import torch

from helpers import load_golden
from oracle import poet_oracle as O
from oracle.make_golden import criterion_case
from poet_b200 import synthetic as S
from test_gpu_model import _StubBackbone

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _args(cfg, **kw):
    a = types.SimpleNamespace(
        hidden_dim=cfg["d_model"], nheads=cfg["nheads"], enc_layers=cfg["enc_layers"], dec_layers=cfg["dec_layers"],
        dim_feedforward=cfg["dim_ff"], dropout=0.1, num_feature_levels=cfg["n_levels"], dec_n_points=cfg["n_points"],
        enc_n_points=cfg["n_points"], num_queries=cfg["num_queries"], n_classes=cfg["n_classes"], bbox_mode="gt",
        reference_points="bbox", query_embedding="bbox", rotation_representation="6d", class_mode=cfg["class_mode"],
        aleatoric=False, aux_loss=True, backbone="maskrcnn", matcher_type="pose", set_cost_class=1, set_cost_bbox=1,
        translation_loss_coef=1.0, rotation_loss_coef=1.0, device=DEV, lr=2e-4, lr_backbone=2e-5, lr_linear_proj_mult=0.1,
        weight_decay=1e-4, clip_max_norm=0.1, lr_drop=2)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _random_rotations(n, gen):
    q = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1)
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).view(n, 3, 3)


def _setup(cfg, seed=11):
    from poet_b200.pose_estimation_transformer import _Nested
    import poet_b200
    inp = S.make_inputs(cfg, pad_columns=True)
    H0, W0 = inp["srcs"][0].shape[-2:]
    bb = _StubBackbone(cfg["d_model"], [f.to(DEV) for f in inp["srcs"][:3]], [m.to(DEV) for m in inp["masks"][:3]])
    model, criterion, matcher = poet_b200.build_model(_args(cfg, backbone_module=bb))
    model.load_state_dict(S.make_params(cfg, with_input_proj=True), strict=True)
    model.to(DEV)
    gen = torch.Generator().manual_seed(seed)
    targets = []
    for b, l in zip(inp["boxes"], inp["labels"]):
        n = b.shape[0]
        targets.append({"boxes": b.to(DEV), "labels": l.to(DEV), "relative_position": torch.randn(n, 3, generator=gen).to(DEV),
                        "relative_rotation": _random_rotations(n, gen).to(DEV)})
    img_mask = torch.zeros(cfg["batch"], H0 * 16, W0 * 16, dtype=torch.bool, device=DEV)
    img_mask[1::2, :, (W0 - max(1, W0 // 8)) * 16:] = True
    samples = _Nested(torch.zeros(cfg["batch"], 3, H0 * 16, W0 * 16, device=DEV), img_mask)
    return model, criterion, matcher, samples, targets


def test_set_criterion_signature_matches_reference_fixture_and_oracle():
    """SetCriterion(matcher, weight_dict, losses)(outputs, targets, n_boxes) with list-of-dict targets: every entry of
    the loss dict vs the unmodified reference (fixture criterion/gt) and each entry separately differentiable."""
    from poet_b200.criterion import SetCriterion, PoseMatcher
    g = load_golden("criterion/gt")
    t_all, R_all, boxes, labels, tgt_t, tgt_R, n_boxes = criterion_case()
    L, B, Q = t_all.shape[:3]
    td, Rd = t_all.to(DEV).requires_grad_(True), R_all.to(DEV).requires_grad_(True)
    pb = torch.full((B, Q, 4), -1.0)
    pc = torch.full((B, Q), -1, dtype=torch.int64)
    for b, n in enumerate(n_boxes):
        pb[b, :n], pc[b, :n] = boxes[b][:n], labels[b][:n]
    layer = lambda l: {"pred_translation": td[l], "pred_rotation": Rd[l], "pred_boxes": pb.to(DEV), "pred_classes": pc.to(DEV)}
    outputs = dict(layer(L - 1), aux_outputs=[layer(l) for l in range(L - 1)])
    targets = [{"boxes": boxes[b], "labels": labels[b], "relative_position": tgt_t[b], "relative_rotation": tgt_R[b]}
               for b in range(B)]                                     # host targets, as the reference's data loader yields them
    weight_dict = dict(g["weights"])
    for i in range(L - 1):
        weight_dict.update({k + f"_{i}": v for k, v in g["weights"].items()})
    crit = SetCriterion(PoseMatcher(bbox_mode="gt"), weight_dict, ["translation", "rotation"])
    loss_dict = crit(outputs, targets, n_boxes)
    assert set(loss_dict) == set(g["losses"])
    for k, v in g["losses"].items():
        assert abs(float(loss_dict[k]) - v) <= 5e-6 * max(1.0, abs(v)), k
    # engine.py:57-58: the caller forms the weighted sum; one entry alone back-propagates into its own layer only
    loss_dict["loss_rot_0"].backward(retain_graph=True)
    assert float(Rd.grad[0].abs().max()) > 0 and float(Rd.grad[1:].abs().max()) == 0
    assert td.grad is None or float(td.grad.abs().max()) == 0
    Rd.grad, td.grad = None, None
    total = sum(loss_dict[k] * crit.weight_dict[k] for k in loss_dict if k in crit.weight_dict)
    total.backward()
    t64, R64 = t_all.double().requires_grad_(True), R_all.double().requires_grad_(True)
    _, tot64 = O.pose_criterion_gt(t64, R64, [t.double() for t in tgt_t], [r.double() for r in tgt_R], n_boxes,
                                   g["weights"]["loss_trans"], g["weights"]["loss_rot"])
    tot64.backward()
    assert abs(float(total) - float(tot64)) <= 5e-6 * abs(float(tot64))
    assert float((td.grad.cpu().double() - t64.grad).abs().max()) < 1e-6
    assert float((Rd.grad.cpu().double() - R64.grad).abs().max()) < 1e-4 * float(R64.grad.abs().max())


def test_engine_style_training_iterations_and_resume():
    from poet_b200 import ops
    from poet_b200.data_parallel import FlatGradReducer
    from poet_b200.optim import FusedClipAdamW
    cfg = dict(S.CONFIGS["tiny16"], batch=4)

    def make():
        model, criterion, matcher, samples, targets = _setup(cfg)
        a = _args(cfg)
        red = FlatGradReducer(model.parameters())
        opt = FusedClipAdamW(model, red, lr=a.lr, weight_decay=a.weight_decay, max_norm=a.clip_max_norm,
                             lr_backbone=a.lr_backbone, lr_linear_proj_mult=a.lr_linear_proj_mult)
        sched = torch.optim.lr_scheduler.StepLR(opt, a.lr_drop)                 # main.py:278
        return model, criterion, samples, targets, opt, sched

    def iteration(model, criterion, samples, targets, opt):                     # engine.py:55-81
        outputs, n_boxes_per_sample = model(samples, targets)
        loss_dict = criterion(outputs, targets, n_boxes_per_sample)
        weight_dict = criterion.weight_dict
        losses = sum(loss_dict[k] * weight_dict[k] for k in loss_dict.keys() if k in weight_dict)
        loss_value = float(losses)
        assert math.isfinite(loss_value)
        opt.zero_grad()
        losses.backward()
        opt.step()                                                               # clip_grad_norm_(0.1) is fused into the step
        return loss_value, float(opt.grad_norm())

    ops.set_dropout_seed(2024)
    model, criterion, samples, targets, opt, sched = make()
    model.train()
    criterion.train()
    w0 = model.transformer.encoder.layers[0].linear1.weight.detach().clone()
    hist = []
    for epoch in range(2):
        for _ in range(2):
            hist.append(iteration(model, criterion, samples, targets, opt))
        sched.step()
    assert all(g > 0 for _, g in hist)
    assert not torch.equal(w0, model.transformer.encoder.layers[0].linear1.weight)
    assert abs(opt.param_groups[0]["lr"] - 2e-4 * 0.1) < 1e-12                  # StepLR(lr_drop=2) decayed every group once
    assert abs(opt.param_groups[2]["lr"] - 2e-5 * 0.1) < 1e-12
    # eval keeps the reference contract and ignores dropout
    model.eval()
    with torch.no_grad():
        out_a, n_a = model(samples, targets)
        out_b, _ = model(samples, targets)
    assert torch.equal(out_a["pred_translation"], out_b["pred_translation"]) and n_a == [t["boxes"].shape[0] for t in targets]
    assert set(out_a) == {"pred_translation", "pred_rotation", "pred_boxes", "pred_classes", "aux_outputs"}

    # checkpoint (main.py:357-369) -> fresh process state -> resume (main.py:287-317) -> same next iteration
    buf = io.BytesIO()
    torch.save({"model": model.state_dict(), "optimizer": opt.state_dict(), "lr_scheduler": sched.state_dict(), "epoch": 1}, buf)
    model.train()
    ops.set_dropout_seed(555)
    next_ref = iteration(model, criterion, samples, targets, opt)
    w_ref = model.transformer.encoder.layers[0].linear1.weight.detach().clone()

    buf.seek(0)
    ckpt = torch.load(buf, map_location="cpu", weights_only=False)
    model2, criterion2, samples2, targets2, opt2, sched2 = make()
    missing, unexpected = model2.load_state_dict(ckpt["model"], strict=False)
    assert not missing and not unexpected
    opt2.load_state_dict(ckpt["optimizer"])
    sched2.load_state_dict(ckpt["lr_scheduler"])
    assert opt2.step_count == 4 and abs(opt2.param_groups[0]["lr"] - 2e-5) < 1e-12
    model2.train()
    ops.set_dropout_seed(555)
    next_res = iteration(model2, criterion2, samples2, targets2, opt2)
    assert abs(next_res[0] - next_ref[0]) <= 1e-5 * max(1.0, abs(next_ref[0]))
    assert abs(next_res[1] - next_ref[1]) <= 1e-4 * next_ref[1]
    w_res = model2.transformer.encoder.layers[0].linear1.weight
    assert float((w_res - w_ref).abs().max()) <= 1e-6
