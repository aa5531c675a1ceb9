"""The two drop-in seams against the UNMODIFIED reference code, CPU only.

Runs where /root/reference exists (the build container; the GPU box has no reference tree and skips):
  B-py1  poet_b200.install_into_reference() -> the reference's own models.deformable_transformer builds its layers
         from OUR MSDeformAttn class; parameter names, shapes and init values match the reference construction,
         checkpoints load strict=True in both directions;
  B-py2  poet_b200.build_model(args) returns (model, criterion, matcher) like models.build_model (models/__init__.py:10);
         our PoET / DeformableTransformer load a reference state_dict strict=True and vice versa; a reference-format
         checkpoint {model, optimizer, lr_scheduler, epoch, args} survives torch.save -> torch.load -> resume the way
         main.py:287-317 does it (strict=False load, optimizer + StepLR restore).
No kernel runs here (no GPU): this is the structural half of the seam; the numerical half is tests/test_gpu_*.py
against fixtures that the same reference classes produced (oracle/make_golden.py).
"""
import copy
import io
import os
import sys
import types

import pytest
import torch

from poet_b200 import synthetic as S

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="no reference tree on this machine")


def _args(cfg, **kw):
    a = types.SimpleNamespace(
        hidden_dim=cfg["d_model"], nheads=cfg["nheads"], enc_layers=cfg["enc_layers"], dec_layers=cfg["dec_layers"],
        dim_feedforward=cfg["dim_ff"], dropout=0.1, num_feature_levels=cfg["n_levels"], dec_n_points=cfg["n_points"],
        enc_n_points=cfg["n_points"], num_queries=cfg["num_queries"], n_classes=cfg["n_classes"], bbox_mode="gt",
        reference_points="bbox", query_embedding="bbox", rotation_representation="6d", class_mode=cfg["class_mode"],
        aleatoric=False, aux_loss=True, backbone="maskrcnn", matcher_type="pose", set_cost_class=1, set_cost_bbox=1,
        translation_loss_coef=1, rotation_loss_coef=1, device="cpu")
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.fixture()
def reference_with_our_op():
    """sys.modules['deformable_attention'] = ours, reference tree importable; restored afterwards."""
    import poet_b200
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "deformable_attention" or k == "models" or
             k.startswith("models.") or k == "util" or k.startswith("util.")}
    for k in saved:
        sys.modules.pop(k, None)
    poet_b200.install_into_reference()
    sys.path.insert(0, REF)
    try:
        yield
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "deformable_attention" or k == "models" or k.startswith("models.") or
                  k == "util" or k.startswith("util.")]:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


class _StubBackbone(torch.nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.strides, self.num_channels = [8, 16, 32], [channels] * 3


def test_module_seam_reference_transformer_builds_on_our_op(reference_with_our_op):
    import deformable_attention
    from poet_b200.deformable_attention import MSDeformAttn as Ours
    from poet_b200.deformable_transformer import DeformableTransformer as OurTransformer
    assert deformable_attention.MSDeformAttn is Ours
    from models.deformable_transformer import build_deforamble_transformer, DeformableTransformer as RefTransformer
    import models.deformable_transformer as ref_mod
    assert ref_mod.MSDeformAttn is Ours                       # reference :24 picked up our class
    cfg = S.CONFIGS["cfg2_b2"]
    torch.manual_seed(42)
    ref = build_deforamble_transformer(_args(cfg))
    assert isinstance(ref, RefTransformer)
    attn = [m for m in ref.modules() if isinstance(m, Ours)]
    assert len(attn) == cfg["enc_layers"] + cfg["dec_layers"]   # :177 and :248 constructed ours
    # reference _reset_parameters (:53-62) found the modules by isinstance and re-initialised them (:58-59)
    for m in attn:
        assert float(m.sampling_offsets.weight.abs().max()) == 0.0 and float(m.attention_weights.weight.abs().max()) == 0.0
        assert float(m.sampling_offsets.bias.abs().max()) > 0.0
    torch.manual_seed(42)
    ours = OurTransformer(cfg["d_model"], cfg["nheads"], cfg["enc_layers"], cfg["dec_layers"], cfg["dim_ff"], 0.1, "relu",
                          True, cfg["n_levels"], cfg["n_points"], cfg["n_points"])
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref) == list(sd_ours)                      # same keys in the same registration order
    assert all(sd_ref[k].shape == sd_ours[k].shape for k in sd_ref)
    for k in sd_ref:                                          # same init procedure, same RNG stream -> same values
        assert torch.equal(sd_ref[k], sd_ours[k]), k
    ours.load_state_dict(sd_ref, strict=True)
    ref.load_state_dict(sd_ours, strict=True)
    # lr-group substring match of main.py:41,267-269 finds the same parameters
    pick = lambda m: sorted(n for n, _ in m.named_parameters() if "sampling_offsets" in n or "reference_points" in n)
    assert pick(ref) == pick(ours) and len(pick(ours)) == 2 * (cfg["enc_layers"] + cfg["dec_layers"]) + 2


def test_model_seam_build_model_and_state_dict_both_ways(reference_with_our_op):
    import poet_b200
    from models.pose_estimation_transformer import PoET as RefPoET, SetCriterion as RefCriterion
    from models.deformable_transformer import build_deforamble_transformer
    from models.matcher import PoseMatcher as RefMatcher
    cfg = S.CONFIGS["cfg2_b2"]
    args = _args(cfg, backbone_module=_StubBackbone(cfg["d_model"]))
    model, criterion, matcher = poet_b200.build_model(args)
    assert type(criterion).__name__ == RefCriterion.__name__ and type(matcher).__name__ == RefMatcher.__name__
    assert criterion.matcher is matcher
    # weight_dict exactly as reference build() makes it (:715-733)
    want = {"loss_trans": 1, "loss_rot": 1}
    for i in range(cfg["dec_layers"] - 1):
        want.update({f"loss_trans_{i}": 1, f"loss_rot_{i}": 1})
    want.update({"loss_trans_enc": 1, "loss_rot_enc": 1})
    assert criterion.weight_dict == want
    ref = RefPoET(_StubBackbone(cfg["d_model"]), build_deforamble_transformer(args), num_queries=cfg["num_queries"],
                  num_feature_levels=cfg["n_levels"], n_classes=cfg["n_classes"], bbox_mode="gt", ref_points_mode="bbox",
                  query_embedding_mode="bbox", rotation_mode="6d", class_mode=cfg["class_mode"], aleatoric=False,
                  aux_loss=True, backbone_type="maskrcnn")
    sd_ref, sd_ours = ref.state_dict(), model.state_dict()
    assert list(sd_ref) == list(sd_ours)          # same registration order: optimizer state is numbered by it (main.py:302)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in model.named_parameters()]
    assert all(sd_ref[k].shape == sd_ours[k].shape for k in sd_ref)
    model.load_state_dict(sd_ref, strict=True)
    ref.load_state_dict(sd_ours, strict=True)
    # attributes read from outside (main.py:340-347, pose_estimation_transformer.py:58,139)
    for attr in ("transformer", "input_proj", "translation_head", "rotation_head", "backbone"):
        assert hasattr(model, attr)
    assert model.transformer.d_model == cfg["d_model"] and model.transformer.decoder.num_layers == cfg["dec_layers"]


def test_checkpoint_round_trip_in_reference_format(reference_with_our_op):
    """main.py:357-369 saves {model, optimizer, lr_scheduler, epoch, args}; main.py:287-317 resumes from it."""
    import poet_b200
    from poet_b200.data_parallel import FlatGradReducer
    from poet_b200.optim import FusedClipAdamW
    from models.pose_estimation_transformer import PoET as RefPoET
    from models.deformable_transformer import build_deforamble_transformer
    cfg = S.CONFIGS["tiny16"]
    args = _args(cfg, backbone_module=_StubBackbone(cfg["d_model"]))
    # a checkpoint written by the REFERENCE model + torch.optim.AdamW with the reference's param_dicts (main.py:253-277)
    torch.manual_seed(1)
    ref = RefPoET(_StubBackbone(cfg["d_model"]), build_deforamble_transformer(args), num_queries=cfg["num_queries"],
                  num_feature_levels=cfg["n_levels"], n_classes=cfg["n_classes"], bbox_mode="gt", ref_points_mode="bbox",
                  query_embedding_mode="bbox", rotation_mode="6d", class_mode=cfg["class_mode"], aleatoric=False,
                  aux_loss=True, backbone_type="maskrcnn")
    bb, lp = ("backbone.0",), ("reference_points", "sampling_offsets")
    match = lambda n, ks: any(k in n for k in ks)
    named = [(n, p) for n, p in ref.named_parameters() if p.requires_grad]
    groups = [{"params": [p for n, p in named if not match(n, bb) and not match(n, lp)], "lr": 2e-4},
              {"params": [p for n, p in named if match(n, bb)], "lr": 2e-5},
              {"params": [p for n, p in named if match(n, lp)], "lr": 2e-5}]
    topt = torch.optim.AdamW(groups, lr=2e-4, weight_decay=1e-4)
    tsched = torch.optim.lr_scheduler.StepLR(topt, 2)
    for _ in range(3):
        for n, p in named:
            p.grad = None if "transformer.reference_points" in n else torch.randn_like(p) * 1e-2
        topt.step()
        tsched.step()
    buf = io.BytesIO()
    torch.save({"model": ref.state_dict(), "optimizer": topt.state_dict(), "lr_scheduler": tsched.state_dict(),
                "epoch": 2, "args": args}, buf)
    buf.seek(0)
    ckpt = torch.load(buf, map_location="cpu", weights_only=False)

    # resume into OUR model + fused optimizer the way main.py:293-317 does
    model, _, _ = poet_b200.build_model(args)
    missing, unexpected = model.load_state_dict(ckpt["model"], strict=False)
    assert not missing and not unexpected
    red = FlatGradReducer(model.parameters())
    opt = FusedClipAdamW(model, red, lr=2e-4, weight_decay=1e-4, max_norm=0.1)
    sched = torch.optim.lr_scheduler.StepLR(opt, 2)                      # main.py:278 works on our optimizer
    p_groups = copy.deepcopy(opt.param_groups)
    opt.load_state_dict(ckpt["optimizer"])
    for pg, pg_old in zip(opt.param_groups, p_groups):                  # main.py:303-306
        pg["lr"] = pg_old["lr"]
        pg["initial_lr"] = pg_old["initial_lr"]
    sched.load_state_dict(ckpt["lr_scheduler"])
    assert opt.step_count == 3 and sched.last_epoch == 3
    # moments landed in the arena slots of the matching parameters
    idx = {id(p): i for i, p in enumerate(p for g in topt.param_groups for p in g["params"])}
    ours_named = dict(model.named_parameters())
    for n, p in named:
        st = topt.state.get(p)
        q = ours_named[n]
        off = red.offsets[[id(x) for x in red.params].index(id(q))]
        got = opt.m[off:off + q.numel()].view_as(q)
        if st:
            assert torch.equal(got, st["exp_avg"]), n
        else:
            assert float(got.abs().max()) == 0.0, n
    # ... and our optimizer's checkpoint loads back into torch.optim.AdamW (vice versa)
    sd = opt.state_dict()
    topt2 = torch.optim.AdamW(groups, lr=2e-4, weight_decay=1e-4)
    topt2.load_state_dict(sd)
    for n, p in named:
        if topt.state.get(p):
            assert torch.equal(topt2.state[p]["exp_avg_sq"], topt.state[p]["exp_avg_sq"]), n
            assert float(topt2.state[p]["step"]) == 3.0
    assert idx                                                            # (numbering helper used above)
