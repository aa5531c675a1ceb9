"""CPU tests: the oracle restatement vs fixtures produced by the unmodified reference
(oracle/make_golden.py), plus self-consistency of the MSDA core restatements."""
import math

import pytest
import torch

from oracle import poet_oracle as O
from poet_b200 import synthetic as S
from helpers import load_golden, sample_indices, oracle_poet_from_feats, same_fingerprint


def test_posenc_matches_reference():
    g = load_golden("posenc")
    for key, rec in g.items():
        if not key.startswith("pos_"):
            continue
        got = O.sine_position_embedding(rec["mask"], 128)
        assert got.shape == rec["pos"].shape
        assert torch.equal(got, rec["pos"]), key                 # same torch ops -> bit exact


def test_bbox_embedding_matches_reference():
    rec = load_golden("posenc")["bbox"]
    assert torch.equal(O.bbox_sine_embedding(rec["boxes"], 32), rec["embed"])


def test_build_queries_padding():
    boxes = [torch.rand(2, 4), torch.rand(4, 4)]
    labels = [torch.tensor([3, 1]), torch.tensor([2, 2, 5, 1])]
    qe, pb, pc, n = O.build_queries(boxes, labels, 4, 256)
    assert qe.shape == (2, 4, 512) and pb.shape == (2, 4, 4) and pc.dtype == torch.int64
    assert n == [2, 4]
    assert torch.all(qe[0, 2:] == -10) and torch.all(pb[0, 2:] == -1) and torch.all(pc[0, 2:] == -1)
    assert torch.equal(qe[0, :2, :256], qe[0, :2, 256:])


@pytest.mark.parametrize("key", ["transformer/tiny/pad0", "transformer/tiny/pad1",
                                 "transformer/tiny16/pad1", "transformer/cfg1/pad0"])
def test_transformer_matches_reference(key):
    g = load_golden(key)
    cfg = S.CONFIGS[g["cfg"]]
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=g["pad"])
    assert same_fingerprint(S.fingerprint(inp["srcs"]), g["fp_inputs"]), "seeded generation is not reproducible here"
    pos = [O.sine_position_embedding(m, cfg["d_model"] // 2) for m in inp["masks"]]
    qe, pb, _, _ = O.build_queries(inp["boxes"], inp["labels"], cfg["num_queries"], cfg["d_model"])
    cap = {}
    with torch.no_grad():
        hs, init_ref, inter = O.transformer_forward(P, cfg, inp["srcs"], inp["masks"], pos, qe, pb[:, :, :2], capture=cap)
    assert (cap["memory"][:, ::g["memory_stride"]] - g["memory_rows"]).abs().max() < 2e-5
    assert (hs - g["hs"]).abs().max() < 2e-5
    assert torch.equal(init_ref, g["init_ref"]) and torch.equal(inter, g["inter_ref"])


@pytest.mark.parametrize("key", ["poet/tiny/pad1", "poet/tiny16/pad0", "poet/cfg1/pad0", "poet/cfg2_b2/pad1"])
def test_poet_forward_backward_matches_reference(key):
    g = load_golden(key)
    cfg = S.CONFIGS[g["cfg"]]
    P, feats, srcs, masks, inp, out, n_boxes, cap = oracle_poet_from_feats(cfg, g["pad"])
    assert n_boxes == g["n_boxes"]
    assert torch.equal(out["pred_boxes"], g["pred_boxes"]) and torch.equal(out["pred_classes"], g["pred_classes"])
    assert (cap["translation_all"] - g["translation"]).abs().max() < 2e-5
    assert (cap["rotation_all"] - g["rotation"]).abs().max() < 2e-5
    assert torch.equal(out["pred_translation"], cap["translation_all"][-1])
    assert len(out["aux_outputs"]) == cfg["dec_layers"] - 1
    g_t, g_R = S.make_cotangents(cfg)
    loss = O.synthetic_loss((cap["translation_all"], cap["rotation_all"]), g_t, g_R)
    assert abs(float(loss) - g["loss"]) < 1e-3 * max(1.0, abs(g["loss"]))
    loss.backward()
    grads = {k: v.grad for k, v in P.items()}
    for l, f in enumerate(feats):
        grads[f"__feat{l}"] = f.grad
    checked = 0
    for name, rec in g["grads"].items():
        got = grads.get(name)
        if rec is None:                                   # transformer.reference_points.*: unused in bbox mode
            assert got is None or float(got.abs().max()) == 0.0, name
            continue
        flat = got.flatten()
        ref = rec["samples"]
        tol = 2e-4 * max(1.0, rec["norm"] / math.sqrt(flat.numel()) * 10)
        assert (flat[sample_indices(flat.numel())] - ref).abs().max() < tol, name
        assert abs(float(flat.double().norm()) - rec["norm"]) < 1e-3 * max(rec["norm"], 1e-3), name
        checked += 1
    assert checked > 20


def _rand_msda(B=2, M=4, D=8, Lq=9, P=3, shapes=((6, 4), (3, 2), (1, 1)), dtype=torch.float64, seed=3):
    g = torch.Generator().manual_seed(seed)
    S_ = sum(h * w for h, w in shapes)
    value = torch.randn(B, S_, M, D, generator=g, dtype=dtype)
    loc = torch.rand(B, Lq, M, len(shapes), P, 2, generator=g, dtype=dtype) * 1.6 - 0.3
    loc[0, 0] = -1.0                                       # dummy-query reference points
    attn = torch.softmax(torch.randn(B, Lq, M, len(shapes) * P, generator=g, dtype=dtype), -1)
    return value, list(shapes), loc, attn.view(B, Lq, M, len(shapes), P)


def test_msda_core_restatements_agree():
    value, shapes, loc, attn = _rand_msda()
    a = O.msda_core(value, shapes, loc, attn)
    b = O.msda_core_direct(value, shapes, loc, attn)
    assert (a - b).abs().max() < 1e-12
    assert float(a[0, 0].abs().max()) == 0.0              # loc = -1 -> exactly zero output


def test_msda_core_vs_transformers_class():
    mdd = pytest.importorskip("transformers.models.deformable_detr.modeling_deformable_detr")
    value, shapes, loc, attn = _rand_msda()
    ss = torch.tensor(shapes)
    lsi = torch.cat((ss.new_zeros(1), ss.prod(1).cumsum(0)[:-1]))
    ref = mdd.MultiScaleDeformableAttention().forward(value, ss, shapes, lsi, loc, attn, 64)
    assert (O.msda_core(value, shapes, loc, attn) - ref).abs().max() == 0.0


def test_msda_core_gradcheck():
    value, shapes, loc, attn = _rand_msda(B=1, M=2, D=2, Lq=2, P=2, shapes=((6, 4), (3, 2)))
    value.requires_grad_(True); loc.requires_grad_(True); attn.requires_grad_(True)
    assert torch.autograd.gradcheck(lambda v, l, a: O.msda_core_direct(v, shapes, l, a), (value, loc, attn),
                                    eps=1e-6, atol=1e-5, nondet_tol=1e-8)


def test_fp32_noise_floor_vs_fp64():
    cfg = S.CONFIGS["tiny16"]
    _, _, _, _, _, _, _, cap32 = oracle_poet_from_feats(cfg, True, torch.float32, need_grad=False)
    _, _, _, _, _, _, _, cap64 = oracle_poet_from_feats(cfg, True, torch.float64, need_grad=False)
    assert (cap32["translation_all"].double() - cap64["translation_all"]).abs().max() < 1e-5
    assert (cap32["rot6d"].double() - cap64["rot6d"]).abs().max() < 1e-5


def test_pose_criterion_oracle_matches_reference_golden():
    """oracle.pose_criterion_gt == the unmodified reference SetCriterion + PoseMatcher('gt') (fixture generated by
    oracle/make_golden.py): every loss term and the gradient of the weighted total."""
    from helpers import load_golden
    from oracle.make_golden import criterion_case
    from oracle import poet_oracle as O
    g = load_golden("criterion/gt")
    t_all, R_all, _boxes, _labels, tgt_t, tgt_R, n_boxes = criterion_case()
    t_all, R_all = t_all.double().requires_grad_(True), R_all.double().requires_grad_(True)
    losses, total = O.pose_criterion_gt(t_all, R_all, [t.double() for t in tgt_t], [r.double() for r in tgt_R], n_boxes,
                                        g["weights"]["loss_trans"], g["weights"]["loss_rot"])
    assert set(losses) == set(g["losses"])
    for k, v in g["losses"].items():
        assert abs(float(losses[k]) - v) <= 2e-6 * max(1.0, abs(v)), k
    assert abs(float(total) - g["total"]) <= 2e-6 * abs(g["total"])
    total.backward()
    assert float((t_all.grad - g["grad_t"].double()).abs().max()) < 1e-6
    assert float((R_all.grad - g["grad_R"].double()).abs().max()) < 2e-4      # fp32 acos' near the clamp in the fixture
