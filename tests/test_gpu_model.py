"""GPU parity tests, module level: our nn.Module mirrors (DeformableTransformer, PoET, MSDeformAttn)
against the CPU oracle and against the golden fixtures produced by the unmodified reference.

Tolerances (BASELINE.json north_star): |translation| <= 1e-4 abs, rot-6D / rotation matrix <= 1e-3."""
import math

import pytest
import torch

from oracle import poet_oracle as O
from poet_b200 import synthetic as S
from helpers import load_golden, sample_indices, oracle_poet_from_feats, same_fingerprint

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_T, TOL_R = 1e-4, 1e-3


# Gradient tolerances per GEMM precision.  The network is only piecewise smooth (ReLU kinks, bilinear
# cell edges): an implementation whose forward differs from the oracle by eps flips a fraction ~eps of
# those units, and every flip changes that unit's gradient by O(1).  fp32 kernels (eps ~1e-7) therefore
# agree with the oracle's gradients to ~1e-3, the split-bf16 tensor-core path (eps ~1e-5) to ~1e-2,
# while the forward outputs of both stay inside the 1e-4 / 1e-3 budget.  See DESIGN.md "gradient parity".
GRAD_TOL = {"fp32": dict(l2=5e-3, elem=1e-3, frac=1e-2, mx=5e-2, samp=5e-3),
            "bf16x3": dict(l2=2e-2, elem=5e-3, frac=2e-2, mx=1e-1, samp=3e-2)}


@pytest.fixture(params=["fp32", "bf16x3"])
def precision(request):
    from poet_b200 import ops
    old = ops.get_gemm_precision()
    ops.set_gemm_precision(request.param)
    yield request.param
    ops.set_gemm_precision(old)


def build_model(cfg, P):
    from poet_b200.deformable_transformer import DeformableTransformer
    from poet_b200.pose_estimation_transformer import PoET
    tr = DeformableTransformer(cfg["d_model"], cfg["nheads"], cfg["enc_layers"], cfg["dec_layers"], cfg["dim_ff"], 0.0,
                               "relu", True, cfg["n_levels"], cfg["n_points"], cfg["n_points"])
    model = PoET(None, tr, cfg["num_queries"], cfg["n_levels"], cfg["n_classes"], class_mode=cfg["class_mode"])
    model.load_state_dict({k: v for k, v in P.items() if not k.startswith("input_proj")}, strict=True)
    return model.to(DEV).train()


def stack_outputs(out):
    t = torch.stack([a["pred_translation"] for a in out["aux_outputs"]] + [out["pred_translation"]])
    R = torch.stack([a["pred_rotation"] for a in out["aux_outputs"]] + [out["pred_rotation"]])
    return t, R


@pytest.mark.parametrize("key", ["transformer/tiny/pad0", "transformer/tiny/pad1", "transformer/tiny16/pad1",
                                 "transformer/cfg1/pad0"])
def test_transformer_vs_reference_golden(key, precision):
    """DeformableTransformer.forward with the reference's call signature (NCHW srcs/pos) vs the fixture."""
    from poet_b200 import ops
    g = load_golden(key)
    cfg = S.CONFIGS[g["cfg"]]
    P = S.make_params(cfg)
    model = build_model(cfg, P)
    inp = S.make_inputs(cfg, pad_columns=g["pad"])
    assert same_fingerprint(S.fingerprint(inp["srcs"]), g["fp_inputs"])
    masks = [m.to(DEV) for m in inp["masks"]]
    pos = [ops.posenc_sine_nchw(m, cfg["d_model"] // 2) for m in masks]
    qe, pb, pc, _ = model.build_queries(inp["boxes"], inp["labels"], DEV)
    with torch.no_grad():
        hs, init_ref, inter, _, _ = model.transformer([s.to(DEV) for s in inp["srcs"]], masks, pos, qe,
                                                      pb[:, :, :2].contiguous())
    assert float((hs.cpu() - g["hs"]).abs().max()) < 1e-4
    assert torch.equal(init_ref.cpu(), g["init_ref"]) and torch.equal(inter.cpu(), g["inter_ref"])


@pytest.mark.parametrize("key", ["poet/tiny/pad1", "poet/tiny16/pad0", "poet/cfg1/pad0", "poet/cfg2_b2/pad1"])
def test_poet_path_vs_reference_golden(key, precision):
    """forward_pyramid + backward vs the fixture from the reference PoET (input_proj done by the oracle,
    which is outside the CUDA path: SURVEY.md §8f N1)."""
    g = load_golden(key)
    cfg = S.CONFIGS[g["cfg"]]
    P, feats, srcs, masks, inp, _, _, _ = oracle_poet_from_feats(cfg, g["pad"], need_grad=False)
    model = build_model(cfg, {k: v.detach() for k, v in P.items()})
    d_srcs = [s.detach().to(DEV).requires_grad_(True) for s in srcs]
    out, n_boxes = model.forward_pyramid(d_srcs, [m.to(DEV) for m in masks], inp["boxes"], inp["labels"])
    assert n_boxes == g["n_boxes"]
    assert torch.equal(out["pred_boxes"].cpu(), g["pred_boxes"]) and torch.equal(out["pred_classes"].cpu(), g["pred_classes"])
    t, R = stack_outputs(out)
    assert float((t.detach().cpu() - g["translation"]).abs().max()) < TOL_T
    assert float((R.detach().cpu() - g["rotation"]).abs().max()) < TOL_R
    g_t, g_R = S.make_cotangents(cfg)
    ((t * g_t.to(DEV)).sum() + (R * g_R.to(DEV)).sum()).backward()
    tight, loose = [], []
    tol = GRAD_TOL[precision]
    for name, p in model.named_parameters():
        rec = g["grads"].get(name)
        if rec is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        flat = p.grad.detach().cpu().flatten()
        scale = max(rec["norm"] / math.sqrt(flat.numel()), 1e-6)                 # RMS of the reference gradient
        err = (flat[sample_indices(flat.numel())] - rec["samples"]).abs()
        norm_err = abs(float(flat.double().norm()) - rec["norm"]) / max(rec["norm"], 1e-3)
        tight.append(float((err > tol["samp"] * scale + 1e-5).double().mean()) <= 0.01 and norm_err < tol["l2"])
        loose.append((name, float(err.max()) < 3.0 * scale + 1e-5 and norm_err < 0.1))
    check_grad_census(tight, loose)


@pytest.mark.parametrize("name,pad", [("tiny16", True), ("cfg1", False), ("cfg2_b2", True)])
def test_poet_path_vs_oracle_all_grads(name, pad, precision):
    """Every output of every decoder layer and the gradient of every parameter and of the input
    pyramid vs the oracle (fp32 CPU) on identical seeded inputs."""
    cfg = S.CONFIGS[name]
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=pad)
    g_t, g_R = S.make_cotangents(cfg)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    r_srcs = [s.clone().requires_grad_(True) for s in inp["srcs"]]
    cap = {}
    O.poet_path_forward(Pr, cfg, r_srcs, inp["masks"], inp["boxes"], inp["labels"], capture=cap)
    O.synthetic_loss((cap["translation_all"], cap["rotation_all"]), g_t, g_R).backward()

    model = build_model(cfg, P)
    d_srcs = [s.to(DEV).requires_grad_(True) for s in inp["srcs"]]
    out, _ = model.forward_pyramid(d_srcs, [m.to(DEV) for m in inp["masks"]], inp["boxes"], inp["labels"])
    t, R = stack_outputs(out)
    ((t * g_t.to(DEV)).sum() + (R * g_R.to(DEV)).sum()).backward()
    assert float((t.detach().cpu() - cap["translation_all"]).abs().max()) < TOL_T
    assert float((R.detach().cpu() - cap["rotation_all"]).abs().max()) < TOL_R
    tight, loose = [], []
    for k, p in model.named_parameters():
        ref = Pr[k].grad
        if ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        ok_t, ok_l = grad_close(p.grad.cpu(), ref, precision)
        tight.append(ok_t)
        loose.append((k, ok_l))
    for l, (s_d, s_r) in enumerate(zip(d_srcs, r_srcs)):
        ok_t, ok_l = grad_close(s_d.grad.cpu(), s_r.grad, precision)
        tight.append(ok_t)
        loose.append((f"srcs[{l}]", ok_l))
    # cfg1 has B*Q = 5 decoder rows: one knife-edge flip in a head / decoder FFN unit then perturbs every
    # upstream gradient by a few percent, so only the loose bound is meaningful there
    check_grad_census(tight, loose, require_tight=(name != "cfg1"))


def grad_close(got, ref, precision):
    """(tight, loose) verdicts for one gradient tensor.  tight = the per-precision budget of GRAD_TOL
    (relative L2, fraction of elements off by more than `elem` of max, worst element); loose = relative L2
    below 10 % (what a single knife-edge flip can do to a gradient that sums over only B*Q = 5 rows)."""
    tol = GRAD_TOL[precision]
    scale = float(ref.abs().max()) + 1e-12
    err = (got.double() - ref.double()).abs()
    rel_l2 = float(err.norm() / (ref.double().norm() + 1e-12))
    tight = rel_l2 < tol["l2"] and float((err > tol["elem"] * scale).double().mean()) < tol["frac"] and \
        float(err.max()) < tol["mx"] * scale + 1e-7
    return tight, rel_l2 < 0.1


def check_grad_census(tight, loose, require_tight=True):
    """A wrong backward kernel is off by O(1) on whole parameter groups; kink flips (GRAD_TOL) perturb a few
    tensors by a few percent.  So: EVERY tensor inside the loose bound, and at least 85 % inside the tight one."""
    bad = [k for k, ok in loose if not ok]
    assert not bad, f"gradients off by more than 10 %: {bad[:8]}"
    assert len(tight) > 20
    if require_tight:
        assert sum(tight) >= 0.85 * len(tight), f"only {sum(tight)}/{len(tight)} gradients inside the tight budget"


def test_backward_matches_forward_directional_derivative(precision):
    """Kink-insensitive end-to-end check of the CUDA backward against the CUDA forward (itself pinned to the
    oracle at ~1e-6): for random parameter directions v, (L(theta + eps v) - L(theta - eps v)) / 2 eps must
    equal <grad L, v>.  Units within eps of a kink contribute O(eps) errors that average out."""
    cfg = S.CONFIGS["cfg2_b2"]
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=True)
    g_t, g_R = (t.to(DEV) for t in S.make_cotangents(cfg))
    srcs, masks = [s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]]
    model = build_model(cfg, P)

    def loss():
        out, _ = model.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
        t, R = stack_outputs(out)
        return ((t * g_t).sum() + (R * g_R).sum()).double()

    loss().backward()
    named = [(k, p) for k, p in model.named_parameters() if p.grad is not None]
    groups = {}
    for k, p in named:
        key = ".".join(k.split(".")[:4]) if k.startswith("transformer.") else k.split(".")[0]
        groups.setdefault(key, []).append(p)
    assert len(groups) >= 12
    for key, ps in groups.items():
        # step along the (normalised) gradient of the group: <grad, v> = |grad|, no cancellation; the step
        # is small enough to stay in the locally-smooth regime (measured: the central difference converges
        # to the analytic value as eps -> 0, but is 10-40 % off at a 2e-3 relative step)
        gnorm = math.sqrt(sum(float((p.grad.double() ** 2).sum()) for p in ps))
        pnorm = math.sqrt(sum(float((p.detach().double() ** 2).sum()) for p in ps))
        if gnorm == 0.0:
            continue
        vs = [p.grad / gnorm for p in ps]
        eps = min(1e-4 * pnorm, 0.02 / gnorm)      # predicted |dL| <= 0.02: well above fp32 loss noise (~1e-4)
        with torch.no_grad():
            for p, v in zip(ps, vs):
                p.add_(v, alpha=eps)
            lp = float(loss())
            for p, v in zip(ps, vs):
                p.add_(v, alpha=-2 * eps)
            lm = float(loss())
            for p, v in zip(ps, vs):
                p.add_(v, alpha=eps)
        numeric = (lp - lm) / (2 * eps)
        assert abs(numeric - gnorm) <= 0.05 * gnorm + 1e-3, (key, numeric, gnorm)


@pytest.mark.parametrize("name", ["cfg3_b2", "cfg5_b1"])
def test_other_baseline_configs_forward_parity(name):
    """BASELINE.json configs 3 (LM-O: 9 class slots) and 5 (1280x960 pyramid S=6380, 6/6 layers, 8 heads of 32
    channels, 25 queries) at reduced batch: every decoder layer's outputs vs the oracle, default precision."""
    cfg = S.CONFIGS[name]
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=(name == "cfg3_b2"))
    cap = {}
    with torch.no_grad():
        O.poet_path_forward(P, cfg, inp["srcs"], inp["masks"], inp["boxes"], inp["labels"], capture=cap)
        model = build_model(cfg, P)
        out, n = model.forward_pyramid([s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]],
                                       inp["boxes"], inp["labels"])
    t, R = stack_outputs(out)
    assert float((t.cpu() - cap["translation_all"]).abs().max()) < TOL_T
    assert float((R.cpu() - cap["rotation_all"]).abs().max()) < TOL_R
    assert n == [max(1, cfg["num_queries"] - (i % 4)) for i in range(cfg["batch"])]


def test_micro_batches_match_single_pass():
    """PoET.micro_batches = 2 (batch slices on separate streams, gradients accumulated atomically into the same
    slots) == the single-pass step: forward bit-identical per image, every parameter gradient to fp32 rounding;
    eager and under the whole-step CUDA graph."""
    from poet_b200 import ops
    from poet_b200.data_parallel import FlatGradReducer
    from poet_b200.graph import GraphedStep
    cfg = dict(S.CONFIGS["tiny16"], batch=4)
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=True)
    g_t, g_R = (t.to(DEV) for t in S.make_cotangents(cfg))
    srcs, masks = [s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]]

    def loss_fn(out):
        t, R = stack_outputs(out)
        return (t * g_t).sum() + (R * g_R).sum()

    results = []
    for mb, graphed in ((1, False), (2, False), (2, True)):
        model = build_model(cfg, P)
        model.micro_batches = mb
        red = FlatGradReducer(model.parameters())
        if graphed:
            step = GraphedStep(model, loss_fn, srcs, masks, inp["boxes"], inp["labels"], reducer=red)
            loss, out = step.run()
        else:
            red.zero()
            out, _ = model.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
            loss = loss_fn(out)
            loss.backward()
        torch.cuda.synchronize()
        t, R = stack_outputs(out)
        results.append((t.detach().clone(), R.detach().clone(), red.flat.detach().clone()))
    t0, R0, g0 = results[0]
    for t, R, g in results[1:]:
        assert torch.equal(t, t0) and torch.equal(R, R0)
        assert float((g - g0).abs().max()) <= 2e-5 * float(g0.abs().max())


@pytest.mark.parametrize("case", ["empty_and_full", "batch1_one_box", "mostly_padded"])
def test_edge_cases_vs_oracle(case):
    """Ragged / degenerate inputs: an image with NO boxes next to one with all Q slots used, a batch of one image
    with a single box, and masks that pad most of every level (dummy queries and padded tokens must behave exactly
    like the reference: dummy reference points of -1 sample nothing, padded value rows are zero)."""
    from poet_b200 import ops
    old = ops.get_gemm_precision()
    ops.set_gemm_precision("bf16x3")
    try:
        cfg = dict(S.CONFIGS["tiny16"], batch=1 if case == "batch1_one_box" else 3)
        P = S.make_params(cfg)
        inp = S.make_inputs(cfg, pad_columns=True)
        Q = cfg["num_queries"]
        g = torch.Generator().manual_seed(11)
        if case == "empty_and_full":
            inp["boxes"][0], inp["labels"][0] = torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64)
            inp["boxes"][1] = torch.rand(Q, 4, generator=g) * 0.4 + 0.3
            inp["labels"][1] = torch.randint(1, cfg["n_classes"] + 1, (Q,), generator=g)
        elif case == "batch1_one_box":
            inp["boxes"][0], inp["labels"][0] = inp["boxes"][0][:1], inp["labels"][0][:1]
        else:
            for m in inp["masks"]:
                m[:, :, max(1, m.shape[2] // 4):] = True            # keep only the left quarter of the columns
                m[0, max(1, m.shape[1] // 2):, :] = True            # and only the top half of image 0
        g_t, g_R = S.make_cotangents(cfg)
        Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        cap = {}
        O.poet_path_forward(Pr, cfg, inp["srcs"], inp["masks"], inp["boxes"], inp["labels"], capture=cap)
        O.synthetic_loss((cap["translation_all"], cap["rotation_all"]), g_t, g_R).backward()
        model = build_model(cfg, P)
        out, n = model.forward_pyramid([s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]],
                                       inp["boxes"], inp["labels"])
        assert n == [min(int(b.shape[0]), Q) for b in inp["boxes"]]
        t, R = stack_outputs(out)
        ((t * g_t.to(DEV)).sum() + (R * g_R.to(DEV)).sum()).backward()
        assert float((t.detach().cpu() - cap["translation_all"]).abs().max()) < TOL_T
        assert float((R.detach().cpu() - cap["rotation_all"]).abs().max()) < TOL_R
        assert torch.isfinite(t).all() and torch.isfinite(R).all()
        loose = []
        for k, p in model.named_parameters():
            ref = Pr[k].grad
            if ref is None or float(ref.abs().max()) == 0.0:
                continue
            assert torch.isfinite(p.grad).all(), k
            loose.append((k, grad_close(p.grad.cpu(), ref, "bf16x3")[1]))
        bad = [k for k, ok in loose if not ok]
        assert len(bad) <= max(1, len(loose) // 20), f"gradients off by more than 10 %: {bad[:8]}"
    finally:
        ops.set_gemm_precision(old)


class _StubBackbone(torch.nn.Module):
    """What PoET reads from a backbone: strides / num_channels to size input_proj, and a forward that returns
    (feature maps with masks, position encodings, predictions); the maps are fixed seeded tensors (the detector
    itself is outside the path)."""

    def __init__(self, channels, feats=None, masks=None):
        super().__init__()
        self.strides, self.num_channels = [8, 16, 32], [channels] * 3
        self.feats, self.masks = feats, masks

    def forward(self, samples):
        from poet_b200.pose_estimation_transformer import _Nested
        return [_Nested(f, m) for f, m in zip(self.feats, self.masks)], [None] * len(self.feats), None


@pytest.mark.parametrize("key", ["poet/tiny/pad1", "poet/tiny16/pad0", "poet/cfg1/pad0", "poet/cfg2_b2/pad1"])
def test_input_proj_and_path_vs_reference_golden(key, precision):
    """backbone feature maps -> input_proj (1x1 conv + GroupNorm, 3x3/s2 conv + GroupNorm) -> hot path,
    all on our kernels, against the fixture produced by the unmodified reference PoET.forward on the same stub
    feature maps: projected tokens vs the oracle, every decoder layer's pose, and the gradients of the input_proj
    parameters."""
    from poet_b200 import ops
    from poet_b200.deformable_transformer import DeformableTransformer
    from poet_b200.pose_estimation_transformer import PoET
    from helpers import image_mask_for
    g = load_golden(key)
    cfg = S.CONFIGS[g["cfg"]]
    P, feats, srcs, masks, inp, _, _, _ = oracle_poet_from_feats(cfg, g["pad"], need_grad=False)
    tr = DeformableTransformer(cfg["d_model"], cfg["nheads"], cfg["enc_layers"], cfg["dec_layers"], cfg["dim_ff"], 0.0,
                               "relu", True, cfg["n_levels"], cfg["n_points"], cfg["n_points"])
    model = PoET(_StubBackbone(cfg["d_model"]), tr, cfg["num_queries"], cfg["n_levels"], cfg["n_classes"],
                 class_mode=cfg["class_mode"])
    model.load_state_dict({k: v.detach() for k, v in P.items()}, strict=True)
    model = model.to(DEV).train()
    d_feats = [f.detach().to(DEV) for f in feats]
    levels = [(ip[0].weight, ip[0].bias, ip[1].weight, ip[1].bias) for ip in model.input_proj]
    tokens = ops.input_proj_tokens(d_feats, levels)
    ref_tokens = torch.cat([s.detach().flatten(2).transpose(1, 2) for s in srcs], 1)
    tol = 2e-5 if precision == "fp32" else 1e-4
    assert float((tokens.detach().cpu() - ref_tokens).abs().max()) < tol * max(1.0, float(ref_tokens.abs().max()))
    out, n_boxes = model.forward_features(d_feats, [m.to(DEV) for m in inp["masks"][:3]],
                                          image_mask_for(cfg, g["pad"]).to(DEV), inp["boxes"], inp["labels"])
    assert n_boxes == g["n_boxes"]
    t, R = stack_outputs(out)
    assert float((t.detach().cpu() - g["translation"]).abs().max()) < TOL_T
    assert float((R.detach().cpu() - g["rotation"]).abs().max()) < TOL_R
    g_t, g_R = S.make_cotangents(cfg)
    ((t * g_t.to(DEV)).sum() + (R * g_R.to(DEV)).sum()).backward()
    checked, bad = 0, []
    for name, p in model.named_parameters():
        if not name.startswith("input_proj"):
            continue
        rec = g["grads"].get(name)
        assert rec is not None and p.grad is not None, name
        flat = p.grad.detach().cpu().flatten()
        norm_err = abs(float(flat.double().norm()) - rec["norm"]) / max(rec["norm"], 1e-6)
        scale = max(rec["norm"] / math.sqrt(flat.numel()), 1e-6)
        err = float((flat[sample_indices(flat.numel())] - rec["samples"]).abs().max())
        checked += 1
        if norm_err > 0.1 or err > 3.0 * scale + 1e-5:
            bad.append((name, norm_err, err / scale))
    assert checked == 4 * cfg["n_levels"] and not bad, bad
    # the reference-facing entry point model(samples, targets) takes the same route
    from poet_b200.pose_estimation_transformer import _Nested
    model.backbone.feats, model.backbone.masks = d_feats, [m.to(DEV) for m in inp["masks"][:3]]
    img_mask = image_mask_for(cfg, g["pad"]).to(DEV)
    samples = _Nested(torch.zeros(img_mask.shape[0], 3, *img_mask.shape[1:], device=DEV), img_mask)
    targets = [{"boxes": b.to(DEV), "labels": l.to(DEV)} for b, l in zip(inp["boxes"], inp["labels"])]
    out2, n2 = model(samples, targets)
    assert n2 == n_boxes
    assert torch.equal(out2["pred_translation"], out["pred_translation"]) and torch.equal(out2["pred_rotation"], out["pred_rotation"])


def test_backbone_mode_inference_vs_reference_golden():
    """SURVEY.md section 8f N4 (inference mode): queries built from detector output (xyxy -> normalised cxcywh, top-Q by score,
    an image with no detections) + input_proj + path, model(samples, None) in eval mode, against the reference run."""
    from poet_b200 import ops
    from poet_b200.deformable_transformer import DeformableTransformer
    from poet_b200.pose_estimation_transformer import PoET, _Nested
    from oracle.make_golden import backbone_predictions
    g = load_golden("poet_backbone_mode/tiny16")
    cfg = dict(S.CONFIGS[g["cfg"]], batch=g["batch"])
    P = S.make_params(cfg, with_input_proj=True)
    inp = S.make_inputs(cfg, pad_columns=False)
    H0, W0 = inp["srcs"][0].shape[-2:]
    preds = backbone_predictions(cfg, (H0 * 16, W0 * 16))
    old = ops.get_gemm_precision()
    ops.set_gemm_precision("bf16x3")
    try:
        tr = DeformableTransformer(cfg["d_model"], cfg["nheads"], cfg["enc_layers"], cfg["dec_layers"], cfg["dim_ff"], 0.0,
                                   "relu", True, cfg["n_levels"], cfg["n_points"], cfg["n_points"])
        bb = _StubBackbone(cfg["d_model"], [f.to(DEV) for f in inp["srcs"][:3]], [m.to(DEV) for m in inp["masks"][:3]])
        bb.forward = lambda samples: ([_Nested(f, m) for f, m in zip(bb.feats, bb.masks)], [None] * 3,
                                      [None if p is None else p.to(DEV) for p in preds])
        model = PoET(bb, tr, cfg["num_queries"], cfg["n_levels"], cfg["n_classes"], bbox_mode="backbone",
                     class_mode=cfg["class_mode"])
        model.load_state_dict(P, strict=True)
        model = model.to(DEV).eval()
        img_mask = torch.zeros(cfg["batch"], H0 * 16, W0 * 16, dtype=torch.bool, device=DEV)
        samples = _Nested(torch.zeros(cfg["batch"], 3, H0 * 16, W0 * 16, device=DEV), img_mask)
        with torch.no_grad():
            out, n_boxes = model(samples, None)
    finally:
        ops.set_gemm_precision(old)
    assert n_boxes == g["n_boxes"]
    assert torch.equal(out["pred_classes"].cpu(), g["pred_classes"])
    assert float((out["pred_boxes"].cpu() - g["pred_boxes"]).abs().max()) < 1e-6
    t, R = stack_outputs(out)
    assert float((t.cpu() - g["translation"]).abs().max()) < TOL_T
    assert float((R.cpu() - g["rotation"]).abs().max()) < TOL_R


def test_throughput_mode_tolerance():
    """BASELINE.json cfg4 (bf16 training mode; no reference counterpart): single-pass bf16 tensor-core
    MMAs on the token-row GEMMs (encoder layers + decoder value projections), split-bf16 elsewhere.  ITS OWN tolerance,
    measured on B200 (profiles/r02_grad_noise.txt: 6.7e-3 / 2.8e-2 forward, <= 0.16 relative L2 on any gradient):
    |translation| <= 2e-2, |rotation| <= 5e-2 vs the fp32 oracle on every decoder layer, every gradient within 0.35
    relative L2 and 90 % of them within 0.2; the default (parity) mode of the same model stays inside 1e-4 / 1e-3."""
    from poet_b200 import ops
    cfg = S.CONFIGS["cfg2_b2"]
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=True)
    g_t, g_R = S.make_cotangents(cfg)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    cap = {}
    O.poet_path_forward(Pr, cfg, inp["srcs"], inp["masks"], inp["boxes"], inp["labels"], capture=cap)
    O.synthetic_loss((cap["translation_all"], cap["rotation_all"]), g_t, g_R).backward()
    old = ops.get_gemm_precision()
    ops.set_gemm_precision("bf16x3")
    try:
        model = build_model(cfg, P)
        srcs, masks = [s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]]
        with torch.no_grad():
            t0, R0 = stack_outputs(model.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])[0])
        model.transformer.set_throughput_mode(True)
        out, _ = model.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
        t, R = stack_outputs(out)
        ((t * g_t.to(DEV)).sum() + (R * g_R.to(DEV)).sum()).backward()
    finally:
        ops.set_gemm_precision(old)
    assert float((t0.cpu() - cap["translation_all"]).abs().max()) < TOL_T and float((R0.cpu() - cap["rotation_all"]).abs().max()) < TOL_R
    dt = float((t.detach().cpu() - cap["translation_all"]).abs().max())
    dR = float((R.detach().cpu() - cap["rotation_all"]).abs().max())
    assert 1e-4 < dt < 2e-2 and dR < 5e-2, (dt, dR)                      # really a different (cheaper) arithmetic, inside its budget
    errs = []
    for k, p in model.named_parameters():
        ref = Pr[k].grad
        if ref is None:
            continue
        assert torch.isfinite(p.grad).all(), k
        errs.append(float((p.grad.cpu().double() - ref.double()).norm() / (ref.double().norm() + 1e-12)))
    assert max(errs) < 0.35 and sum(e < 0.2 for e in errs) >= 0.9 * len(errs), sorted(errs)[-5:]
