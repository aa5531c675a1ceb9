"""Train-mode dropout (reference nn.Dropout sites deformable_transformer.py:178-286, default 0.1 main.py:94).

PyTorch's Philox stream cannot be reproduced (trap 2), so the checks are:
  * statistical: keep rate = 1 - p within binomial tolerance, kept values scaled by exactly 1/(1-p);
  * exact: the mask recovered from one output (or from the backward) reproduces the forward with plain torch math,
    the backward's mask equals the forward's (torch autograd on the explicit-mask formula), the fused epilogue
    path and the fallback kernel draw the same mask, same seed -> same mask, new step -> new mask;
  * model level: eval() is bit-identical to the dropout-0 model, train() under the whole-step CUDA graph draws a
    fresh mask every replay, and the end-to-end backward matches finite differences of the forward at a fixed seed.
"""
import math

import pytest
import torch

from poet_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _binomial_ok(kept_fraction, p, n, sigmas=5.0):
    return abs(kept_fraction - (1.0 - p)) <= sigmas * math.sqrt(p * (1 - p) / n) + 1e-4


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_layernorm_residual_dropout(p):
    from poet_b200 import ops
    torch.manual_seed(0)
    R, C = 4096, 256
    x = torch.randn(R, C, device=DEV, requires_grad=True)
    r = torch.randn(R, C, device=DEV, requires_grad=True)
    gamma = (1 + 0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    beta = (0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    pos = torch.randn(R, C, device=DEV)
    gy, gy2 = torch.randn(R, C, device=DEV), torch.randn(R, C, device=DEV)
    ops.set_dropout_seed(1234)
    ops.begin_dropout_forward(DEV)
    y, y2 = ops.add_layernorm(x, r, gamma, beta, pos=pos, drop_p=p, drop_site=7)
    (y * gy + y2 * gy2).sum().backward()
    # mask from the backward: dr = dz * mask / (1-p)
    ratio = r.grad / x.grad
    kept = ratio.abs() > 0.5
    assert _binomial_ok(float(kept.float().mean()), p, R * C)
    assert float((ratio[kept] - 1.0 / (1.0 - p)).abs().max()) < 1e-5
    mk = kept.float() / (1.0 - p)
    # the forward used the same mask: explicit-mask formula in torch, forward and all gradients
    xr, rr, gr, br = (t.detach().clone().requires_grad_(True) for t in (x, r, gamma, beta))
    yr = torch.nn.functional.layer_norm(xr + rr * mk, (C,), gr, br, 1e-5)
    (yr * gy + (yr + pos) * gy2).sum().backward()
    assert float((y - yr).abs().max()) < 2e-5 and float((y2 - (yr + pos)).abs().max()) < 2e-5
    for a, b in ((x.grad, xr.grad), (r.grad, rr.grad), (gamma.grad, gr.grad), (beta.grad, br.grad)):
        assert float((a - b).abs().max()) <= 2e-4 * max(1.0, float(b.abs().max()))
    # same seed + site -> same mask; another site or the next forward -> another mask
    with torch.no_grad():
        y_same = ops.add_layernorm(x, r, gamma, beta, drop_p=p, drop_site=7)
        y_site = ops.add_layernorm(x, r, gamma, beta, drop_p=p, drop_site=8)
        ops.begin_dropout_forward(DEV)
        y_next = ops.add_layernorm(x, r, gamma, beta, drop_p=p, drop_site=7)
    assert torch.equal(y_same, y.detach())
    assert not torch.equal(y_site, y.detach()) and not torch.equal(y_next, y.detach())
    # p = 0 is exactly the parity path
    with torch.no_grad():
        assert torch.equal(ops.add_layernorm(x, r, gamma, beta, drop_p=0.0), ops.add_layernorm(x, r, gamma, beta))


@pytest.mark.parametrize("R", [2048, 160])            # 2048: tensor-core epilogue path (mask in the ReLU bitmask); 160: fallback kernel
def test_ffn_hidden_dropout(R):
    from poet_b200 import ops
    torch.manual_seed(1)
    p, C, F = 0.1, 256, 1024
    x = torch.randn(R, C, device=DEV, requires_grad=True)
    W1 = (torch.randn(F, C, device=DEV) / 16).requires_grad_(True)
    b1 = (0.1 * torch.randn(F, device=DEV)).requires_grad_(True)
    W2 = (torch.randn(C, F, device=DEV) / 32).requires_grad_(True)
    b2 = (0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    gy = torch.randn(R, C, device=DEV)
    old = ops.get_gemm_precision()
    ops.set_gemm_precision("bf16x3")
    try:
        ops.set_dropout_seed(99)
        seed = ops.begin_dropout_forward(DEV)
        out = ops.mlp(x, ((W1, b1), (W2, b2)), drop_p=p, drop_site=0x203)
        (out * gy).sum().backward()
        # the mask, from the fallback kernel with the same (seed, site): h_drop = dropout(relu(x W1^T + b1))
        with torch.no_grad():
            h = torch.relu(x.double() @ W1.double().t() + b1.double()).float()
            hd = h.clone()
            ops._call("poet_dropout", hd.data_ptr(), hd.numel(), seed.data_ptr(), 0x203, p, ops._stream(hd))
            pos = h > 1e-3
            kept = (hd != 0) & pos
            scale = ops.dropout_scale(p, pair_scheme=True)
            assert _binomial_ok(float(kept.sum()) / float(pos.sum()), p, int(pos.sum()))
            assert float((hd[kept] / h[kept] - scale).abs().max()) < 1e-5
            mk = ((hd != 0) | (h <= 0)).double() * scale        # where relu kills the unit the mask does not matter
    finally:
        ops.set_gemm_precision(old)
    xr, W1r, b1r, W2r, b2r = (t.detach().double().requires_grad_(True) for t in (x, W1, b1, W2, b2))
    ref = (torch.relu(xr @ W1r.t() + b1r) * mk) @ W2r.t() + b2r
    (ref * gy.double()).sum().backward()
    assert float((out.double() - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
    for name, a, b in (("x", x.grad, xr.grad), ("W1", W1.grad, W1r.grad), ("b1", b1.grad, b1r.grad), ("W2", W2.grad, W2r.grad),
                       ("b2", b2.grad, b2r.grad)):
        err = float((a.double() - b).norm() / b.norm())
        assert err < 5e-3, (name, err)        # bf16x3 pre-activations within ~1e-5 of zero gate differently than fp64


def test_attention_probability_dropout():
    from poet_b200 import ops
    torch.manual_seed(2)
    B, Q, M, D, p = 4, 10, 16, 16, 0.3
    C = M * D
    qk = torch.randn(B, Q, 2 * C, device=DEV, requires_grad=True)
    v_eye = torch.zeros(B, Q, M, D, device=DEV)
    for j in range(Q):
        v_eye[:, j, :, j] = 1.0                               # out[b,i,m,j] = P_ij * mask_ij / (1-p)
    ops.set_dropout_seed(5)
    ops.begin_dropout_forward(DEV)
    with torch.no_grad():
        pm = ops.mha_smallq(qk, v_eye.view(B, Q, C), M, drop_p=p, drop_site=0x10100).view(B, Q, M, D)[..., :Q]
        q, k = qk[..., :C].view(B, Q, M, D), qk[..., C:].view(B, Q, M, D)
        P = torch.softmax(torch.einsum("bimd,bjmd->bmij", q, k) / math.sqrt(D), -1)       # [B,M,Q,Q]
        ratio = pm.permute(0, 2, 1, 3) / P
    kept = ratio > 0.5
    assert _binomial_ok(float(kept.float().mean()), p, kept.numel())
    assert float((ratio[kept] - 1.0 / (1.0 - p)).abs().max()) < 1e-4
    mk = kept.float() / (1.0 - p)
    v = torch.randn(B, Q, C, device=DEV, requires_grad=True)
    go = torch.randn(B, Q, C, device=DEV)
    out = ops.mha_smallq(qk, v, M, drop_p=p, drop_site=0x10100)                          # same seed, same site: same mask
    (out * go).sum().backward()
    qkr, vr = qk.detach().clone().requires_grad_(True), v.detach().clone().requires_grad_(True)
    q, k = qkr[..., :C].view(B, Q, M, D), qkr[..., C:].view(B, Q, M, D)
    Pr = torch.softmax(torch.einsum("bimd,bjmd->bmij", q, k) / math.sqrt(D), -1) * mk
    ref = torch.einsum("bmij,bjmd->bimd", Pr, vr.view(B, Q, M, D)).reshape(B, Q, C)
    (ref * go).sum().backward()
    assert float((out - ref).abs().max()) < 2e-5
    assert float((qk.grad - qkr.grad).abs().max()) < 2e-4 and float((v.grad - vr.grad).abs().max()) < 2e-5


def _model(cfg, P, dropout):
    from poet_b200.deformable_transformer import DeformableTransformer
    from poet_b200.pose_estimation_transformer import PoET
    tr = DeformableTransformer(cfg["d_model"], cfg["nheads"], cfg["enc_layers"], cfg["dec_layers"], cfg["dim_ff"], dropout,
                               "relu", True, cfg["n_levels"], cfg["n_points"], cfg["n_points"])
    model = PoET(None, tr, cfg["num_queries"], cfg["n_levels"], cfg["n_classes"], class_mode=cfg["class_mode"])
    model.load_state_dict({k: v for k, v in P.items() if not k.startswith("input_proj")}, strict=True)
    return model.to(DEV)


def _stack(out):
    t = torch.stack([a["pred_translation"] for a in out["aux_outputs"]] + [out["pred_translation"]])
    R = torch.stack([a["pred_rotation"] for a in out["aux_outputs"]] + [out["pred_rotation"]])
    return t, R


def test_model_train_mode_dropout_eval_parity_and_graph():
    """The reference default (dropout 0.1, model.train(), engine.py:38) runs; eval() of the same model is bit-identical
    to the dropout-0 model; a CUDA-graph replay draws a new mask every step."""
    from poet_b200 import ops
    from poet_b200.data_parallel import FlatGradReducer
    from poet_b200.graph import GraphedStep
    cfg = dict(S.CONFIGS["cfg2_b2"])
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=True)
    g_t, g_R = (x.to(DEV) for x in S.make_cotangents(cfg))
    srcs, masks = [s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]]
    m0, m1 = _model(cfg, P, 0.0), _model(cfg, P, 0.1)
    with torch.no_grad():
        ref_t, ref_R = _stack(m0.eval().forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])[0])
        ev_t, ev_R = _stack(m1.eval().forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])[0])
    assert torch.equal(ref_t, ev_t) and torch.equal(ref_R, ev_R)

    def loss_fn(out):
        t, R = _stack(out)
        return (t * g_t).sum() + (R * g_R).sum()

    m1.train()
    ops.set_dropout_seed(321)
    red = FlatGradReducer(m1.parameters())
    step = GraphedStep(m1, loss_fn, srcs, masks, inp["boxes"], inp["labels"], reducer=red, warmup=1)
    outs = []
    for _ in range(3):
        loss, out = step.run()
        outs.append((float(loss), _stack(out)[0].clone(), red.flat.clone()))
    torch.cuda.synchronize()
    assert all(math.isfinite(l) for l, _, _ in outs)
    assert not torch.equal(outs[0][1], outs[1][1]) and not torch.equal(outs[1][1], outs[2][1])      # fresh masks per replay
    assert not torch.equal(outs[0][1], ref_t)
    # dropout perturbs, it does not destroy: the train-mode poses stay near the eval poses
    assert float((outs[0][1] - ref_t).abs().mean()) < 0.5 * float(ref_t.abs().mean()) + 0.5
    assert all(torch.isfinite(g).all() and float(g.abs().max()) > 0 for _, _, g in outs)


def test_model_backward_matches_forward_with_dropout():
    """Directional derivative of the whole train-mode model at a FIXED dropout seed (every site: LayerNorm residual
    branches, FFN hidden bitmask, attention probabilities): the backward must regenerate exactly the forward's masks."""
    from poet_b200 import ops
    cfg = dict(S.CONFIGS["cfg2_b2"])
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=True)
    g_t, g_R = (x.to(DEV) for x in S.make_cotangents(cfg))
    srcs, masks = [s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]]
    model = _model(cfg, P, 0.1).train()

    def loss():
        ops.set_dropout_seed(777)                                 # same masks at every evaluation
        out, _ = model.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
        t, R = _stack(out)
        return ((t * g_t).sum() + (R * g_R).sum()).double()

    l0 = loss()
    assert float(l0) == float(loss())                            # deterministic given the seed
    l0.backward()
    groups = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        key = ".".join(k.split(".")[:4]) if k.startswith("transformer.") else k.split(".")[0]
        groups.setdefault(key, []).append(p)
    checked = 0
    for key, ps in groups.items():
        gnorm = math.sqrt(sum(float((p.grad.double() ** 2).sum()) for p in ps))
        pnorm = math.sqrt(sum(float((p.detach().double() ** 2).sum()) for p in ps))
        if gnorm == 0.0:
            continue
        vs = [p.grad / gnorm for p in ps]
        eps = min(1e-4 * pnorm, 0.02 / gnorm)
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
        checked += 1
    assert checked >= 12
