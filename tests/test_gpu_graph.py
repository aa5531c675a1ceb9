"""GPU tests of the whole-step CUDA graph (poet_b200.graph.GraphedStep) on the BENCHMARKED shape, of the
graph + fused-optimizer training loop, and of the opt-in rule for direct gradient accumulation.

Tolerances (BASELINE.json north_star): |translation| <= 1e-4 abs, rotation <= 1e-3 on every decoder layer."""
import copy

import pytest
import torch

from oracle import poet_oracle as O
from poet_b200 import synthetic as S
from test_gpu_model import build_model, stack_outputs, grad_close, check_grad_census, TOL_T, TOL_R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _loss_fn(g_t, g_R):
    def loss_fn(out):
        t, R = stack_outputs(out)
        return (t * g_t).sum() + (R * g_R).sum()
    return loss_fn


def test_benchmarked_shape_graph_replay_vs_oracle():
    """bench.py's default line: cfg2 at B=16 (S = Lq = 1600, M=16: the mode-1 slab forward with its query split for
    B*M = 256, the dense-tile MSDA backward, 256-wide GEMM tiles), bf16x3, replayed from the whole-step CUDA graph.
    All five decoder layers' poses vs the oracle forward; every parameter gradient and the pyramid gradients of the
    first level vs the oracle backward on the same 16 images."""
    from poet_b200 import ops
    from poet_b200.data_parallel import FlatGradReducer
    from poet_b200.graph import GraphedStep
    cfg = S.CONFIGS["cfg2"]
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg)
    g_t, g_R = S.make_cotangents(cfg)
    old = ops.get_gemm_precision()
    ops.set_gemm_precision("bf16x3")
    try:
        model = build_model(cfg, P)
        red = FlatGradReducer(model.parameters())
        srcs, masks = [s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]]
        step = GraphedStep(model, _loss_fn(g_t.to(DEV), g_R.to(DEV)), srcs, masks, inp["boxes"], inp["labels"], reducer=red)
        step.run()
        loss, out = step.run()                                     # a REPLAY, like every timed step of the bench
        torch.cuda.synchronize()
        t, R = (x.detach().cpu() for x in stack_outputs(out))
        grads = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()}
    finally:
        ops.set_gemm_precision(old)

    torch.set_num_threads(max(1, torch.get_num_threads()))
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    cap = {}
    O.poet_path_forward(Pr, cfg, inp["srcs"], inp["masks"], inp["boxes"], inp["labels"], capture=cap)
    ref_loss = O.synthetic_loss((cap["translation_all"], cap["rotation_all"]), g_t, g_R)
    ref_loss.backward()
    assert t.shape[0] == cfg["dec_layers"]
    for l in range(cfg["dec_layers"]):
        assert float((t[l] - cap["translation_all"][l]).abs().max()) < TOL_T, f"layer {l} translation"
        assert float((R[l] - cap["rotation_all"][l]).abs().max()) < TOL_R, f"layer {l} rotation"
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * max(1.0, abs(float(ref_loss)))
    tight, loose = [], []
    for k, g in grads.items():
        ref = Pr[k].grad
        if ref is None:
            assert float(g.abs().max()) == 0.0, k
            continue
        ok_t, ok_l = grad_close(g, ref, "bf16x3")
        tight.append(ok_t)
        loose.append((k, ok_l))
    assert len(tight) >= 100
    check_grad_census(tight, loose)


def test_graph_with_fused_optimizer_matches_eager_loop():
    """The documented 'training at speed' recipe: GraphedStep(optimizer=FusedClipAdamW) replayed for several steps
    must follow the same parameter trajectory as the eager loop (forward, backward, optimizer.step()).  The captured
    forward reads weight planes written by the optimizer and a flat copy of the 1-D parameters refreshed inside the
    graph: a stale copy (biases of the fused [offsets | logits] projection) makes the two loops diverge."""
    from poet_b200 import ops
    from poet_b200.data_parallel import FlatGradReducer
    from poet_b200.graph import GraphedStep
    from poet_b200.optim import FusedClipAdamW
    cfg = dict(S.CONFIGS["tiny16"], batch=4)
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=True)
    g_t, g_R = (x.to(DEV) for x in S.make_cotangents(cfg))
    srcs, masks = [s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]]
    loss_fn = _loss_fn(g_t, g_R)
    kw = dict(lr=5e-3, weight_decay=1e-2, max_norm=10.0, eps=1e-3)   # large lr: a stale bias shows after one step; large eps:
    # gradients that are pure rounding noise (softmax-invariant key bias) must not become +-lr steps of random sign
    n_steps = 4

    old = ops.get_gemm_precision()
    ops.set_gemm_precision("bf16x3")
    try:
        eager = build_model(cfg, P)
        red_e = FlatGradReducer(eager.parameters())
        opt_e = FusedClipAdamW(eager, red_e, **kw)
        losses_e = []
        for _ in range(n_steps):
            opt_e.zero_grad()
            out, _ = eager.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
            loss = loss_fn(out)
            loss.backward()
            opt_e.step()
            losses_e.append(float(loss))

        graphed = build_model(cfg, P)
        red_g = FlatGradReducer(graphed.parameters())
        opt_g = FusedClipAdamW(graphed, red_g, **kw)
        step = GraphedStep(graphed, loss_fn, srcs, masks, inp["boxes"], inp["labels"], reducer=red_g, optimizer=opt_g, warmup=1)
        losses_g = []
        for _ in range(n_steps):
            loss, _ = step.run()
            losses_g.append(float(loss))
            opt_g.step()
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_precision(old)
    assert losses_e[0] != losses_e[-1]                            # the parameters really moved
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(a)), (losses_e, losses_g)
    for (k, p), (_, q) in zip(eager.named_parameters(), graphed.named_parameters()):
        err = float((p - q).abs().max())
        assert err <= 1e-4 * max(1.0, float(p.abs().max())), f"{k}: {err:.3e}"


def test_direct_grad_accumulation_is_opt_in():
    """A parameter whose .grad is an ordinary tensor goes through autograd's AccumulateGrad (tensor hooks and
    post-accumulate hooks fire, torch.autograd.grad returns the gradient); only FlatGradReducer's registered arena
    views are written directly by the backward kernels.  Both routes give the same numbers."""
    from poet_b200 import ops
    from poet_b200.data_parallel import FlatGradReducer
    cfg = S.CONFIGS["tiny16"]
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg)
    g_t, g_R = (x.to(DEV) for x in S.make_cotangents(cfg))
    srcs, masks = [s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]]
    loss_fn = _loss_fn(g_t, g_R)

    plain = build_model(cfg, P)
    fired = []
    w = plain.transformer.encoder.layers[0].linear1.weight
    w.register_hook(lambda g: fired.append("tensor"))
    w.register_post_accumulate_grad_hook(lambda p: fired.append("post"))
    for p in plain.parameters():
        p.grad = torch.zeros_like(p)                              # pre-existing, NOT registered: must not be written directly
    out, _ = plain.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
    loss_fn(out).backward()
    assert fired == ["tensor", "post"]
    out, _ = plain.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
    (gw,) = torch.autograd.grad(loss_fn(out), [w])
    assert gw is not None and float((gw - w.grad).abs().max()) <= 1e-5 * float(gw.abs().max())

    arena = build_model(cfg, P)
    red = FlatGradReducer(arena.parameters())
    red.zero()
    out, _ = arena.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
    loss_fn(out).backward()
    torch.cuda.synchronize()
    for (k, p), (_, q) in zip(plain.named_parameters(), arena.named_parameters()):
        if "reference_points" in k:
            continue
        scale = max(float(p.grad.abs().max()), 1e-12)
        assert float((p.grad - q.grad).abs().max()) <= 2e-5 * scale, k
    ops.unregister_direct_grad_slots([p.grad for p in arena.parameters()])
