"""Gradient agreement of the CUDA path with the fp32 CPU oracle, per GEMM precision (GPU box).
Prints, per config/precision, forward errors and the parameters with the largest relative-L2 gradient error."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import poet_oracle as O  # noqa: E402
from poet_b200 import ops, synthetic as S  # noqa: E402
from test_gpu_model import build_model, stack_outputs  # noqa: E402

DEV = "cuda:0"
for name, pad in (("cfg1", False), ("cfg2_b2", True)):
    cfg = S.CONFIGS[name]
    P = S.make_params(cfg)
    inp = S.make_inputs(cfg, pad_columns=pad)
    g_t, g_R = S.make_cotangents(cfg)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    cap = {}
    O.poet_path_forward(Pr, cfg, inp["srcs"], inp["masks"], inp["boxes"], inp["labels"], capture=cap)
    O.synthetic_loss((cap["translation_all"], cap["rotation_all"]), g_t, g_R).backward()
    for prec in ("fp32", "bf16x3", "mixed", "bf16"):
        ops.set_gemm_precision("bf16x3" if prec == "mixed" else prec)
        model = build_model(cfg, P)
        if prec == "mixed":                      # cfg4 throughput mode: single-pass bf16 on the token-row GEMMs only
            model.transformer.set_throughput_mode(True)
        out, _ = model.forward_pyramid([s.to(DEV) for s in inp["srcs"]], [m.to(DEV) for m in inp["masks"]],
                                       inp["boxes"], inp["labels"])
        t, R = stack_outputs(out)
        ((t * g_t.to(DEV)).sum() + (R * g_R.to(DEV)).sum()).backward()
        dt = float((t.detach().cpu() - cap["translation_all"]).abs().max())
        dR = float((R.detach().cpu() - cap["rotation_all"]).abs().max())
        rows = []
        for k, p in model.named_parameters():
            ref = Pr[k].grad
            if ref is None:
                continue
            err = (p.grad.cpu().double() - ref.double()).abs()
            scale = float(ref.abs().max()) + 1e-12
            rows.append((float(err.norm() / (ref.double().norm() + 1e-12)), float((err > 1e-3 * scale).double().mean()),
                         float(err.max()) / scale, k))
        rows.sort(reverse=True)
        print(f"{name} {prec}: fwd |dt|={dt:.2e} |dR|={dR:.2e}; worst grads (rel_l2, frac>1e-3, max/scale):")
        for r in rows[:5]:
            print(f"    {r[0]:.2e} {r[1]:.3f} {r[2]:.2e} {r[3]}")
