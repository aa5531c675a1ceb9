"""One eager forward+backward step of the benchmark workload between cudaProfilerStart/Stop, for
    ncu --profile-from-start off ... python tools/profile_step.py [--config cfg2] [--precision bf16x3]
(see profiles/README.md for the exact commands).  Not a benchmark: numbers printed under ncu are never bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    from bench import build_gpu_model
    from poet_b200 import ops, synthetic as S
    from poet_b200.data_parallel import FlatGradReducer
    ops.set_gemm_precision(args.precision)
    dev = torch.device("cuda:0")
    cfg = S.CONFIGS[args.config]
    model = build_gpu_model(cfg, dev)
    red = FlatGradReducer(model.parameters())
    inp = S.make_inputs(cfg)
    g_t, g_R = (t.to(dev) for t in S.make_cotangents(cfg))
    srcs, masks = [s.to(dev) for s in inp["srcs"]], [m.to(dev) for m in inp["masks"]]

    def step():
        red.zero()
        out, _ = model.forward_pyramid(srcs, masks, inp["boxes"], inp["labels"])
        t = torch.stack([a["pred_translation"] for a in out["aux_outputs"]] + [out["pred_translation"]])
        R = torch.stack([a["pred_rotation"] for a in out["aux_outputs"]] + [out["pred_rotation"]])
        ((t * g_t).sum() + (R * g_R).sum()).backward()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
