"""MSDA encoder-block micro-benchmark only (forward + backward, cfg2 or cfg5 shape), current env knobs.
usage: python tools/msda_micro.py [tag] [cfg2|cfg5] [offset_scale]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from poet_b200 import ops  # noqa: E402
from kernel_micro import timed  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "default"
    which = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    if which == "cfg5":
        shapes, B, M = ((60, 80), (30, 40), (15, 20), (8, 10)), 8, 8
    elif which == "ref8":                                   # REF pyramid, 8 heads of 32 channels (cfg1 geometry at batch 16)
        shapes, B, M = ((30, 40), (15, 20), (8, 10), (4, 5)), 16, 8
    else:
        shapes, B, M = ((30, 40), (15, 20), (8, 10), (4, 5)), 16, 16
    S = sum(h * w for h, w in shapes)
    D, L, P = 256 // M, 4, 4
    DEV = "cuda:0"
    value = torch.randn(B, S, M * D, device=DEV)
    oa = torch.randn(B, S, M * L * P * 3, device=DEV)
    oa[..., : M * L * P * 2] *= scale
    ref = torch.rand(B, S, L, 2, device=DEV)
    n_off = M * L * P * 2
    us = timed(lambda: ops.msda_fwd_raw(value, oa, oa.shape[2], oa.view(-1)[n_off:], oa.shape[2], ref, shapes, B, S, S, M, D, L, P, 1))
    by = 4.0 * B * (S * M * D + 3 * S * M * L * P + S * M * D)
    env = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("POET_"))
    print(f"[{tag}] {which} scale={scale} {env}: msda fwd {us:8.1f} us {by / us / 1e3:7.0f} GB/s", end="  |  ")
    go = torch.randn(B, S, M * D, device=DEV)
    gv = torch.zeros_like(value)
    goa = torch.empty_like(oa)
    sh = ops.shapes_array(shapes)
    us = timed(lambda: ops._call("poet_msda_bwd", value.data_ptr(), oa.data_ptr(), oa.shape[2], oa.view(-1)[n_off:].data_ptr(),
                                 oa.shape[2], ref.data_ptr(), go.data_ptr(), gv.data_ptr(), goa.data_ptr(),
                                 goa.view(-1)[n_off:].data_ptr(), sh, B, S, S, M, D, L, P, 1, ops._stream(value)))
    by = 4.0 * B * (2 * S * M * D + 6 * S * M * L * P + S * M * D)
    print(f"msda bwd {us:8.1f} us {by / us / 1e3:7.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
