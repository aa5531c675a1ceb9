"""Micro-benchmark of the hot kernels at the cfg2 shapes (one process, current env knobs):
CUDA-graph timing of 20 back-to-back launches per shape, L2-warm operands (as inside a step, where each
operand was just written by the producing kernel).  usage: python tools/kernel_micro.py [tag]
Prints one line per shape: us per launch, effective TFLOP/s (2MNK) and algorithmic GB/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from poet_b200 import ops  # noqa: E402

DEV = "cuda:0"


def timed(fn, n=20, reps=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (n * reps) * 1e3          # us


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "default"
    ops.set_gemm_precision("bf16x3")
    R = 25600
    print(f"# kernel_micro [{tag}] env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("POET_")))
    # forward / dgrad GEMMs with pre-split weights
    for (M, N, K, b_k, relu, what) in [(R, 256, 256, True, False, "proj fwd"), (R, 768, 256, True, False, "offs|logits fwd"),
                                       (R, 1024, 256, True, True, "FFN1 fwd +relu+bits"), (R, 256, 1024, True, False, "FFN2 fwd"),
                                       (R, 1024, 256, False, False, "FFN2 dgrad (gate bits)"), (R, 256, 1024, False, False, "FFN1 dgrad"),
                                       (R, 256, 768, False, False, "offs|logits dgrad"), (160, 256, 256, True, False, "decoder row"),
                                       (160, 1024, 256, True, True, "decoder FFN1")]:
        A = torch.randn(M, K, device=DEV)
        W = torch.randn((N, K) if b_k else (K, N), device=DEV)
        b = torch.randn(N, device=DEV)
        hi = torch.empty(W.shape, device=DEV, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        ops._call("poet_split_bf16", W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), ops._stream(W))
        out = torch.empty(M, N, device=DEV)
        bits = ops.relu_bits_buffer(M, N, K, DEV) if (relu or "gate" in what) else None
        if bits is not None:
            bits.random_(-2 ** 31, 2 ** 31 - 1)
        kw = {}
        if relu and bits is not None:
            kw["relu_bits"] = bits
        if "gate" in what:
            if bits is not None:
                kw["gate_bits"] = bits
            else:
                kw["gate"] = torch.randn(M, N, device=DEV)
        us = timed(lambda: ops.gemm(A, W, M, N, K, b_kcontig=b_k, bias=b, relu=relu, out=out, b_split=(hi, lo), **kw))
        fl, by = 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)
        print(f"gemm {what:26s} {M}x{N}x{K:<5d} {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s(2MNK) {3 * fl / us / 1e6:7.1f} TF/s issued  {by / us / 1e3:7.0f} GB/s")
    # weight gradients (accumulate into an existing gradient)
    for (Mo, No, what) in [(256, 256, "proj wgrad"), (512, 256, "offsets wgrad"), (1024, 256, "FFN1 wgrad"), (256, 1024, "FFN2 wgrad")]:
        dY, X = torch.randn(R, Mo, device=DEV), torch.randn(R, No, device=DEV)
        out = torch.zeros(Mo, No, device=DEV)
        us = timed(lambda: ops.gemm(dY, X, Mo, No, R, a_kcontig=False, b_kcontig=False, out=out, accumulate=True))
        fl, by = 2.0 * Mo * No * R, 4.0 * (R * Mo + R * No + Mo * No)
        print(f"gemm {what:26s} {Mo}x{No}x{R:<5d} {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s(2MNK) {3 * fl / us / 1e6:7.1f} TF/s issued  {by / us / 1e3:7.0f} GB/s")
    # MSDA block (encoder shape), forward and backward
    shapes = ((30, 40), (15, 20), (8, 10), (4, 5))
    B, S, M, D, L, P = 16, 1600, 16, 16, 4, 4
    value = torch.randn(B, S, M * D, device=DEV)
    oa = torch.randn(B, S, M * L * P * 3, device=DEV)
    ref = torch.rand(B, S, L, 2, device=DEV)
    n_off = M * L * P * 2
    us = timed(lambda: ops.msda_fwd_raw(value, oa, oa.shape[2], oa.view(-1)[n_off:], oa.shape[2], ref, shapes, B, S, S, M, D, L, P, 1))
    by = 4.0 * B * (S * M * D + 3 * S * M * L * P + S * M * D)
    print(f"msda fwd  encoder block       {us:8.1f} us  {by / us / 1e3:7.0f} GB/s algorithmic")
    go = torch.randn(B, S, M * D, device=DEV)
    gv = torch.zeros_like(value)
    goa = torch.empty_like(oa)
    sh = ops.shapes_array(shapes)
    us = timed(lambda: ops._call("poet_msda_bwd", value.data_ptr(), oa.data_ptr(), oa.shape[2], oa.view(-1)[n_off:].data_ptr(),
                                 oa.shape[2], ref.data_ptr(), go.data_ptr(), gv.data_ptr(), goa.data_ptr(),
                                 goa.view(-1)[n_off:].data_ptr(), sh, B, S, S, M, D, L, P, 1, ops._stream(value)))
    by = 4.0 * B * (2 * S * M * D + 6 * S * M * L * P + S * M * D)
    print(f"msda bwd  encoder block       {us:8.1f} us  {by / us / 1e3:7.0f} GB/s algorithmic")
    # residual + LayerNorm
    x, r = torch.randn(R, 256, device=DEV), torch.randn(R, 256, device=DEV)
    gam, bet = torch.ones(256, device=DEV), torch.zeros(256, device=DEV)
    y, xh, rs = torch.empty_like(x), torch.empty_like(x), torch.empty(R, device=DEV)
    us = timed(lambda: ops._call("poet_add_layernorm_fwd", x.data_ptr(), r.data_ptr(), gam.data_ptr(), bet.data_ptr(), None,
                                 y.data_ptr(), None, xh.data_ptr(), rs.data_ptr(), R, 256, 1e-5, None, 0, 0.0, ops._stream(x)))
    print(f"add+LN fwd                    {us:8.1f} us  {4.0 * R * 256 * 4 / us / 1e3:7.0f} GB/s")


if __name__ == "__main__":
    main()
