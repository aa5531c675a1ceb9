"""Pipeline bisection of the tcgen05 GEMM (POET_GEMM_DEBUG knobs): times each shape in a child process per knob.
Needs a bisection build of the library (POET_GEMM_BISECT=1 python -m poet_b200.build --force): the knobs are compiled out of the product.
bits: 1 no A loads, 2 no A smem stores, 4 no TMA (B), 8 no epilogue stores, 16 no epilogue math, 32 no MMA issue."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from poet_b200 import ops
    dev = "cuda:0"
    ops.set_gemm_precision(os.environ.get("POET_BISECT_PREC", "bf16x3"))
    for shape in sys.argv[2:]:
        M, N, K = (int(v) for v in shape.split("x"))
        A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
        hi = torch.empty(N, K, device=dev, dtype=torch.bfloat16); lo = torch.empty_like(hi)
        ops._call("poet_split_bf16", W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), ops._stream(W))
        out = torch.empty(M, N, device=dev)
        def run():
            ops.gemm(A, W, M, N, K, bias=b, out=out, b_split=(hi, lo if ops.get_gemm_precision() == 'bf16x3' else None))
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3): run()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): run()
        g.replay(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5): g.replay()
        e.record(); torch.cuda.synchronize()
        print(f"debug={os.environ.get('POET_GEMM_DEBUG','0'):>2s} epi={os.environ.get('POET_GEMM_TMA_EPI','1')} {M}x{N}x{K}: {s.elapsed_time(e)/100*1e3:8.1f} us", flush=True)
else:
    shapes = ["25600x1024x256", "25600x256x256", "25600x256x1024"]
    for dbg in [int(v) for v in os.environ.get('POET_BISECT', '0,8,16,24,1,3,4,7,15,31,32,40,47,63').split(',')]:
        env = dict(os.environ, POET_GEMM_DEBUG=str(dbg))
        subprocess.run([sys.executable, __file__, "child", *shapes], env=env)
