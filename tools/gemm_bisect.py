"""Pipeline bisection of the tcgen05 GEMM (POET_GEMM_DEBUG knobs): times one shape per process."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from poet_b200 import ops
    M, N, K = (int(v) for v in sys.argv[2:5])
    dev = "cuda:0"
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    hi = torch.empty(N, K, device=dev, dtype=torch.bfloat16); lo = torch.empty_like(hi)
    ops._call("poet_split_bf16", W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), ops._stream(W))
    out = torch.empty(M, N, device=dev)
    def run():
        ops.gemm(A, W, M, N, K, bias=b, out=out, b_split=(hi, lo))
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
    print(f"debug={os.environ.get('POET_GEMM_DEBUG','0'):>2s} {M}x{N}x{K}: {s.elapsed_time(e)/100*1e3:8.1f} us")
else:
    if True:
        for shape in ((25600, 256, 256), (25600, 1024, 256), (160, 256, 256)):
            for dbg in (0, 1, 2, 3, 4, 8, 7, 15):
                env = dict(os.environ, POET_GEMM_DEBUG=str(dbg))
                subprocess.run([sys.executable, __file__, "child", *map(str, shape)], env=env)
