"""Gradients of two half batches summed vs the whole batch in one process, next to the rerun noise of the whole batch
(one GPU).  Shows that sharding changes single gradient entries by per-cent amounts through ReLU / bilinear-cell flips
 although every kernel is deterministic up to the order of the scatter atomics.
usage (GPU box): python tools/shard_check.py"""
import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from poet_b200 import synthetic as S, ops
from poet_b200.data_parallel import FlatGradReducer
from test_gpu_distributed import _build, _step  # noqa: E402
dev = torch.device("cuda:0")
ops.set_gemm_precision("bf16x3")
cfg = dict(S.CONFIGS["cfg2_b2"], batch=4)
P = S.make_params(cfg); inp = S.make_inputs(cfg, pad_columns=True); g_t, g_R = S.make_cotangents(cfg)
def run(lo, hi):
    m = _build(cfg, P, dev); r = FlatGradReducer(m.parameters())
    _step(m, r, cfg, inp, g_t, g_R, lo, hi, dev, False)
    return r.flat.detach().clone(), r.offsets, [n for n, _ in m.named_parameters()]
full1, offs, names = run(0, 4)
full2, _, _ = run(0, 4)
a, _, _ = run(0, 2); b, _, _ = run(2, 4)
sh = a + b
bounds = list(offs) + [full1.numel()]
rows = []
for n, x, y in zip(names, bounds[:-1], bounds[1:]):
    f = full1[x:y]; sc = float(f.abs().max())
    if sc == 0: continue
    rows.append((float((sh[x:y] - f).abs().max()) / sc, float((full2[x:y] - f).abs().max()) / sc, n))
rows.sort(reverse=True)
for r in rows[:12]: print(f"shard-vs-full {r[0]:.2e}   rerun-noise {r[1]:.2e}   {r[2]}")
print("n over 2e-5:", sum(1 for r in rows if r[0] > 2e-5), "of", len(rows))
