"""GPU parity tests, op level: every C-ABI kernel against the CPU oracle / fp64 torch math on the
same seeded inputs.  Tolerances are written next to each check (fp32 kernels vs fp64 truth)."""
import math

import pytest
import torch

from oracle import poet_oracle as O
from poet_b200 import synthetic as S
from helpers import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ops():
    from poet_b200 import ops as _ops
    return _ops


def rel_err(got, ref):
    ref = ref.double()
    return float((got.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def bad_fraction(got, ref, tol):
    """Fraction of elements off by more than tol*max|ref|.  Bilinear-location gradients jump at integer
    pixel coordinates, so an fp32 kernel and the fp64 oracle may legitimately disagree on the handful of
    samples that sit within fp32 rounding of a pixel edge."""
    ref = ref.double()
    err = (got.double().cpu() - ref).abs() / ref.abs().max().clamp_min(1e-30)
    return float((err > tol).double().mean())


# ------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(160, 256, 256), (160, 66, 256), (160, 256, 66), (1600, 768, 256),
                                   (3200, 1024, 256), (37, 50, 19), (25600, 256, 256)])
@pytest.mark.parametrize("a_k,b_k", [(True, True), (True, False), (False, False), (False, True)])
def test_gemm_fp32_layouts(M, N, K, a_k, b_k):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K) if a_k else (K, M), generator=g)
    Bm = torch.randn((N, K) if b_k else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = (A.double() if a_k else A.double().t()) @ (Bm.double().t() if b_k else Bm.double()) + bias.double()
    out = ops().gemm(A.to(DEV), Bm.to(DEV), M, N, K, a_kcontig=a_k, b_kcontig=b_k, bias=bias.to(DEV),
                     precision=ops().GEMM_FP32)
    assert rel_err(out, ref) < 2e-6 * math.sqrt(K)


def test_gemm_epilogues_and_splitk():
    o = ops()
    g = torch.Generator().manual_seed(5)
    M, N, K = 512, 256, 4096
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    gate = torch.randn(M, N, generator=g)
    mask = (torch.rand(M, generator=g) < 0.2).to(torch.uint8)
    ref = A.double() @ W.double().t() + b.double()
    relu = o.gemm(A.to(DEV), W.to(DEV), M, N, K, bias=b.to(DEV), relu=True, precision=o.GEMM_FP32)
    assert rel_err(relu, ref.clamp_min(0)) < 1e-4
    gated = o.gemm(A.to(DEV), W.to(DEV), M, N, K, bias=b.to(DEV), gate=gate.to(DEV), precision=o.GEMM_FP32)
    assert rel_err(gated, ref * (gate > 0)) < 1e-4
    masked = o.gemm(A.to(DEV), W.to(DEV), M, N, K, bias=b.to(DEV), row_mask=mask.to(DEV), precision=o.GEMM_FP32)
    assert rel_err(masked, ref * (mask == 0)[:, None]) < 1e-4
    # wgrad shape: tiny output, long K -> split-K path; with and without accumulate
    X, dY = torch.randn(25600, 64, generator=g), torch.randn(25600, 96, generator=g)
    ref_w = dY.double().t() @ X.double()
    dW = o.gemm(dY.to(DEV), X.to(DEV), 96, 64, 25600, a_kcontig=False, b_kcontig=False, precision=o.GEMM_FP32)
    assert rel_err(dW, ref_w) < 1e-4
    base = torch.randn(96, 64, generator=g)
    acc = o.gemm(dY.to(DEV), X.to(DEV), 96, 64, 25600, a_kcontig=False, b_kcontig=False, out=base.to(DEV).clone(),
                 accumulate=True, precision=o.GEMM_FP32)
    assert rel_err(acc, ref_w + base.double()) < 1e-4
    cs = o.colsum(dY.to(DEV), 25600, 96)
    assert rel_err(cs, dY.double().sum(0)) < 1e-5


def test_linear_and_mlp_autograd():
    o = ops()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 50, 256, generator=g)
    Ws = [torch.randn(256, 256, generator=g) * 0.06, torch.randn(256, 256, generator=g) * 0.06,
          torch.randn(66, 256, generator=g) * 0.06]
    bs = [torch.randn(256, generator=g) * 0.1, torch.randn(256, generator=g) * 0.1, torch.randn(66, generator=g) * 0.1]
    cot = torch.randn(3, 50, 66, generator=g)

    def run(dev, dt):
        xs = x.to(dev, dt).requires_grad_(True)
        ps = [(W.to(dev, dt).requires_grad_(True), b.to(dev, dt).requires_grad_(True)) for W, b in zip(Ws, bs)]
        if dev == "cpu":
            h = xs
            for i, (W, b) in enumerate(ps):
                h = torch.nn.functional.linear(h, W, b)
                if i < 2:
                    h = h.relu()
        else:
            h = o.mlp(xs, ps)
        (h * cot.to(dev, dt)).sum().backward()
        return [h.detach(), xs.grad] + [t.grad for pr in ps for t in pr]

    ref, got = run("cpu", torch.float64), run(DEV, torch.float32)
    for r, t in zip(ref, got):
        assert rel_err(t, r) < 2e-5


# ------------------------------------------------------------------------------- MSDA
def _msda_inputs(B, M, D, Lq, P, shapes, seed):
    g = torch.Generator().manual_seed(seed)
    S_ = sum(h * w for h, w in shapes)
    L = len(shapes)
    value = torch.randn(B, S_, M, D, generator=g)
    loc = torch.rand(B, Lq, M, L, P, 2, generator=g) * 1.6 - 0.3
    loc[0, 0] = -1.0
    attn = torch.softmax(torch.randn(B, Lq, M, L * P, generator=g), -1).view(B, Lq, M, L, P)
    cot = torch.randn(B, Lq, M * D, generator=g)
    return value, loc, attn, cot


@pytest.mark.parametrize("B,M,D,Lq,P,shapes", [
    (1, 2, 8, 2, 2, [(6, 4), (3, 2)]),                                   # upstream test.py shapes (D rounded to 8)
    (2, 16, 16, 37, 4, [(30, 40), (15, 20), (8, 10), (4, 5)]),           # YCB-V head geometry on the REF pyramid (few queries: warp-per-(q,m) kernels)
    (2, 8, 32, 50, 4, [(30, 40), (15, 20), (8, 10), (4, 5)]),            # cfg1 head geometry
    (1, 4, 64, 9, 4, [(6, 8), (3, 4), (2, 2), (1, 1)]),
    (2, 3, 16, 1700, 4, [(6, 8), (3, 4), (2, 2), (1, 1)]),               # smem-slab forward: S=65 (TMA box tail, OOB fill)
    (3, 8, 8, 21, 4, [(6, 8), (3, 4), (2, 2), (1, 1)]),                  # warp-per-(q,m) kernels, D=8
    (2, 16, 16, 500, 4, [(30, 40), (15, 20), (8, 10), (4, 5)]),          # smem-slab forward on the REF pyramid; tile backward, levels 2-3 dense
    (2, 16, 16, 333, 4, [(30, 40), (15, 20), (10, 12), (8, 10)]),        # tile backward: only the last level is dense (cfg5-like), ragged last tile
    (2, 16, 16, 300, 4, [(12, 16), (6, 8), (4, 5), (2, 3)]),             # tile backward: levels 1-3 dense
    (2, 8, 32, 700, 4, [(30, 40), (15, 20), (8, 10), (4, 5)]),           # tile backward, D = 32 (8 lanes per (q,m), 32-query tiles)
    (1, 8, 32, 1250, 4, [(30, 40), (15, 20), (10, 12), (8, 10)]),        # tile backward, D = 32, only the last level dense (cfg5-like)
])
def test_msda_core_fwd_bwd(B, M, D, Lq, P, shapes):
    o = ops()
    value, loc, attn, cot = _msda_inputs(B, M, D, Lq, P, shapes, seed=B + M + D)
    v64, l64, a64 = (t.double().requires_grad_(True) for t in (value, loc, attn))
    ref = O.msda_core(v64, shapes, l64, a64)
    (ref * cot.double()).sum().backward()
    v, l, a = (t.to(DEV).requires_grad_(True) for t in (value, loc, attn))
    out = o.msda_core(v, shapes, l, a)
    (out * cot.to(DEV)).sum().backward()
    assert float(out[0, 0].abs().max()) == 0.0                            # dummy reference point -> exact zero
    assert rel_err(out, ref) < 5e-6
    assert rel_err(v.grad, v64.grad) < 5e-6
    assert rel_err(a.grad, a64.grad) < 5e-6
    assert bad_fraction(l.grad, l64.grad, 5e-5) < 1e-3


@pytest.mark.parametrize("Lq", [64, 480])                                # 480: forward served by the smem-slab kernel
def test_msda_block_matches_module_math(Lq):
    """mode 1 (fused softmax + ref + off/(W,H)) == oracle module math, forward and all gradients."""
    o = ops()
    shapes = [(30, 40), (15, 20), (8, 10), (4, 5)]
    B, M, D, L, P = 2, 16, 16, 4, 4
    g = torch.Generator().manual_seed(21)
    S_ = sum(h * w for h, w in shapes)
    value = torch.randn(B, S_, M * D, generator=g)
    off = torch.randn(B, Lq, M * L * P * 2, generator=g) * 3
    logit = torch.randn(B, Lq, M * L * P, generator=g)
    ref_pts = torch.rand(B, Lq, L, 2, generator=g)
    ref_pts[1, 3] = -1.0
    cot = torch.randn(B, Lq, M * D, generator=g)
    v64, o64, g64 = (t.double().requires_grad_(True) for t in (value, off, logit))
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float64)
    loc = ref_pts.double()[:, :, None, :, None, :] + o64.view(B, Lq, M, L, P, 2) / wh[None, None, None, :, None, :]
    a = torch.softmax(g64.view(B, Lq, M, L * P), -1).view(B, Lq, M, L, P)
    ref = O.msda_core(v64.view(B, S_, M, D), shapes, loc, a)
    (ref * cot.double()).sum().backward()
    v = value.to(DEV).requires_grad_(True)
    oa = torch.cat((off, logit), -1).to(DEV).requires_grad_(True)
    out = o.msda_block(v, oa, ref_pts.to(DEV), shapes, M, L, P)
    (out * cot.to(DEV)).sum().backward()
    assert rel_err(out, ref) < 5e-6
    assert rel_err(v.grad, v64.grad) < 5e-6
    assert bad_fraction(oa.grad[..., : M * L * P * 2], o64.grad, 5e-5) < 1e-3
    assert rel_err(oa.grad[..., M * L * P * 2:], g64.grad) < 5e-5


def test_msda_full_size_properties():
    """BASELINE cfg2 size (B=16, S=Lq=1600, M=16, D=16): linearity in value and a per-head checksum
    against uniform attention on a constant map, which no CPU oracle run is needed for."""
    o = ops()
    shapes = [(30, 40), (15, 20), (8, 10), (4, 5)]
    B, M, D, L, P = 16, 16, 16, 4, 4
    S_ = Lq = 1600
    g = torch.Generator().manual_seed(77)
    v1 = torch.randn(B, S_, M, D, generator=g).to(DEV)
    v2 = torch.randn(B, S_, M, D, generator=g).to(DEV)
    loc = (torch.rand(B, Lq, M, L, P, 2, generator=g) * 0.7 + 0.15).to(DEV)    # strictly interior on every level
    attn = torch.softmax(torch.randn(B, Lq, M, L * P, generator=g), -1).view(B, Lq, M, L, P).to(DEV)
    a, b_, c = (o.msda_core(v, shapes, loc, attn) for v in (v1, v2, 2.0 * v1 - 3.0 * v2))
    assert float((c - (2.0 * a - 3.0 * b_)).abs().max()) < 5e-5
    const = torch.ones(B, S_, M, D, device=DEV) * torch.arange(1, M + 1, device=DEV).view(1, 1, M, 1)
    out = o.msda_core(const, shapes, loc, attn).view(B, Lq, M, D)         # interior samples of a constant map
    assert float((out - torch.arange(1, M + 1, device=DEV).view(1, 1, M, 1)).abs().max()) < 1e-4


@pytest.mark.parametrize("R,n_alias,with_pos", [(160, 1, False), (160, 2, True), (3200, 3, True), (160, 4, True)])
def test_add_layernorm_reader_handles(R, n_alias, with_pos):
    """One autograd handle per reader of a LayerNorm result (add_layernorm(n_alias=...)): the readers' gradients reach
    poet_layernorm_bwd as separate pointers (up to four; more are pre-added) and must sum to what autograd's own
    accumulation gives for a tensor that is read n_alias + 1 times."""
    o = ops()
    C = 256
    g = torch.Generator().manual_seed(R + 7 * n_alias)
    x, r, pos = (torch.randn(R, C, generator=g) for _ in range(3))
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    cots = [torch.randn(R, C, generator=g) for _ in range(n_alias + 2)]

    xs, rs, ga, be = (t.double().requires_grad_(True) for t in (x, r, gamma, beta))
    y = torch.nn.functional.layer_norm(xs + rs, (C,), ga, be, 1e-5)
    loss = sum((y * c.double()).sum() for c in cots[:n_alias + 1])
    if with_pos:
        loss = loss + ((y + pos.double()) * cots[-1].double()).sum()
    loss.backward()
    ref = [xs.grad, rs.grad, ga.grad, be.grad]

    xd, rd, gd, bd = (t.to(DEV).requires_grad_(True) for t in (x, r, gamma, beta))
    res = o.add_layernorm(xd, rd, gd, bd, pos=pos.to(DEV) if with_pos else None, n_alias=n_alias)
    handles = [res[0]] + list(res[2 if with_pos else 1:])
    assert len(handles) == n_alias + 1
    loss = sum((h * c.to(DEV)).sum() for h, c in zip(handles, cots))
    if with_pos:
        loss = loss + (res[1] * cots[-1].to(DEV)).sum()
    loss.backward()
    for got, want in zip((xd.grad, rd.grad, gd.grad, bd.grad), ref):
        assert rel_err(got, want) < 2e-5


# ------------------------------------------------------------------------------- block-level entry points
@pytest.mark.parametrize("R", [160, 3200])                              # query rows (latency kernel) / token rows (tcgen05)
def test_block_entry_points_ffn_fused_and_linear_epilogue(R):
    """poet_ffn_fused / poet_linear_epilogue (one C-ABI call per reference sub-block, inference semantics) against fp64
    torch: reference deformable_transformer.py:193-197 + 205-206 (FFN + norm2) and :201-204 (output_proj + norm1)."""
    o = ops()
    g = torch.Generator().manual_seed(R)
    Cc, F = 256, 1024
    x, res = torch.randn(R, Cc, generator=g), torch.randn(R, Cc, generator=g)
    W1, b1 = torch.randn(F, Cc, generator=g) / 16, torch.randn(F, generator=g)
    W2, b2 = torch.randn(Cc, F, generator=g) / 32, torch.randn(Cc, generator=g)
    gam, bet = torch.randn(Cc, generator=g), torch.randn(Cc, generator=g)
    d = lambda t: t.double()
    h = torch.relu(d(x) @ d(W1).t() + d(b1))
    ref_ffn = torch.nn.functional.layer_norm(d(x) + h @ d(W2).t() + d(b2), (Cc,), d(gam), d(bet), 1e-5)
    c = lambda t: t.to(DEV)
    got = o.ffn_fused(c(x), c(W1), c(b1), c(W2), c(b2), c(gam), c(bet))
    assert rel_err(got, ref_ffn) < 2e-5
    Wp, bp = W2[:, :Cc].contiguous(), b2
    ref_lin = d(x) @ d(Wp).t() + d(bp)
    assert rel_err(o.linear_epilogue(c(x), c(Wp), c(bp)), ref_lin) < 2e-5
    assert rel_err(o.linear_epilogue(c(x), c(Wp), c(bp), relu=True), ref_lin.clamp_min(0)) < 2e-5
    ref_ln = torch.nn.functional.layer_norm(d(res) + ref_lin, (Cc,), d(gam), d(bet), 1e-5)
    assert rel_err(o.linear_epilogue(c(x), c(Wp), c(bp), residual=c(res), gamma=c(gam), beta=c(bet)), ref_ln) < 2e-5
    ref_ln0 = torch.nn.functional.layer_norm(ref_lin, (Cc,), d(gam), d(bet), 1e-5)
    assert rel_err(o.linear_epilogue(c(x), c(Wp), c(bp), gamma=c(gam), beta=c(bet)), ref_ln0) < 2e-5


# ------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("R,C,with_r,with_pos", [(160, 256, True, True), (3200, 256, True, False), (77, 128, False, False),
                                                 (50, 1024, True, True)])
def test_add_layernorm_fwd_bwd(R, C, with_r, with_pos):
    o = ops()
    g = torch.Generator().manual_seed(R + C)
    x, r, pos = (torch.randn(R, C, generator=g) for _ in range(3))
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    c1, c2 = torch.randn(R, C, generator=g), torch.randn(R, C, generator=g)

    def run(dev, dt):
        xs, rs, ga, be = (t.to(dev, dt).requires_grad_(True) for t in (x, r, gamma, beta))
        ps = pos.to(dev, dt).requires_grad_(True)
        if dev == "cpu":
            y = torch.nn.functional.layer_norm(xs + rs if with_r else xs, (C,), ga, be, 1e-5)
            y2 = y + ps
        else:
            res = o.add_layernorm(xs, rs if with_r else None, ga, be, pos=ps if with_pos else None)
            y, y2 = res if with_pos else (res, None)
        loss = (y * c1.to(dev, dt)).sum()
        if with_pos:
            loss = loss + (y2 * c2.to(dev, dt)).sum()
        loss.backward()
        outs = [y.detach(), xs.grad, ga.grad, be.grad]
        if with_r:
            outs.append(rs.grad)
        if with_pos:
            outs += [y2.detach(), ps.grad]
        return outs

    for r_, t_ in zip(run("cpu", torch.float64), run(DEV, torch.float32)):
        assert rel_err(t_, r_) < 2e-5


# ------------------------------------------------------------------------------- decoder self-attention
@pytest.mark.parametrize("B,Q,M,D", [(16, 10, 16, 16), (3, 5, 8, 32), (2, 25, 8, 32), (1, 32, 4, 64)])
def test_mha_smallq_fwd_bwd(B, Q, M, D):
    o = ops()
    C = M * D
    g = torch.Generator().manual_seed(B * Q)
    qk, v, cot = torch.randn(B, Q, 2 * C, generator=g), torch.randn(B, Q, C, generator=g), torch.randn(B, Q, C, generator=g)
    qk64, v64 = qk.double().requires_grad_(True), v.double().requires_grad_(True)
    q_, k_ = qk64[..., :C].view(B, Q, M, D).transpose(1, 2), qk64[..., C:].view(B, Q, M, D).transpose(1, 2)
    p = torch.softmax(q_ @ k_.transpose(-1, -2) / math.sqrt(D), -1)
    ref = (p @ v64.view(B, Q, M, D).transpose(1, 2)).transpose(1, 2).reshape(B, Q, C)
    (ref * cot.double()).sum().backward()
    qkd, vd = qk.to(DEV).requires_grad_(True), v.to(DEV).requires_grad_(True)
    out = o.mha_smallq(qkd, vd, M)
    (out * cot.to(DEV)).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(qkd.grad, qk64.grad) < 2e-5
    assert rel_err(vd.grad, v64.grad) < 1e-5


# ------------------------------------------------------------------------------- heads tail
@pytest.mark.parametrize("n_slots", [1, 9, 22])
def test_heads_select_rot6d(n_slots):
    o = ops()
    g = torch.Generator().manual_seed(n_slots)
    R = 160
    rot, tr = torch.randn(R, n_slots * 6, generator=g), torch.randn(R, n_slots * 3, generator=g)
    cls = torch.randint(-1, n_slots, (R,), generator=g)
    ct, cR = torch.randn(R, 3, generator=g), torch.randn(R, 3, 3, generator=g)
    slot = cls.clamp_min(0) if n_slots > 1 else torch.zeros(R, dtype=torch.long)
    r64, t64 = rot.double().requires_grad_(True), tr.double().requires_grad_(True)
    rows = torch.arange(R)
    Rm = O.rotation_6d_to_matrix(r64.view(R, n_slots, 6)[rows, slot])
    ts = t64.view(R, n_slots, 3)[rows, slot]
    ((Rm * cR.double()).sum() + (ts * ct.double()).sum()).backward()
    rd, td = rot.to(DEV).requires_grad_(True), tr.to(DEV).requires_grad_(True)
    t, Rg, r6 = o.heads_select_rot6d(rd, td, cls.to(DEV) if n_slots > 1 else None, n_slots)
    ((Rg * cR.to(DEV)).sum() + (t * ct.to(DEV)).sum()).backward()
    assert rel_err(t, ts) == 0.0
    assert rel_err(Rg, Rm) < 2e-6
    assert rel_err(rd.grad, r64.grad) < 2e-5
    assert rel_err(td.grad, t64.grad) == 0.0
    # orthonormality of the produced rotations
    eye = (Rg.transpose(1, 2) @ Rg - torch.eye(3, device=DEV)).abs().max()
    assert float(eye) < 1e-5


# ------------------------------------------------------------------------------- position encodings / layout
def test_posenc_matches_reference_golden():
    o = ops()
    for key, rec in load_golden("posenc").items():
        if not key.startswith("pos_"):
            continue
        got = o.posenc_sine_nchw(rec["mask"].to(DEV), 128)
        # same fp32 argument; CUDA sinf/cosf vs CPU libm differ by <= 2 ulp at |arg| <= 2*pi
        assert float((got.cpu() - rec["pos"]).abs().max()) < 1e-6, key
        B, H, W = rec["mask"].shape
        le = torch.randn(256)
        tok = torch.zeros(B, H * W + 5, 256, device=DEV)
        o.posenc_sine_tokens_(tok, rec["mask"].to(DEV), le.to(DEV), 3, F=128)
        ref_tok = rec["pos"].flatten(2).transpose(1, 2) + le
        assert float((tok[:, 3:3 + H * W].cpu() - ref_tok).abs().max()) < 1e-6
        assert float(tok[:, :3].abs().max()) == 0.0 and float(tok[:, 3 + H * W:].abs().max()) == 0.0


def test_bbox_embedding_matches_reference_golden():
    from poet_b200.position_encoding import BoundingBoxEmbeddingSine
    rec = load_golden("posenc")["bbox"]
    got = BoundingBoxEmbeddingSine(32)(rec["boxes"].to(DEV)).cpu()
    # arguments reach c * 2^31: needs the accurate (Payne-Hanek) sinf/cosf slow path
    assert float((got - rec["embed"]).abs().max()) < 5e-7
    qe = ops().bbox_embed_pad(torch.cat((rec["boxes"], -torch.ones(3, 4))).view(1, -1, 4).to(DEV),
                              torch.tensor([rec["boxes"].shape[0]], dtype=torch.int32, device=DEV), 32)
    assert torch.equal(qe[0, :, :256], qe[0, :, 256:])
    assert float((qe[0, -3:] + 10).abs().max()) == 0.0


def test_flatten_levels_and_reference_points():
    o = ops()
    cfg = S.CONFIGS["tiny16"]
    inp = S.make_inputs(cfg, pad_columns=True)
    le = torch.randn(4, 256)
    srcs = [s.to(DEV).requires_grad_(True) for s in inp["srcs"]]
    led = le.to(DEV).requires_grad_(True)
    tok = o.flatten_levels(srcs, led)
    ref = torch.cat([s.flatten(2).transpose(1, 2) + le[l] for l, s in enumerate(inp["srcs"])], 1)
    assert torch.equal(tok.cpu(), ref)
    cot = torch.randn(ref.shape)
    (tok * cot.to(DEV)).sum().backward()
    off = 0
    for l, s in enumerate(inp["srcs"]):
        hw = s.shape[2] * s.shape[3]
        gref = cot[:, off:off + hw].transpose(1, 2).reshape(s.shape)
        assert torch.equal(srcs[l].grad.cpu(), gref)
        assert rel_err(led.grad[l], cot[:, off:off + hw].double().sum((0, 1))) < 1e-5
        off += hw
    vr = torch.stack([O.valid_ratio(m) for m in inp["masks"]], 1)
    shapes = S.pyramid_of(cfg)
    got = o.enc_reference_points(vr.to(DEV), shapes)
    assert float((got.cpu() - O.encoder_reference_points(shapes, vr)).abs().max()) < 1e-6


# ------------------------------------------------------------------------------- tcgen05 GEMM
@pytest.mark.parametrize("prec,tol", [("bf16x3", 3e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("M,N,K,a_k,b_k", [
    (25600, 256, 256, True, True),      # encoder projection, forward (NT)
    (3200, 1024, 256, True, True),      # FFN1 forward
    (3000, 256, 1024, True, True),      # FFN2 forward, ragged M (tile tail)
    (25600, 256, 768, True, False),     # dgrad of the fused [offsets|logits] projection (NN, MN-major B)
    (3200, 256, 1024, True, False),     # dgrad FFN1
    (1024, 256, 25600, False, False),   # wgrad FFN1 (TN: both MN-major, split-K)
    (256, 256, 3200, False, False),     # wgrad value_proj
    (768, 256, 6400, False, False),     # wgrad fused projection (N tile 256, M 6 tiles)
    (1600, 128, 200, True, True),       # BN=128 path, K tail (200 = 3*64 + 8)
])
def test_gemm_tcgen05(M, N, K, a_k, b_k, prec, tol):
    o = ops()
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = torch.randn((M, K) if a_k else (K, M), generator=g)
    Bm = torch.randn((N, K) if b_k else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = (A.double() if a_k else A.double().t()) @ (Bm.double().t() if b_k else Bm.double()) + bias.double()
    out = o.gemm(A.to(DEV), Bm.to(DEV), M, N, K, a_kcontig=a_k, b_kcontig=b_k, bias=bias.to(DEV),
                 precision=o._PRECISION[prec])
    err = rel_err(out, ref)
    assert err < tol, f"{prec} {M}x{N}x{K} a_k={a_k} b_k={b_k}: rel err {err:.3e}"


def test_gemm_tcgen05_epilogues():
    o = ops()
    g = torch.Generator().manual_seed(9)
    M, N, K = 3200, 256, 256
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    gate = torch.randn(M, N, generator=g)
    mask = (torch.rand(M, generator=g) < 0.2).to(torch.uint8)
    base = torch.randn(M, N, generator=g)
    ref = A.double() @ W.double().t() + b.double()
    P = o.GEMM_BF16X3
    assert rel_err(o.gemm(A.to(DEV), W.to(DEV), M, N, K, bias=b.to(DEV), relu=True, precision=P), ref.clamp_min(0)) < 3e-5
    assert rel_err(o.gemm(A.to(DEV), W.to(DEV), M, N, K, bias=b.to(DEV), gate=gate.to(DEV), precision=P), ref * (gate > 0)) < 3e-5
    assert rel_err(o.gemm(A.to(DEV), W.to(DEV), M, N, K, bias=b.to(DEV), row_mask=mask.to(DEV), precision=P),
                   ref * (mask == 0)[:, None]) < 3e-5
    acc = o.gemm(A.to(DEV), W.to(DEV), M, N, K, bias=b.to(DEV), out=base.to(DEV).clone(), accumulate=True, alpha=0.5, precision=P)
    assert rel_err(acc, 0.5 * (ref - b.double()) + b.double() + base.double()) < 3e-5


def test_gemm_tcgen05_relu_bitmask_and_views():
    """poet_gemm_ex: the forward GEMM's ReLU sign bitmask gates the dgrad GEMM (no fp32 re-read); TMA-store
    epilogue on a ragged M, on a column-block view of a wider matrix (ldc > N) and with beta = 1."""
    o = ops()
    g = torch.Generator().manual_seed(31)
    R, C, F = 3000, 256, 1024                                     # ragged M: 23 full tiles + 56 rows
    x, W1, b1 = torch.randn(R, C, generator=g), torch.randn(F, C, generator=g) * 0.06, torch.randn(F, generator=g) * 0.1
    gy = torch.randn(R, C, generator=g)
    W2 = torch.randn(C, F, generator=g) * 0.06
    old = o.get_gemm_precision()
    o.set_gemm_precision("bf16x3")
    try:
        bits = o.relu_bits_buffer(R, F, C, DEV)
        if bits is None:
            pytest.skip("TMA epilogue disabled (POET_GEMM_TMA_EPI=0)")
        h = o.gemm(x.to(DEV), W1.to(DEV), R, F, C, bias=b1.to(DEV), relu=True, relu_bits=bits)
        pre = x.double() @ W1.double().t() + b1.double()
        assert rel_err(h, pre.clamp_min(0)) < 3e-5
        # bit c%32 of word [r, c/32] == (h[r,c] > 0), judged on the kernel's own output
        got_bits = ((bits.view(R, F // 32, 1) >> torch.arange(32, device=DEV).view(1, 1, 32)) & 1).view(R, F).bool()
        assert torch.equal(got_bits, h > 0)
        dh = o.gemm(gy.to(DEV), W2.to(DEV), R, F, C, b_kcontig=False, gate_bits=bits)
        ref_dh = (gy.double() @ W2.double()) * (h.cpu() > 0)
        assert rel_err(dh, ref_dh) < 3e-5
        # column-block view: write N = 256 columns at offset 128 of a 640-wide matrix, beta = 1
        wide = torch.randn(R, 640, generator=g).to(DEV)
        keep = wide.clone()
        view = wide[:, 128:384]
        W3 = torch.randn(256, C, generator=g)
        o.gemm(x.to(DEV), W3.to(DEV), R, 256, C, out=view, accumulate=True)
        ref = keep.cpu().double()
        ref[:, 128:384] += x.double() @ W3.double().t()
        assert rel_err(wide, ref) < 3e-5
        assert torch.equal(wide[:, :128], keep[:, :128]) and torch.equal(wide[:, 384:], keep[:, 384:])
    finally:
        o.set_gemm_precision(old)


@pytest.mark.parametrize("Mo,No,R", [(256, 256, 25600), (1024, 256, 25600), (256, 1024, 6400), (128, 384, 3208),
                                     (512, 256, 160)])
def test_gemm_tcgen05_wgrad_accumulate(Mo, No, R):
    """Weight-gradient shape (TN, both operands fp32 activations): BK=32 pipeline, 128x256 tiles when N allows,
    split-K reduced by TMA reduce-add straight into an existing gradient (beta = 1), K tails (3208 = 100*32 + 8)."""
    o = ops()
    g = torch.Generator().manual_seed(Mo + No + R)
    dY, X = torch.randn(R, Mo, generator=g), torch.randn(R, No, generator=g)
    base = torch.randn(Mo, No, generator=g)
    ref = dY.double().t() @ X.double()
    P = o.GEMM_BF16X3
    out = o.gemm(dY.to(DEV), X.to(DEV), Mo, No, R, a_kcontig=False, b_kcontig=False, precision=P)
    assert rel_err(out, ref) < 3e-5
    acc = o.gemm(dY.to(DEV), X.to(DEV), Mo, No, R, a_kcontig=False, b_kcontig=False, out=base.to(DEV).clone(),
                 accumulate=True, precision=P)
    assert rel_err(acc, ref + base.double()) < 3e-5
    # fused bias gradient: b += colsum(dY) summed from the dY tiles inside the weight-gradient GEMM
    old = o.get_gemm_precision()
    o.set_gemm_precision("bf16x3")
    try:
        w_acc, b_acc = base.to(DEV).clone(), torch.arange(Mo, dtype=torch.float32, device=DEV)
        o.wgrad_bias(dY.to(DEV), X.to(DEV), Mo, No, R, w_acc, b_acc)
        assert rel_err(w_acc, ref + base.double()) < 3e-5
        assert rel_err(b_acc, dY.double().sum(0) + torch.arange(Mo, dtype=torch.float64)) < 1e-5
    finally:
        o.set_gemm_precision(old)
    # operands that are column blocks of wider activations (fused projection gradients): lda / ldb > extent
    wideY = torch.randn(R, Mo + 128, generator=g).to(DEV)
    out2 = o.gemm(wideY[:, 128:], X.to(DEV), Mo, No, R, a_kcontig=False, b_kcontig=False, lda=Mo + 128, precision=P)
    assert rel_err(out2, wideY[:, 128:].cpu().double().t() @ X.double()) < 3e-5


@pytest.mark.parametrize("prec,tol", [("bf16x3", 3e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("M,N,K,b_k", [
    (25600, 256, 256, True),      # forward: weight [N,K] K-major via TMA
    (3000, 1024, 256, True),      # ragged M, 4 N tiles
    (3200, 256, 1024, True),      # long K
    (25600, 256, 768, False),     # dgrad: weight stored [K,N] (MN-major B) via TMA
    (3200, 1024, 256, False),
    (1600, 128, 200, True),       # BN=128, K tail handled by TMA zero fill
])
def test_gemm_tcgen05_presplit_weights(M, N, K, b_k, prec, tol):
    """B operand = bf16 hi/lo planes produced once by poet_split_bf16 and fetched by TMA."""
    o = ops()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn((N, K) if b_k else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ (W.double().t() if b_k else W.double()) + bias.double()
    Wd = W.to(DEV)
    old = o.get_gemm_precision()
    o.set_gemm_precision(prec)
    try:
        hi = torch.empty(Wd.shape, device=DEV, dtype=torch.bfloat16)
        lo = torch.empty(Wd.shape, device=DEV, dtype=torch.bfloat16) if prec == "bf16x3" else None
        o._call("poet_split_bf16", Wd.data_ptr(), hi.data_ptr(), None if lo is None else lo.data_ptr(), Wd.numel(),
                o._stream(Wd))
        assert torch.equal(hi, Wd.to(torch.bfloat16))
        if lo is not None:
            assert torch.equal(lo, (Wd - hi.float()).to(torch.bfloat16))
        out = o.gemm(A.to(DEV), Wd, M, N, K, b_kcontig=b_k, bias=bias.to(DEV), b_split=(hi, lo))
    finally:
        o.set_gemm_precision(old)
    err = rel_err(out, ref)
    assert err < tol, f"{prec} {M}x{N}x{K} b_k={b_k}: rel err {err:.3e}"


def test_msda_backward_tile_vs_thread_kernels():
    """Encoder-size backward (B=16, S=Lq=1600: the benchmarked shape, mode 1): the tile kernel (dense low-resolution levels
    on mma.sync) against the one-thread-per-channel-group scatter kernel of the same library (POET_MSDA_TILE=0 in a child
    process), whose results the op-level fp64 checks pin."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, sys\n"
        "from poet_b200 import ops\n"
        "shapes=((30,40),(15,20),(8,10),(4,5)); B,S,M,D,L,P=16,1600,16,16,4,4\n"
        "g=torch.Generator().manual_seed(3)\n"
        "value=torch.randn(B,S,M*D,generator=g).cuda(); oa=torch.randn(B,S,M*L*P*3,generator=g).cuda()\n"
        "oa[...,:M*L*P*2]*=2.0\n"
        "ref=torch.rand(B,S,L,2,generator=g).cuda(); go=torch.randn(B,S,M*D,generator=g).cuda()\n"
        "value.requires_grad_(True); oa.requires_grad_(True)\n"
        "out=ops.msda_block(value,oa,ref,shapes,M,L,P); (out*go).sum().backward(); torch.cuda.synchronize()\n"
        "torch.save({'gv':value.grad.cpu(),'goa':oa.grad.cpu()}, sys.argv[1])\n")
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for knob in ("1", "0"):
        path = f"/tmp/poet_tile_{knob}.pt"
        r = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, POET_MSDA_TILE=knob), cwd=here,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res[knob] = torch.load(path)
    n_off = 16 * 16 * 2
    assert rel_err(res["1"]["gv"], res["0"]["gv"]) < 5e-6
    assert rel_err(res["1"]["goa"][..., n_off:], res["0"]["goa"][..., n_off:]) < 5e-5
    assert bad_fraction(res["1"]["goa"][..., :n_off], res["0"]["goa"][..., :n_off], 5e-5) < 1e-3


def test_msda_dense_lowres_backward_path_in_subprocess():
    """The dense tensor-core grad_value path of the MSDA backward (POET_MSDA_DENSE=1, read once per process) against the
    same fp64 oracle checks: run the MSDA tests of this file in a child process with the knob set."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, POET_MSDA_DENSE="1", POET_MSDA_TILE="0")
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-m", "gpu", "-q", "-x", "-k",
                        "test_msda_core_fwd_bwd or test_msda_block_matches_module_math"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(here)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


# ------------------------------------------------------------------------------- small-row GEMM (csrc/gemm_small.cu)
@pytest.mark.parametrize("M,N,K", [(160, 256, 256), (160, 1024, 256), (160, 256, 1024), (160, 768, 256), (400, 132, 256),
                                   (36, 264, 72), (160, 66, 256),
                                   (160, 256, 768), (160, 256, 640), (96, 128, 1152)])   # cluster split-K: 3 / 2 (uneven) / 4 CTAs (one idle)
@pytest.mark.parametrize("a_k,b_k", [(True, True), (True, False), (False, False), (False, True)])
def test_gemm_small_rows_exact_fp32(M, N, K, a_k, b_k):
    """Decoder / pose-head shapes in the default (bf16x3) precision mode are served by the latency kernel in exact fp32:
    all four operand layouts, ragged N, bias + ReLU gate, beta = 1 accumulation."""
    o = ops()
    if (not a_k and M % 4) or (not b_k and N % 4):
        pytest.skip("row-contiguous operands need 16-byte chunks (falls to the SIMT kernel)")
    g = torch.Generator().manual_seed(M * 5 + N * 3 + K + 2 * a_k + b_k)
    A = torch.randn((M, K) if a_k else (K, M), generator=g)
    Bm = torch.randn((N, K) if b_k else (K, N), generator=g)
    bias, gate, base = torch.randn(N, generator=g), torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    ref = (A.double() if a_k else A.double().t()) @ (Bm.double().t() if b_k else Bm.double()) + bias.double()
    tol = 2e-6 * math.sqrt(K)
    out = o.gemm(A.to(DEV), Bm.to(DEV), M, N, K, a_kcontig=a_k, b_kcontig=b_k, bias=bias.to(DEV))
    assert rel_err(out, ref) < tol
    out = o.gemm(A.to(DEV), Bm.to(DEV), M, N, K, a_kcontig=a_k, b_kcontig=b_k, bias=bias.to(DEV), gate=gate.to(DEV),
                 alpha=0.5)
    assert rel_err(out, (ref - 0.5 * (ref - bias.double())) * (gate > 0)) < tol
    out = o.gemm(A.to(DEV), Bm.to(DEV), M, N, K, a_kcontig=a_k, b_kcontig=b_k, relu=True, accumulate=True,
                 out=base.to(DEV).clone())
    assert rel_err(out, base.double() + (ref - bias.double()).clamp_min(0)) < tol


@pytest.mark.parametrize("M,N,K,b_k", [(160, 256, 256, True), (160, 768, 256, True), (160, 256, 768, False),
                                       (160, 256, 1024, False), (250, 1024, 256, True)])
def test_gemm_small_rows_weight_planes(M, N, K, b_k):
    """Weights as the step's bf16 hi/lo planes (hi + lo = the 2^-17 operand of the tensor-core path)."""
    o = ops()
    g = torch.Generator().manual_seed(M + N + K)
    A, W, bias = torch.randn(M, K, generator=g), torch.randn((N, K) if b_k else (K, N), generator=g), torch.randn(N, generator=g)
    Wd = W.to(DEV)
    hi = torch.empty(Wd.shape, device=DEV, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    o._call("poet_split_bf16", Wd.data_ptr(), hi.data_ptr(), lo.data_ptr(), Wd.numel(), o._stream(Wd))
    ref = A.double() @ (W.double().t() if b_k else W.double()) + bias.double()
    out = o.gemm(A.to(DEV), Wd, M, N, K, b_kcontig=b_k, bias=bias.to(DEV), relu=True, b_split=(hi, lo))
    assert rel_err(out, ref.clamp_min(0)) < 3e-5


@pytest.mark.parametrize("No,Ko,R", [(256, 256, 160), (512, 256, 160), (1024, 256, 400), (132, 256, 160)])
def test_gemm_small_rows_wgrad_colsum(No, Ko, R):
    """Weight gradient of a query-row layer: dW += dY^T X with the bias gradient summed from the dY tiles."""
    o = ops()
    g = torch.Generator().manual_seed(No + Ko + R)
    dY, X = torch.randn(R, No, generator=g), torch.randn(R, Ko, generator=g)
    w0, b0 = torch.randn(No, Ko, generator=g), torch.randn(No, generator=g)
    w, b = w0.to(DEV).clone(), b0.to(DEV).clone()
    o.wgrad_bias(dY.to(DEV), X.to(DEV), No, Ko, R, w, b)
    assert rel_err(w, w0.double() + dY.double().t() @ X.double()) < 2e-6 * math.sqrt(R)
    assert rel_err(b, b0.double() + dY.double().sum(0)) < 1e-5
