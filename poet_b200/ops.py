"""torch.autograd.Function bindings over the C ABI (raw wrappers, GEMM / Linear / MLP, LayerNorm, MSDA, attention, heads,
position encodings, input_proj).  Runtime state lives in _runtime.py, the weight-plane arena in _planes.py; both are
re-exported here, so `poet_b200.ops` stays the one import for callers, tests and tools.

PyTorch is plumbing here: it owns device memory, the CUDA stream and the autograd tape; every
numeric operation below is a call into libpoet_b200.so.  Nothing in this file has a CPU or
eager-PyTorch fallback: a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os as _os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._runtime import *  # noqa: F401,F403
from ._runtime import (_ABLATE, _B_STABLE, _PRECISION, _GradMarker, _call, _chk, _drop_args, _drop_counters, _marker_cb, _p,  # noqa: F401
                       _pending_joins, _side_streams, _state, _stream, _stream_ns, _timing, _touched_side_streams)
from ._planes import *  # noqa: F401,F403
from ._planes import _active_planes, _eligible  # noqa: F401

# ------------------------------------------------------------------------------------------
# raw ops (no autograd)
# ------------------------------------------------------------------------------------------
def gemm(A: torch.Tensor, Bm: torch.Tensor, M: int, N: int, K: int, *, a_kcontig: bool = True, b_kcontig: bool = True,
         lda: Optional[int] = None, ldb: Optional[int] = None, bias=None, relu=False, gate=None, row_mask=None,
         out: Optional[torch.Tensor] = None, accumulate=False, alpha: float = 1.0,
         precision: Optional[int] = None, b_split=None, relu_bits: Optional[torch.Tensor] = None,
         gate_bits: Optional[torch.Tensor] = None, a_colsum: Optional[torch.Tensor] = None,
         a_row_mask: Optional[torch.Tensor] = None, drop=None, b_stable: bool = False) -> torch.Tensor:
    """out[M,N] = epi(alpha * op(A) @ op(B)); see include/poet_b200.h poet_gemm / poet_gemm_ex.
    relu_bits (out) / gate_bits (in): int32 [M, N/32] sign bitmask of a ReLU (tensor-core path only)."""
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    lda = (K if a_kcontig else M) if lda is None else lda
    ldb = (K if b_kcontig else N) if ldb is None else ldb
    prec = _state["precision"] if precision is None else precision
    ws, ws_bytes = None, 0
    if prec != GEMM_FP32:
        ws_bytes = _lib.lib().poet_gemm_workspace_bytes(M, N, K, int(a_kcontig), int(b_kcontig), prec)
        if ws_bytes:
            ws = torch.empty(ws_bytes, device=A.device, dtype=torch.uint8)
    # b_stable: B is a parameter (or its planes), not written by the kernels just before this call (POET_GEMM_B_STABLE)
    flags = ((1 if relu else 0) | (2 if accumulate else 0) | (4 if (b_stable and _B_STABLE) else 0) |
             (8 if _state.get("gemm_background") else 0))
    tag = (f"{M}x{N}x{K}" + ("" if a_kcontig else ",At") + ("" if b_kcontig else ",Bt")) if _timing["on"] else None
    work = (4 * (M * K + N * K + M * N), 2 * M * N * K)
    if relu_bits is not None or gate_bits is not None or a_colsum is not None or a_row_mask is not None or drop is not None:
        assert gate is None and prec != GEMM_FP32
        bs = b_split if (b_split is not None and a_kcontig) else (None, None)
        d_seed, d_site, d_p = drop if drop is not None else (None, 0, 0.0)        # (seed tensor, site, p): epilogue dropout
        _call("poet_gemm_ex", _p(A), lda, int(a_kcontig), _p(Bm), _p(bs[0]), _p(bs[1]), ldb, int(b_kcontig), _p(out),
              out.stride(0), M, N, K, alpha, _p(bias), _p(row_mask), _p(relu_bits), _p(gate_bits), _p(a_colsum),
              _p(a_row_mask), flags, prec, _p(d_seed), int(d_site), float(d_p), _stream(A), tag=tag, work=work)
        return out
    if b_split is not None and prec != GEMM_FP32 and a_kcontig:
        _call("poet_gemm_bsplit", _p(A), lda, int(a_kcontig), _p(Bm), _p(b_split[0]), _p(b_split[1]), ldb,
              int(b_kcontig), _p(out), out.stride(0), M, N, K, alpha, _p(bias), _p(gate), _p(row_mask), flags, prec,
              _stream(A), tag=tag, work=work)
        return out
    _call("poet_gemm", _p(A), lda, int(a_kcontig), _p(Bm), ldb, int(b_kcontig), _p(out), out.stride(0), M, N, K,
          alpha, _p(bias), _p(gate), _p(row_mask), flags, prec, _p(ws), ws_bytes, _stream(A), tag=tag, work=work)
    return out


def relu_bits_buffer(R: int, N: int, K: int, device) -> Optional[torch.Tensor]:
    """int32 [R, N/32] buffer for a ReLU sign bitmask if the GEMM [R,N,K] supports it (large tensor-core shapes:
    the dgrad then reads 1 bit instead of one fp32 activation per element), else None."""
    prec = _state["precision"]
    if prec == GEMM_FP32 or R < 1024 or N % 32:
        return None
    if not _lib.lib().poet_gemm_relu_bits_supported(R, N, K, prec):
        return None
    return torch.empty((R, N // 32), device=device, dtype=torch.int32)


def wgrad_bias(gy2: torch.Tensor, x2: torch.Tensor, N: int, K: int, R: int, w_out: torch.Tensor, b_out: Optional[torch.Tensor],
               lda: Optional[int] = None, row_mask: Optional[torch.Tensor] = None) -> None:
    """w_out[N,K] += gy2[R,N]^T x2[R,K] and, if given, b_out[N] += colsum(gy2): one launch when the weight-gradient GEMM
    runs on the tensor-core path (the bias gradient is summed from the dY tiles as they stream through the
    producers), else the GEMM plus a separate column-sum kernel.  gy2 may be a column block (lda = row stride)."""
    lda = N if lda is None else lda
    prec = _state["precision"]
    # Long reductions are split-K over ~148 CTAs whose per-warp partial sums would meet on the same N addresses
    # (measured: +54 us of serialized L2 atomics on [256 x 256 x 25600]); those keep the streaming colsum kernel.
    fused = (b_out is not None and prec != GEMM_FP32 and R <= 2048 and
             _lib.lib().poet_gemm_tc_eligible(N, K, R, lda, K, w_out.stride(0)))
    gemm(gy2, x2, N, K, R, a_kcontig=False, b_kcontig=False, lda=lda, out=w_out, accumulate=True,
         a_colsum=b_out if fused else None, a_row_mask=row_mask)
    if b_out is not None and not fused:
        _call("poet_colsum_masked", _p(gy2), lda, _p(row_mask), _p(b_out), R, N, 1, _stream(gy2), work=(4 * R * N, R * N))


def colsum(X: torch.Tensor, M: int, N: int, out: Optional[torch.Tensor] = None, accumulate=False) -> torch.Tensor:
    if out is None:
        out = torch.empty(N, device=X.device, dtype=torch.float32)
    _call("poet_colsum", _p(X), N, _p(out), M, N, int(accumulate), _stream(X))
    return out


def mask_rows_(x2d: torch.Tensor, mask_u8: torch.Tensor) -> torch.Tensor:
    _call("poet_mask_rows", _p(x2d), _p(mask_u8), x2d.shape[0], x2d.shape[1], _stream(x2d))
    return x2d


def add(a: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    out = torch.empty_like(a)
    _call("poet_add", _p(a), _p(b), _p(out), a.numel(), _stream(a))
    return out


# ------------------------------------------------------------------------------------------
# autograd: Linear / FFN / MLP
# ------------------------------------------------------------------------------------------
_direct_slots = {}          # data_ptr of a registered gradient slot -> numel


def register_direct_grad_slots(grads) -> None:
    """Opt in: backward kernels may accumulate straight into these tensors (FlatGradReducer registers the views of
    its arena).  A .grad that was not registered is left to autograd's AccumulateGrad, so tensor hooks,
    post-accumulate hooks and DistributedDataParallel's reducer hooks keep firing for it."""
    for g in grads:
        _direct_slots[g.data_ptr()] = g.numel()


def unregister_direct_grad_slots(grads=None) -> None:
    if grads is None:
        _direct_slots.clear()
        return
    for g in grads:
        _direct_slots.pop(g.data_ptr(), None)


def _grad_slot(param) -> Optional[torch.Tensor]:
    """The .grad of a leaf parameter if it is a registered arena slot (FlatGradReducer), else None.
    Backward kernels then accumulate straight into it (beta = 1 epilogue) and hand autograd `None`,
    which removes one zero-fill and one add kernel per parameter per step.  Direct accumulation bypasses
    AccumulateGrad (no hooks, torch.autograd.grad() sees None): it is therefore opt-in per tensor, and
    incompatible with DistributedDataParallel -- FlatGradReducer is the data-parallel path that goes with it."""
    if not _state["direct_grads"] or not _direct_slots or param is None or not isinstance(param, torch.Tensor):
        return None
    if not (param.is_leaf and param.requires_grad):
        return None
    g = param.grad
    if g is None or _direct_slots.get(g.data_ptr()) != g.numel():
        return None
    if not g.is_contiguous() or g.dtype != torch.float32 or g.shape != param.shape:
        return None
    return g


def set_direct_param_grads(on: bool) -> None:
    _state["direct_grads"] = bool(on)


def _linear_bwd(gy2: torch.Tensor, x2: torch.Tensor, W: torch.Tensor, need_x: bool, need_w: bool, need_b: bool,
                gate: Optional[torch.Tensor] = None, w_split=None, w_param=None, b_param=None,
                gate_bits: Optional[torch.Tensor] = None, gy_row_mask: Optional[torch.Tensor] = None, alpha: float = 1.0):
    """gy2 [R,N], x2 [R,K], W [N,K] -> (dx [R,K] (gated by `gate`>0 if given), dW [N,K], db [N]).
    dW / db come back as None when they were accumulated directly into the parameters' .grad."""
    R, N = gy2.shape
    K = x2.shape[1]
    if gate_bits is not None:
        gate = None
    # gy_row_mask: rows of gy2 to treat as zero (value masked_fill backward).  Zero rows of dY give zero rows of dX,
    # so the dgrad applies it as an output row mask; the weight / bias gradients read dY through the masked producer.
    dx = gemm(gy2, W, R, K, N, a_kcontig=True, b_kcontig=False, gate=gate, b_split=w_split, b_stable=True,
              gate_bits=gate_bits, row_mask=gy_row_mask, alpha=alpha) if need_x else None
    dW = db = None
    w_slot = _grad_slot(w_param) if need_w else None
    b_slot = _grad_slot(b_param) if need_b else None
    # Parameter gradients are leaves of the backward graph: when they go straight into .grad slots nobody
    # downstream waits for them, so they are issued on a side stream and only the dgrad stays on the
    # critical path (joined at the end of the backward pass).
    side = None
    if (w_slot is not None or b_slot is not None) and parallel_streams_enabled():
        side = fork(_WGRAD_STREAM, gy2.device)
        side.uses(gy2, x2)
        _join_at_end_of_backward(side)
        side.__enter__()
    try:
        if w_slot is not None:
            wgrad_bias(gy2, x2, N, K, R, w_slot, b_slot, row_mask=gy_row_mask)
        elif b_slot is not None:
            _call("poet_colsum_masked", _p(gy2), N, _p(gy_row_mask), _p(b_slot), R, N, 1, _stream(gy2), work=(4 * R * N, R * N))
    finally:
        if side is not None:
            side.__exit__(None, None, None)
    if need_w and w_slot is None:
        dW = gemm(gy2, x2, N, K, R, a_kcontig=False, b_kcontig=False, a_row_mask=gy_row_mask)
    if need_b and b_slot is None:
        db = torch.zeros(N, device=gy2.device, dtype=torch.float32)
        _call("poet_colsum_masked", _p(gy2), N, _p(gy_row_mask), _p(db), R, N, 1, _stream(gy2), work=(4 * R * N, R * N))
    return dx, dW, db


_WGRAD_STREAM = 8


def _join_at_end_of_backward(f: "fork") -> None:
    """Make the stream that called backward() wait for side stream `f.side` once the backward pass is over."""
    key = id(f.side)
    if key in _pending_joins:
        return
    _pending_joins[key] = f.side

    def _cb():
        side = _pending_joins.pop(key, None)
        if side is not None:
            torch.cuda.current_stream(side.device).wait_stream(side)

    torch.autograd.Variable._execution_engine.queue_callback(_cb)


class _JoinAfterBackward(torch.autograd.Function):
    """Identity whose backward (the first node of the slice's backward to run) asks the autograd engine to make the
    stream that called backward() wait for `stream` when the pass is over: a micro-batch's backward runs on that
    micro-batch's stream, and its tail (parameter gradients written straight into .grad) has no consumer the engine
    would otherwise synchronise with."""

    @staticmethod
    def forward(ctx, x, stream):
        ctx.stream = stream
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        side = ctx.stream
        key = id(side)
        if key not in _pending_joins:
            _pending_joins[key] = side

            def _cb():
                st = _pending_joins.pop(key, None)
                if st is not None:
                    torch.cuda.current_stream(st.device).wait_stream(st)

            torch.autograd.Variable._execution_engine.queue_callback(_cb)
        return g, None


def join_after_backward(x: torch.Tensor, stream) -> torch.Tensor:
    return _JoinAfterBackward.apply(x, stream) if x.requires_grad else x


class _Linear(torch.autograd.Function):
    """y = x W^T + b, rows with row_mask zeroed (MSDeformAttn value_proj + masked_fill)."""

    @staticmethod
    def forward(ctx, x, W, b, row_mask, mask_grad_inplace, bias_grad_elsewhere=False):
        ctx.prec = _state["precision"]
        ctx.background = bool(_state.get("gemm_background"))
        ctx.mask_grad_inplace = mask_grad_inplace
        ctx.bias_grad_elsewhere = bias_grad_elsewhere
        x2 = _chk(x).view(-1, x.shape[-1])
        W = _chk(W)
        R, K = x2.shape
        N = W.shape[0]
        ctx.w_param, ctx.b_param = W, b
        ctx.w_split = split_weight(W, R)
        y = gemm(x2, W, R, N, K, bias=b, row_mask=row_mask, b_split=ctx.w_split, b_stable=True)
        ctx.save_for_backward(x2, W)
        ctx.row_mask = row_mask
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, *grads):
        with precision_scope(ctx.prec), background_gemms(ctx.background):
            return _Linear._backward_impl(ctx, *grads)

    @staticmethod
    def _backward_impl(ctx, gy):
        x2, W = ctx.saved_tensors
        gy2 = _chk(gy).view(-1, gy.shape[-1])
        gy_mask = None
        if ctx.row_mask is not None:
            R, N, K = gy2.shape[0], gy2.shape[1], x2.shape[1]
            if (_state["precision"] != GEMM_FP32 and _os.environ.get("POET_FUSE_ROW_MASK", "1") != "0" and
                    _lib.lib().poet_gemm_tc_eligible(N, K, R, N, K, K)):
                gy_mask = ctx.row_mask            # applied inside the dgrad epilogue / wgrad producers / masked colsum
            else:
                gy2 = mask_rows_(gy2 if ctx.mask_grad_inplace else gy2.clone(), ctx.row_mask)
        dx, dW, db = _linear_bwd(gy2, x2, W, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                 ctx.has_bias and ctx.needs_input_grad[2] and not ctx.bias_grad_elsewhere, w_split=ctx.w_split,
                                 w_param=ctx.w_param, b_param=ctx.b_param, gy_row_mask=gy_mask)
        return (dx.view(ctx.xshape) if dx is not None else None), dW, db, None, None, None


def linear(x, W, b=None, row_mask=None, mask_grad_inplace=False, bias_grad_elsewhere=False):
    """bias_grad_elsewhere: b's gradient is produced by the add_layernorm(..., r_bias=b) that consumes the result."""
    if _os.environ.get("POET_LN_RBIAS", "1") == "0":
        bias_grad_elsewhere = False
    return _Linear.apply(x, W, b, row_mask, mask_grad_inplace, bool(bias_grad_elsewhere))


class _MLP(torch.autograd.Function):
    """Chain of Linear layers with ReLU between them (FFN blocks: deformable_transformer.py:193-197,
    267-271; pose heads MLP: pose_estimation_transformer.py:677-689).  The ReLU lives in the
    producing GEMM's epilogue; its backward is the `gate` epilogue of the dgrad GEMM."""

    @staticmethod
    def forward(ctx, x, drop_p, drop_site, last_bias_elsewhere, *wb):
        ctx.prec = _state["precision"]
        ctx.last_bias_elsewhere = bool(last_bias_elsewhere)
        n = len(wb) // 2
        x2 = _chk(x).view(-1, x.shape[-1])
        acts = [x2]
        ctx.params = wb
        ctx.w_splits = []
        ctx.relu_bits = []                       # sign bitmask of layer i's ReLU output (None: gate on the fp32 activation)
        # nn.Dropout after every ReLU (reference FFN: dropout2 / dropout3 on relu(linear1(x))).  On the tensor-core
        # epilogue path the mask is folded into the ReLU bitmask; otherwise the activation is dropped in place and the
        # backward gates on it.  Either way the dgrad of the next layer is scaled by 1/(1-p) through its alpha.
        seed, drop_p = _drop_args(drop_p)
        ctx.drop_alpha = dropout_scale(drop_p, pair_scheme=True)
        for i in range(n):
            W, b = _chk(wb[2 * i]), wb[2 * i + 1]
            h = acts[-1]
            ctx.w_splits.append(split_weight(W, h.shape[0]))
            bits = relu_bits_buffer(h.shape[0], W.shape[0], W.shape[1], h.device) if i < n - 1 else None
            ctx.relu_bits.append(bits)
            dropping = drop_p > 0.0 and i < n - 1
            fused = dropping and bits is not None and W.shape[0] % 32 == 0
            acts.append(gemm(h, W, h.shape[0], W.shape[0], W.shape[1], bias=b, relu=(i < n - 1), b_stable=True,
                             b_split=ctx.w_splits[-1], relu_bits=bits,
                             drop=(seed, drop_site + i, drop_p) if fused else None))
            if dropping and not fused:
                if acts[-1].numel() % 4:
                    raise NotImplementedError("dropout needs hidden activations with a multiple of 4 elements")
                _call("poet_dropout", _p(acts[-1]), acts[-1].numel(), _p(seed), int(drop_site + i), float(drop_p), _stream(h))
                if bits is not None:             # the sign bitmask does not know the mask: gate on the dropped activation
                    ctx.relu_bits[-1] = None
        ctx.save_for_backward(*acts[:-1], *[wb[2 * i] for i in range(n)])
        ctx.n = n
        ctx.xshape = x.shape
        return acts[-1].view(*x.shape[:-1], acts[-1].shape[1])

    @staticmethod
    def backward(ctx, *grads):
        with precision_scope(ctx.prec):
            return _MLP._backward_impl(ctx, *grads)

    @staticmethod
    def _backward_impl(ctx, gy):
        n = ctx.n
        acts, Ws = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        g = _chk(gy).view(-1, gy.shape[-1])
        grads = [None] * (2 * n)
        for i in range(n - 1, -1, -1):
            need_x = i > 0 or ctx.needs_input_grad[0]
            need_b = ctx.needs_input_grad[5 + 2 * i] and not (ctx.last_bias_elsewhere and i == n - 1)
            dx, dW, db = _linear_bwd(g, acts[i], Ws[i], need_x, ctx.needs_input_grad[4 + 2 * i],
                                     need_b, gate=acts[i] if i > 0 else None,
                                     w_split=ctx.w_splits[i], w_param=ctx.params[2 * i], b_param=ctx.params[2 * i + 1],
                                     gate_bits=ctx.relu_bits[i - 1] if i > 0 else None,
                                     alpha=ctx.drop_alpha if i > 0 else 1.0)
            grads[2 * i], grads[2 * i + 1] = dW, db
            g = dx
        return (g.view(ctx.xshape) if g is not None else None, None, None, None, *grads)


def mlp(x, layers: Sequence[Tuple[torch.Tensor, torch.Tensor]], drop_p: float = 0.0, drop_site: int = 0,
        last_bias_grad_elsewhere: bool = False):
    """Linear/ReLU chain; drop_p > 0: nn.Dropout after every ReLU (sites drop_site, drop_site + 1, ...).
    last_bias_grad_elsewhere: the last layer's bias gradient comes from add_layernorm(..., r_bias=) on the result."""
    flat = []
    for W, b in layers:
        flat += [W, b]
    if _os.environ.get("POET_LN_RBIAS", "1") == "0":
        last_bias_grad_elsewhere = False
    return _MLP.apply(x, float(drop_p), int(drop_site), bool(last_bias_grad_elsewhere), *flat)


class _ProjPair(torch.autograd.Function):
    """Two projections of (possibly different) inputs against row-blocks of parameters, without any autograd
    slicing / concatenation of the parameters:

      mode 'cat'   (MSDeformAttn, one input): out = x @ [W0; W1]^T + [b0; b1]  -> [R, N0+N1]
                   (sampling_offsets | attention_weights: one fused GEMM, the gather reads one row)
      mode 'split' (decoder self-attention in_proj, W = in_proj_weight [3C, C]):
                   out0 = x0 @ W[:2C]^T + b[:2C]  (q | k of tgt+pos),  out1 = x1 @ W[2C:]^T + b[2C:]  (v of tgt)

    Weight gradients are produced by GEMMs on row / column blocks addressed by pointer offset and leading
    dimension, accumulated straight into the parameters' .grad slots when those exist."""

    @staticmethod
    def forward(ctx, mode, x0, x1, W0, b0, W1, b1):
        ctx.prec = _state["precision"]
        ctx.mode = mode
        x0_2 = _chk(x0).view(-1, x0.shape[-1])
        R, K = x0_2.shape
        if mode == "cat":
            N0, N1 = W0.shape[0], W1.shape[0]
            with torch.no_grad():
                bc = None
                for pl in reversed(_active_planes):
                    bc = pl.bias_pair(b0, b1)
                    if bc is not None:
                        break
                if bc is None:
                    bc = torch.cat((b0, b1), 0)
                ctx.w_split = split_weight_pair(W0, W1, R)
                if ctx.w_split is not None:
                    Wc = None                       # with planes the GEMMs never read the fp32 matrix
                else:
                    Wc = torch.cat((W0, W1), 0)
                    ctx.w_split = split_weight(Wc, R)
            ctx.w_planes = Wc is None                # B = the step's weight planes (a torch.cat copy is written just before)
            out = gemm(x0_2, Wc, R, N0 + N1, K, bias=bc, b_split=ctx.w_split, b_stable=ctx.w_planes)
            ctx.save_for_backward(x0_2, Wc)
            ctx.params = (W0, b0, W1, b1)
            ctx.dims = (R, K, N0, N1)
            ctx.xshape = x0.shape
            return out.view(*x0.shape[:-1], N0 + N1)
        # 'split': W0 is the full in_proj_weight [3C, C], b0 the full bias; W1/b1 unused
        x1_2 = _chk(x1).view(-1, x1.shape[-1])
        C = K
        W = _chk(W0)
        Wqk, Wv = W[: 2 * C], W[2 * C:]
        ctx.w_split = (split_weight(Wqk, R), split_weight(Wv, R))
        qk = gemm(x0_2, Wqk, R, 2 * C, C, bias=b0[: 2 * C], b_split=ctx.w_split[0], b_stable=True)
        v = gemm(x1_2, Wv, R, C, C, bias=b0[2 * C:], b_split=ctx.w_split[1], b_stable=True)
        ctx.save_for_backward(x0_2, x1_2, W)
        ctx.params = (W0, b0)
        ctx.dims = (R, C)
        ctx.xshape = x0.shape
        return qk.view(*x0.shape[:-1], 2 * C), v.view(*x0.shape[:-1], C)

    @staticmethod
    def backward(ctx, *grads):
        with precision_scope(ctx.prec):
            return _ProjPair._backward_impl(ctx, *grads)

    @staticmethod
    def _backward_impl(ctx, g0, g1=None):
        if ctx.mode == "cat":
            x2, Wc = ctx.saved_tensors
            W0, b0, W1, b1 = ctx.params
            R, K, N0, N1 = ctx.dims
            g = _chk(g0).view(R, N0 + N1)
            dx = gemm(g, Wc, R, K, N0 + N1, a_kcontig=True, b_kcontig=False, b_split=ctx.w_split, b_stable=ctx.w_planes) if ctx.needs_input_grad[1] else None
            outs = []
            for W, b, col, n in ((W0, b0, 0, N0), (W1, b1, N0, N1)):
                gcols = g[:, col:col + n]                           # column block: pointer offset + ld = N0+N1
                slot = _grad_slot(W)
                bslot = _grad_slot(b)
                if slot is not None and bslot is not None:
                    wgrad_bias(gcols, x2, n, K, R, slot, bslot, lda=N0 + N1)
                    outs += [None, None]
                    continue
                dW = gemm(gcols, x2, n, K, R, a_kcontig=False, b_kcontig=False, lda=N0 + N1, out=slot, accumulate=slot is not None)
                if bslot is not None:
                    _call("poet_colsum", _p(gcols), N0 + N1, _p(bslot), R, n, 1, _stream(g))
                    db = None
                else:
                    db = torch.empty(n, device=g.device, dtype=torch.float32)
                    _call("poet_colsum", _p(gcols), N0 + N1, _p(db), R, n, 0, _stream(g))
                outs += [None if slot is not None else dW, db]
            return None, (dx.view(ctx.xshape) if dx is not None else None), None, outs[0], outs[1], outs[2], outs[3]
        x0_2, x1_2, W = ctx.saved_tensors
        Wp, bp = ctx.params
        R, C = ctx.dims
        gqk, gv = _chk(g0).view(R, 2 * C), _chk(g1).view(R, C)
        dx0 = gemm(gqk, W[: 2 * C], R, C, 2 * C, a_kcontig=True, b_kcontig=False, b_split=ctx.w_split[0], b_stable=True) if ctx.needs_input_grad[1] else None
        dx1 = gemm(gv, W[2 * C:], R, C, C, a_kcontig=True, b_kcontig=False, b_split=ctx.w_split[1], b_stable=True) if ctx.needs_input_grad[2] else None
        slot, bslot = _grad_slot(Wp), _grad_slot(bp)
        dW = slot if slot is not None else torch.zeros_like(W)
        db = bslot if bslot is not None else torch.zeros(3 * C, device=W.device, dtype=torch.float32)
        wgrad_bias(gqk, x0_2, 2 * C, C, R, dW[: 2 * C], db[: 2 * C])
        wgrad_bias(gv, x1_2, C, C, R, dW[2 * C:], db[2 * C:])
        return (None, dx0.view(ctx.xshape) if dx0 is not None else None, dx1.view(ctx.xshape) if dx1 is not None else None,
                None if slot is not None else dW, None if bslot is not None else db, None, None)


def proj_cat(x, W0, b0, W1, b1):
    """x @ [W0; W1]^T + [b0; b1] (MSDeformAttn sampling_offsets | attention_weights)."""
    return _ProjPair.apply("cat", x, None, W0, b0, W1, b1)


def in_proj_qk_v(qk_in, v_in, in_proj_weight, in_proj_bias):
    """nn.MultiheadAttention in_proj on (tgt+pos, tgt+pos, tgt): returns (qk [.., 2C], v [.., C])."""
    return _ProjPair.apply("split", qk_in, v_in, in_proj_weight, in_proj_bias, None, None)


# ------------------------------------------------------------------------------------------
# autograd: residual + LayerNorm
# ------------------------------------------------------------------------------------------
class _AddLayerNorm(torch.autograd.Function):
    """y = LN(x + dropout(r)); optionally also y2 = y + pos (the next layer's query), one pass over HBM."""

    @staticmethod
    def forward(ctx, x, r, gamma, beta, pos, eps, drop_p=0.0, drop_site=0, r_bias=None, n_alias=0):
        """n_alias > 0: besides y [, y2] the op returns n_alias further autograd handles of y (same storage).  A tensor read
        by several ops (FFN input + next residual; value input + residual + heads) is handed out as one handle per reader,
        so that the readers' gradients reach THIS op's backward kernel as separate pointers and are summed there in
        registers, instead of by accumulation kernels inserted between the kernels of the decoder's dependent chain."""
        ctx.n_alias = int(n_alias)
        ctx.r_bias = r_bias if (r is not None and r_bias is not None) else None
        if _os.environ.get("POET_LN_RBIAS", "1") == "0":
            ctx.r_bias = None
        x2 = _chk(x).view(-1, x.shape[-1])
        r2 = None if r is None else _chk(r).view(-1, x.shape[-1])
        R, Cc = x2.shape
        y = torch.empty_like(x2)
        need_grad = any(ctx.needs_input_grad[:4])
        xhat = torch.empty_like(x2) if need_grad else None
        rstd = torch.empty(R, device=x.device, dtype=torch.float32) if need_grad else None
        p2 = None if pos is None else _chk(pos).view(-1, Cc)
        y2 = torch.empty_like(x2) if pos is not None else None
        n_streams = 2 + (r is not None) + 2 * (pos is not None) + (1 if need_grad else 0)     # x, y [, r] [, pos, y2] [, xhat]
        seed, drop_p = _drop_args(drop_p if r is not None else 0.0)
        ctx.drop = (seed, int(drop_site), drop_p)
        _call("poet_add_layernorm_fwd", _p(x2), _p(r2), _p(gamma), _p(beta), _p(p2), _p(y), _p(y2), _p(xhat), _p(rstd),
              R, Cc, eps, _p(seed), int(drop_site), drop_p, _stream(x), work=(4 * R * Cc * n_streams + 4 * R, 8 * R * Cc))
        ctx.gb_params = (gamma, beta)
        if need_grad:
            ctx.save_for_backward(xhat, rstd, gamma)
        ctx.has_r, ctx.has_pos, ctx.shape = r is not None, pos is not None, x.shape
        aliases = tuple(y.view(x.shape) for _ in range(ctx.n_alias))
        if pos is not None:
            return (y.view(x.shape), y2.view(x.shape)) + aliases
        return ((y.view(x.shape),) + aliases) if aliases else y.view(x.shape)

    @staticmethod
    def backward(ctx, gy, *more):
        xhat, rstd, gamma = ctx.saved_tensors
        R, Cc = xhat.shape
        gy2_out = more[0] if (ctx.has_pos and more) else None            # gradient of y2 = y + pos (also pos's own gradient)
        gs = [g for g in (gy,) + tuple(more) if g is not None]
        gs = [_chk(g).view(R, Cc) for g in gs]
        if not gs:
            raise RuntimeError("add_layernorm backward without any incoming gradient")
        while len(gs) > 4:                                                # the kernel takes four gradient pointers
            gs = gs[:3] + [add(gs[3], gs[4])] + gs[5:]
        gy, gy2, gy3, gy4 = (gs + [None] * 4)[:4]
        dz = torch.empty_like(xhat)
        g_slot, b_slot = _grad_slot(ctx.gb_params[0]), _grad_slot(ctx.gb_params[1])
        if g_slot is not None and b_slot is not None:            # accumulate straight into gamma.grad / beta.grad
            dgb = (None, None)
            dg_ptr, db_ptr = _p(g_slot), _p(b_slot)
        else:
            dgb = torch.zeros(2, Cc, device=xhat.device, dtype=torch.float32)
            dg_ptr, db_ptr = _p(dgb[0]), _p(dgb[1])
        seed, site, drop_p = ctx.drop
        dr = torch.empty_like(xhat) if (drop_p > 0.0 and ctx.has_r) else None     # gradient of the dropped branch
        # bias gradient of the Linear that produced r: the column sums of r's gradient, taken while it is in registers
        d_rb, rb_ptr = None, None
        if ctx.r_bias is not None and ctx.needs_input_grad[8]:
            rb_slot = _grad_slot(ctx.r_bias)
            if rb_slot is None:
                d_rb = torch.zeros(Cc, device=xhat.device, dtype=torch.float32)
            rb_ptr = _p(rb_slot if rb_slot is not None else d_rb)
        _call("poet_layernorm_bwd", _p(gy), _p(gy2), _p(gy3), _p(gy4), _p(xhat), _p(rstd), _p(gamma), _p(dz), dg_ptr, db_ptr,
              R, Cc, _p(dr), rb_ptr, _p(seed), site, drop_p if dr is not None else 0.0, _stream(xhat),
              work=(4 * R * Cc * (2 + len(gs) + (dr is not None)) + 4 * R, 10 * R * Cc))   # gradients, xhat -> dz [, dr]
        dz = dz.view(ctx.shape)
        gpos = None
        if ctx.has_pos and ctx.needs_input_grad[4]:
            gpos = gy2_out.view(ctx.shape) if gy2_out is not None else None
        g_r = None if not ctx.has_r else (dr.view(ctx.shape) if dr is not None else dz)
        return dz, g_r, dgb[0], dgb[1], gpos, None, None, None, d_rb, None


def add_layernorm(x, r, gamma, beta, pos=None, eps: float = 1e-5, drop_p: float = 0.0, drop_site: int = 0, r_bias=None,
                  n_alias: int = 0):
    """LN(x + dropout_p(r)) [, + pos]; drop_p > 0 only in training (the residual branch's nn.Dropout).
    r_bias: the bias of the nn.Linear whose output is r, when that Linear was called with bias_grad_elsewhere=True:
    its gradient (the column sums of r's gradient) is then produced by this op's backward kernel.
    n_alias: extra autograd handles of y, one per additional reader (returns a tuple: y [, y2], alias_1 .. alias_n)."""
    return _AddLayerNorm.apply(x, r, gamma, beta, pos, eps, float(drop_p), int(drop_site), r_bias, int(n_alias))


def ffn_block(x, W1, b1, W2, b2, gamma, beta, pos=None, eps: float = 1e-5, drop_p: float = 0.0, site_hidden: int = 0,
              site_res: int = 0, x_mlp=None, n_alias: int = 0):
    """LN(x + dropout(linear2(dropout(relu(linear1(x)))))) [, + pos]: the FFN half of an encoder / decoder layer (reference
    deformable_transformer.py:193-197,205-206 and :267-271,289-290).  The linear2 bias gradient is summed by the LayerNorm
    backward kernel (r_bias) instead of a separate pass over the gradient rows.
    Measured and rejected (profiles/r02_fusion_ab.txt): one autograd node whose linear1 dgrad adds into the LayerNorm
    backward's dz through the beta = 1 TMA-reduce epilogue (no autograd accumulation pass) -- 7.53 vs 7.45 ms/step: the
    26 MB reduce-add store costs more than the 11 us add kernel it replaces."""
    # x_mlp: a second autograd handle of x for the MLP input (x itself feeds the residual), see add_layernorm(n_alias)
    f = mlp(x if x_mlp is None else x_mlp, ((W1, b1), (W2, b2)), drop_p=drop_p, drop_site=site_hidden, last_bias_grad_elsewhere=True)
    return add_layernorm(x, f, gamma, beta, pos=pos, eps=eps, drop_p=drop_p, drop_site=site_res, r_bias=b2, n_alias=n_alias)


def _planes_or_none(W: torch.Tensor, rows: int):
    """(hi, lo) plane TENSORS (the caller keeps them alive until its launch is enqueued) or (None, None)."""
    sp = split_weight(W, rows)
    return (None, None) if sp is None else sp


def linear_epilogue(x, W, b=None, residual=None, gamma=None, beta=None, relu: bool = False, eps: float = 1e-5):
    """Inference-only block entry point `poet_linear_epilogue` (no autograd): act(x W^T + b), or with gamma / beta
    LN(residual + x W^T + b) -- nn.Linear + residual + LayerNorm of a reference layer behind ONE C-ABI call."""
    x2 = _chk(x).view(-1, x.shape[-1])
    R, K = x2.shape
    N = W.shape[0]
    W = _chk(W)
    y = torch.empty((R, N), device=x.device, dtype=torch.float32)
    ln = gamma is not None
    ws_bytes = _lib.lib().poet_linear_epilogue_workspace_bytes(R, N, K, int(ln))
    ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8) if ws_bytes else None
    hi, lo = _planes_or_none(W, R)
    r2 = None if residual is None else _chk(residual).view(R, N)
    _call("poet_linear_epilogue", _p(x2), x2.stride(0), _p(W), _p(hi), _p(lo), _p(b), _p(r2), _p(gamma), _p(beta), _p(y), R, N, K,
          1 if relu else 0, eps, _state["precision"], _p(ws), ws_bytes, _stream(x))
    return y.view(*x.shape[:-1], N)


def ffn_fused(x, W1, b1, W2, b2, gamma, beta, eps: float = 1e-5):
    """Inference-only block entry point `poet_ffn_fused` (no autograd): LN(x + linear2(relu(linear1(x)))), the FFN half of
    an encoder / decoder layer behind ONE C-ABI call (three launches inside).  Training goes through `ffn_block`."""
    x2 = _chk(x).view(-1, x.shape[-1])
    R, Cc = x2.shape
    F = W1.shape[0]
    W1, W2 = _chk(W1), _chk(W2)
    y = torch.empty_like(x2)
    ws_bytes = _lib.lib().poet_ffn_fused_workspace_bytes(R, Cc, F)
    ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    h1, l1 = _planes_or_none(W1, R)
    h2, l2 = _planes_or_none(W2, R)
    _call("poet_ffn_fused", _p(x2), _p(W1), _p(h1), _p(l1), _p(b1), _p(W2), _p(h2), _p(l2), _p(b2), _p(gamma), _p(beta), _p(y), R, Cc, F, eps,
          _state["precision"], _p(ws), ws_bytes, _stream(x))
    _state["launches"] += 2                       # three kernels behind the one call
    return y.view(x.shape)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return add(_chk(a), _chk(b))

    @staticmethod
    def backward(ctx, g):
        return g, g


def add_tensors(a, b):
    return _Add.apply(a, b)


# ------------------------------------------------------------------------------------------
# autograd: multi-scale deformable attention
# ------------------------------------------------------------------------------------------
def msda_fwd_raw(value, a, lda, w, ldw, ref, shapes, B, S, Lq, M, D, L, P, mode):
    out = torch.empty((B, Lq, M * D), device=value.device, dtype=torch.float32)
    _call("poet_msda_fwd", _p(value), _p(a), lda, _p(w), ldw, _p(ref), _p(out), shapes_array(shapes),
          B, S, Lq, M, D, L, P, mode, _stream(value), tag=f"Lq={Lq}",
          work=(4 * B * (S * M * D + 3 * Lq * M * L * P + Lq * M * D), 10 * B * Lq * M * D * L * P))
    return out


class _MsdaCore(torch.autograd.Function):
    """mode 0: upstream MSDeformAttnFunction semantics (value, sampling_locations, attention_weights)."""

    @staticmethod
    def forward(ctx, value, shapes, loc, attn):
        value, loc, attn = _chk(value), _chk(loc), _chk(attn)
        B, S, M, D = value.shape
        _, Lq, _, L, P, _ = loc.shape
        ctx.dims = (B, S, Lq, M, D, L, P)
        ctx.shapes = shapes
        ctx.save_for_backward(value, loc, attn)
        return msda_fwd_raw(value, loc, M * L * P * 2, attn, M * L * P, None, shapes, B, S, Lq, M, D, L, P, 0)

    @staticmethod
    def backward(ctx, go):
        value, loc, attn = ctx.saved_tensors
        B, S, Lq, M, D, L, P = ctx.dims
        go = _chk(go)
        gv = torch.zeros_like(value)
        gl, ga = torch.empty_like(loc), torch.empty_like(attn)
        _call("poet_msda_bwd", _p(value), _p(loc), M * L * P * 2, _p(attn), M * L * P, None, _p(go), _p(gv), _p(gl),
              _p(ga), shapes_array(ctx.shapes), B, S, Lq, M, D, L, P, 0, _stream(value))
        return gv, None, gl, ga


def msda_core(value, shapes, loc, attn):
    return _MsdaCore.apply(value, tuple(tuple(s) for s in shapes), loc, attn)


class _MsdaBlock(torch.autograd.Function):
    """mode 1: `oa` is the fused projection output [B,Lq, M*L*P*2 + M*L*P] = [offsets | logits];
    softmax and loc = ref + off/(W,H) happen inside the gather kernel."""

    @staticmethod
    def forward(ctx, value, oa, ref, shapes, M, L, P, gv_buf=None):
        """gv_buf (optional): a zero-filled tensor shaped like `value` that the backward accumulates grad_value into and
        returns; the caller fills it in the forward pass on a side stream, which takes the 26 MB zero-fill out of the
        backward's dependent chain."""
        value, oa, ref = _chk(value), _chk(oa), _chk(ref)
        ctx.gv_buf = gv_buf if (gv_buf is not None and gv_buf.shape == value.shape and gv_buf.is_contiguous()) else None
        B, S, Cc = value.shape
        Lq = oa.shape[1]
        D = Cc // M
        n_off = M * L * P * 2
        ld = oa.shape[2]
        ctx.dims = (B, S, Lq, M, D, L, P, n_off, ld)
        ctx.shapes = shapes
        ctx.save_for_backward(value, oa, ref)
        logits = oa.view(-1)[n_off:]
        return msda_fwd_raw(value, oa, ld, logits, ld, ref, shapes, B, S, Lq, M, D, L, P, 1)

    @staticmethod
    def backward(ctx, go):
        value, oa, ref = ctx.saved_tensors
        B, S, Lq, M, D, L, P, n_off, ld = ctx.dims
        go = _chk(go)
        gv, ctx.gv_buf = (ctx.gv_buf, None) if ctx.gv_buf is not None else (torch.zeros_like(value), None)   # single use: a second backward re-zeros
        goa = torch.empty_like(oa)
        _call("poet_msda_bwd", _p(value), _p(oa), ld, _p(oa.view(-1)[n_off:]), ld, _p(ref), _p(go), _p(gv), _p(goa),
              _p(goa.view(-1)[n_off:]), shapes_array(ctx.shapes), B, S, Lq, M, D, L, P, 1, _stream(value),
              tag=f"Lq={Lq}", work=(4 * B * (2 * S * M * D + 6 * Lq * M * L * P + Lq * M * D), 30 * B * Lq * M * D * L * P))
        return gv, goa, None, None, None, None, None, None


def msda_block(value, oa, ref, shapes, M, L, P, grad_value_buf=None):
    return _MsdaBlock.apply(value, oa, ref, tuple(tuple(s) for s in shapes), M, L, P, grad_value_buf)


def grad_value_buffer(value: torch.Tensor) -> Optional[torch.Tensor]:
    """Zero-filled grad_value buffer for msda_block(grad_value_buf=...), or None when no gradient will be asked for."""
    if not (torch.is_grad_enabled() and value.requires_grad):
        return None
    return torch.zeros_like(value)


# ------------------------------------------------------------------------------------------
# autograd: decoder self-attention core
# ------------------------------------------------------------------------------------------
class _MhaSmallQ(torch.autograd.Function):
    """qk [B,Q,2C] (q | k projections of tgt+pos), v [B,Q,C] -> softmax(q k^T / sqrt(D)) v  [B,Q,C]."""

    @staticmethod
    def forward(ctx, qk, v, M, drop_p=0.0, drop_site=0):
        qk, v = _chk(qk), _chk(v)
        B, Q, C2 = qk.shape
        Cc = C2 // 2
        D = Cc // M
        out = torch.empty((B, Q, Cc), device=qk.device, dtype=torch.float32)
        probs = torch.empty((B, M, Q, Q), device=qk.device, dtype=torch.float32)
        scale = 1.0 / math.sqrt(D)
        seed, drop_p = _drop_args(drop_p)
        ctx.drop = (seed, int(drop_site), drop_p)
        _call("poet_mha_smallq_fwd", _p(qk), C2, _p(qk.view(-1)[Cc:]), C2, _p(v), Cc, _p(out), _p(probs), B, Q, M, D,
              scale, _p(seed), int(drop_site), drop_p, _stream(qk))
        ctx.save_for_backward(qk, v, probs)
        ctx.dims = (B, Q, M, D, Cc, scale)
        return out

    @staticmethod
    def backward(ctx, go):
        qk, v, probs = ctx.saved_tensors
        B, Q, M, D, Cc, scale = ctx.dims
        go = _chk(go)
        gqk, gv = torch.empty_like(qk), torch.empty_like(v)
        seed, site, drop_p = ctx.drop
        _call("poet_mha_smallq_bwd", _p(qk), 2 * Cc, _p(qk.view(-1)[Cc:]), 2 * Cc, _p(v), Cc, _p(probs), _p(go),
              _p(gqk), 2 * Cc, _p(gqk.view(-1)[Cc:]), 2 * Cc, _p(gv), Cc, B, Q, M, D, scale, _p(seed), site, drop_p, _stream(qk))
        return gqk, gv, None, None, None


def mha_smallq(qk, v, M, drop_p: float = 0.0, drop_site: int = 0):
    """softmax(q k^T / sqrt(D)) v; drop_p > 0: nn.MultiheadAttention's dropout on the attention probabilities."""
    return _MhaSmallQ.apply(qk, v, M, float(drop_p), int(drop_site))


# ------------------------------------------------------------------------------------------
# autograd: heads tail
# ------------------------------------------------------------------------------------------
class _HeadsSelect(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rot_all, trans_all, classes, n_slots):
        rot_all, trans_all = _chk(rot_all), _chk(trans_all)
        lead = rot_all.shape[:-1]
        R = rot_all.numel() // rot_all.shape[-1]
        cls = None if classes is None else _chk(classes.reshape(-1), torch.int64)
        dev = rot_all.device
        trans = torch.empty((R, 3), device=dev, dtype=torch.float32)
        rot6d = torch.empty((R, 6), device=dev, dtype=torch.float32)
        rotmat = torch.empty((R, 9), device=dev, dtype=torch.float32)
        _call("poet_heads_select_rot6d_fwd", _p(rot_all), _p(trans_all), _p(cls), _p(trans), _p(rot6d), _p(rotmat), R,
              n_slots, _stream(rot_all))
        ctx.save_for_backward(rot6d, cls) if cls is not None else ctx.save_for_backward(rot6d)
        ctx.meta = (R, n_slots, cls is not None, rot_all.shape, trans_all.shape)
        ctx.mark_non_differentiable(rot6d)
        return trans.view(*lead, 3), rotmat.view(*lead, 3, 3), rot6d.view(*lead, 6)

    @staticmethod
    def backward(ctx, g_trans, g_rot, _g6):
        R, n_slots, has_cls, rshape, tshape = ctx.meta
        saved = ctx.saved_tensors
        rot6d, cls = saved[0], (saved[1] if has_cls else None)
        dev = rot6d.device
        g_trans = torch.zeros((R, 3), device=dev) if g_trans is None else _chk(g_trans)
        g_rot = torch.zeros((R, 9), device=dev) if g_rot is None else _chk(g_rot)
        gr = torch.empty(rshape, device=dev, dtype=torch.float32)
        gt = torch.empty(tshape, device=dev, dtype=torch.float32)
        _call("poet_heads_select_rot6d_bwd", _p(rot6d), _p(cls), _p(g_trans), _p(g_rot), _p(gr), _p(gt), R, n_slots,
              _stream(rot6d))
        return gr, gt, None, None


def heads_select_rot6d(rot_all, trans_all, classes, n_slots):
    return _HeadsSelect.apply(rot_all, trans_all, classes, n_slots)


# ------------------------------------------------------------------------------------------
# position encodings / layout (no parameters except level_embed)
# ------------------------------------------------------------------------------------------
_dim_t_cache = {}


def _dim_t(F: int, temperature: float, device) -> torch.Tensor:
    key = (F, float(temperature), str(device))
    if key not in _dim_t_cache:
        k = torch.arange(F, dtype=torch.float32)
        _dim_t_cache[key] = (temperature ** (2 * (k // 2) / F)).to(device)     # position_encoding.py:52-53
    return _dim_t_cache[key]


def posenc_sine_nchw(mask: torch.Tensor, F: int = 128, temperature: float = 10000.0, normalize: bool = True,
                     scale: float = 2 * math.pi) -> torch.Tensor:
    """mask [B,H,W] bool -> [B,2F,H,W] (reference layout, position_encoding.py:40-60)."""
    B, H, W = mask.shape
    m8 = _chk(as_u8(mask), torch.uint8)
    out = torch.empty((B, 2 * F, H, W), device=mask.device, dtype=torch.float32)
    _call("poet_posenc_sine", _p(m8), _p(_dim_t(F, temperature, mask.device)), None, _p(out), B, H, W, F, scale,
          int(normalize), 0, 0, 0, _stream(out))
    return out


def posenc_sine_tokens_(out: torch.Tensor, mask: torch.Tensor, level_embed_row: Optional[torch.Tensor], row_offset: int,
                        F: int = 128, temperature: float = 10000.0, normalize: bool = True,
                        scale: float = 2 * math.pi) -> None:
    """Writes pos (+level_embed) for one level straight into the token-major [B,S,2F] buffer."""
    B, H, W = mask.shape
    m8 = _chk(as_u8(mask), torch.uint8)
    _call("poet_posenc_sine", _p(m8), _p(_dim_t(F, temperature, mask.device)), _p(level_embed_row), _p(out), B, H, W, F,
          scale, int(normalize), 1, out.shape[1], row_offset, _stream(out))


def bbox_embed_pad(boxes_padded: torch.Tensor, n_boxes: torch.Tensor, F: int) -> torch.Tensor:
    """boxes [B,Q,4] (dummies -1), n_boxes [B] int32 -> query_embeds [B,Q,16F]."""
    B, Q, _ = boxes_padded.shape
    out = torch.empty((B, Q, 16 * F), device=boxes_padded.device, dtype=torch.float32)
    _call("poet_bbox_embed_pad", _p(_chk(boxes_padded)), _p(_chk(n_boxes, torch.int32)), _p(out), B, Q, F, _stream(out))
    return out


def as_u8(mask: torch.Tensor) -> torch.Tensor:
    """bool -> uint8 without a copy (same bytes)."""
    if mask.dtype == torch.uint8:
        return mask
    return mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)


def mask_prep(masks: Sequence[torch.Tensor]):
    """Per-level masks [B,H_l,W_l] -> (pad [B,S] uint8, valid_ratios [B,L,2] fp32) in one launch."""
    m8 = [_chk(as_u8(m), torch.uint8) for m in masks]
    B, L = m8[0].shape[0], len(m8)
    shapes = [(int(m.shape[1]), int(m.shape[2])) for m in m8]
    S = sum(h * w for h, w in shapes)
    pad = torch.empty((B, S), device=m8[0].device, dtype=torch.uint8)
    vr = torch.empty((B, L, 2), device=m8[0].device, dtype=torch.float32)
    ptrs = (C.c_void_p * L)(*[m.data_ptr() for m in m8])
    _call("poet_mask_prep", ptrs, shapes_array(shapes), _p(pad), _p(vr), B, L, _stream(pad))
    return pad, vr


def enc_reference_points(valid_ratios: torch.Tensor, shapes) -> torch.Tensor:
    B, L, _ = valid_ratios.shape
    S = sum(h * w for h, w in shapes)
    out = torch.empty((B, S, L, 2), device=valid_ratios.device, dtype=torch.float32)
    _call("poet_enc_reference_points", _p(_chk(valid_ratios)), _p(out), shapes_array(shapes), B, L, _stream(out))
    return out


class _FlattenLevels(torch.autograd.Function):
    """L x [B,C,H,W] -> [B,S,C] tokens (deformable_transformer.py:124-140); optional per-level vector
    add (level_embed) whose gradient is the per-level column sum."""

    @staticmethod
    def forward(ctx, level_embed, *maps):
        B, Cc = maps[0].shape[:2]
        hws = [m.shape[2] * m.shape[3] for m in maps]
        S = sum(hws)
        out = torch.empty((B, S, Cc), device=maps[0].device, dtype=torch.float32)
        off = 0
        for l, m in enumerate(maps):
            vec = None if level_embed is None else level_embed[l]
            _call("poet_nchw_to_tokens", _p(_chk(m)), _p(vec), _p(out), B, Cc, hws[l], S, off, _stream(out))
            off += hws[l]
        ctx.meta = (B, Cc, hws, S, [m.shape for m in maps], level_embed is not None)
        ctx.le_param = level_embed
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cc, hws, S, shapes, has_le = ctx.meta
        g = _chk(g)
        gle, gle_ret = None, None
        if has_le and ctx.needs_input_grad[0]:
            gle = _grad_slot(ctx.le_param)                 # accumulate straight into level_embed.grad when it exists
            if gle is None:
                gle = gle_ret = torch.zeros((len(hws), Cc), device=g.device, dtype=torch.float32)
        outs = []
        off = 0
        for l, hw in enumerate(hws):
            need = ctx.needs_input_grad[1 + l]
            gm = torch.empty(shapes[l], device=g.device, dtype=torch.float32) if need else None
            if need or gle is not None:
                _call("poet_tokens_to_nchw", _p(g), _p(gm), _p(gle[l]) if gle is not None else None, B, Cc, hw, S, off,
                      _stream(g))
            outs.append(gm)
            off += hw
        return (gle_ret, *outs)


def flatten_levels(maps: Sequence[torch.Tensor], level_embed: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _FlattenLevels.apply(level_embed, *maps)


# ------------------------------------------------------------------------------------------
# autograd: input_proj
# ------------------------------------------------------------------------------------------
class _InputProj(torch.autograd.Function):
    """Backbone feature maps -> the transformer's token matrix [B, S, C]: per level Conv2d(1x1) + GroupNorm(32), and
    for the extra level Conv2d(3x3, stride 2, padding 1) + GroupNorm(32) on the last feature map (reference
    models/pose_estimation_transformer.py:100-135, 313-335).  The convolutions are tensor-core GEMMs on token rows
    (NCHW -> rows by poet_nchw_to_tokens / poet_im2col_3x3s2), GroupNorm writes each level straight into its
    slice of the token matrix, so the NCHW projections and their flatten / transpose / cat copies never exist.
    Gradients: conv weights / biases and GroupNorm affine parameters (the backbone is frozen: no input gradient)."""

    @staticmethod
    def forward(ctx, n_feats, groups, eps, *tensors):
        ctx.prec = _state["precision"]
        feats = [_chk(t) for t in tensors[:n_feats]]
        params = tensors[n_feats:]                           # per level: conv weight, conv bias, gn weight, gn bias
        L = len(params) // 4
        if L > n_feats + 1:
            raise NotImplementedError("more than one extra feature level")
        B = feats[0].shape[0]
        C = params[0].shape[0]
        dims = []                                            # (H, W) per level
        for f in feats:
            dims.append((int(f.shape[2]), int(f.shape[3])))
        if L == n_feats + 1:
            H, W = dims[-1]
            dims.append(((H - 1) // 2 + 1, (W - 1) // 2 + 1))
        S = sum(h * w for h, w in dims)
        tokens = torch.empty((B, S, C), device=feats[0].device, dtype=torch.float32)
        saved, off = [], 0
        for l in range(L):
            Wc, bc, gw, gb = params[4 * l: 4 * l + 4]
            HW = dims[l][0] * dims[l][1]
            if l < n_feats:
                Cin = feats[l].shape[1]
                x2 = torch.empty((B * HW, Cin), device=tokens.device, dtype=torch.float32)
                _call("poet_nchw_to_tokens", _p(feats[l]), None, _p(x2), B, Cin, HW, HW, 0, _stream(x2))
            else:
                f = feats[-1]
                Cin = f.shape[1] * 9
                x2 = torch.empty((B * HW, Cin), device=tokens.device, dtype=torch.float32)
                _call("poet_im2col_3x3s2", _p(f), _p(x2), B, f.shape[1], f.shape[2], f.shape[3], _stream(x2))
            W2 = _chk(Wc).view(C, Cin)
            y = gemm(x2, W2, B * HW, C, Cin, bias=bc, b_split=split_weight(W2, B * HW))
            stats = torch.empty((B, groups, 2), device=tokens.device, dtype=torch.float64)
            _call("poet_groupnorm_tokens_fwd", _p(y), _p(gw), _p(gb), _p(tokens), _p(stats), B, HW, C, groups, S, off,
                  float(eps), _stream(y))
            saved += [x2, y, stats]
            off += HW
        ctx.save_for_backward(*saved)
        ctx.params, ctx.meta = params, (n_feats, groups, eps, B, C, S, dims)
        return tokens

    @staticmethod
    def backward(ctx, *grads):
        with precision_scope(ctx.prec):
            return _InputProj._backward_impl(ctx, *grads)

    @staticmethod
    def _backward_impl(ctx, g):
        n_feats, groups, eps, B, C, S, dims = ctx.meta
        g = _chk(g)
        saved, params = ctx.saved_tensors, ctx.params
        grads, off = [], 0
        for l in range(len(params) // 4):
            Wc, bc, gw, gb = params[4 * l: 4 * l + 4]
            x2, y, stats = saved[3 * l: 3 * l + 3]
            HW = dims[l][0] * dims[l][1]
            Cin = x2.shape[1]
            gw_slot, gb_slot = _grad_slot(gw), _grad_slot(gb)
            dgw = gw_slot if gw_slot is not None else torch.zeros(C, device=g.device, dtype=torch.float32)
            dgb = gb_slot if gb_slot is not None else torch.zeros(C, device=g.device, dtype=torch.float32)
            dy = torch.empty_like(y)
            ws = torch.empty((B, groups, 2), device=g.device, dtype=torch.float64)
            _call("poet_groupnorm_tokens_bwd", _p(g), _p(y), _p(stats), _p(gw), _p(dy), _p(dgw), _p(dgb), _p(ws), B, HW, C,
                  groups, S, off, float(eps), _stream(g))
            w_slot, b_slot = _grad_slot(Wc), _grad_slot(bc)
            dW = w_slot.view(C, Cin) if w_slot is not None else torch.zeros((C, Cin), device=g.device, dtype=torch.float32)
            db = b_slot if b_slot is not None else torch.zeros(C, device=g.device, dtype=torch.float32)
            wgrad_bias(dy, x2, C, Cin, B * HW, dW, db)
            grads += [None if w_slot is not None else dW.view_as(Wc), None if b_slot is not None else db,
                      None if gw_slot is not None else dgw, None if gb_slot is not None else dgb]
            off += HW
        return (None, None, None, *([None] * n_feats), *grads)


def input_proj_tokens(feats: Sequence[torch.Tensor], levels: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]],
                      groups: int = 32, eps: float = 1e-5) -> torch.Tensor:
    """levels[l] = (conv weight, conv bias, GroupNorm weight, GroupNorm bias); returns the token matrix [B, S, C]."""
    flat = [t for lv in levels for t in lv]
    return _InputProj.apply(len(feats), groups, eps, *feats, *flat)
