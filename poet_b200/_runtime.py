"""Runtime state shared by the op bindings: GEMM precision mode, side-stream forking, gradient markers, the dropout
counter, launch accounting / per-call timing and the ctypes call helper.  Split out of ops.py (round 2); ops.py re-exports
every name, so `from poet_b200 import ops; ops.set_gemm_precision(...)` keeps working.

PyTorch is plumbing here: it owns device memory, the CUDA stream and the autograd tape; every
numeric operation is a call into libpoet_b200.so.  Nothing has a CPU or eager-PyTorch fallback: a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

GEMM_FP32, GEMM_BF16X3, GEMM_BF16 = 0, 1, 2
_PRECISION = {"fp32": GEMM_FP32, "bf16x3": GEMM_BF16X3, "bf16": GEMM_BF16}
import os as _os

# default: tcgen05 split-bf16 (fp32-grade) for the large contractions; the SIMT fp32 kernel serves the small ones
_state = {"precision": _PRECISION[_os.environ.get("POET_GEMM_PRECISION", "bf16x3")], "launches": 0, "direct_grads": True}


def set_gemm_precision(name: str) -> None:
    """'fp32' (SIMT FFMA), 'bf16x3' (tcgen05 split-bf16, fp32-grade) or 'bf16' (tcgen05 single pass)."""
    _state["precision"] = _PRECISION[name]


def get_gemm_precision() -> str:
    return {v: k for k, v in _PRECISION.items()}[_state["precision"]]


class precision_scope:
    """with precision_scope('bf16'): ... -- GEMMs issued inside use that precision (None: leave as is); the backward
    of every op recorded inside runs at the precision of its forward (each autograd Function stores it).  Used by the
    mixed throughput mode (BASELINE.json cfg4): single-pass bf16 MMAs for the encoder layers, bf16x3 elsewhere.
    The weight planes are shared: a bf16 GEMM simply ignores the lo plane."""

    def __init__(self, name: Optional[str]):
        self.value = None if name is None else (_PRECISION[name] if isinstance(name, str) else int(name))

    def __enter__(self):
        self.saved = _state["precision"]
        if self.value is not None:
            if self.saved == GEMM_FP32 and self.value != GEMM_FP32:
                raise RuntimeError("precision_scope cannot enable tensor-core GEMMs under the global 'fp32' mode (no weight planes)")
            _state["precision"] = self.value
        return self

    def __exit__(self, *exc):
        _state["precision"] = self.saved
        return False


# ------------------------------------------------------------------------------------------
# stream forking: the decoder / head chain is a sequence of launch-latency-bound kernels that leaves the
# GPU mostly idle, so independent work (the decoder layers' value projections of `memory`, the per-layer
# pose heads and their backward) is issued on side streams.  Under CUDA-graph capture the fork/join
# events become graph edges and the branches run concurrently; autograd replays each op's backward on
# the stream of its forward, so the backward overlaps the same way.
# ------------------------------------------------------------------------------------------
_side_streams = {}
_stream_ns = [0]          # namespace of the side streams: each micro-batch forks onto its own set


class stream_namespace:
    """Side streams requested inside the block are private to namespace `ns` (micro-batch index)."""

    def __init__(self, ns: int):
        self.ns = ns

    def __enter__(self):
        _stream_ns.append(self.ns)
        return self

    def __exit__(self, *exc):
        _stream_ns.pop()
        return False


def side_stream(idx: int, device) -> "torch.cuda.Stream":
    key = (str(device), _stream_ns[-1], idx)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def parallel_streams_enabled() -> bool:
    return _state.get("parallel_streams", True)


def set_parallel_streams(on: bool) -> None:
    _state["parallel_streams"] = bool(on)


class fork:
    """with fork(idx, device) as f: ... work issued on side stream idx ...; f.join() makes the caller's stream wait."""

    def __init__(self, idx: int, device):
        self.main = torch.cuda.current_stream(device)
        self.side = side_stream(idx, device)
        self._ctx = None

    def __enter__(self):
        self.side.wait_stream(self.main)
        _touched_side_streams[id(self.side)] = self.side
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._ctx.__exit__(*exc)
        return False

    def join(self, *tensors) -> None:
        """Caller's stream waits for the side stream; `tensors` produced there are marked as used on it."""
        self.main.wait_stream(self.side)
        for t in tensors:
            if isinstance(t, torch.Tensor):
                t.record_stream(self.main)

    def uses(self, *tensors) -> None:
        """Tensors allocated on the caller's stream that the side stream reads."""
        for t in tensors:
            if isinstance(t, torch.Tensor):
                t.record_stream(self.side)

    def checkpoint(self) -> "torch.cuda.Event":
        """Event marking the side-stream work issued so far (call inside the `with` block)."""
        ev = torch.cuda.Event()
        ev.record(self.side)
        return ev

    def wait(self, ev, *tensors) -> None:
        """Caller's stream waits for a checkpoint; `tensors` produced before it are marked as used there."""
        self.main.wait_event(ev)
        for t in tensors:
            if isinstance(t, torch.Tensor):
                t.record_stream(self.main)


# Side streams that received work since the last reset_touched_side_streams(): what a gradient all-reduce issued in the
# middle of the backward pass has to wait for besides the calling stream (data_parallel.FlatGradReducer.on_marker).
_touched_side_streams = {}


def reset_touched_side_streams() -> None:
    _touched_side_streams.clear()


def touched_side_streams():
    return list(_touched_side_streams.values())


# Gradient-ready markers: identity in forward; in backward they tell a registered callback that every gradient kernel
# downstream of this point of the forward graph has been ISSUED (autograd has finished all nodes created after it).
_marker_cb = [None]


def set_grad_marker_callback(cb) -> None:
    _marker_cb[0] = cb


class _GradMarker(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, key):
        ctx.key = key
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        cb = _marker_cb[0]
        if cb is not None:
            cb(ctx.key)
        return g, None


def grad_marker(x: torch.Tensor, key):
    """No-op unless a callback is registered (overlapped gradient all-reduce) and x carries a gradient."""
    if _marker_cb[0] is None or not (torch.is_grad_enabled() and x.requires_grad):
        return x
    return _GradMarker.apply(x, key)


class background_gemms:
    """with background_gemms(): ... -- the large GEMMs issued inside (and the backward of the ops recorded inside) carry
    POET_GEMM_BACKGROUND: the persistent tensor-core kernel leaves some SMs free for a latency-bound chain that runs on
    another stream at the same time (the decoder's value projections of `memory` next to the query-row chain)."""

    def __init__(self, on: bool = True):
        self.on = bool(on)

    def __enter__(self):
        self.saved = _state.get("gemm_background", False)
        _state["gemm_background"] = self.on
        return self

    def __exit__(self, *exc):
        _state["gemm_background"] = self.saved
        return False


def launch_count() -> int:
    """Number of libpoet_b200 kernel-launching calls issued so far (bench.py's gpu_launches)."""
    return _state["launches"]


# ------------------------------------------------------------------------------------------
# train-mode dropout: counter-based, no mask tensors (include/poet_b200.h "Train-mode dropout")
# ------------------------------------------------------------------------------------------
# One int64 counter per device.  Every training forward bumps it IN PLACE (so the bump is a node of a captured CUDA
# graph and every replay draws new masks) and takes a private snapshot; all dropout sites of that forward -- and
# their backward kernels, which regenerate the masks -- read the snapshot through its device pointer.
_drop_counters = {}


def set_dropout_seed(seed: int, device=None) -> None:
    """Deterministic mask sequence from here on (the analogue of torch.manual_seed for the dropout of this library)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    c = _drop_counters.get(str(dev))
    if c is None:
        _drop_counters[str(dev)] = torch.full((1,), int(seed), dtype=torch.int64, device=dev)
    else:
        c.fill_(int(seed))


def begin_dropout_forward(device) -> torch.Tensor:
    """Called once per training forward with dropout > 0: returns this forward's seed tensor (int64 [1], device)."""
    key = str(torch.device(device))
    c = _drop_counters.get(key)
    if c is None:
        c = _drop_counters[key] = torch.full((1,), int(torch.initial_seed()) & 0x7FFFFFFFFFFF, dtype=torch.int64, device=device)
    c.add_(1)
    cur = c.clone()
    _state["drop_seed"] = cur
    return cur


def _drop_args(p: float):
    """(seed tensor, p) for an op called with dropout probability p; the seed must have been set by the model."""
    if p <= 0.0:
        return None, 0.0
    if not 0.0 < p < 1.0:
        raise ValueError(f"dropout probability must be in [0, 1), got {p}")
    seed = _state.get("drop_seed")
    if seed is None:
        raise RuntimeError("dropout > 0 needs ops.begin_dropout_forward() at the start of the forward pass")
    return seed, float(p)


def dropout_scale(p: float, pair_scheme: bool = False) -> float:
    return float(_lib.lib().poet_dropout_scale(float(p), int(pair_scheme))) if p > 0.0 else 1.0


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.PoetLibraryError("poet_b200 ops run on CUDA tensors only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _stream(t: torch.Tensor):
    _lib.require_b200(t.device.index if t.device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(t.device).cuda_stream


_timing = {"on": False, "events": {}}


def kernel_timing(enable: bool) -> None:
    """bench.py: bracket every library call with CUDA events on the launching stream."""
    _timing["on"] = enable
    if enable:
        _timing["events"] = {}
        _timing["work"] = {}


def kernel_times_ms() -> dict:
    """name -> (total ms, launches, algorithmic bytes, flops); call after a device synchronize."""
    work = _timing.get("work", {})
    return {k: (sum(s.elapsed_time(e) for s, e in v), len(v), *work.get(k, (0, 0))) for k, v in _timing["events"].items()}


_B_STABLE = _os.environ.get("POET_GEMM_B_STABLE", "1") != "0"
_ABLATE = frozenset(x for x in _os.environ.get("POET_ABLATE_CALLS", "").split(",") if x)


def _call(name: str, *args, tag: Optional[str] = None, work=None) -> None:
    """`tag` / `work` = (algorithmic bytes, flops) only feed bench.py's per-kernel roofline table."""
    if _ABLATE and name in _ABLATE:              # timing experiments only (tools/gpu_ab.sh): results are wrong
        return
    _state["launches"] += 1
    if _timing["on"]:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = getattr(_lib.lib(), name)(*args)
        e.record()
        key = name if tag is None else f"{name}[{tag}]"
        _timing["events"].setdefault(key, []).append((s, e))
        if work is not None:
            acc = _timing.setdefault("work", {}).setdefault(key, [0, 0])
            acc[0] += work[0]
            acc[1] += work[1]
    else:
        rc = getattr(_lib.lib(), name)(*args)
    if rc != 0:
        _lib.check(rc, name)


def shapes_array(shapes: Sequence[Tuple[int, int]]):
    flat = [int(v) for hw in shapes for v in hw]
    return (C.c_int32 * len(flat))(*flat)


# forks whose join is deferred to the end of the backward pass (ops._join_at_end_of_backward); cleared by planes_scope
_pending_joins = {}
