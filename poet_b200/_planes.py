"""bf16 hi/lo weight planes: one arena per module tree, refreshed by one launch per step (or written by the fused
optimizer), looked up by the GEMM bindings.  Split out of ops.py (round 2); ops.py re-exports every name."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib
from ._runtime import GEMM_BF16X3, GEMM_FP32, _call, _p, _pending_joins, _state, _stream

# Weights are the B operand of the forward (NT) and of the dgrad (NN) GEMM of a layer: they are split into
# bf16 hi/lo planes once per step and both GEMMs fetch the planes by TMA.  `WeightPlanes` does that for ALL
# weight matrices of a module tree in one launch into one arena (the per-layer autograd Functions then only
# look their planes up); a weight that is not covered falls back to its own poet_split_bf16 launch.
class WeightPlanes:
    """bf16 hi/lo planes of every 2-D parameter of `module`, refreshed by one poet_split_bf16_multi launch.
    Parameters keep their registration order in the arena, so row-blocks of one matrix and consecutive
    matrices (sampling_offsets | attention_weights) are contiguous plane views as well."""

    @staticmethod
    def select(module: torch.nn.Module):
        # matrices and conv kernels (input_proj: [out, in, kh, kw] is the GEMM weight [out, in*kh*kw])
        return [p for p in module.parameters() if p.dim() in (2, 4) and p.dtype == torch.float32 and p.is_cuda
                and p.is_contiguous() and p.numel() % 8 == 0]

    def __init__(self, module: torch.nn.Module):
        params = self.select(module)
        self.params = params
        self.device = params[0].device if params else None
        total = sum(p.numel() for p in params)
        self.hi = torch.empty(total, device=self.device, dtype=torch.bfloat16) if params else None
        self.lo = torch.empty(total, device=self.device, dtype=torch.bfloat16) if params else None
        self.ranges = []                      # (data_ptr, nbytes, element offset in the arena)
        off = 0
        for p in params:
            self.ranges.append((p.data_ptr(), p.numel() * 4, off))
            off += p.numel()
        self._table_key, self._table, self._chunks = None, None, 0
        # all 1-D parameters (biases, norm affine) in registration order: refresh() concatenates them with ONE launch, so
        # [sampling_offsets.bias | attention_weights.bias] of the fused projection is a view instead of a cat per layer
        self.vec_params = [p for p in module.parameters() if p.dim() == 1 and p.dtype == torch.float32 and p.is_cuda]
        self.vec_flat, self.vec_off = None, {}
        o = 0
        for p in self.vec_params:
            self.vec_off[p.data_ptr()] = (o, p.numel())
            o += p.numel()
        self.with_lo = True
        self._fresh_versions = None           # parameter versions for which the planes are known to be current

    def _build_table(self, with_lo: bool):
        import struct
        raw, chunk, off = bytearray(), 0, 0
        for p in self.params:
            n4 = p.numel() // 4
            raw += struct.pack("<QQQqq", p.data_ptr(), self.hi.data_ptr() + 2 * off,
                               (self.lo.data_ptr() + 2 * off) if with_lo else 0, n4, chunk)
            chunk += (n4 + 1023) // 1024
            off += p.numel()
        self._table = torch.frombuffer(raw, dtype=torch.uint8).clone().to(self.device)
        self._chunks = chunk

    def mark_fresh(self) -> None:
        """The planes were just written from the current parameter values by someone else (the fused optimizer
        step): the next refresh() is a no-op unless a parameter is modified in between."""
        self._fresh_versions = [p._version for p in self.params]

    def bias_pair(self, b0: torch.Tensor, b1: torch.Tensor):
        """cat(b0, b1) as a view of the per-step flat copy of the 1-D parameters, or None."""
        if self.vec_flat is None:
            return None
        r0, r1 = self.vec_off.get(b0.data_ptr()), self.vec_off.get(b1.data_ptr())
        if r0 is None or r1 is None or r1[0] != r0[0] + r0[1] or r0[1] != b0.numel() or r1[1] != b1.numel() or r0[0] % 4:
            return None
        return self.vec_flat[r0[0]: r0[0] + r0[1] + r1[1]]

    def refresh_vectors(self) -> None:
        """Re-copy the 1-D parameters into the flat buffer, IN PLACE: the buffer's address is baked into captured
        CUDA graphs (bias_pair views), and this copy is part of every forward -- also of one replayed from a graph
        whose planes are written by the fused optimizer -- so a bias updated by optimizer.step() or load_state_dict
        is what the next forward reads."""
        if not self.vec_params:
            return
        with torch.no_grad():
            srcs = [p.detach().reshape(-1) for p in self.vec_params]
            if self.vec_flat is None or self.vec_flat.numel() != sum(t.numel() for t in srcs):
                self.vec_flat = torch.empty(sum(t.numel() for t in srcs), device=self.device, dtype=torch.float32)
            torch.cat(srcs, out=self.vec_flat)

    def refresh(self) -> None:
        """Re-derive all planes from the current parameter values (call once per forward)."""
        prec = _state["precision"]
        if not self.params or prec == GEMM_FP32:
            return
        self.refresh_vectors()
        if (self._fresh_versions is not None and self.with_lo == (prec == GEMM_BF16X3) and
                self._fresh_versions == [p._version for p in self.params]):
            return
        key = (tuple(p.data_ptr() for p in self.params), prec == GEMM_BF16X3)
        if key != self._table_key:
            self.ranges, off = [], 0
            for p in self.params:
                self.ranges.append((p.data_ptr(), p.numel() * 4, off))
                off += p.numel()
            self._build_table(prec == GEMM_BF16X3)
            self._table_key = key
        self.with_lo = prec == GEMM_BF16X3
        _call("poet_split_bf16_multi", _p(self._table), len(self.params), self._chunks,
              torch.cuda.current_stream(self.device).cuda_stream)

    def lookup(self, ptr: int, numel: int):
        """(hi, lo) flat plane views for the fp32 range [ptr, ptr + 4*numel) if it lies inside the arena's
        parameters (a whole matrix, a row block, or consecutive matrices), else None."""
        if getattr(self, "_by_base_src", None) is not self.ranges:       # rebuilt whenever the ranges list is replaced
            self._by_base = {base: (nbytes, off) for base, nbytes, off in self.ranges}
            self._by_base_src = self.ranges
        hit = self._by_base.get(ptr)
        if hit is not None and 4 * numel <= hit[0]:
            off = hit[1]
            return self.hi[off:off + numel], (self.lo[off:off + numel] if self.with_lo else None)
        for base, nbytes, off in self.ranges:
            if base <= ptr < base + nbytes:
                e0 = off + (ptr - base) // 4
                # consecutive parameters are consecutive in the arena only if they are consecutive in memory too
                if ptr + 4 * numel > base + nbytes:
                    return None
                hi = self.hi[e0:e0 + numel]
                lo = self.lo[e0:e0 + numel] if self.with_lo else None
                return hi, lo
        return None

    def lookup_pair(self, W0: torch.Tensor, W1: torch.Tensor):
        """Planes of cat(W0, W1) when the two matrices follow each other in the arena (no fp32 cat needed)."""
        r0 = r1 = None
        for base, nbytes, off in self.ranges:
            if base == W0.data_ptr() and nbytes == W0.numel() * 4:
                r0 = off
            if base == W1.data_ptr() and nbytes == W1.numel() * 4:
                r1 = off
        if r0 is None or r1 is None or r1 != r0 + W0.numel():
            return None
        n = W0.numel() + W1.numel()
        return self.hi[r0:r0 + n], (self.lo[r0:r0 + n] if self.with_lo else None)


_active_planes: List[WeightPlanes] = []


class planes_scope:
    """with planes_scope(module): ... -- inside, split_weight() is served from the module's refreshed arena.
    The arena object is cached on the module; nested scopes whose parameters are already covered are no-ops."""

    def __init__(self, module: torch.nn.Module, refresh: bool = True):
        """refresh=False: trust the arena as it is (the fused optimizer step wrote the planes of the new weights;
        used when the forward is replayed from a CUDA graph that must not contain the split pass)."""
        self.module, self.pushed, self.do_refresh = module, False, refresh

    def __enter__(self):
        if not _active_planes:
            _pending_joins.clear()                        # a backward that raised must not suppress the next one's joins
        first = next((p for p in self.module.parameters() if p.dim() in (2, 4)), None)
        if first is None or not first.is_cuda or _state["precision"] == GEMM_FP32:
            return self
        if any(pl.lookup(first.data_ptr(), first.numel()) is not None for pl in _active_planes):
            return self                                   # an enclosing scope already covers this module
        pl = getattr(self.module, "_poet_weight_planes", None)
        if pl is None or [p.data_ptr() for p in pl.params] != [p.data_ptr() for p in WeightPlanes.select(self.module)]:
            if not self.do_refresh:
                raise RuntimeError("planes_scope(refresh=False) needs planes written by FusedClipAdamW for this module")
            pl = WeightPlanes(self.module)
            object.__setattr__(self.module, "_poet_weight_planes", pl)
        if self.do_refresh:
            pl.refresh()
        else:
            pl.refresh_vectors()                          # the optimizer writes the matrix planes, not the bias copy
        _active_planes.append(pl)
        self.pushed = True
        return self

    def __exit__(self, *exc):
        if self.pushed:
            _active_planes.pop()
        return False


def _eligible(M_rows: int, N: int, K: int) -> bool:
    return not (K % 8 or N % 8) and bool(_lib.lib().poet_gemm_tc_eligible(M_rows, N, K, K, K, N))


def split_weight(W: torch.Tensor, M_rows: int):
    """(hi, lo) bf16 planes of W [N,K] if the GEMMs that will use it are tensor-core eligible, else None."""
    prec = _state["precision"]
    if prec == GEMM_FP32:
        return None
    N, K = W.shape
    if not _eligible(M_rows, N, K):
        return None
    if W.is_contiguous():
        for pl in reversed(_active_planes):
            v = pl.lookup(W.data_ptr(), W.numel())
            if v is not None:
                return v[0].view(N, K), (v[1].view(N, K) if v[1] is not None else None)
    hi = torch.empty(W.shape, device=W.device, dtype=torch.bfloat16)
    lo = torch.empty(W.shape, device=W.device, dtype=torch.bfloat16) if prec == GEMM_BF16X3 else None
    _call("poet_split_bf16", _p(W), _p(hi), _p(lo), W.numel(), _stream(W))
    return hi, lo


def split_weight_pair(W0: torch.Tensor, W1: torch.Tensor, M_rows: int):
    """Planes of cat(W0, W1) [N0+N1, K] straight from the arena, or None (caller concatenates and splits)."""
    N, K = W0.shape[0] + W1.shape[0], W0.shape[1]
    # the planes serve the forward [R,N,K] AND the dgrad [R,K,N] GEMM (the fp32 matrix is then never built): both
    # must be tensor-core eligible, else the caller concatenates and keeps the fp32 copy for the SIMT path
    if _state["precision"] == GEMM_FP32 or not _eligible(M_rows, N, K) or not _eligible(M_rows, K, N):
        return None
    for pl in reversed(_active_planes):
        v = pl.lookup_pair(W0, W1)
        if v is not None:
            N, K = W0.shape[0] + W1.shape[0], W0.shape[1]
            return v[0].view(N, K), (v[1].view(N, K) if v[1] is not None else None)
    return None


