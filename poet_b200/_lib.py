"""ctypes binding of libpoet_b200.so (C ABI declared in include/poet_b200.h).

There is no fallback of any kind: if the shared library is missing, or a call returns non-zero,
this module raises.  `lib()` loads lazily so that importing the package (e.g. to build it) works
before the first build.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libpoet_b200.so")

_vp, _i, _i64, _f, _sz, _u32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint32

# name -> (restype, argtypes); mirrors include/poet_b200.h exactly (tests/test_abi.py parses the header)
SIGNATURES = {
    "poet_version": (_i, []),
    "poet_sm": (_i, []),
    "poet_check_device": (_i, [_i]),
    "poet_error_string": (C.c_char_p, [_i]),
    "poet_posenc_sine": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _i, _i, _vp]),
    "poet_bbox_embed_pad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "poet_nchw_to_tokens": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "poet_tokens_to_nchw": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "poet_mask_prep": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "poet_enc_reference_points": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "poet_msda_fwd": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "poet_msda_bwd": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "poet_gemm_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "poet_gemm": (_i, [_vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "poet_split_bf16": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "poet_split_bf16_multi": (_i, [_vp, _i, _i64, _vp]),
    "poet_gemm_tc_eligible": (_i, [_i, _i, _i, _i64, _i64, _i64]),
    "poet_gemm_bsplit": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp]),
    "poet_gemm_relu_bits_supported": (_i, [_i, _i, _i, _i]),
    "poet_gemm_ex": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i,
                          _vp, _u32, _f, _vp]),
    "poet_dropout": (_i, [_vp, _i64, _vp, _u32, _f, _vp]),
    "poet_dropout_scale": (_f, [_f, _i]),
    "poet_colsum_masked": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _vp]),
    "poet_colsum": (_i, [_vp, _i64, _vp, _i, _i, _i, _vp]),
    "poet_mask_rows": (_i, [_vp, _vp, _i, _i, _vp]),
    "poet_add_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _u32, _f, _vp]),
    "poet_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _u32, _f, _vp]),
    "poet_add": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "poet_mha_smallq_fwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _f, _vp, _u32, _f, _vp]),
    "poet_mha_smallq_bwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64,
                                 _i, _i, _i, _i, _f, _vp, _u32, _f, _vp]),
    "poet_heads_select_rot6d_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "poet_im2col_3x3s2": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "poet_groupnorm_tokens_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "poet_groupnorm_tokens_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "poet_pose_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp]),
    "poet_sumsq": (_i, [_vp, _i64, _vp, _vp]),
    "poet_grad_sumsq_multi": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "poet_adamw_clip_multi": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _f, _f, _f, _f, _i64, _vp]),
    "poet_heads_select_rot6d_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "poet_linear_epilogue_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "poet_linear_epilogue": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _sz, _vp]),
    "poet_ffn_fused_workspace_bytes": (_sz, [_i, _i, _i]),
    "poet_ffn_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _sz, _vp]),
}

_lock = threading.Lock()
_lib = None


class PoetLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise PoetLibraryError(
                        f"{LIB_PATH} is missing: build it with `python -m poet_b200.build` "
                        "(poet_b200 has no CPU or PyTorch fallback)")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)          # AttributeError if the symbol is not exported
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().poet_error_string(rc)
        raise PoetLibraryError(f"{what} failed with status {rc}: {msg.decode() if msg else '?'}")


_device_ok = set()


def require_b200(device_index: int) -> None:
    """The library is built for sm_100a only; refuse anything else instead of mis-executing."""
    if device_index not in _device_ok:
        check(lib().poet_check_device(device_index), f"poet_check_device({device_index})")
        _device_ok.add(device_index)
