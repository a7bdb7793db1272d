"""ctypes binding of ``libacan_b200.so`` (the C ABI declared in ``include/acan_b200.h``).

There is no fallback: if the shared library is missing or a call returns non-zero, the caller gets
an exception.  Build with ``make`` (or ``python __graft_entry__.py build``) at the repo root.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libacan_b200.so")

ABI_VERSION = 3


class AcanLibraryError(RuntimeError):
    pass


# name -> (restype, argtypes); mirrors include/acan_b200.h one to one
SIGNATURES = {
    "acan_abi_version": (c_int, []),
    "acan_last_error": (c_char_p, []),
    "acan_device_check": (c_int, []),
    "acan_launch_count": (ctypes.c_longlong, []),
    "acan_profile_begin": (c_int, [c_int]),
    "acan_profile_begin_coarse": (c_int, [c_int]),
    "acan_profile_end": (c_int, [c_void_p, c_void_p]),
    "acan_decode_ord_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    "acan_decode_ord_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    "acan_head_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_double, c_double, c_int, c_int, c_void_p]),
    "acan_tail_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_double, c_int, c_void_p]),
    "acan_ordinal_loss_ws_floats": (c_int, []),
    "acan_ordinal_loss_fwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int64, c_int, c_void_p]),
    "acan_ordinal_loss_bwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int64, c_int, c_void_p]),
    "acan_ce_loss_fwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int64, c_int, c_float, c_int, c_void_p]),
    "acan_ce_loss_bwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int64, c_int, c_float, c_int, c_void_p]),
    "acan_tail_loss_ws_floats": (c_size_t, [c_int, c_int, c_int]),
    "acan_tail_loss_fwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "acan_tail_loss_bwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "acan_bn_ws_floats": (c_size_t, [c_int]),
    "acan_bn_set_sm_limit": (c_int, [c_int]),
    "acan_bn_train_fwd": (c_int, [c_void_p, c_int] + [c_void_p] * 8 + [c_float, c_float, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "acan_bn_train_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int] + [c_void_p] * 8 + [c_int64, c_int, c_void_p]),
    "acan_depth_metrics_ws_floats": (c_int, []),
    "acan_depth_metrics": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "acan_continuous2discrete": (c_int, [c_void_p, c_void_p, c_int64, c_double, c_double, c_int, c_void_p]),
    "acan_pool_branch_fwd": (c_int, [c_void_p] * 9 + [c_int] * 5 + [c_void_p]),
    "acan_pool_nhwc_scratch_floats": (c_size_t, [c_int, c_int, c_int]),
    "acan_pool_branch_fwd_nhwc": (c_int, [c_void_p, c_int] + [c_void_p] * 9 + [c_int] * 5 + [c_void_p]),
    "acan_bias_relu_nhwc": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "acan_maxpool3x3s2_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "acan_attn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "acan_attn_fwd": (c_int, [c_void_p] * 6 + [c_size_t] + [c_int] * 4 + [c_float, c_void_p]),
    "acan_attn_fwd_loss": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p, c_void_p, c_size_t] + [c_int] * 4 + [c_float, c_void_p]),
    "acan_attn_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "acan_attn_bwd": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t] + [c_int] * 4 + [c_float, c_void_p]),
    "acan_attn_fwd_nhwc_f16": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p, c_void_p, c_size_t] + [c_int] * 4 + [c_float, c_void_p]),
    "acan_attn_fwd_nhwc": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t]
                           + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "acan_attn_bwd_nhwc_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "acan_attn_bwd_nhwc": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                   c_void_p, c_int, c_void_p, c_int] + [c_int] * 4 + [c_float, c_void_p]),
    "acan_attn_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p] + [c_int] * 4 + [c_float, c_void_p]),
    "acan_attention_loss": (c_int, [c_void_p] * 5 + [c_float, c_int, c_int, c_int, c_void_p]),
    "acan_gt_sim_map": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "acan_sgd_flat_chunk": (c_int, []),
    "acan_sgd_flat": (c_int, [c_void_p] * 7 + [c_int, c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_void_p]),
    "acan_attention_loss_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 5 + [c_float, c_void_p]),
    "acan_cast_to_f32_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the library once; raise AcanLibraryError when it is absent or its ABI differs."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AcanLibraryError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `make` in the repo root). "
            "acan_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise AcanLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    got = lib.acan_abi_version()
    if got != ABI_VERSION:
        raise AcanLibraryError(f"libacan_b200 ABI {got} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().acan_last_error()
        raise AcanLibraryError(f"{what} failed with code {code}: {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
