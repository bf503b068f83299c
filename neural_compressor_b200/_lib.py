"""ctypes binding of libb200woq.so (include/b200woq.h).  No torch types cross this boundary:
tensors are passed as raw device pointers + sizes + a cudaStream_t.

The library is REQUIRED: importing a product op without it raises, there is no CPU fallback.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200woq.so")

F32, F16, BF16 = 0, 1, 2
_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

_SIGS = {
    "b200woq_version": (c_int, []),
    "b200woq_last_error": (c_char_p, []),
    "b200woq_device_arch": (c_int, [c_char_p, c_int]),
    "b200woq_launch_count": (c_int64, []),
    "b200woq_rtn_params": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_int, c_float,
                                   c_void_p, c_void_p, c_void_p]),
    "b200woq_rtn_quant_pack": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "b200woq_rtn_fake_quant": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_int, c_float,
                                       c_void_p, c_void_p, c_void_p]),
    "b200woq_pack_codes": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "b200woq_pack_params": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "b200woq_unpack": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "b200woq_dequantize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                   c_void_p, c_void_p]),
    "b200woq_f4_quantize": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_float, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "b200woq_pack_rows": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "b200woq_f4_dequantize": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "b200woq_linear_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64, c_int, c_int]),
    "b200woq_linear_forward": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_void_p, c_int64, c_int, c_void_p]),
    "b200woq_stream_layout_bytes": (c_int64, [c_int64, c_int64, c_int, c_int]),
    "b200woq_build_stream_layout": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p,
                                            c_void_p]),
    "b200woq_linear_forward_stream": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int,
                                              c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200woq_hessian_accumulate": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "b200woq_hessian_finalize": (c_int, [c_void_p, c_int64, c_double, c_float, c_void_p, c_void_p, c_void_p]),
    "b200woq_cholinv_workspace_bytes": (c_int64, [c_int64]),
    "b200woq_cholinv_upper": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "b200woq_hessian_finalize_cholinv_upper": (c_int, [c_void_p, c_int64, c_double, c_float, c_void_p, c_void_p, c_void_p,
                                                       c_void_p, c_int64, c_void_p, c_void_p]),
    "b200woq_gptq_workspace_bytes": (c_int64, [c_int64, c_int64, c_int]),
    "b200woq_gptq_fasterquant": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                         c_void_p]),
    "b200woq_gptq_rebuild_q": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "b200woq_awq_weight_scale": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "b200woq_abs_colsum_accumulate": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "b200woq_mse_accumulate": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "b200woq_w8a8_padded_k": (c_int64, [c_int64]),
    "b200woq_sq_smooth_quant_weight": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p]),
    "b200woq_w8a8_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "b200woq_w8a8_workspace_zeroed_offset": (c_int64, [c_int64, c_int64]),
    "b200woq_w8a8_linear_forward": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                            c_int64, c_void_p]),
    "b200woq_minmax_cols_accumulate": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                               c_void_p]),
}

_lib = None


class B200WOQError(RuntimeError):
    pass


def exported_symbols():
    return list(_SIGS)


def load():
    """Load the shared library (no GPU needed to load; calls need one)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200WOQError(
                f"{LIB_PATH} is missing: build it with `python -m neural_compressor_b200._build` "
                "(nvcc, sm_100a). neural_compressor_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError => header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().b200woq_last_error()
        raise B200WOQError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise B200WOQError(f"unsupported dtype {t.dtype}") from None


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name="tensor"):
    if not t.is_cuda:
        raise B200WOQError(f"{name} must be a CUDA tensor (neural_compressor_b200 has no CPU path)")
    if not t.is_contiguous():
        raise B200WOQError(f"{name} must be contiguous")
