"""ctypes binding of libosufusion_sm100.so (the C-ABI declared in include/osufusion_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

from .build import LIB_PATH, build_native

_lib = None


class NativeError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("b_mn_major", C.c_int), ("batch", C.c_int), ("rows", C.c_int),
        ("N", C.c_int), ("K", C.c_int), ("taps", C.c_int), ("shift0", C.c_int), ("shift_step", C.c_int),
        ("a", C.c_void_p), ("a_ld", C.c_longlong), ("a_batch_stride", C.c_longlong),
        ("b", C.c_void_p), ("b_ld", C.c_longlong), ("b_tap_stride", C.c_longlong),
        ("bias", C.c_void_p),
        ("aux_f32", C.c_void_p), ("aux_f32_ld", C.c_longlong), ("aux_f32_batch_stride", C.c_longlong),
        ("aux_bf16", C.c_void_p), ("aux_bf16_ld", C.c_longlong), ("aux_bf16_batch_stride", C.c_longlong),
        ("aux_is_dsilu", C.c_int), ("act", C.c_int),
        ("pre_bf16", C.c_void_p),
        ("out_bf16", C.c_void_p), ("out_bf16_ld", C.c_longlong), ("out_bf16_batch_stride", C.c_longlong),
        ("out_f32", C.c_void_p), ("out_f32_ld", C.c_longlong), ("out_f32_batch_stride", C.c_longlong),
        ("stats", C.c_void_p),
        ("block_n", C.c_int), ("split_k", C.c_int),
    ]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = Path(LIB_PATH)
        if not path.exists():
            build_native()
        if not path.exists():
            raise NativeError(f"{path} is missing: the sm_100a CUDA library must be built (python -m osufusion_b200.build)")
        _lib = C.CDLL(str(path))
        _lib.of_last_error.restype = C.c_char_p
        _lib.of_launch_count.restype = C.c_longlong
        _lib.of_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
        _lib.of_attn_fwd.argtypes = [C.POINTER(AttnArgs), C.c_void_p]
        _lib.of_attn_bwd.argtypes = [C.POINTER(AttnArgs), C.c_void_p]
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what} failed (rc={rc}): {lib().of_last_error().decode()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


class AttnArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("H", C.c_int), ("KVH", C.c_int), ("L", C.c_int), ("D", C.c_int),
        ("scale", C.c_float), ("variant", C.c_int),
        ("q", C.c_void_p), ("q_ld", C.c_longlong), ("q_batch_stride", C.c_longlong),
        ("k", C.c_void_p), ("v", C.c_void_p), ("kv_ld", C.c_longlong), ("kv_batch_stride", C.c_longlong),
        ("out", C.c_void_p), ("out_ld", C.c_longlong), ("out_batch_stride", C.c_longlong),
        ("lse", C.c_void_p),
        ("dout", C.c_void_p), ("dout_ld", C.c_longlong), ("dout_batch_stride", C.c_longlong),
        ("delta", C.c_void_p),
        ("dq", C.c_void_p), ("dq_ld", C.c_longlong), ("dq_batch_stride", C.c_longlong),
        ("dk", C.c_void_p), ("dv", C.c_void_p), ("dkv_ld", C.c_longlong), ("dkv_batch_stride", C.c_longlong),
    ]
