"""ctypes binding of libosufusion_sm100.so (the C-ABI declared in include/osufusion_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

from .build import LIB_PATH, build_native

_lib = None

P, I, LL, F = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class NativeError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("mode", I), ("b_mn_major", I), ("batch", I), ("rows", I),
        ("N", I), ("K", I), ("taps", I), ("shift0", I), ("shift_step", I),
        ("a", P), ("a_ld", LL), ("a_batch_stride", LL),
        ("b", P), ("b_ld", LL), ("b_tap_stride", LL),
        ("bias", P),
        ("aux_f32", P), ("aux_f32_ld", LL), ("aux_f32_batch_stride", LL),
        ("aux_bf16", P), ("aux_bf16_ld", LL), ("aux_bf16_batch_stride", LL),
        ("aux_is_dsilu", I), ("act", I),
        ("pre_bf16", P),
        ("out_bf16", P), ("out_bf16_ld", LL), ("out_bf16_batch_stride", LL),
        ("out_f32", P), ("out_f32_ld", LL), ("out_f32_batch_stride", LL),
        ("stats", P),
        ("block_n", I), ("split_k", I),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("B", I), ("H", I), ("KVH", I), ("L", I), ("D", I),
        ("scale", F), ("variant", I),
        ("q", P), ("q_ld", LL), ("q_batch_stride", LL),
        ("k", P), ("v", P), ("kv_ld", LL), ("kv_batch_stride", LL),
        ("out", P), ("out_ld", LL), ("out_batch_stride", LL),
        ("lse", P),
        ("dout", P), ("dout_ld", LL), ("dout_batch_stride", LL),
        ("delta", P),
        ("dq", P), ("dq_ld", LL), ("dq_batch_stride", LL),
        ("dk", P), ("dv", P), ("dkv_ld", LL), ("dkv_batch_stride", LL),
        ("zero_grads", I),
    ]


class RbArgs(C.Structure):
    _fields_ = [
        ("B", I), ("L", I), ("C", I), ("eps", F), ("mode", I),
        ("y", P), ("y_ld", LL), ("y_bs", LL),
        ("stats", P), ("gamma", P), ("beta", P), ("ss", P),
        ("vec", P), ("vec_bs", LL), ("vec_bias", P),
        ("p", P), ("pooled", P), ("gate", P),
        ("res_f32", P), ("res_f32_ld", LL), ("res_f32_bs", LL),
        ("res_bf16", P), ("res_bf16_ld", LL), ("res_bf16_bs", LL),
        ("out_f32", P), ("out_f32_ld", LL), ("out_f32_bs", LL),
        ("out_bf16", P), ("out_bf16_ld", LL), ("out_bf16_bs", LL),
        ("out_rows", P), ("acc_bc", P),
        ("dout_f32", P), ("dout_f32_ld", LL), ("dout_f32_bs", LL),
        ("dh_bf16", P), ("dh_ld", LL), ("dh_bs", LL),
        ("dpooled", P), ("da", P), ("wk", P),
        ("dstats", P), ("dgamma", P), ("dbeta", P), ("dwk", P), ("dbk", P), ("dss", P),
        ("dxhat_bf16", P), ("dxhat_ld", LL), ("dxhat_bs", LL),
        ("dout_bf16", P), ("dout_bf16_ld", LL), ("dout_bf16_bs", LL),
        ("dy_bf16", P), ("dy_ld", LL), ("dy_bs", LL),
        ("dbias", P),
    ]


class FilmGroup(C.Structure):
    _fields_ = [("W", P), ("bias", P), ("dW", P), ("dbias", P), ("out_off", LL), ("N", I), ("row_start", I)]


class PackSeg(C.Structure):
    _fields_ = [("src", P), ("dst", P), ("Cout", I), ("Cin", I), ("k", I), ("cin_pad", I), ("cta_begin", I), ("scale", F)]


class LoraFinishSeg(C.Structure):
    _fields_ = [("dBraw", P), ("rowscale", P), ("gB", P), ("dm", P), ("mag", P), ("gmag", P), ("Cout", I), ("r", I), ("cta_begin", I), ("_pad", I)]


class OptTensor(C.Structure):
    _fields_ = [("param", P), ("arena_off", LL), ("numel", LL), ("operand_bf16", P), ("cta_begin", I), ("Cout", I), ("Cin", I), ("k", I)]


# name -> argtypes (the trailing stream pointer included); mirrors include/osufusion_b200.h
_SIGS = {
    "of_gemm": [C.POINTER(GemmArgs), P],
    "of_attn_fwd": [C.POINTER(AttnArgs), P],
    "of_attn_bwd": [C.POINTER(AttnArgs), P],
    "of_rb_apply_fwd": [C.POINTER(RbArgs), P],
    "of_rb_rowdot": [C.POINTER(RbArgs), P],
    "of_rb_pool": [C.POINTER(RbArgs), P],
    "of_rb_logit_pool": [C.POINTER(RbArgs), P, P, P],
    "of_rb_gate_fwd": [C.POINTER(RbArgs), P],
    "of_rb_gate_bwd_reduce": [C.POINTER(RbArgs), P],
    "of_rb_bwd_pass1": [C.POINTER(RbArgs), P],
    "of_rb_bwd_apply": [C.POINTER(RbArgs), P],
    "of_softmax_rows": [P, I, I, P],
    "of_softmax_bwd_rows": [P, P, I, I, P],
    "of_layernorm_fwd": [P, LL, I, I, P, P, F, P, P, LL, P, P],
    "of_layernorm_bwd": [P, LL, P, LL, I, I, P, P, P, P, LL, P, P, P],
    "of_rope_fwd": [P, LL, LL, I, I, I, I, I, P, P, I, P],
    "of_rope_bwd": [P, LL, LL, P, P, LL, LL, P, LL, LL, I, I, I, I, I, P, P, I, P],
    "of_linear_small_fwd": [P, LL, I, I, I, P, LL, P, I, I, P, LL, P, P],
    "of_linear_small_bwd": [P, LL, P, I, P, LL, I, I, I, P, LL, I, P, P, P, LL, I, P],
    "of_colsum_bf16": [P, LL, LL, I, P, P],
    "of_coldot_bf16": [P, LL, P, LL, LL, I, P, P, P],
    "of_pack_input": [P, P, P, P, I, I, I, P, I, I, F, P],
    "of_unpack_output": [P, LL, LL, I, I, I, P, P],
    "of_upsample2x_fwd": [P, LL, LL, I, I, I, P, LL, LL, P],
    "of_upsample2x_bwd": [P, LL, LL, I, I, I, P, P, LL, LL, P],
    "of_cast_copy": [P, P, LL, LL, I, I, I, P, P, LL, LL, I, P],
    "of_time_embed": [P, I, I, F, P, P],
    "of_silu_small": [P, P, P, LL, P],
    "of_mse_fwd": [P, LL, LL, P, P, F, F, P, I, I, I, P, P, P],
    "of_mse_bwd": [P, LL, LL, P, P, F, F, P, I, I, I, I, I, P, P, P, P],
    "of_sampler_update": [P, P, P, LL, LL, F, I, F, F, F, F, I, I, I, P, P, I, I, F, P],
    "of_sampler_update_dev": [P, P, P, LL, LL, F, I, P, I, I, I, P, P, I, I, F, P],
    "of_pack_conv_weight": [P, I, I, I, P, I, I, I, P],
    "of_unpack_conv_wgrad": [P, I, I, I, I, I, P, I, I, P],
    "of_cast_f32_bf16": [P, P, LL, P],
    "of_film_fwd": [P, I, I, P, I, I, P, P],
    "of_film_bwd": [P, P, I, P, P, I, I, P, P],
    "of_pack_weights": [P, I, I, P],
    "of_grad_sumsq": [P, LL, P, P],
    "of_adamw_step": [P, I, I, P, P, P, P, F, F, F, F, F, F, I, P],
    "of_dora_rankr_prep": [P, P, P, F, I, I, P, P, P],
    "of_dora_rankr_finish": [P, P, P, P, P, P, I, I, P],
    "of_dora_scale_pack": [P, P, I, I, I, P, P, I, LL, P],
    "of_scale_cast_f32_bf16": [P, F, P, LL, P],
    "of_dora_scale_pack_prep": [P, P, I, I, I, P, P, I, LL, P, F, I, P, P, P],
    "of_lora_finish_all": [P, I, I, P],
    "of_dora_merge": [P, P, P, P, F, I, I, I, I, P, P, I, LL, P, P],
    "of_dora_grad": [P, P, P, P, F, I, I, I, I, P, P, I, LL, P, P, P, P],
    "of_gate_residual_fwd": [P, P, LL, LL, P, LL, LL, P, LL, I, I, I, I, P, LL, LL, P],
    "of_gate_mul_bwd": [P, LL, LL, P, LL, I, I, I, I, P, P, LL, LL, P],
    "of_headnorm_fwd": [P, LL, LL, I, I, I, I, I, I, P, P, F, P, LL, LL, I, P],
    "of_headnorm_bwd": [P, LL, LL, P, P, LL, LL, P, LL, LL, I, I, I, I, I, I, P, P, F, P, LL, LL, P, P, I, P],
    "of_row_mean_std": [P, I, I, I, P, P],
    "of_adaln_fwd": [P, LL, LL, I, I, I, P, LL, P, LL, F, P, LL, LL, P, P],
    "of_adaln_bwd": [P, LL, LL, P, LL, LL, I, I, I, P, LL, P, P, LL, LL, P, LL, LL, P, LL, P, LL, P],
    "of_gate_bwd": [P, LL, LL, P, LL, P, LL, LL, I, I, I, I, P, LL, LL, P, LL, P],
}
_PLAIN = {"of_set_sm_limit": [I], "of_rb_pool_parts": [C.POINTER(RbArgs)], "of_film_chunk_rows": [], "of_pack_seg_ctas": [I, I, I, I], "of_opt_tensor_ctas": [LL], "of_opt_tensor_ctas2": [LL, I, I, I]}    # host-side helpers: no stream argument, return a value
EXPORTS = ["of_last_error", "of_version", "of_launch_count", "of_reset_launch_count", *_SIGS.keys(), *_PLAIN.keys()]


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
        for name, sig in _SIGS.items():
            fn = getattr(_lib, name)
            fn.argtypes = sig
            fn.restype = C.c_int
        for name, sig in _PLAIN.items():
            fn = getattr(_lib, name)
            fn.argtypes = sig
            fn.restype = C.c_int
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what} failed (rc={rc}): {lib().of_last_error().decode()}")


# When set to a list, every entry point is bracketed by CUDA events on the launching stream and
# (family, algorithmic_flops, start_event, end_event) is appended (bench.py roofline section).
PROFILE = None


def call(name: str, *args, flops: float = 0.0, family: str | None = None, tag: str = "") -> None:
    """Invoke an entry point with the current torch CUDA stream appended; raises on a non-zero return."""
    fn = getattr(lib(), name)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args, stream_ptr())
        e1.record()
        PROFILE.append((family or name, flops, e0, e1, tag))
    else:
        rc = fn(*args, stream_ptr())
    if rc != 0:
        raise NativeError(f"{name} failed (rc={rc}): {lib().of_last_error().decode()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
