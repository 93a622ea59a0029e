"""Raw (non-autograd) Python wrappers over the C-ABI.  Tensors are torch CUDA tensors used purely as device
memory handles; all arithmetic happens in libosufusion_sm100.so."""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N

FWD, WGRAD = 0, 1
ACT_NONE, ACT_SILU = 0, 1


def _bl(t: torch.Tensor):
    """(batch_stride, ld) of a (B, L, C) channels-last view with unit channel stride."""
    assert t.dim() == 3 and t.stride(2) == 1, (t.shape, t.stride())
    return t.stride(0), t.stride(1)


def gemm_fwd(a, b, *, N_out, K, taps=1, shift0=0, shift_step=0, b_mn_major=False, b_ld=None, b_tap_stride=None,
             bias=None, aux_f32=None, aux_bf16=None, aux_is_dsilu=False, act=ACT_NONE, pre_bf16=None, out_bf16=None,
             out_f32=None, stats=None, block_n=0):
    """a: (B, L, >=K) bf16 view; b: packed weight bf16 ([taps][N][K] or [taps][K][N])."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    g = N.GemmArgs()
    g.mode = FWD
    g.b_mn_major = int(b_mn_major)
    g.batch, g.rows = a.shape[0], a.shape[1]
    g.N, g.K, g.taps = N_out, K, taps
    g.shift0, g.shift_step = shift0, shift_step
    g.a = a.data_ptr()
    g.a_batch_stride, g.a_ld = _bl(a)
    g.b = b.data_ptr()
    if b_ld is None:
        b_ld = N_out if b_mn_major else K
    if b_tap_stride is None:
        b_tap_stride = (K * b_ld) if b_mn_major else (N_out * b_ld)
    g.b_ld, g.b_tap_stride = b_ld, b_tap_stride
    g.bias = N.ptr(bias)
    if aux_f32 is not None:
        assert aux_f32.dtype == torch.float32
        g.aux_f32 = aux_f32.data_ptr()
        g.aux_f32_batch_stride, g.aux_f32_ld = _bl(aux_f32)
    if aux_bf16 is not None:
        assert aux_bf16.dtype == torch.bfloat16
        g.aux_bf16 = aux_bf16.data_ptr()
        g.aux_bf16_batch_stride, g.aux_bf16_ld = _bl(aux_bf16)
    g.aux_is_dsilu = int(aux_is_dsilu)
    g.act = act
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16
        g.out_bf16 = out_bf16.data_ptr()
        g.out_bf16_batch_stride, g.out_bf16_ld = _bl(out_bf16)
    if pre_bf16 is not None:
        assert pre_bf16.dtype == torch.bfloat16
        g.pre_bf16 = pre_bf16.data_ptr()
        if out_bf16 is not None:
            assert _bl(pre_bf16) == _bl(out_bf16)
        else:
            g.out_bf16_batch_stride, g.out_bf16_ld = _bl(pre_bf16)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32
        g.out_f32 = out_f32.data_ptr()
        g.out_f32_batch_stride, g.out_f32_ld = _bl(out_f32)
    if stats is not None:
        assert stats.dtype == torch.float64
        g.stats = stats.data_ptr()
    g.block_n = block_n
    N.call("of_gemm", C.byref(g), flops=2.0 * g.batch * g.rows * N_out * K * taps, family="gemm_kernel (tcgen05 GEMM/conv)",
           tag=f"{'wgrad' if g.mode else ('dgrad' if g.b_mn_major else 'fwd')} B{g.batch} L{g.rows} N{g.N} K{g.K} T{g.taps}")


def gemm_wgrad(dy, x, out_f32, *, M, N_out, taps=1, shift0=0, shift_step=0, split_k=0, block_n=0):
    """out_f32[t][m][n] += sum_{b,l} dy[b,l,m] * x[b,l+shift0+t*shift_step,n].  out_f32: (taps, M, >=N) fp32."""
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and out_f32.dtype == torch.float32
    g = N.GemmArgs()
    g.mode = WGRAD
    g.batch, g.rows = dy.shape[0], dy.shape[1]
    assert x.shape[0] == dy.shape[0] and x.shape[1] == dy.shape[1]
    g.N, g.K, g.taps = N_out, M, taps
    g.shift0, g.shift_step = shift0, shift_step
    g.a = dy.data_ptr()
    g.a_batch_stride, g.a_ld = _bl(dy)
    g.b = x.data_ptr()
    g.b_tap_stride, g.b_ld = _bl(x)
    g.out_f32 = out_f32.data_ptr()
    assert out_f32.dim() == 3 and out_f32.stride(2) == 1
    g.out_f32_batch_stride, g.out_f32_ld = out_f32.stride(0), out_f32.stride(1)
    g.split_k = split_k
    g.block_n = block_n
    N.call("of_gemm", C.byref(g), flops=2.0 * g.batch * g.rows * N_out * M * taps, family="gemm_kernel (tcgen05 GEMM/conv)",
           tag=f"{'wgrad' if g.mode else ('dgrad' if g.b_mn_major else 'fwd')} B{g.batch} L{g.rows} N{g.N} K{g.K} T{g.taps}")


def _attn_common(q, k, v, H, KVH, D, variant):
    g = N.AttnArgs()
    g.B, g.L = q.shape[0], q.shape[1]
    g.H, g.KVH, g.D = H, KVH, D
    g.scale = 0.0
    g.variant = variant
    g.q = q.data_ptr()
    g.q_batch_stride, g.q_ld = _bl(q)
    g.k, g.v = k.data_ptr(), v.data_ptr()
    assert _bl(k) == _bl(v)
    g.kv_batch_stride, g.kv_ld = _bl(k)
    return g


def attn_fwd(q, k, v, out, lse, *, H, KVH, D, variant=0):
    """q (B,L,H*D), k/v (B,L,KVH*D) bf16 views; out (B,L,H*D) bf16; lse (B,H,L) fp32."""
    g = _attn_common(q, k, v, H, KVH, D, variant)
    g.out = out.data_ptr()
    g.out_batch_stride, g.out_ld = _bl(out)
    g.lse = N.ptr(lse)
    N.call("of_attn_fwd", C.byref(g), flops=4.0 * g.B * H * g.L * g.L * D, family="attn_fwd_kernel (tcgen05 MQA flash)")


def attn_bwd(q, k, v, out, lse, dout, delta, dq, dk, dv, *, H, KVH, D, variant=0, zero_grads=False):
    """dq / dk / dv are accumulated into; zero_grads=True makes the delta pre-pass zero-fill them (uninitialised buffers are fine)."""
    g = _attn_common(q, k, v, H, KVH, D, variant)
    g.zero_grads = int(zero_grads)
    g.out = out.data_ptr()
    g.out_batch_stride, g.out_ld = _bl(out)
    g.lse = lse.data_ptr()
    g.dout = dout.data_ptr()
    g.dout_batch_stride, g.dout_ld = _bl(dout)
    g.delta = delta.data_ptr()
    g.dq = dq.data_ptr()
    g.dq_batch_stride, g.dq_ld = _bl(dq)
    g.dk, g.dv = dk.data_ptr(), dv.data_ptr()
    assert _bl(dk) == _bl(dv)
    g.dkv_batch_stride, g.dkv_ld = _bl(dk)
    # algorithmic FLOPs: backward counted as 2x forward = 8 B H L^2 D; the S = QK^T recompute (another 2 B H L^2 D
    # executed by the kernel) is not credited
    N.call("of_attn_bwd", C.byref(g), flops=8.0 * g.B * H * g.L * g.L * D, family="attn_bwd_kernel (tcgen05 MQA flash)")
