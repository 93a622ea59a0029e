"""The reference's two transformer backbones on the B200 engine — drop-ins for `osu_fusion.modules.dit.DiT` (dit.py:162-292) and
`osu_fusion.modules.mmdit.MMDiT` (mmdit.py:236-389): same constructor signatures, attribute names and state_dict keys/shapes,
same call `net(x, a, t, c, cond_drop_prob)` / `forward_with_cond_scale`.

Like the UNet (modules.py) the nn.Linear / nn.Conv1d / nn.LayerNorm leaves are PARAMETER CONTAINERS: every FLOP runs in
libosufusion_sm100.so through the engine tape, torch.autograd sees the whole backbone as one node, there is no CPU path.

Kernel mapping (activations channels-last (B, L, C)):
  * q/k/v, attention-out and feed-forward projections, CrossEmbed / patch embedding, output convs   -> of_gemm (tcgen05)
  * (joint) attention incl. grouped-query head mapping `j % kv_heads` (mmdit.py:116-117)             -> of_attn_fwd / of_attn_bwd
  * adaLN `modulate(norm(x), shift, scale)` (dit.py:13-15)   -> of_layernorm_fwd/bwd per sample (gamma = 1 + scale_b, beta = shift_b)
  * `x + gate * branch` and its backward                      -> of_gate_residual_fwd, of_gate_mul_bwd + of_coldot_bf16
  * MultiHeadRMSNorm on q, k (dit.py:62-69)                   -> of_headnorm_fwd/bwd (writes both modality streams into the joint sequence)
  * audio statistics pooling, conditioning MLPs, adaLN heads  -> of_row_mean_std, of_time_embed, of_linear_small_*, of_silu_small

Numerics: the reference under CUDA bf16 autocast — bf16 GEMM / attention operands, fp32 accumulation, fp32 LayerNorm, bf16-rounded
modulation vectors and gate products.  One deliberate difference: the residual stream is kept in fp32 (under autocast the reference's
stream silently becomes bf16 after the first `x + gate * ...`), which is closer to the fp32 truth.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _native as N
from . import engine as E
from . import ops_raw as R
from .engine import BF16, F32, Act, Ctx, ParamStore, Tape, _bl, _p
from .modules import (A_PAD_VALUE, X_PAD_VALUE, CrossEmbedLayer, SinusoidalPositionEmbedding, UNetFunction, _Container, _pack,
                      prob_mask_like)


# adaLN / gate backward as ONE launch over all samples (of_adaln_fwd/bwd, of_gate_bwd) vs the per-sample composition of the UNet
# path's LayerNorm kernels + of_gate_mul_bwd + of_coldot_bf16 (kept as the cross-check: OF_BACKBONE_BATCHED=0)
BATCHED = os.environ.get("OF_BACKBONE_BATCHED", "1") != "0"
# bf16 operand copies of all projection weights by ONE grouped of_pack_weights launch (the U-Net's mechanism, engine.ParamStore) instead of
# one of_cast_f32_bf16 launch per weight; every adaLN head (`modulation[1]` of all blocks + the final layer; their common input is SiLU(c))
# by ONE grouped launch forward and ONE backward (of_film_fwd / of_film_bwd, the U-Net's FiLM-head kernels) instead of one small-M linear
# per head.  Both parity-checked on the B200 (tests/test_zz_backbones_gpu.py with the switches on; profiles/r01_backbones_gpu_tests.log);
# their effect on the step time has not been measured yet (round-1 GPU budget).  =0 restores the per-tensor launches.
GROUPED_PACK = os.environ.get("OF_BACKBONE_GROUPED_PACK", "1") != "0"
GROUPED_MOD = os.environ.get("OF_BACKBONE_GROUPED_MOD", "1") != "0"
# of_headnorm_fwd/bwd kernel variant (see include/osufusion_b200.h): 1 = thread per head vector, 2 = thread per 16-byte vector, 0 = auto
HEADNORM_VARIANT = int(os.environ.get("OF_HEADNORM_VARIANT", "0"))


# ------------------------------------------------------------------------------------------------ parameter containers
class FeedForward(nn.Sequential):  # dit.py:52-59, mmdit.py:34-41
    def __init__(self, dim: int, dim_mult: int = 4) -> None:
        super().__init__(nn.Linear(dim, dim * dim_mult), nn.SiLU(), nn.Linear(dim * dim_mult, dim))


class MultiHeadRMSNorm(_Container):  # dit.py:62-69
    def __init__(self, dim: int, heads: int) -> None:
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(heads, 1, dim))


class FinalLayer(_Container):  # dit.py:72-86, mmdit.py:219-233
    def __init__(self, dim_h: int, dim_out: int) -> None:
        super().__init__()
        self.norm = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim_h, dim_h * 2, bias=True))
        self.linear = nn.Linear(dim_h, dim_out)


class Attend(_Container):
    pass


class DiTAttention(_Container):  # dit.py:89-118
    def __init__(self, dim: int, heads: int, dim_head: int, qk_norm: bool = True, context_len: int = 4096) -> None:
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.to_qkv = nn.Linear(dim, dim_head * heads * 3, bias=False)
        self.q_norm = MultiHeadRMSNorm(dim_head, heads=heads) if qk_norm else nn.Identity()
        self.k_norm = MultiHeadRMSNorm(dim_head, heads=heads) if qk_norm else nn.Identity()
        self.attn = Attend()


class DiTBlock(_Container):  # dit.py:121-159
    def __init__(self, dim_h: int, dim_h_mult: int = 4, attn_heads: int = 8, attn_dim_head: int = 64, attn_qk_norm: bool = True,
                 attn_context_len: int = 4096) -> None:
        super().__init__()
        if attn_heads * attn_dim_head != dim_h:      # the reference fails later, at `x + gate * attn(...)` (no output projection)
            raise ValueError(f"DiTBlock: attn_heads * attn_dim_head ({attn_heads * attn_dim_head}) must equal dim_h ({dim_h})")
        self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim_h, dim_h * 6, bias=True))
        self.norm1 = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.attn = DiTAttention(dim_h, heads=attn_heads, dim_head=attn_dim_head, qk_norm=attn_qk_norm, context_len=attn_context_len)
        self.norm2 = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(dim_h, dim_h_mult)
        self.gradient_checkpointing = False


class PatchEmbedding(_Container):  # mmdit.py:44-52
    def __init__(self, dim_in: int, dim_emb: int, patch_size: int) -> None:
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv1d(dim_in, dim_emb, patch_size, stride=patch_size)


class JointAttention(_Container):  # mmdit.py:65-130
    def __init__(self, dim: int, dim_head: int, heads: int, kv_heads: int, qk_norm: bool = True, context_len: int = 4096) -> None:
        super().__init__()
        self.heads, self.kv_heads, self.dim_head, self.qk_norm = heads, kv_heads, dim_head, qk_norm
        for s in ("x", "a"):
            setattr(self, f"to_q_{s}", nn.Linear(dim, dim_head * heads, bias=False))
            setattr(self, f"to_k_{s}", nn.Linear(dim, dim_head * kv_heads, bias=False))
            setattr(self, f"to_v_{s}", nn.Linear(dim, dim_head * kv_heads, bias=False))
            setattr(self, f"q_{s}_norm", MultiHeadRMSNorm(dim_head, heads) if qk_norm else nn.Identity())
            setattr(self, f"k_{s}_norm", MultiHeadRMSNorm(dim_head, kv_heads) if qk_norm else nn.Identity())
        self.attn = Attend()


class MMDiTBlock(_Container):  # mmdit.py:133-216
    def __init__(self, dim_h: int, dim_h_mult: int = 4, attn_dim_head: int = 64, attn_heads: int = 8, attn_kv_heads: int = 2,
                 attn_qk_norm: bool = True, attn_context_len: int = 4096) -> None:
        super().__init__()
        self.modulation_x = nn.Sequential(nn.SiLU(), nn.Linear(dim_h, dim_h * 6, bias=True))
        self.modulation_a = nn.Sequential(nn.SiLU(), nn.Linear(dim_h, dim_h * 6, bias=True))
        self.norm1_x = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.attn_out_x = nn.Linear(dim_h, dim_h, bias=False)
        self.norm2_x = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.mlp_x = FeedForward(dim_h, dim_mult=dim_h_mult)
        self.norm1_a = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.attn_out_a = nn.Linear(dim_h, dim_h, bias=False)
        self.norm2_a = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.mlp_a = FeedForward(dim_h, dim_mult=dim_h_mult)
        if attn_dim_head * attn_heads != dim_h:      # attn_out_{x,a} are Linear(dim_h, dim_h) applied to the merged heads
            raise ValueError(f"MMDiTBlock: attn_heads * attn_dim_head ({attn_heads * attn_dim_head}) must equal dim_h ({dim_h})")
        self.attn = JointAttention(dim_h, attn_dim_head, attn_heads, attn_kv_heads, qk_norm=attn_qk_norm, context_len=attn_context_len)
        self.gradient_checkpointing = False


# ------------------------------------------------------------------------------------------------ engine helpers
class Vec:
    """(B, n) fp32 conditioning vector + its gradient (small-M side of the model)."""
    __slots__ = ("val", "grad")

    def __init__(self, val: torch.Tensor) -> None:
        self.val, self.grad = val, None


def small_chain(ctx: Ctx, xin: torch.Tensor, layers) -> Vec:
    """y = L_n(...act(L_1(x))) over [(nn.Linear, act)] with M <= 16 rows (act: 0 none, 1 SiLU); bf16 rounding as under autocast."""
    st, dev = ctx.store, ctx.device
    saved = []
    h = xin
    for lin, act in layers:
        y, pre = E.linear_small_fwd(h, lin.weight, lin.bias, act=act, want_pre=(act != 0 and ctx.tape is not None))
        saved.append((h, pre, act, lin))
        h = y
    out = Vec(h)
    if ctx.tape is not None:
        def backward():
            d = out.grad
            out.grad = None
            if d is None:
                return
            for i in range(len(saved) - 1, -1, -1):
                x_i, pre, act, lin = saved[i]
                dx = E.zeros(tuple(x_i.shape), F32, dev) if i > 0 else None
                E.linear_small_bwd_param(st, d, pre, act, x_i, lin.weight, lin.bias, dx)
                d = dx
        ctx.tape.push(backward)
    return out


def ada_ln_fwd(x32: torch.Tensor, shift: torch.Tensor, scale1p: torch.Tensor, eps: float):
    """modulate(LayerNorm_noaffine(x), shift, scale) -> bf16 GEMM operand; per-sample gamma = 1 + scale, beta = shift."""
    B, L, Cc = x32.shape
    bs, ld = _bl(x32)
    h16 = E.empty((B, L, Cc), BF16, x32.device)
    mr = E.empty((B, L, 2), F32, x32.device)
    if BATCHED:
        N.call("of_adaln_fwd", _p(x32), ld, bs, B, L, Cc, _p(scale1p), scale1p.stride(0), _p(shift), shift.stride(0), eps, _p(h16), Cc,
               L * Cc, _p(mr))
        return h16, mr
    for b in range(B):
        N.call("of_layernorm_fwd", x32.data_ptr() + 4 * b * bs, ld, L, Cc, scale1p[b].data_ptr(), shift[b].data_ptr(), eps, None,
               h16[b].data_ptr(), Cc, mr[b].data_ptr())
    return h16, mr


def ada_ln_bwd(dh32: torch.Tensor, x32: torch.Tensor, scale1p: torch.Tensor, mr: torch.Tensor, dshift: torch.Tensor,
               dscale: torch.Tensor, dres: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns dx (+ dres, the stream gradient that bypasses the branch) as a fresh fp32 tensor; accumulates d shift / d scale into the
    given (B, C) row views of the modulation gradient."""
    B, L, Cc = x32.shape
    bs, ld = _bl(x32)
    d_bs, d_ld = _bl(dh32)
    dx = E.empty((B, L, Cc), F32, x32.device)
    if BATCHED:
        r_bs, r_ld = _bl(dres) if dres is not None else (0, 0)
        N.call("of_adaln_bwd", _p(dh32), d_ld, d_bs, _p(x32), ld, bs, B, L, Cc, _p(scale1p), scale1p.stride(0), _p(mr), _p(dres), r_ld, r_bs,
               _p(dx), Cc, L * Cc, _p(dscale), dscale.stride(0), _p(dshift), dshift.stride(0))
        return dx
    for b in range(B):
        N.call("of_layernorm_bwd", dh32.data_ptr() + 4 * b * d_bs, d_ld, x32.data_ptr() + 4 * b * bs, ld, L, Cc, scale1p[b].data_ptr(),
               mr[b].data_ptr(), dx[b].data_ptr(), None, Cc, dscale[b].data_ptr(), dshift[b].data_ptr())
    if dres is not None:
        E.cast_copy(dres, dx, accumulate=True)
    return dx


def one_plus(scale: torch.Tensor) -> torch.Tensor:
    """`1 + scale` on a bf16 tensor (dit.py:15): rounded to bf16, kept as fp32 values."""
    return (1.0 + scale).to(BF16).to(F32).contiguous()


def gate_residual(x: torch.Tensor, y16: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """x (fp32 or bf16 (B,L,C)) + gate[b,:] * y16 -> fp32; gate is a (B, C) row view of the modulation output."""
    B, L, Cc = y16.shape
    out = E.empty((B, L, Cc), F32, y16.device)
    x_bs, x_ld = _bl(x)
    y_bs, y_ld = _bl(y16)
    N.call("of_gate_residual_fwd", _p(x) if x.dtype == F32 else None, _p(x) if x.dtype == BF16 else None, x_ld, x_bs, _p(y16), y_ld, y_bs,
           _p(gate), gate.stride(0), 1, B, L, Cc, _p(out), Cc, L * Cc)
    return out


def gate_backward(d32: torch.Tensor, gate: torch.Tensor, y16: torch.Tensor, dgate: torch.Tensor) -> torch.Tensor:
    """Backward of `gate * y`: returns dy (bf16) and accumulates d gate[b, c] = sum_l d * y into the (B, C) row view `dgate`."""
    B, L, Cc = d32.shape
    dev = d32.device
    d_bs, d_ld = _bl(d32)
    y_bs, y_ld = _bl(y16)
    dy16 = E.empty((B, L, Cc), BF16, dev)
    if BATCHED:
        N.call("of_gate_bwd", _p(d32), d_ld, d_bs, _p(gate), gate.stride(0), _p(y16), y_ld, y_bs, 1, B, L, Cc, _p(dy16), Cc, L * Cc,
               _p(dgate), dgate.stride(0))
        return dy16
    dx16 = E.empty((B, L, Cc), BF16, dev)
    N.call("of_gate_mul_bwd", _p(d32), d_ld, d_bs, _p(gate), gate.stride(0), 1, B, L, Cc, _p(dx16), _p(dy16), Cc, L * Cc)
    for b in range(B):
        N.call("of_coldot_bf16", dx16[b].data_ptr(), Cc, y16[b].data_ptr(), y_ld, L, Cc, None, dgate[b].data_ptr())
    return dy16


def _gamma(norm):
    return None if isinstance(norm, nn.Identity) else norm.gamma


def headnorm_fwd(raw: torch.Tensor, out: torch.Tensor, Hq: int, Hk: int, D: int, qn, kn) -> None:
    """q/k RMS norm of one stream's fused [q | k | v] projection (B, L, (Hq + 2 Hk) D) into `out` (a row view of the joint sequence)."""
    B, L, _ = raw.shape
    if _gamma(qn) is None:
        E.cast_copy(raw, out)
        return
    i_bs, i_ld = _bl(raw)
    o_bs, o_ld = _bl(out)
    N.call("of_headnorm_fwd", _p(raw), i_ld, i_bs, B, L, Hq, Hk, Hk, D, _p(qn.gamma), _p(kn.gamma), float(qn.scale), _p(out), o_ld, o_bs,
           HEADNORM_VARIANT)


def headnorm_bwd(st: ParamStore, dq, dk, dv, raw: torch.Tensor, Hq: int, Hk: int, D: int, qn, kn) -> torch.Tensor:
    B, L, W = raw.shape
    dqkv = E.empty((B, L, W), BF16, raw.device)
    if _gamma(qn) is None:
        E.cast_copy(dq, dqkv[:, :, :Hq * D])
        E.cast_copy(dk, dqkv[:, :, Hq * D:(Hq + Hk) * D])
        E.cast_copy(dv, dqkv[:, :, (Hq + Hk) * D:])
        return dqkv
    dq_bs, dq_ld = _bl(dq)
    dk_bs, dk_ld = _bl(dk)
    assert _bl(dv) == (dk_bs, dk_ld)
    i_bs, i_ld = _bl(raw)
    N.call("of_headnorm_bwd", _p(dq), dq_ld, dq_bs, _p(dk), _p(dv), dk_ld, dk_bs, _p(raw), i_ld, i_bs, B, L, Hq, Hk, Hk, D, _p(qn.gamma),
           _p(kn.gamma), float(qn.scale), _p(dqkv), W, L * W, st.grad(qn.gamma).data_ptr(), st.grad(kn.gamma).data_ptr(), HEADNORM_VARIANT)
    return dqkv


def ff_fwd(ctx: Ctx, ff, h16: torch.Tensor):
    st, dev = ctx.store, ctx.device
    B, L, Cc = h16.shape
    l1, l2 = ff[0], ff[2]
    Ci = l1.weight.shape[0]
    w1, w2 = st.linear_w(l1.weight), st.linear_w(l2.weight)
    u16 = E.empty((B, L, Ci), BF16, dev) if ctx.tape is not None else None
    s16 = E.empty((B, L, Ci), BF16, dev)
    R.gemm_fwd(h16, w1, N_out=Ci, K=Cc, bias=l1.bias, act=R.ACT_SILU, pre_bf16=u16, out_bf16=s16)
    f16 = E.empty((B, L, l2.weight.shape[0]), BF16, dev)
    R.gemm_fwd(s16, w2, N_out=l2.weight.shape[0], K=Ci, bias=l2.bias, out_bf16=f16)
    return f16, (h16, u16, s16, w1, w2)


def ff_bwd(ctx: Ctx, ff, df16: torch.Tensor, saved) -> torch.Tensor:
    """Returns the fp32 gradient of the feed-forward input."""
    st, dev = ctx.store, ctx.device
    h16, u16, s16, w1, w2 = saved
    B, L, Cc = h16.shape
    l1, l2 = ff[0], ff[2]
    Ci = l1.weight.shape[0]
    dU = E.empty((B, L, Ci), BF16, dev)
    R.gemm_fwd(df16, w2, N_out=Ci, K=l2.weight.shape[0], b_mn_major=True, aux_bf16=u16, aux_is_dsilu=True, out_bf16=dU)
    E._wgrad_linear(st, l2.weight, df16, s16)
    E._bias_grad(st, l2.bias, df16)
    hact = Act(None, None)
    E._dgrad_into(hact, dU, w1, N_out=Cc, K=Ci)
    E._wgrad_linear(st, l1.weight, dU, h16)
    E._bias_grad(st, l1.bias, dU)
    return hact.grad


def to_stream(ctx: Ctx, h: Act) -> Act:
    """bf16 embedding output -> fp32 residual stream."""
    x32 = E.empty(tuple(h.bf16.shape), F32, ctx.device)
    E.cast_copy(h.bf16, x32)
    out = Act(x32, None)
    if ctx.tape is not None:
        def backward():
            g = out.grad
            out.grad = None
            if g is not None:
                h.add_grad(g)
        ctx.tape.push(backward)
    return out


def _mod_views(mod: torch.Tensor, Cc: int, n: int):
    return [mod[:, i * Cc:(i + 1) * Cc] for i in range(n)]


def modulation_fwd(ctx: Ctx, lin: nn.Linear, cact: torch.Tensor) -> torch.Tensor:
    """adaLN head `Linear(SiLU(c))` -> (B, n) fp32 with bf16-rounded values (precomputed for all heads when GROUPED_MOD)."""
    if ctx.film is not None:
        return ctx.film[id(lin)]
    return E.linear_small_fwd(cact, lin.weight, lin.bias)[0]


def modulation_grad_buffer(ctx: Ctx, lin: nn.Linear, B: int) -> torch.Tensor:
    """Zero-initialised (B, n) accumulator for d modulation (a slice of the grouped gradient buffer when GROUPED_MOD)."""
    if ctx.film_dss is not None:
        return ctx.film_dss[id(lin)]
    return E.zeros((B, lin.weight.shape[0]), F32, ctx.device)


def modulation_bwd(ctx: Ctx, lin: nn.Linear, dmod: torch.Tensor, cact: torch.Tensor) -> None:
    if ctx.film_dss is not None:
        return      # consumed by the grouped of_film_bwd launch in the conditioning backward
    E.linear_small_bwd_param(ctx.store, dmod, None, 0, cact, lin.weight, lin.bias, ctx.d_emb_act)


# ------------------------------------------------------------------------------------------------ blocks
def dit_block(ctx: Ctx, m: DiTBlock, x: Act, cact: torch.Tensor) -> Act:
    """DiTBlock.forward_body (dit.py:147-153)."""
    st, dev = ctx.store, ctx.device
    B, L, Cc = x.f32.shape
    at = m.attn
    H, D = at.heads, at.dim_head
    HD = H * D
    lin = m.modulation[1]
    mod = modulation_fwd(ctx, lin, cact)
    sh1, sc1, g1, sh2, sc2, g2 = _mod_views(mod, Cc, 6)
    s1p1, s1p2 = one_plus(sc1), one_plus(sc2)
    x0 = x.f32
    h1, mr1 = ada_ln_fwd(x0, sh1, s1p1, m.norm1.eps)
    wqkv = st.linear_w(at.to_qkv.weight)
    raw = E.empty((B, L, 3 * HD), BF16, dev)
    R.gemm_fwd(h1, wqkv, N_out=3 * HD, K=Cc, out_bf16=raw)
    qkv = E.empty((B, L, 3 * HD), BF16, dev)
    headnorm_fwd(raw, qkv, H, H, D, at.q_norm, at.k_norm)
    q, k, v = qkv[:, :, :HD], qkv[:, :, HD:2 * HD], qkv[:, :, 2 * HD:]
    o16 = E.empty((B, L, HD), BF16, dev)
    lse = E.empty((B, H, L), F32, dev)
    R.attn_fwd(q, k, v, o16, lse, H=H, KVH=H, D=D, variant=ctx.attn_variant)
    x1 = gate_residual(x0, o16, g1)
    h2, mr2 = ada_ln_fwd(x1, sh2, s1p2, m.norm2.eps)
    f16, ff_saved = ff_fwd(ctx, m.ff, h2)
    x2 = gate_residual(x1, f16, g2)
    out = Act(x2, None)

    if ctx.tape is not None:
        def backward():
            d2 = out.grad
            out.grad = None
            dmod = modulation_grad_buffer(ctx, lin, B)
            dsh1, dsc1, dg1, dsh2, dsc2, dg2 = _mod_views(dmod, Cc, 6)
            # feed-forward branch
            df16 = gate_backward(d2, g2, f16, dg2)
            dh2 = ff_bwd(ctx, m.ff, df16, ff_saved)
            d1 = ada_ln_bwd(dh2, x1, s1p2, mr2, dsh2, dsc2, dres=d2)
            # attention branch
            dO = gate_backward(d1, g1, o16, dg1)
            delta = E.empty((B, H, L), F32, dev)
            dq = E.zeros((B, L, HD), F32, dev)
            dkv = E.zeros((B, L, 2 * HD), F32, dev)
            R.attn_bwd(q, k, v, o16, lse, dO, delta, dq, dkv[:, :, :HD], dkv[:, :, HD:], H=H, KVH=H, D=D)
            dqkv = headnorm_bwd(st, dq, dkv[:, :, :HD], dkv[:, :, HD:], raw, H, H, D, at.q_norm, at.k_norm)
            hact = Act(None, None)
            E._dgrad_into(hact, dqkv, wqkv, N_out=Cc, K=3 * HD)
            E._wgrad_linear(st, at.to_qkv.weight, dqkv, h1)
            d0 = ada_ln_bwd(hact.grad, x0, s1p1, mr1, dsh1, dsc1, dres=d1)
            x.add_grad(d0)
            modulation_bwd(ctx, lin, dmod, cact)
        ctx.tape.push(backward)
    return out


def mmdit_block(ctx: Ctx, m: MMDiTBlock, x: Act, a: Act, cact: torch.Tensor, last: bool):
    """MMDiTBlock.forward_body (mmdit.py:174-206).  `last`: the audio stream's output of the final block has no consumer, so
    (as in the reference's autograd graph) its attention-out / feed-forward half gets no gradient."""
    st, dev = ctx.store, ctx.device
    B, Lx, Cc = x.f32.shape
    La = a.f32.shape[1]
    Lt = La + Lx
    at = m.attn
    H, KVH, D = at.heads, at.kv_heads, at.dim_head
    HD, KD = H * D, KVH * D
    W = HD + 2 * KD
    joint = E.empty((B, Lt, W), BF16, dev)
    S = {}
    for s, src, r0, r1 in (("a", a, 0, La), ("x", x, La, Lt)):
        lin = getattr(m, f"modulation_{s}")[1]
        mod = modulation_fwd(ctx, lin, cact)
        sh1, sc1, g1, sh2, sc2, g2 = _mod_views(mod, Cc, 6)
        d = dict(lin=lin, mod=mod, sh1=sh1, g1=g1, sh2=sh2, g2=g2, s1p1=one_plus(sc1), s1p2=one_plus(sc2), r0=r0, r1=r1, src=src,
                 x0=src.f32, qn=getattr(at, f"q_{s}_norm"), kn=getattr(at, f"k_{s}_norm"),
                 projs=(getattr(at, f"to_q_{s}"), getattr(at, f"to_k_{s}"), getattr(at, f"to_v_{s}")),
                 out_lin=getattr(m, f"attn_out_{s}"), ff=getattr(m, f"mlp_{s}"),
                 eps1=getattr(m, f"norm1_{s}").eps, eps2=getattr(m, f"norm2_{s}").eps)
        d["h1"], d["mr1"] = ada_ln_fwd(d["x0"], sh1, d["s1p1"], d["eps1"])
        d["w"] = st.linear_w(*[p.weight for p in d["projs"]])
        L = r1 - r0
        d["raw"] = E.empty((B, L, W), BF16, dev)
        R.gemm_fwd(d["h1"], d["w"], N_out=W, K=Cc, out_bf16=d["raw"])
        headnorm_fwd(d["raw"], joint[:, r0:r1], H, KVH, D, d["qn"], d["kn"])
        S[s] = d
    q, k, v = joint[:, :, :HD], joint[:, :, HD:HD + KD], joint[:, :, HD + KD:]
    o16 = E.empty((B, Lt, HD), BF16, dev)
    lse = E.empty((B, H, Lt), F32, dev)
    R.attn_fwd(q, k, v, o16, lse, H=H, KVH=KVH, D=D, variant=ctx.attn_variant)
    outs = {}
    for s in ("a", "x"):
        d = S[s]
        L = d["r1"] - d["r0"]
        d["o"] = o16[:, d["r0"]:d["r1"]]
        d["wout"] = st.linear_w(d["out_lin"].weight)
        d["y"] = E.empty((B, L, Cc), BF16, dev)
        R.gemm_fwd(d["o"], d["wout"], N_out=Cc, K=HD, out_bf16=d["y"])
        d["x1"] = gate_residual(d["x0"], d["y"], d["g1"])
        d["h2"], d["mr2"] = ada_ln_fwd(d["x1"], d["sh2"], d["s1p2"], d["eps2"])
        d["f"], d["ff_saved"] = ff_fwd(ctx, d["ff"], d["h2"])
        outs[s] = Act(gate_residual(d["x1"], d["f"], d["g2"]), None)

    if ctx.tape is not None:
        def backward():
            dO = E.zeros((B, Lt, HD), BF16, dev)
            for s in ("x", "a"):
                d = S[s]
                d2 = outs[s].grad
                outs[s].grad = None
                d["dmod"] = modulation_grad_buffer(ctx, d["lin"], B)
                dsh1, dsc1, dg1, dsh2, dsc2, dg2 = _mod_views(d["dmod"], Cc, 6)
                d["d1"] = None
                if d2 is None:
                    assert last and s == "a", "only the audio stream of the final block may lack a gradient"
                    continue
                df = gate_backward(d2, d["g2"], d["f"], dg2)
                dh2 = ff_bwd(ctx, d["ff"], df, d["ff_saved"])
                d1 = ada_ln_bwd(dh2, d["x1"], d["s1p2"], d["mr2"], dsh2, dsc2, dres=d2)
                dy = gate_backward(d1, d["g1"], d["y"], dg1)
                R.gemm_fwd(dy, d["wout"], N_out=HD, K=Cc, b_mn_major=True, out_bf16=dO[:, d["r0"]:d["r1"]])
                E._wgrad_linear(st, d["out_lin"].weight, dy, d["o"])
                d["d1"] = d1
            delta = E.empty((B, H, Lt), F32, dev)
            dq = E.zeros((B, Lt, HD), F32, dev)
            dkv = E.zeros((B, Lt, 2 * KD), F32, dev)
            R.attn_bwd(q, k, v, o16, lse, dO, delta, dq, dkv[:, :, :KD], dkv[:, :, KD:], H=H, KVH=KVH, D=D)
            for s in ("x", "a"):
                d = S[s]
                r0, r1 = d["r0"], d["r1"]
                dsh1, dsc1 = _mod_views(d["dmod"], Cc, 2)
                dqkv = headnorm_bwd(st, dq[:, r0:r1], dkv[:, r0:r1, :KD], dkv[:, r0:r1, KD:], d["raw"], H, KVH, D, d["qn"], d["kn"])
                hact = Act(None, None)
                E._dgrad_into(hact, dqkv, d["w"], N_out=Cc, K=W)
                c0 = 0
                for p in d["projs"]:
                    n_ = p.weight.shape[0]
                    E._wgrad_linear(st, p.weight, dqkv[:, :, c0:c0 + n_], d["h1"])
                    c0 += n_
                d0 = ada_ln_bwd(hact.grad, d["x0"], d["s1p1"], d["mr1"], dsh1, dsc1, dres=d["d1"])
                d["src"].add_grad(d0)
                modulation_bwd(ctx, d["lin"], d["dmod"], cact)
        ctx.tape.push(backward)
    return outs["x"], outs["a"]


def final_layer(ctx: Ctx, m: FinalLayer, x: Act, cact: torch.Tensor) -> Act:
    """FinalLayer.forward (dit.py:82-86, mmdit.py:229-233) -> bf16 (B, L, dim_out)."""
    st, dev = ctx.store, ctx.device
    B, L, Cc = x.f32.shape
    lin = m.modulation[1]
    mod = modulation_fwd(ctx, lin, cact)
    sh, sc = _mod_views(mod, Cc, 2)
    s1p = one_plus(sc)
    h16, mr = ada_ln_fwd(x.f32, sh, s1p, m.norm.eps)
    w = st.linear_w(m.linear.weight)
    No = m.linear.weight.shape[0]
    y = E.empty((B, L, No), BF16, dev)
    R.gemm_fwd(h16, w, N_out=No, K=Cc, bias=m.linear.bias, out_bf16=y)
    out = Act(None, y)

    if ctx.tape is not None:
        def backward():
            dy16 = out.grad          # bf16 (B, L, dim_out), produced by the output conv's dgrad
            out.grad = None
            dmod = modulation_grad_buffer(ctx, lin, B)
            dsh, dsc = _mod_views(dmod, Cc, 2)
            E._wgrad_linear(st, m.linear.weight, dy16, h16)
            E._bias_grad(st, m.linear.bias, dy16)
            hact = Act(None, None)
            E._dgrad_into(hact, dy16, w, N_out=Cc, K=No)
            x.add_grad(ada_ln_bwd(hact.grad, x.f32, s1p, mr, dsh, dsc))
            modulation_bwd(ctx, lin, dmod, cact)
        ctx.tape.push(backward)
    return out


# ------------------------------------------------------------------------------------------------ backbones
class _Backbone(nn.Module):
    """Shared host side of DiT / MMDiT: conditioning vector, output conv, autograd plumbing (one node: modules.UNetFunction)."""

    def _init_engine(self) -> None:
        self._store = ParamStore()
        self.attn_variant = 0
        self.grad_sync = None
        self.grad_finish = None
        self.grad_prescale = 1.0
        self._mod_plans = {}     # adaLN-head descriptor tables per (with gradients?) — device tensors, not part of the state_dict

    # ---- reference API
    def set_gradient_checkpointing(self, value: bool) -> None:
        """dit.py:252-256.  Accepted for API compatibility; nothing is recomputed (activations fit in 180 GB)."""
        for _, m in self.named_modules():
            if hasattr(m, "gradient_checkpointing"):
                m.gradient_checkpointing = value

    def forward_with_cond_scale(self, *args, cond_scale: float = 1.0, **kwargs) -> torch.Tensor:
        """dit.py:258-265 / mmdit.py:335-342.  At inference the conditional and the null pass run as ONE forward over the batch
        [cond ; null] (keep mask 1..1 0..0): every op of the backbones is per-sample, so the result equals two separate passes."""
        if cond_scale != 1.0 and not kwargs and len(args) == 4 and not torch.is_grad_enabled() and 2 * args[0].shape[0] <= 16:
            x, a, t, c = args
            if not x.is_cuda:
                raise RuntimeError(f"osufusion_b200.{type(self).__name__} runs only on CUDA (sm_100a); there is no CPU path")
            return self._cfg_batched(x, a, t, c, cond_scale)
        cond = self(*args, **kwargs)
        if cond_scale == 1.0:
            return cond
        null = self(*args, **kwargs, cond_drop_prob=1.0)
        return null + (cond - null) * cond_scale

    def _cfg_batched(self, x, a, t, c, cond_scale: float) -> torch.Tensor:
        B = x.shape[0]
        keep = torch.cat([torch.ones(B, dtype=torch.bool, device=x.device), torch.zeros(B, dtype=torch.bool, device=x.device)])
        out16, _ = self.run(None, torch.cat([x, x]), torch.cat([a, a]), torch.cat([t, t]), torch.cat([c, c]), keep)
        y = self.unpack(out16, x.shape[-1])
        cond, null = y[:B], y[B:]
        return null + (cond - null) * cond_scale

    def forward(self, x, a, t, c, cond_drop_prob: float = 0.0, cond_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError(f"osufusion_b200.{type(self).__name__} runs only on CUDA (sm_100a); there is no CPU path")
        if x.shape[0] > 16:
            raise ValueError("batch size > 16 per call is not supported by the small-M conditioning kernels")
        if cond_mask is None:
            cond_mask = prob_mask_like((x.shape[0],), 1.0 - cond_drop_prob, x.device)
        params = list(self.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return UNetFunction.apply(self, x, a, t, c, cond_mask, *params)
        out16, _ = self.run(None, x, a, t, c, cond_mask)
        return self.unpack(out16, x.shape[-1])

    # ---- engine plumbing (the protocol modules.UNetFunction drives)
    def backward_param_order(self) -> List[nn.Parameter]:
        """Gradient-arena order ~ the order gradients become final in backward: output head first, embeddings last."""
        return list(reversed(list(self.parameters())))

    def unpack(self, out16: torch.Tensor, n: int) -> torch.Tensor:
        B = out16.shape[0]
        y = torch.empty((B, self.dim_in_x, n), dtype=F32, device=out16.device)
        bs, ld = _bl(out16)
        N.call("of_unpack_output", out16.data_ptr(), ld, bs, B, self.dim_in_x, n, y.data_ptr())
        return y

    def _begin(self, tape: Optional[Tape], device) -> Ctx:
        ctx = Ctx(device, self._store, tape)
        ctx.attn_variant = self.attn_variant
        self._store.begin_forward(tape is not None, self if GROUPED_PACK else None)
        if tape is not None:
            self._store.ensure_arena(self)
        return ctx

    def _conditioning(self, ctx: Ctx, a: torch.Tensor, t: torch.Tensor, c: torch.Tensor, keep: torch.Tensor, time_layers, cond_layers,
                      audio_layers, theta: float) -> torch.Tensor:
        """c = where(keep, mlp_cond(c), null_cond) + mlp_time(t) + mlp_audio(feature_extractor_a([mean_n a, std_n a]))
        (dit.py:278-286, mmdit.py:352-376); returns SiLU(c), the shared input of every adaLN head."""
        st, dev = ctx.store, ctx.device
        B, Ca, n = a.shape
        dim = self.null_cond.shape[0]
        a32 = a.contiguous().float()
        stats = E.empty((B, 2 * Ca), F32, dev)
        N.call("of_row_mean_std", a32.data_ptr(), B, Ca, n, stats.data_ptr())
        temb = E.empty((B, dim), F32, dev)
        tf = t.to(F32).contiguous()
        N.call("of_time_embed", tf.data_ptr(), B, dim, float(theta), temb.data_ptr())
        tv = small_chain(ctx, temb, time_layers)
        cv = small_chain(ctx, c.to(F32).contiguous(), cond_layers)
        av = small_chain(ctx, stats, audio_layers)
        csel = torch.where(keep[:, None], cv.val, self.null_cond.detach()[None, :].to(F32))
        cfull = (csel + tv.val + av.val).contiguous()
        cact = E.empty((B, dim), F32, dev)
        N.call("of_silu_small", cfull.data_ptr(), None, cact.data_ptr(), cfull.numel())
        plan = self._modulation_plan(st, B, ctx.tape is not None) if GROUPED_MOD else None
        emb_r = None
        if plan is not None:
            emb_r = cact.to(BF16).to(F32)          # the heads' Linear sees its input in bf16 under autocast
            mod_all = E.empty((B * plan["rows"],), F32, dev)
            N.call("of_film_fwd", plan["groups"].data_ptr(), plan["num_groups"], plan["rows"], emb_r.data_ptr(), B, dim, mod_all.data_ptr())
            ctx.film = {k: mod_all[o:o + B * n].view(B, n) for k, (o, n) in plan["slices"].items()}
        if ctx.tape is not None:
            ctx.d_emb_act = E.zeros((B, dim), F32, dev)
            dmod_all = None
            if plan is not None:
                dmod_all = torch.zeros((B * plan["rows"],), dtype=F32, device=dev)
                ctx.film_dss = {k: dmod_all[o:o + B * n].view(B, n) for k, (o, n) in plan["slices"].items()}

            def backward():
                if plan is not None:
                    N.call("of_film_bwd", plan["groups"].data_ptr(), plan["chunks"].data_ptr(), plan["num_chunks"], dmod_all.data_ptr(),
                           emb_r.data_ptr(), B, dim, ctx.d_emb_act.data_ptr())
                    for lin in plan["heads"]:
                        for p_ in lin.parameters():
                            if p_.requires_grad:
                                st.touch(p_)
                dc = E.empty((B, dim), F32, dev)
                N.call("of_silu_small", cfull.data_ptr(), ctx.d_emb_act.data_ptr(), dc.data_ptr(), cfull.numel())
                if self.null_cond.requires_grad:
                    st.set_grad(self.null_cond, (dc * (~keep)[:, None]).sum(0))
                cv.grad = (dc * keep[:, None]).contiguous()
                tv.grad = dc
                av.grad = dc
            ctx.tape.push(backward)
        return cact

    def _modulation_plan(self, st: ParamStore, B: int, with_grads: bool):
        """Device-resident descriptor tables (of_film_group) of every adaLN head for batch size B; rebuilt when a pointer moves."""
        heads = [m.modulation[1] for m in self.modules() if isinstance(getattr(m, "modulation", None), nn.Sequential)]
        for blk in self.blocks:
            heads += [getattr(blk, f"modulation_{s_}")[1] for s_ in ("a", "x") if hasattr(blk, f"modulation_{s_}")]
        ptrs = tuple(h.weight.data_ptr() for h in heads) + (st.arena.data_ptr() if (with_grads and st.arena is not None) else 0, B)
        cache = self._mod_plans
        plan = cache.get(with_grads)
        if plan is not None and plan["ptrs"] == ptrs:
            return plan
        dev = heads[0].weight.device
        groups, chunks, slices, row = [], [], {}, 0
        ch = N.lib().of_film_chunk_rows()
        for gi, lin in enumerate(heads):
            Nn = lin.weight.shape[0]
            assert Nn % 4 == 0 and lin.weight.shape[1] == heads[0].weight.shape[1]
            dW = st.arena_views[id(lin.weight)].data_ptr() if (with_grads and lin.weight.requires_grad) else 0
            db = st.arena_views[id(lin.bias)].data_ptr() if (with_grads and lin.bias is not None and lin.bias.requires_grad) else 0
            groups.append(N.FilmGroup(lin.weight.data_ptr(), lin.bias.data_ptr() if lin.bias is not None else 0, dW, db, B * row, Nn, row))
            for n0 in range(0, Nn, ch):
                chunks += [gi, n0]
            slices[id(lin)] = (B * row, Nn)
            row += Nn
        arr = (N.FilmGroup * len(groups))(*groups)
        plan = {"ptrs": ptrs, "heads": heads, "rows": row, "num_groups": len(groups), "slices": slices,
                "groups": torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev),
                "chunks": torch.tensor(chunks, dtype=torch.int32).to(dev), "num_chunks": len(chunks) // 2}
        cache[with_grads] = plan
        return plan

    def _output_conv(self, ctx: Ctx, conv: nn.Conv1d, y16: torch.Tensor) -> torch.Tensor:
        """1x1 conv dim_h -> dim_in_x on a (B, L, dim_h) bf16 view, padded to 8 output channels."""
        st, dev = ctx.store, ctx.device
        B, L, Cc = y16.shape
        wf = st.padded_rows_w(conv.weight, 8)
        bf = None
        if conv.bias is not None:
            bf = torch.zeros(8, dtype=F32, device=dev)
            bf[:self.dim_in_x] = conv.bias.detach()
        out16 = E.empty((B, L, 8), BF16, dev)
        R.gemm_fwd(y16, wf.view(1, 8, Cc), N_out=8, K=Cc, bias=bf, out_bf16=out16)
        return out16

    def _output_conv_backward(self, ctx: Ctx, conv: nn.Conv1d, y16: torch.Tensor, dY16: torch.Tensor) -> torch.Tensor:
        """Parameter gradients of the output conv; returns d y as bf16 (B, L, dim_h)."""
        st, dev = ctx.store, ctx.device
        Cc = y16.shape[2]
        if conv.weight.requires_grad:
            tmp = E.zeros((1, 8, Cc), F32, dev)
            R.gemm_wgrad(dY16, y16, tmp, M=8, N_out=Cc)
            st.set_grad(conv.weight, tmp[0, :self.dim_in_x].reshape(conv.weight.shape).clone())
        if conv.bias is not None and conv.bias.requires_grad:
            db = E.zeros((8,), F32, dev)
            E.colsum(dY16, db)
            st.set_grad(conv.bias, db[:self.dim_in_x].clone())
        wf = st.padded_rows_w(conv.weight, 8)
        hold = Act(None, None)
        return E._dgrad_into(hold, dY16, wf.view(1, 8, Cc), N_out=Cc, K=8, want_bf16=True, want_f32=False)

    def backward_from(self, ctx: Ctx, xf, dY16: torch.Tensor, params):
        st = ctx.store
        E.use_pool(ctx.zpool)
        st.begin_backward(self, film_overwritten=False)
        if self.grad_sync is not None:
            self.grad_sync(len(ctx.tape.ops) + 1)
        self.final_backward(ctx, xf, dY16)
        ctx.tape.run_backward(self.grad_sync)
        if self.grad_finish is not None:
            self.grad_finish()
        return st.take_grads(params)


class DiT(_Backbone):
    """Drop-in for osu_fusion.modules.dit.DiT (dit.py:162-292)."""

    def __init__(self, dim_in_x: int, dim_in_a: int, dim_in_c: int, dim_h: int, dim_h_mult: int = 4, depth: int = 12,
                 cross_embed_kernel_sizes: Sequence[int] = (3, 7, 15), attn_heads: int = 8, attn_dim_head: int = 64,
                 attn_qk_norm: bool = True, attn_context_len: int = 4096) -> None:
        super().__init__()
        self.dim_in_x, self.dim_in_a = dim_in_x, dim_in_a
        self.preprocess = CrossEmbedLayer(dim_in_x + dim_in_a, dim_h, cross_embed_kernel_sizes)
        self.postprocess = nn.Conv1d(dim_h, dim_in_x, 1, bias=False)
        self.mlp_time = nn.Sequential(SinusoidalPositionEmbedding(dim_h), nn.Linear(dim_h, dim_h, bias=False), nn.SiLU(),
                                      nn.Linear(dim_h, dim_h, bias=False))
        self.mlp_cond = nn.Sequential(nn.Linear(dim_in_c, dim_h), nn.SiLU(), nn.Linear(dim_h, dim_h))
        self.null_cond = nn.Parameter(torch.randn(dim_h))
        self.feature_extractor_a = nn.Linear(dim_in_a * 2, dim_h)
        self.mlp_audio = nn.Sequential(nn.Linear(dim_h, dim_h), nn.SiLU(), nn.Linear(dim_h, dim_h))
        self.blocks = nn.ModuleList([DiTBlock(dim_h, dim_h_mult, attn_heads, attn_dim_head, attn_qk_norm, attn_context_len)
                                     for _ in range(depth)])
        self.final = FinalLayer(dim_h, dim_h)
        self.initialize_weights()
        self._init_engine()
        self._perm_cache = {}

    def initialize_weights(self) -> None:
        """dit.py:223-250."""
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv1d)):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        for lin in (self.mlp_time[1], self.mlp_time[3], self.mlp_cond[0], self.mlp_cond[2], self.mlp_audio[0], self.mlp_audio[2]):
            nn.init.normal_(lin.weight, std=0.02)
        for head in [b.modulation[1] for b in self.blocks] + [self.final.modulation[1]]:
            nn.init.zeros_(head.weight)
            nn.init.zeros_(head.bias)
        nn.init.zeros_(self.postprocess.weight)

    # wide input buffer: [a (dim_in_a) | x (dim_in_x) | zero pad] so that both slices start 16-byte aligned; the CrossEmbed weights'
    # input channels are permuted accordingly (the reference concatenates [x, a], dit.py:275)
    def _perm(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._perm_cache:     # built once per device (no host-to-device copy inside a captured step)
            self._perm_cache[key] = torch.cat([torch.arange(self.dim_in_x, self.dim_in_x + self.dim_in_a),
                                               torch.arange(self.dim_in_x)]).to(device)
        return self._perm_cache[key]

    def _cross_embed(self, ctx: Ctx, wide: torch.Tensor) -> Act:
        st, dev = ctx.store, ctx.device
        convs = self.preprocess.convs
        ws = tuple(c.weight for c in convs)
        B, L, Cp = wide.shape
        Cin = ws[0].shape[1]
        perm = self._perm(dev)
        kmax = max(w.shape[2] for w in ws)
        Cout = sum(w.shape[0] for w in ws)

        def build():
            out = E.zeros((kmax, Cout, Cp), F32, dev)
            r = 0
            for w in ws:
                k = w.shape[2]
                o = kmax // 2 - k // 2
                out[o:o + k, r:r + w.shape[0], :Cin] = w.detach()[:, perm, :].permute(2, 0, 1)
                r += w.shape[0]
            return out.to(BF16)
        w = st._cached(("crossperm",) + tuple(id(p) for p in ws), ws, build)
        bias = torch.cat([c.bias.detach() for c in convs])
        y = E.empty((B, L, Cout), BF16, dev)
        R.gemm_fwd(wide, w, N_out=Cout, K=Cp, taps=kmax, shift0=-(kmax // 2), shift_step=1, bias=bias, out_bf16=y)
        out = Act(None, y)
        if ctx.tape is not None:
            def backward():
                dy = out.grad
                out.grad = None
                if dy is None:
                    return
                dy16 = E.empty(tuple(y.shape), BF16, dev)
                E.cast_copy(dy, dy16)
                tmp = E.zeros((kmax, Cout, Cp), F32, dev)
                R.gemm_wgrad(dy16, wide, tmp, M=Cout, N_out=Cp, taps=kmax, shift0=-(kmax // 2), shift_step=1)
                db = E.zeros((Cout,), F32, dev)
                E.colsum(dy16, db)
                r = 0
                for c in convs:
                    co, ci, k = c.weight.shape
                    o = kmax // 2 - k // 2
                    if c.weight.requires_grad:
                        g = torch.empty((co, ci, k), dtype=F32, device=dev)
                        g[:, perm, :] = tmp[o:o + k, r:r + co, :ci].permute(1, 2, 0)
                        st.set_grad(c.weight, g)
                    if c.bias is not None and c.bias.requires_grad:
                        st.set_grad(c.bias, db[r:r + co].clone())
                    r += co
            ctx.tape.push(backward)
        return out

    def run(self, tape: Optional[Tape], x, a, t, c, keep):
        """DiT.forward (dit.py:267-292).  Returns (out16 (B, n, 8) bf16, (ctx, final-layer output Act))."""
        assert x.shape[-1] == a.shape[-1] and x.shape[1] == self.dim_in_x and a.shape[1] == self.dim_in_a
        n = x.shape[-1]
        B = x.shape[0]
        ctx = self._begin(tape, x.device)
        dev = ctx.device
        Ca, Cx = self.dim_in_a, self.dim_in_x
        assert Ca % 8 == 0 and Cx <= 8
        a16 = _pack(a, Ca, n, A_PAD_VALUE)
        x8 = _pack(x, 8, n, X_PAD_VALUE)
        wide = E.empty((B, n, Ca + 8), BF16, dev)
        E.cast_copy(a16, wide[:, :, :Ca])
        E.cast_copy(x8, wide[:, :, Ca:])
        cact = self._conditioning(ctx, a, t, c, keep, [(self.mlp_time[1], 1), (self.mlp_time[3], 0)],
                                  [(self.mlp_cond[0], 1), (self.mlp_cond[2], 0)],
                                  [(self.feature_extractor_a, 0), (self.mlp_audio[0], 1), (self.mlp_audio[2], 0)], self.mlp_time[0].theta)
        h = to_stream(ctx, self._cross_embed(ctx, wide))
        for blk in self.blocks:
            h = dit_block(ctx, blk, h, cact)
        yf = final_layer(ctx, self.final, h, cact)
        out16 = self._output_conv(ctx, self.postprocess, yf.bf16)
        return out16, (ctx, yf)

    def final_backward(self, ctx: Ctx, yf: Act, dY16: torch.Tensor) -> None:
        yf.grad = self._output_conv_backward(ctx, self.postprocess, yf.bf16, dY16)


class MMDiT(_Backbone):
    """Drop-in for osu_fusion.modules.mmdit.MMDiT (mmdit.py:236-389)."""

    def __init__(self, dim_in_x: int, dim_in_a: int, dim_in_c: int, dim_h: int, dim_h_mult: int = 4, patch_size: int = 4,
                 depth: int = 12, attn_dim_head: int = 64, attn_heads: int = 8, attn_kv_heads: int = 2, attn_qk_norm: bool = True,
                 attn_context_len: int = 4096) -> None:
        super().__init__()
        self.dim_h, self.dim_in_x, self.dim_in_a, self.patch_size = dim_h, dim_in_x, dim_in_a, patch_size
        self.attn_context_len = (attn_context_len // patch_size) * 2
        self.emb_x = PatchEmbedding(dim_in_x, dim_h, patch_size)
        self.emb_a = PatchEmbedding(dim_in_a, dim_h, patch_size)
        self.feature_extractor_a = nn.Linear(dim_in_a * 2, dim_h)
        self.mlp_a = FeedForward(dim_h, dim_mult=dim_h_mult)
        self.mlp_time = nn.Sequential(SinusoidalPositionEmbedding(dim_h), FeedForward(dim_h, dim_mult=dim_h_mult))
        self.mlp_cond = nn.Sequential(nn.Linear(dim_in_c, dim_h), FeedForward(dim_h, dim_mult=dim_h_mult))
        self.null_cond = nn.Parameter(torch.randn(dim_h))
        self.blocks = nn.ModuleList([MMDiTBlock(dim_h, dim_h_mult, attn_dim_head, attn_heads, attn_kv_heads, attn_qk_norm,
                                                self.attn_context_len) for _ in range(depth)])
        self.final_layer = FinalLayer(dim_h, patch_size * dim_h)
        self.out = nn.Conv1d(dim_h, dim_in_x, 1)
        self.initialize_weights()
        self._init_engine()

    def initialize_weights(self) -> None:
        """mmdit.py:296-327."""
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv1d)):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        for lin in (self.mlp_a[0], self.mlp_a[2], self.mlp_time[1][0], self.mlp_time[1][2], self.mlp_cond[1][0], self.mlp_cond[1][2]):
            nn.init.normal_(lin.weight, std=0.02)
        zero = [self.final_layer.modulation[1], self.final_layer.linear, self.out]
        for b in self.blocks:
            zero += [b.modulation_x[1], b.modulation_a[1]]
        for m in zero:
            nn.init.zeros_(m.weight)
            nn.init.zeros_(m.bias)

    def _patch_embed(self, ctx: Ctx, m: PatchEmbedding, packed: torch.Tensor, Cin: int) -> Act:
        """Conv1d(kernel = stride = p) == a Linear over the (B, L/p, p*Cp) view of the channels-last input (mmdit.py:44-52)."""
        st, dev = ctx.store, ctx.device
        B, Lp, Cp = packed.shape
        p = m.patch_size
        w, bias = m.proj.weight, m.proj.bias
        Cout = w.shape[0]

        def build():
            out = E.zeros((Cout, p, Cp), F32, dev)
            out[:, :, :Cin] = w.detach().permute(0, 2, 1)
            return out.view(Cout, p * Cp).to(BF16)
        w2 = st._cached(("patch", id(w), Cp), (w,), build)
        xv = packed.view(B, Lp // p, p * Cp)
        y = E.empty((B, Lp // p, Cout), BF16, dev)
        R.gemm_fwd(xv, w2, N_out=Cout, K=p * Cp, bias=bias, out_bf16=y)
        out = Act(None, y)
        if ctx.tape is not None:
            def backward():
                dy = out.grad
                out.grad = None
                if dy is None:
                    return
                dy16 = E.empty(tuple(y.shape), BF16, dev)
                E.cast_copy(dy, dy16)
                E._bias_grad(st, bias, dy16)
                if w.requires_grad:
                    tmp = E.zeros((1, Cout, p * Cp), F32, dev)
                    R.gemm_wgrad(dy16, xv, tmp, M=Cout, N_out=p * Cp)
                    st.set_grad(w, tmp.view(Cout, p, Cp)[:, :, :Cin].permute(0, 2, 1).contiguous())
            ctx.tape.push(backward)
        return out

    def run(self, tape: Optional[Tape], x, a, t, c, keep):
        """MMDiT.forward (mmdit.py:344-389).  Returns (out16 (B, Lp, 8) bf16, (ctx, final-layer output Act))."""
        assert x.shape[-1] == a.shape[-1] and x.shape[1] == self.dim_in_x and a.shape[1] == self.dim_in_a
        n = x.shape[-1]
        B = x.shape[0]
        p = self.patch_size
        Lp = n + ((-n) % p)
        ctx = self._begin(tape, x.device)
        Ca, Cx = self.dim_in_a, self.dim_in_x
        assert Ca % 8 == 0 and Cx <= 8 and self.dim_h % 8 == 0
        x8 = _pack(x, 8, Lp, X_PAD_VALUE)
        a16 = _pack(a, Ca, Lp, A_PAD_VALUE)
        ff_t, ff_c = self.mlp_time[1], self.mlp_cond[1]
        cact = self._conditioning(ctx, a, t, c, keep, [(ff_t[0], 1), (ff_t[2], 0)], [(self.mlp_cond[0], 0), (ff_c[0], 1), (ff_c[2], 0)],
                                  [(self.feature_extractor_a, 0), (self.mlp_a[0], 1), (self.mlp_a[2], 0)], self.mlp_time[0].theta)
        hx = to_stream(ctx, self._patch_embed(ctx, self.emb_x, x8, Cx))
        ha = to_stream(ctx, self._patch_embed(ctx, self.emb_a, a16, Ca))
        for i, blk in enumerate(self.blocks):
            hx, ha = mmdit_block(ctx, blk, hx, ha, cact, last=(i == len(self.blocks) - 1))
        yf = final_layer(ctx, self.final_layer, hx, cact)          # (B, Lp/p, p * dim_h)
        # unpatchify "b n (p d) -> b d (n p)" is a plain view in the channels-last layout (mmdit.py:388)
        out16 = self._output_conv(ctx, self.out, yf.bf16.view(B, Lp, self.dim_h))
        return out16, (ctx, yf)

    def final_backward(self, ctx: Ctx, yf: Act, dY16: torch.Tensor) -> None:
        B, m, Wd = yf.bf16.shape
        g = self._output_conv_backward(ctx, self.out, yf.bf16.view(B, m * self.patch_size, self.dim_h), dY16)
        yf.grad = g.view(B, m, Wd)
