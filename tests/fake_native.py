"""TEST INFRASTRUCTURE: a CPU stand-in for the C-ABI (include/osufusion_b200.h), written from the header's contracts in plain torch
over raw host pointers: of_gemm (forward / dgrad / wgrad incl. the epilogue options), attention, the ResidualBlock / LayerNorm / RoPE /
FiLM / loss / sampler kernels of the U-Net path, LoRA / DoRA, the fused optimizer, grouped packing, and the DiT / MMDiT kernels.

Purpose: exercise the HOST side of the engine (argument order and counts, strides and views, tape order, gradient routing, arena
plumbing, DDP bucket scheduling) on the GPU-less build box.  It is installed by monkeypatching `osufusion_b200._native.call` inside a
test; the product never imports it, and it says nothing about the CUDA kernels themselves (those are checked on the B200 by the
`-m gpu` tests — the new backbone kernels also against these very restatements).  Rounding points follow the kernels where that is
cheap (bf16 operands / outputs); results agree with the kernels to bf16 tolerance, not bit for bit.
"""
from __future__ import annotations

import ctypes
import math

import torch

BF16, F32 = torch.bfloat16, torch.float32
_SIZE = {BF16: 2, F32: 4, torch.float64: 8, torch.int32: 4, torch.int64: 8}


def _mem(ptr, n, dtype):
    buf = (ctypes.c_char * (int(n) * _SIZE[dtype])).from_address(int(ptr))
    return torch.frombuffer(buf, dtype=dtype)


def v3(ptr, dtype, B, L, Cc, bs, ld):
    if not ptr:
        return None
    n = (B - 1) * bs + (L - 1) * ld + Cc
    return _mem(ptr, n, dtype).as_strided((B, L, Cc), (bs, ld, 1))


def v2(ptr, dtype, rows, cols, ld=None):
    if not ptr:
        return None
    ld = cols if ld is None else ld
    return _mem(ptr, (rows - 1) * ld + cols, dtype).as_strided((rows, cols), (ld, 1))


def rb(x):
    return x.to(BF16).to(F32)


def silu(x):
    return x * torch.sigmoid(x)


def dsilu(x):
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


def _struct(arg):
    return arg._obj if hasattr(arg, "_obj") else arg


def of_gemm(g):
    g = _struct(g)
    Bn, rows, Nn, K, T = g.batch, g.rows, g.N, g.K, g.taps
    if g.mode == 0:
        A = v3(g.a, BF16, Bn, rows, K, g.a_batch_stride, g.a_ld).float()
        acc = torch.zeros(Bn, rows, Nn)
        for t in range(T):
            sh = g.shift0 + t * g.shift_step
            if g.b_mn_major:
                Bt = v2(g.b + 2 * t * g.b_tap_stride, BF16, K, Nn, g.b_ld).float()          # [K][N]
            else:
                Bt = v2(g.b + 2 * t * g.b_tap_stride, BF16, Nn, K, g.b_ld).float().t()      # [N][K] -> (K, N)
            As = torch.zeros_like(A)
            lo, hi = max(0, -sh), min(rows, rows - sh)
            if hi > lo:
                As[:, lo:hi] = A[:, lo + sh:hi + sh]
            acc += As @ Bt
        v = acc
        if g.bias:
            v = v + v2(g.bias, F32, 1, Nn)[0]
        if g.aux_f32:
            v = v + v3(g.aux_f32, F32, Bn, rows, Nn, g.aux_f32_batch_stride, g.aux_f32_ld)
        aux16 = v3(g.aux_bf16, BF16, Bn, rows, Nn, g.aux_bf16_batch_stride, g.aux_bf16_ld)
        if aux16 is not None and not g.aux_is_dsilu:
            v = v + aux16.float()
        if g.pre_bf16:
            v3(g.pre_bf16, BF16, Bn, rows, Nn, g.out_bf16_batch_stride, g.out_bf16_ld).copy_(v.to(BF16))
        if g.act == 1:
            v = silu(v)
        if aux16 is not None and g.aux_is_dsilu:
            v = v * dsilu(aux16.float())
        if g.out_bf16:
            v3(g.out_bf16, BF16, Bn, rows, Nn, g.out_bf16_batch_stride, g.out_bf16_ld).copy_(v.to(BF16))
        if g.out_f32:
            v3(g.out_f32, F32, Bn, rows, Nn, g.out_f32_batch_stride, g.out_f32_ld).copy_(v)
        if g.stats:
            vr = rb(v).double()
            st = _mem(g.stats, 2 * Bn, torch.float64).view(Bn, 2)
            st[:, 0] += vr.sum((1, 2))
            st[:, 1] += (vr * vr).sum((1, 2))
    else:
        M = K
        dY = v3(g.a, BF16, Bn, rows, M, g.a_batch_stride, g.a_ld).float()
        X = v3(g.b, BF16, Bn, rows, Nn, g.b_tap_stride, g.b_ld).float()
        out = v3(g.out_f32, F32, T, M, Nn, g.out_f32_batch_stride, g.out_f32_ld)
        for t in range(T):
            sh = g.shift0 + t * g.shift_step
            Xs = torch.zeros_like(X)
            lo, hi = max(0, -sh), min(rows, rows - sh)
            if hi > lo:
                Xs[:, lo:hi] = X[:, lo + sh:hi + sh]
            out[t] += torch.einsum("blm,bln->mn", dY, Xs)


def _attn_views(g):
    B, H, KVH, L, D = g.B, g.H, g.KVH, g.L, g.D
    q = v3(g.q, BF16, B, L, H * D, g.q_batch_stride, g.q_ld).float().view(B, L, H, D).transpose(1, 2)
    k = v3(g.k, BF16, B, L, KVH * D, g.kv_batch_stride, g.kv_ld).float().view(B, L, KVH, D).transpose(1, 2)
    v = v3(g.v, BF16, B, L, KVH * D, g.kv_batch_stride, g.kv_ld).float().view(B, L, KVH, D).transpose(1, 2)
    idx = torch.arange(H) % KVH      # q head i uses kv head i % KVH
    return q, k[:, idx], v[:, idx], idx


def of_attn_fwd(g):
    g = _struct(g)
    B, H, L, D = g.B, g.H, g.L, g.D
    q, k, v, _ = _attn_views(g)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(D)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, L, H * D)
    v3(g.out, BF16, B, L, H * D, g.out_batch_stride, g.out_ld).copy_(o.to(BF16))
    if g.lse:
        v3(g.lse, F32, B, H, L, H * L, L).copy_(torch.logsumexp(s, dim=-1) / math.log(2.0))


def of_attn_bwd(g):
    g = _struct(g)
    B, H, KVH, L, D = g.B, g.H, g.KVH, g.L, g.D
    q, k, v, idx = _attn_views(g)
    dO = v3(g.dout, BF16, B, L, H * D, g.dout_batch_stride, g.dout_ld).float().view(B, L, H, D).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(D)
    p = torch.softmax(s, dim=-1)
    dv = p.transpose(-1, -2) @ dO
    dp = dO @ v.transpose(-1, -2)
    ds = p * (dp - (dp * p).sum(-1, keepdim=True)) / math.sqrt(D)
    dq = ds @ k
    dk = ds.transpose(-1, -2) @ q
    if g.zero_grads:
        v3(g.dq, F32, B, L, H * D, g.dq_batch_stride, g.dq_ld).zero_()
        v3(g.dk, F32, B, L, KVH * D, g.dkv_batch_stride, g.dkv_ld).zero_()
        v3(g.dv, F32, B, L, KVH * D, g.dkv_batch_stride, g.dkv_ld).zero_()
    v3(g.dq, F32, B, L, H * D, g.dq_batch_stride, g.dq_ld).add_(dq.transpose(1, 2).reshape(B, L, H * D))
    dkk = torch.zeros(B, KVH, L, D)
    dvv = torch.zeros(B, KVH, L, D)
    dkk.index_add_(1, idx, dk)
    dvv.index_add_(1, idx, dv)
    v3(g.dk, F32, B, L, KVH * D, g.dkv_batch_stride, g.dkv_ld).add_(dkk.transpose(1, 2).reshape(B, L, KVH * D))
    v3(g.dv, F32, B, L, KVH * D, g.dkv_batch_stride, g.dkv_ld).add_(dvv.transpose(1, 2).reshape(B, L, KVH * D))


def of_layernorm_fwd(x, x_ld, rows, Cc, gamma, beta, eps, out_f32, out_bf16, out_ld, mean_rstd):
    X = v2(x, F32, rows, Cc, x_ld)
    mean = X.mean(1, keepdim=True)
    rstd = torch.rsqrt(((X - mean) ** 2).mean(1, keepdim=True) + eps)
    y = (X - mean) * rstd * v2(gamma, F32, 1, Cc) + v2(beta, F32, 1, Cc)
    if out_f32:
        v2(out_f32, F32, rows, Cc, out_ld).copy_(y)
    if out_bf16:
        v2(out_bf16, BF16, rows, Cc, out_ld).copy_(y.to(BF16))
    if mean_rstd:
        v2(mean_rstd, F32, rows, 2).copy_(torch.cat([mean, rstd], 1))


def of_layernorm_bwd(dy, dy_ld, x, x_ld, rows, Cc, gamma, mean_rstd, dx_f32, dx_bf16, dx_ld, dgamma, dbeta):
    DY, X = v2(dy, F32, rows, Cc, dy_ld), v2(x, F32, rows, Cc, x_ld)
    mr = v2(mean_rstd, F32, rows, 2)
    xh = (X - mr[:, :1]) * mr[:, 1:]
    dxh = DY * v2(gamma, F32, 1, Cc)
    dx = mr[:, 1:] * (dxh - dxh.mean(1, keepdim=True) - xh * (dxh * xh).mean(1, keepdim=True))
    if dx_f32:
        v2(dx_f32, F32, rows, Cc, dx_ld).copy_(dx)
    if dx_bf16:
        v2(dx_bf16, BF16, rows, Cc, dx_ld).copy_(dx.to(BF16))
    v2(dgamma, F32, 1, Cc)[0] += (DY * xh).sum(0)
    v2(dbeta, F32, 1, Cc)[0] += DY.sum(0)


def of_linear_small_fwd(x, x_ld, M, Nn, K, W, w_ld, bias, act, rnd, y, y_ld, ypre):
    X, Wm = v2(x, F32, M, K, x_ld), v2(W, F32, Nn, K, w_ld)
    if rnd:
        X, Wm = rb(X), rb(Wm)
    v = X @ Wm.t()
    if bias:
        v = v + v2(bias, F32, 1, Nn)
    if rnd:
        v = rb(v)
    if ypre:
        v2(ypre, F32, M, Nn, y_ld).copy_(v)
    if act == 1:
        v = silu(v)
    elif act == 2:
        v = torch.sigmoid(v)
    if rnd and act:
        v = rb(v)
    v2(y, F32, M, Nn, y_ld).copy_(v)


def of_linear_small_bwd(dy, dy_ld, ypre, act, x, x_ld, M, Nn, K, W, w_ld, rnd, dW, dbias, dx, dx_ld, accumulate):
    d = v2(dy, F32, M, Nn, dy_ld).clone()
    if act == 1:
        d = d * dsilu(v2(ypre, F32, M, Nn, dy_ld))
    elif act == 2:
        s = torch.sigmoid(v2(ypre, F32, M, Nn, dy_ld))
        d = d * s * (1 - s)
    X, Wm = v2(x, F32, M, K, x_ld), v2(W, F32, Nn, K, w_ld)
    if rnd:
        X, Wm = rb(X), rb(Wm)
    if dW:
        g = v2(dW, F32, Nn, K, w_ld)
        g.copy_(g + d.t() @ X if accumulate else d.t() @ X)
    if dbias:
        b = v2(dbias, F32, 1, Nn)[0]
        b.copy_(b + d.sum(0) if accumulate else d.sum(0))
    if dx:
        v2(dx, F32, M, K, dx_ld).add_(d @ Wm)


def of_cast_copy(s32, s16, s_ld, s_bs, B, L, Cc, d32, d16, d_ld, d_bs, accumulate):
    assert Cc % 8 == 0
    src = v3(s32, F32, B, L, Cc, s_bs, s_ld) if s32 else v3(s16, BF16, B, L, Cc, s_bs, s_ld).float()
    val = src.clone()
    if d32:
        dst = v3(d32, F32, B, L, Cc, d_bs, d_ld)
        if accumulate:
            val = val + dst
        dst.copy_(val)
    if d16:
        v3(d16, BF16, B, L, Cc, d_bs, d_ld).copy_(val.to(BF16))


def of_colsum_bf16(dy, ld, rows, Nn, db):
    v2(db, F32, 1, Nn)[0] += v2(dy, BF16, rows, Nn, ld).float().sum(0)


def of_coldot_bf16(dy, dy_ld, y, y_ld, rows, Nn, bias, out):
    Y = v2(y, BF16, rows, Nn, y_ld).float()
    if bias:
        Y = Y - v2(bias, F32, 1, Nn)
    v2(out, F32, 1, Nn)[0] += (v2(dy, BF16, rows, Nn, dy_ld).float() * Y).sum(0)


def of_pack_input(x, noise, ca, cb, B, Cc, n, out, Lp, Cp, pad):
    X = _mem(x, B * Cc * n, F32).view(B, Cc, n)
    val = X.clone()
    if ca:
        val = val * _mem(ca, B, F32).view(B, 1, 1)
    if noise:
        val = val + _mem(cb, B, F32).view(B, 1, 1) * _mem(noise, B * Cc * n, F32).view(B, Cc, n)
    o = torch.zeros(B, Lp, Cp)
    o[:, :n, :Cc] = val.transpose(1, 2)
    o[:, n:, :Cc] = pad
    _mem(out, B * Lp * Cp, BF16).view(B, Lp, Cp).copy_(o.to(BF16))


def of_unpack_output(y, ld, bs, B, Cc, n, out):
    Y = v3(y, BF16, B, n, Cc, bs, ld).float()
    _mem(out, B * Cc * n, F32).view(B, Cc, n).copy_(Y.transpose(1, 2))


def of_time_embed(t, B, dim, theta, out):
    half = dim // 2
    f = torch.exp(torch.arange(half).float() * -(math.log(theta) / (half - 1)))
    ang = _mem(t, B, F32)[:, None] * f[None, :]
    v2(out, F32, B, dim).copy_(torch.cat([ang.sin(), ang.cos()], 1))


def of_silu_small(x, dy, out, n):
    X = _mem(x, n, F32)
    _mem(out, n, F32).copy_(_mem(dy, n, F32) * dsilu(X) if dy else silu(X))


def of_cast_f32_bf16(src, dst, n):
    _mem(dst, n, BF16).copy_(_mem(src, n, F32).to(BF16))


def of_gate_residual_fwd(x32, x16, x_ld, x_bs, y16, y_ld, y_bs, gate, gate_ld, rnd, B, L, Cc, out32, o_ld, o_bs):
    X = v3(x32, F32, B, L, Cc, x_bs, x_ld) if x32 else v3(x16, BF16, B, L, Cc, x_bs, x_ld).float()
    p = v2(gate, F32, B, Cc, gate_ld)[:, None, :] * v3(y16, BF16, B, L, Cc, y_bs, y_ld).float()
    v3(out32, F32, B, L, Cc, o_bs, o_ld).copy_(X + (rb(p) if rnd else p))


def of_gate_mul_bwd(d32, d_ld, d_bs, gate, gate_ld, rnd, B, L, Cc, dx16, dy16, o_ld, o_bs):
    d = v3(d32, F32, B, L, Cc, d_bs, d_ld)
    v3(dx16, BF16, B, L, Cc, o_bs, o_ld).copy_(d.to(BF16))
    v3(dy16, BF16, B, L, Cc, o_bs, o_ld).copy_((v2(gate, F32, B, Cc, gate_ld)[:, None, :] * (rb(d) if rnd else d)).to(BF16))


def of_headnorm_fwd(in16, in_ld, in_bs, B, L, Hq, Hk, Hv, D, gq, gk, scale, out16, o_ld, o_bs, variant=0):
    W = (Hq + Hk + Hv) * D
    X = v3(in16, BF16, B, L, W, in_bs, in_ld).float()
    out = v3(out16, BF16, B, L, W, o_bs, o_ld)
    g = torch.cat([_mem(gq, Hq * D, F32), _mem(gk, Hk * D, F32)]).view(Hq + Hk, D)
    qk = X[:, :, :(Hq + Hk) * D].reshape(B, L, Hq + Hk, D)
    n = rb(qk.norm(dim=-1, keepdim=True)).clamp_min(1e-12)
    y = (rb(qk / n) * g) * scale
    out[:, :, :(Hq + Hk) * D] = y.reshape(B, L, -1).to(BF16)
    out[:, :, (Hq + Hk) * D:] = X[:, :, (Hq + Hk) * D:].to(BF16)


def of_headnorm_bwd(dq, dq_ld, dq_bs, dk, dv, dkv_ld, dkv_bs, in16, in_ld, in_bs, B, L, Hq, Hk, Hv, D, gq, gk, scale, dqkv16, o_ld, o_bs,
                    dgq, dgk, variant=0):
    W = (Hq + Hk + Hv) * D
    X = v3(in16, BF16, B, L, W, in_bs, in_ld).float()
    out = v3(dqkv16, BF16, B, L, W, o_bs, o_ld)
    dQ = v3(dq, F32, B, L, Hq * D, dq_bs, dq_ld)
    dK = v3(dk, F32, B, L, Hk * D, dkv_bs, dkv_ld)
    d = torch.cat([dQ, dK], dim=2).reshape(B, L, Hq + Hk, D)
    g = torch.cat([_mem(gq, Hq * D, F32), _mem(gk, Hk * D, F32)]).view(Hq + Hk, D)
    qk = X[:, :, :(Hq + Hk) * D].reshape(B, L, Hq + Hk, D)
    n = qk.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    xh = qk / n
    dg = (d * xh * scale).sum((0, 1))
    _mem(dgq, Hq * D, F32).add_(dg[:Hq].reshape(-1))
    _mem(dgk, Hk * D, F32).add_(dg[Hq:].reshape(-1))
    dxh = d * g * scale
    dx = (dxh - xh * (dxh * xh).sum(-1, keepdim=True)) / n
    out[:, :, :(Hq + Hk) * D] = dx.reshape(B, L, -1).to(BF16)
    if Hv:
        out[:, :, (Hq + Hk) * D:] = v3(dv, F32, B, L, Hv * D, dkv_bs, dkv_ld).to(BF16)


def of_row_mean_std(a, B, Cc, n, out):
    A = _mem(a, B * Cc * n, F32).view(B, Cc, n)
    v2(out, F32, B, 2 * Cc).copy_(torch.cat([A.mean(-1), A.std(-1)], 1))


def of_pack_weights(table, num_segs, total_ctas):
    from osufusion_b200._native import PackSeg
    segs = (PackSeg * num_segs).from_address(int(table))
    ctas = 0
    for sg in segs:
        assert sg.cta_begin == ctas
        sc = sg.scale if sg.scale != 0.0 else 1.0
        if sg.k == 1 and sg.cin_pad == sg.Cin:
            n = sg.Cout * sg.Cin
            _mem(sg.dst, n, BF16).copy_((sc * _mem(sg.src, n, F32)).to(BF16))
            ctas += (n + 4095) // 4096
        else:
            w = sc * _mem(sg.src, sg.Cout * sg.Cin * sg.k, F32).view(sg.Cout, sg.Cin, sg.k)
            out = torch.zeros(sg.k, sg.Cout, sg.cin_pad)
            out[:, :, :sg.Cin] = w.permute(2, 0, 1)
            _mem(sg.dst, sg.k * sg.Cout * sg.cin_pad, BF16).copy_(out.reshape(-1).to(BF16))
            ctas += sg.Cout * ((sg.cin_pad + 1023) // 1024)
    assert ctas == total_ctas, (ctas, total_ctas)


def of_adaln_fwd(x, x_ld, x_bs, B, L, Cc, s1p, sc_ld, shift, sh_ld, eps, out16, o_ld, o_bs, mean_rstd):
    X = v3(x, F32, B, L, Cc, x_bs, x_ld)
    mean = X.mean(2, keepdim=True)
    rstd = torch.rsqrt(((X - mean) ** 2).mean(2, keepdim=True) + eps)
    y = (X - mean) * rstd * v2(s1p, F32, B, Cc, sc_ld)[:, None, :] + v2(shift, F32, B, Cc, sh_ld)[:, None, :]
    v3(out16, BF16, B, L, Cc, o_bs, o_ld).copy_(y.to(BF16))
    v2(mean_rstd, F32, B * L, 2).copy_(torch.cat([mean, rstd], 2).view(B * L, 2))


def of_adaln_bwd(dy, dy_ld, dy_bs, x, x_ld, x_bs, B, L, Cc, s1p, sc_ld, mean_rstd, dres, r_ld, r_bs, dx, dx_ld, dx_bs, dscale, ds_ld,
                 dshift, dsh_ld):
    DY, X = v3(dy, F32, B, L, Cc, dy_bs, dy_ld), v3(x, F32, B, L, Cc, x_bs, x_ld)
    mr = v2(mean_rstd, F32, B * L, 2).view(B, L, 2)
    xh = (X - mr[:, :, :1]) * mr[:, :, 1:]
    dxh = DY * v2(s1p, F32, B, Cc, sc_ld)[:, None, :]
    out = mr[:, :, 1:] * (dxh - dxh.mean(2, keepdim=True) - xh * (dxh * xh).mean(2, keepdim=True))
    if dres:
        out = out + v3(dres, F32, B, L, Cc, r_bs, r_ld)
    v3(dx, F32, B, L, Cc, dx_bs, dx_ld).copy_(out)
    v2(dscale, F32, B, Cc, ds_ld).add_((DY * xh).sum(1))
    v2(dshift, F32, B, Cc, dsh_ld).add_(DY.sum(1))


def of_gate_bwd(d32, d_ld, d_bs, gate, gate_ld, y16, y_ld, y_bs, rnd, B, L, Cc, dy16, o_ld, o_bs, dgate, dg_ld):
    d = v3(d32, F32, B, L, Cc, d_bs, d_ld)
    dr = rb(d) if rnd else d
    v3(dy16, BF16, B, L, Cc, o_bs, o_ld).copy_((v2(gate, F32, B, Cc, gate_ld)[:, None, :] * dr).to(BF16))
    v2(dgate, F32, B, Cc, dg_ld).add_((dr * v3(y16, BF16, B, L, Cc, y_bs, y_ld).float()).sum(1))


# ------------------------------------------------------------------------------------------------ U-Net path (of_rb_*, RoPE, FiLM, ...)
def _gn(a):
    """GroupNorm(1, C) + FiLM + SiLU recompute shared by the of_rb_* kernels: returns y, xhat, z, f, h, rstd, sp1."""
    B, L, Cc = a.B, a.L, a.C
    y = v3(a.y, BF16, B, L, Cc, a.y_bs, a.y_ld).float()
    st = _mem(a.stats, 2 * B, torch.float64).view(B, 2)
    n = float(L * Cc)
    mean = (st[:, 0] / n)
    var = (st[:, 1] / n - mean * mean).clamp_min(0)
    rstd = (1.0 / torch.sqrt(var + a.eps)).float().view(B, 1, 1)
    mean = mean.float().view(B, 1, 1)
    gamma, beta = _mem(a.gamma, Cc, F32), _mem(a.beta, Cc, F32)
    xhat = (y - mean) * rstd
    z = xhat * gamma + beta
    sp1 = None
    f = z
    if a.ss:
        ss = _mem(a.ss, B * 2 * Cc, F32).view(B, 2 * Cc)
        sp1 = rb(ss[:, :Cc] + 1.0)[:, None, :]
        f = z * sp1 + ss[:, None, Cc:]
    return y, xhat, z, f, silu(f), rstd, sp1


def of_rb_apply_fwd(a):
    a = _struct(a)
    h = _gn(a)[4]
    v3(a.out_bf16, BF16, a.B, a.L, a.C, a.out_bf16_bs, a.out_bf16_ld).copy_(h.to(BF16))


def of_rb_rowdot(a):
    a = _struct(a)
    B, L, Cc = a.B, a.L, a.C
    h = rb(_gn(a)[4])
    out = _mem(a.out_rows, B * L, F32).view(B, L)
    if a.mode == 0:
        w = rb(_mem(a.vec, Cc, F32))
        bias = _mem(a.vec_bias, 1, F32)[0] if a.vec_bias else 0.0
        out.copy_(rb((h * w).sum(-1) + bias))
    else:
        w = v2(a.vec, F32, B, Cc, a.vec_bs)
        out.copy_((h * w[:, None, :]).sum(-1))


def of_softmax_rows(rows, B, L):
    r = _mem(rows, B * L, F32).view(B, L)
    r.copy_(torch.softmax(r, dim=-1))


def of_softmax_bwd_rows(p, rd, B, L):
    P, R = _mem(p, B * L, F32).view(B, L), _mem(rd, B * L, F32).view(B, L)
    R.copy_(P * (R - (P * R).sum(-1, keepdim=True)))


def of_rb_pool(a):
    a = _struct(a)
    B, L, Cc = a.B, a.L, a.C
    h = rb(_gn(a)[4])
    P = rb(_mem(a.p, B * L, F32).view(B, L))
    _mem(a.acc_bc, B * Cc, F32).view(B, Cc).add_((h * P[:, :, None]).sum(1))


def of_rb_logit_pool(a, part, pooled):
    a = _struct(a)
    B, L, Cc = a.B, a.L, a.C
    h = rb(_gn(a)[4])
    w = rb(_mem(a.vec, Cc, F32))
    bias = _mem(a.vec_bias, 1, F32)[0] if a.vec_bias else 0.0
    P = torch.softmax(rb((h * w).sum(-1) + bias), dim=-1)
    _mem(a.out_rows, B * L, F32).view(B, L).copy_(P)
    _mem(pooled, B * Cc, F32).view(B, Cc).copy_((h * P[:, :, None]).sum(1))


def of_rb_gate_fwd(a):
    a = _struct(a)
    B, L, Cc = a.B, a.L, a.C
    h = _gn(a)[4]
    res = v3(a.res_f32, F32, B, L, Cc, a.res_f32_bs, a.res_f32_ld) if a.res_f32 else \
        v3(a.res_bf16, BF16, B, L, Cc, a.res_bf16_bs, a.res_bf16_ld).float()
    o = h * _mem(a.gate, B * Cc, F32).view(B, 1, Cc) + res
    if a.out_f32:
        v3(a.out_f32, F32, B, L, Cc, a.out_f32_bs, a.out_f32_ld).copy_(o)
    if a.out_bf16:
        v3(a.out_bf16, BF16, B, L, Cc, a.out_bf16_bs, a.out_bf16_ld).copy_(o.to(BF16))


def of_rb_gate_bwd_reduce(a):
    a = _struct(a)
    B, L, Cc = a.B, a.L, a.C
    h = _gn(a)[4]
    d = v3(a.dout_f32, F32, B, L, Cc, a.dout_f32_bs, a.dout_f32_ld)
    _mem(a.acc_bc, B * Cc, F32).view(B, Cc).add_((d * h).sum(1))


def of_rb_bwd_pass1(a):
    a = _struct(a)
    B, L, Cc = a.B, a.L, a.C
    y, xhat, z, f, h, rstd, sp1 = _gn(a)
    gamma = _mem(a.gamma, Cc, F32)
    if a.mode == 0:
        d = v3(a.dout_f32, F32, B, L, Cc, a.dout_f32_bs, a.dout_f32_ld)
        gate = _mem(a.gate, B * Cc, F32).view(B, 1, Cc)
        dpool = _mem(a.dpooled, B * Cc, F32).view(B, 1, Cc)
        P = rb(_mem(a.p, B * L, F32).view(B, L))[:, :, None]
        da = _mem(a.da, B * L, F32).view(B, L)[:, :, None]
        wk = rb(_mem(a.wk, Cc, F32))
        dh = d * gate + dpool * P + da * wk
        _mem(a.dwk, Cc, F32).add_((da * rb(h)).sum((0, 1)))
        if a.dbk:
            _mem(a.dbk, 1, F32).add_(da.sum())
        if a.dout_bf16:
            v3(a.dout_bf16, BF16, B, L, Cc, a.dout_bf16_bs, a.dout_bf16_ld).copy_(d.to(BF16))
    else:
        dh = v3(a.dh_bf16, BF16, B, L, Cc, a.dh_bs, a.dh_ld).float()
    df = dh * dsilu(f)
    dz = df
    if a.ss:
        dss = _mem(a.dss, B * 2 * Cc, F32).view(B, 2 * Cc)
        dss[:, :Cc] += (df * z).sum(1)
        dss[:, Cc:] += df.sum(1)
        dz = df * sp1
    _mem(a.dgamma, Cc, F32).add_((dz * xhat).sum((0, 1)))
    _mem(a.dbeta, Cc, F32).add_(dz.sum((0, 1)))
    dx = dz * gamma
    v3(a.dxhat_bf16, BF16, B, L, Cc, a.dxhat_bs, a.dxhat_ld).copy_(dx.to(BF16))
    dxr = rb(dx)
    ds = _mem(a.dstats, 2 * B, torch.float64).view(B, 2)
    ds[:, 0] += dxr.sum((1, 2)).double()
    ds[:, 1] += (dxr * xhat).sum((1, 2)).double()


def of_rb_bwd_apply(a):
    a = _struct(a)
    B, L, Cc = a.B, a.L, a.C
    y, xhat, z, f, h, rstd, sp1 = _gn(a)
    n = float(L * Cc)
    ds = _mem(a.dstats, 2 * B, torch.float64).view(B, 2)
    m1, m2 = (ds[:, 0] / n).float().view(B, 1, 1), (ds[:, 1] / n).float().view(B, 1, 1)
    dx = v3(a.dxhat_bf16, BF16, B, L, Cc, a.dxhat_bs, a.dxhat_ld).float()
    o = rstd * (dx - m1 - xhat * m2)
    v3(a.dy_bf16, BF16, B, L, Cc, a.dy_bs, a.dy_ld).copy_(o.to(BF16))
    if a.dbias:
        _mem(a.dbias, Cc, F32).add_(rb(o).sum((0, 1)))


def _rope_tabs(cos_tab, sin_tab, L, D, table_f32):
    dt = F32 if table_f32 else BF16
    return _mem(cos_tab, L * D, dt).view(L, D).float(), _mem(sin_tab, L * D, dt).view(L, D).float()


def of_rope_fwd(qkv, ld, bs, B, L, H, KVH, D, cos_tab, sin_tab, table_f32):
    W = (H + KVH) * D
    t = v3(qkv, BF16, B, L, W, bs, ld)
    x = t.float().view(B, L, H + KVH, D)
    cos, sin = _rope_tabs(cos_tab, sin_tab, L, D, table_f32)
    cos, sin = cos[None, :, None, :], sin[None, :, None, :]
    half = D // 2
    rot = torch.cat([-x[..., half:], x[..., :half]], dim=-1)
    o = (x * cos + rot * sin) if table_f32 else (rb(x * cos) + rb(rot * sin))
    t.copy_(o.reshape(B, L, W).to(BF16))


def of_rope_bwd(dq, dq_ld, dq_bs, dk, dv, dkv_ld, dkv_bs, out16, o_ld, o_bs, B, L, H, KVH, D, cos_tab, sin_tab, table_f32):
    g = torch.cat([v3(dq, F32, B, L, H * D, dq_bs, dq_ld), v3(dk, F32, B, L, KVH * D, dkv_bs, dkv_ld)], dim=2).view(B, L, H + KVH, D)
    cos, sin = _rope_tabs(cos_tab, sin_tab, L, D, table_f32)
    cos, sin = cos[None, :, None, :], sin[None, :, None, :]
    half = D // 2
    gs = g * sin
    dx = g * cos + torch.cat([gs[..., half:], -gs[..., :half]], dim=-1)      # transpose of the rotation
    out = v3(out16, BF16, B, L, (H + 2 * KVH) * D, o_bs, o_ld)
    out[:, :, :(H + KVH) * D] = dx.reshape(B, L, -1).to(BF16)
    out[:, :, (H + KVH) * D:] = v3(dv, F32, B, L, KVH * D, dkv_bs, dkv_ld).to(BF16)


def of_upsample2x_fwd(x, x_ld, x_bs, B, L, Cc, out, o_ld, o_bs):
    X = v3(x, BF16, B, L, Cc, x_bs, x_ld)
    v3(out, BF16, B, 2 * L, Cc, o_bs, o_ld).copy_(X.repeat_interleave(2, dim=1))


def of_upsample2x_bwd(d, d_ld, d_bs, B, L, Cc, out32, out16, o_ld, o_bs):
    D2 = v3(d, F32, B, 2 * L, Cc, d_bs, d_ld)
    o = D2[:, 0::2] + D2[:, 1::2]
    if out32:
        v3(out32, F32, B, L, Cc, o_bs, o_ld).copy_(o)
    if out16:
        v3(out16, BF16, B, L, Cc, o_bs, o_ld).copy_(o.to(BF16))


def of_pack_conv_weight(w, Cout, Cin, k, out, cin_pad, tap_offset, taps_total):
    Wt = _mem(w, Cout * Cin * k, F32).view(Cout, Cin, k)
    o = _mem(out, taps_total * Cout * cin_pad, BF16).view(taps_total, Cout, cin_pad)
    o[tap_offset:tap_offset + k, :, :Cin] = Wt.permute(2, 0, 1).to(BF16)
    o[tap_offset:tap_offset + k, :, Cin:] = 0


def of_unpack_conv_wgrad(packed, Cout, Cin, k, cin_pad, tap_offset, dw, accumulate, rezero):
    Pk = _mem(packed, (tap_offset + k) * Cout * cin_pad, F32).view(tap_offset + k, Cout, cin_pad)
    g = Pk[tap_offset:tap_offset + k, :, :Cin].permute(1, 2, 0)
    D = _mem(dw, Cout * Cin * k, F32).view(Cout, Cin, k)
    D.copy_(D + g if accumulate else g)
    if rezero:
        Pk[tap_offset:tap_offset + k, :, :Cin] = 0


def _film_groups(groups, num_groups):
    from osufusion_b200._native import FilmGroup
    return (FilmGroup * num_groups).from_address(int(groups))


def of_film_fwd(groups, num_groups, total_rows, x, M, K, out):
    X = v2(x, F32, M, K)
    for g in _film_groups(groups, num_groups):
        W = rb(v2(g.W, F32, g.N, K))
        y = X @ W.t()
        if g.bias:
            y = y + _mem(g.bias, g.N, F32)
        _mem(out + 4 * g.out_off, M * g.N, F32).view(M, g.N).copy_(rb(y))


def of_film_bwd(groups, chunks, num_chunks, dss, x, M, K, d_emb):
    X = v2(x, F32, M, K)
    idx = sorted(set(_mem(chunks, 2 * num_chunks, torch.int32)[0::2].tolist()))    # a launch may cover a sub-range of the heads
    all_groups = _film_groups(groups, num_groups=max(idx) + 1)
    for g in (all_groups[i] for i in idx):
        d = _mem(dss + 4 * g.out_off, M * g.N, F32).view(M, g.N)
        if g.dW:
            v2(g.dW, F32, g.N, K).copy_(d.t() @ X)
        if g.dbias:
            _mem(g.dbias, g.N, F32).copy_(d.sum(0))
        v2(d_emb, F32, M, K).add_(d @ rb(v2(g.W, F32, g.N, K)))


def of_mse_fwd(pred, ld, bs, x, noise, ta, tb, orig_len, B, Cc, n, accum2, loss):
    P = v3(pred, BF16, B, n, Cc, bs, ld).float().transpose(1, 2)
    tgt = tb * _mem(noise, B * Cc * n, F32).view(B, Cc, n)
    if ta != 0.0:
        tgt = tgt + ta * _mem(x, B * Cc * n, F32).view(B, Cc, n)
    mask = torch.ones(B, 1, n)
    if orig_len:
        ol = _mem(orig_len, B, torch.int64)
        mask = (torch.arange(n)[None, :] < ol[:, None]).float()[:, None, :]
    se, cnt = (mask * (P - tgt) ** 2).sum(), (mask.expand(B, Cc, n)).sum()
    _mem(accum2, 2, F32).copy_(torch.stack([se, cnt]))
    _mem(loss, 1, F32).copy_((se / cnt).view(1))


def of_mse_bwd(pred, ld, bs, x, noise, ta, tb, orig_len, B, Cc, n, Lp, Cp, accum2, gscale, dpred):
    P = v3(pred, BF16, B, n, Cc, bs, ld).float().transpose(1, 2)
    tgt = tb * _mem(noise, B * Cc * n, F32).view(B, Cc, n)
    if ta != 0.0:
        tgt = tgt + ta * _mem(x, B * Cc * n, F32).view(B, Cc, n)
    mask = torch.ones(B, 1, n)
    if orig_len:
        ol = _mem(orig_len, B, torch.int64)
        mask = (torch.arange(n)[None, :] < ol[:, None]).float()[:, None, :]
    g = 2.0 * mask * (P - tgt) / _mem(accum2, 2, F32)[1] * (_mem(gscale, 1, F32)[0] if gscale else 1.0)
    o = torch.zeros(B, Lp, Cp)
    o[:, :n, :Cc] = g.transpose(1, 2)
    _mem(dpred, B * Lp * Cp, BF16).view(B, Lp, Cp).copy_(o.to(BF16))


def of_sampler_update(xin, cond, null_, ld, bs, s, mode, c_eps, c_div, c_x0, c_dir, B, Cc, n, xout, packed, Lp, Cp, pad):
    e = v3(cond, BF16, B, n, Cc, bs, ld).float().transpose(1, 2)
    if null_:
        nl = v3(null_, BF16, B, n, Cc, bs, ld).float().transpose(1, 2)
        e = rb(nl + rb(rb(e - nl) * s))
    X = _mem(xin, B * Cc * n, F32).view(B, Cc, n)
    if mode == 0:
        x0 = ((X - rb(c_eps * e)) / c_div).clamp(-1.0, 1.0)
        o = c_x0 * x0 + rb(c_dir * e)
    else:
        o = X + rb(c_eps * e)
    _mem(xout, B * Cc * n, F32).view(B, Cc, n).copy_(o)
    if packed:
        pk = torch.zeros(B, Lp, Cp)
        pk[:, :n, :Cc] = o.transpose(1, 2)
        pk[:, n:, :Cc] = pad
        _mem(packed, B * Lp * Cp, BF16).view(B, Lp, Cp).copy_(pk.to(BF16))


def of_sampler_update_dev(xin, cond, null_, ld, bs, s, mode, coef, B, Cc, n, xout, packed, Lp, Cp, pad):
    c = _mem(coef, 4, F32).tolist()
    of_sampler_update(xin, cond, null_, ld, bs, s, mode, c[0], c[1], c[2], c[3], B, Cc, n, xout, packed, Lp, Cp, pad)


# ------------------------------------------------------------------------------------------------ LoRA / DoRA, optimizer
def _packed_view(ptr, dtype, k, Cout, cin_pad, tap_stride):
    return _mem(ptr, (k - 1) * tap_stride + Cout * cin_pad, dtype).as_strided((k, Cout, cin_pad), (tap_stride, cin_pad, 1))


def of_scale_cast_f32_bf16(src, scale, dst, n):
    _mem(dst, n, BF16).copy_((_mem(src, n, F32) * scale).to(BF16))


def of_dora_scale_pack(V, mag, Cout, Cin, k, n2_out, packed, cin_pad, tap_stride):
    Vm = _mem(V, Cout * Cin * k, F32).view(Cout, Cin * k)
    n2 = (Vm * Vm).sum(1)
    _mem(n2_out, Cout, F32).copy_(n2)
    sc = _mem(mag, Cout, F32) / n2.sqrt() if mag else torch.ones(Cout)
    out = _packed_view(packed, BF16, k, Cout, cin_pad, tap_stride)
    out[:, :, :Cin] = (sc[:, None] * Vm).view(Cout, Cin, k).permute(2, 0, 1).to(BF16)


def of_dora_scale_pack_prep(V, mag, Cout, Cin, k, n2_out, packed, cin_pad, tap_stride, Bm, scaling, r, Bst, rowscale):
    of_dora_scale_pack(V, mag, Cout, Cin, k, n2_out, packed, cin_pad, tap_stride)
    of_dora_rankr_prep(Bm, mag, n2_out, scaling, Cout, r, Bst, rowscale)


def of_lora_finish_all(table, num_segs, total_ctas):
    from osufusion_b200._native import LoraFinishSeg
    segs = (LoraFinishSeg * num_segs).from_address(int(table))
    ctas = 0
    for sg in segs:
        assert sg.cta_begin == ctas
        of_dora_rankr_finish(sg.dBraw, sg.rowscale, sg.gB, sg.dm, sg.mag, sg.gmag, sg.Cout, sg.r)
        ctas += (sg.Cout * sg.r + 255) // 256
    assert ctas == total_ctas, (ctas, total_ctas)


def of_dora_merge(W, A, Bm, mag, scaling, Cout, Cin, k, r, n2_ws, packed, cin_pad, tap_stride, s_out):
    E = Cin * k
    Vm = _mem(W, Cout * E, F32).view(Cout, E) + scaling * (_mem(Bm, Cout * r, F32).view(Cout, r) @ _mem(A, r * E, F32).view(r, E))
    n2 = (Vm * Vm).sum(1)
    _mem(n2_ws, Cout, F32).copy_(n2)
    sc = _mem(mag, Cout, F32) / n2.sqrt() if mag else torch.ones(Cout)
    if s_out:
        _mem(s_out, Cout, F32).copy_(sc)
    out = _packed_view(packed, BF16, k, Cout, cin_pad, tap_stride)
    out[:, :, :Cin] = (sc[:, None] * Vm).view(Cout, Cin, k).permute(2, 0, 1).to(BF16)


def of_dora_rankr_prep(Bm, mag, n2, scaling, Cout, r, Bst, rowscale):
    sc = _mem(mag, Cout, F32) / _mem(n2, Cout, F32).sqrt() if mag else torch.ones(Cout)
    rs = scaling * sc
    _mem(rowscale, Cout, F32).copy_(rs)
    _mem(Bst, r * Cout, BF16).view(r, Cout).copy_((rs[:, None] * _mem(Bm, Cout * r, F32).view(Cout, r)).t().to(BF16))


def of_dora_rankr_finish(dBraw, rowscale, gB, dm, mag, gmag, Cout, r):
    _mem(gB, Cout * r, F32).view(Cout, r).add_(_mem(rowscale, Cout, F32)[:, None] * _mem(dBraw, Cout * r, F32).view(Cout, r))
    if gmag:
        _mem(gmag, Cout, F32).add_(_mem(dm, Cout, F32) / _mem(mag, Cout, F32))


def of_dora_grad(W, A, Bm, mag, scaling, Cout, Cin, k, r, n2, dWp, cin_pad, tap_stride, dA, dB, dmag):
    E = Cin * k
    Am, Bmat = _mem(A, r * E, F32).view(r, E), _mem(Bm, Cout * r, F32).view(Cout, r)
    Vm = _mem(W, Cout * E, F32).view(Cout, E) + scaling * (Bmat @ Am)
    dWeff = _packed_view(dWp, F32, k, Cout, cin_pad, tap_stride)[:, :, :Cin].permute(1, 2, 0).reshape(Cout, E)
    if mag:
        nrm = _mem(n2, Cout, F32).sqrt()
        sc = _mem(mag, Cout, F32) / nrm
        _mem(dmag, Cout, F32).add_((dWeff * Vm).sum(1) / nrm)
    else:
        sc = torch.ones(Cout)
    dV = sc[:, None] * dWeff
    _mem(dA, r * E, F32).view(r, E).add_(scaling * (Bmat.t() @ dV))
    _mem(dB, Cout * r, F32).view(Cout, r).add_(scaling * (dV @ Am.t()))


def of_grad_sumsq(grads, n, out):
    g = _mem(grads, n, F32).double()
    _mem(out, 1, torch.float64).copy_((g * g).sum().view(1))


def of_adamw_step(table, num, total_ctas, grads, exp_avg, exp_avg_sq, sumsq, max_norm, lr, b1, b2, eps, wd, step):
    from osufusion_b200._native import OptTensor
    clip = 1.0
    if sumsq and max_norm > 0:
        clip = min(1.0, max_norm / (float(_mem(sumsq, 1, torch.float64)[0]) ** 0.5 + 1e-6))
    bc1, bc2s = 1.0 - b1 ** step, (1.0 - b2 ** step) ** 0.5
    for t in (OptTensor * num).from_address(int(table)):
        p = _mem(t.param, t.numel, F32)
        g = _mem(grads + 4 * t.arena_off, t.numel, F32) * clip
        m, v = _mem(exp_avg + 4 * t.arena_off, t.numel, F32), _mem(exp_avg_sq + 4 * t.arena_off, t.numel, F32)
        if t.k > 1:     # packed conv tensor: gradient / moments / operand [k][Cout][Cin], parameter (Cout, Cin, k)
            p = p.view(t.Cout, t.Cin, t.k).permute(2, 0, 1)
            g, m, v = (x.view(t.k, t.Cout, t.Cin) for x in (g, m, v))
        p.mul_(1.0 - lr * wd)
        m.mul_(b1).add_((1.0 - b1) * g)
        v.mul_(b2).add_((1.0 - b2) * g * g)
        p.sub_((lr / bc1) * m / (v.sqrt() / bc2s + eps))
        if t.operand_bf16:
            _mem(t.operand_bf16, t.numel, BF16).view(p.shape).copy_(p.to(BF16))


_TABLE = {k: v for k, v in globals().items() if k.startswith("of_")}
CALLS = []


def call(name, *args, flops=0.0, family=None, tag=""):
    """Drop-in for osufusion_b200._native.call on CPU tensors (no stream argument)."""
    fn = _TABLE.get(name)
    if fn is None:
        raise NotImplementedError(f"fake_native: {name} is not emulated")
    from osufusion_b200._native import _SIGS
    assert len(args) == len(_SIGS[name]) - 1, f"{name}: {len(args)} arguments passed, the binding declares {len(_SIGS[name]) - 1} (+ stream)"
    CALLS.append(name)
    fn(*[0 if a is None else a for a in args])
