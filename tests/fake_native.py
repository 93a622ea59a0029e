"""TEST INFRASTRUCTURE: a CPU stand-in for the subset of the C-ABI (include/osufusion_b200.h) that osufusion_b200/backbones.py
drives, written from the header's contracts in plain torch over raw host pointers.

Purpose: exercise the HOST side of the engine (argument order, strides and views, tape order, gradient routing, arena plumbing)
on the GPU-less build box.  It is installed by monkeypatching `osufusion_b200._native.call` inside a test; the product never
imports it, and it says nothing about the CUDA kernels themselves (those are checked on the B200 by the `-m gpu` tests).
"""
from __future__ import annotations

import ctypes
import math

import torch

BF16, F32 = torch.bfloat16, torch.float32
_SIZE = {BF16: 2, F32: 4, torch.float64: 8}


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
        assert not g.stats, "stats epilogue not emulated"
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
        if sg.k == 1 and sg.cin_pad == sg.Cin:
            n = sg.Cout * sg.Cin
            _mem(sg.dst, n, BF16).copy_(_mem(sg.src, n, F32).to(BF16))
            ctas += (n + 4095) // 4096
        else:
            w = _mem(sg.src, sg.Cout * sg.Cin * sg.k, F32).view(sg.Cout, sg.Cin, sg.k)
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


_TABLE = {k: v for k, v in globals().items() if k.startswith("of_")}
CALLS = []


def call(name, *args, flops=0.0, family=None, tag=""):
    """Drop-in for osufusion_b200._native.call on CPU tensors (no stream argument)."""
    fn = _TABLE.get(name)
    if fn is None:
        raise NotImplementedError(f"fake_native: {name} is not emulated")
    CALLS.append(name)
    fn(*[0 if a is None else a for a in args])
