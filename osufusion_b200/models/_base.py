"""Shared machinery of the two model wrappers: fused noising + denoiser + masked-MSE training step as ONE autograd node,
and the CFG-batched, audio-cached sampling loop."""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import nn

from .. import _native as N
from .. import engine as E
from ..engine import BF16, F32, Act, Ctx, Tape
from ..modules import A_PAD_VALUE, AUDIO_DIM, CONTEXT_DIM, TOTAL_DIM, X_PAD_VALUE, UNet, _pack


class TrainStepFunction(torch.autograd.Function):
    """loss = masked_mse(unet(ca*x + cb*noise, a, t, c), ta*x + tb*noise)  (diffusion.py:96-111, rectified_flow.py:94-111).

    The noising is fused into the input packing kernel, the loss reads the bf16 prediction in its channels-last layout and
    backward starts from the fused loss-gradient kernel: no (B, 6, N) intermediate is materialised on the way.
    """

    @staticmethod
    def forward(fctx, unet: UNet, x, a, t_cond, c, keep, noise, ca, cb, ta, tb, orig_len, *params):
        tape = Tape()
        x = x.contiguous().float()
        noise = noise.contiguous().float()
        out16, (ctx, xf) = unet.run(tape, x, a, t_cond, c, keep, noise=noise, ca=ca, cb=cb)
        B, Cc, n = x.shape
        accum = torch.empty(2, dtype=F32, device=x.device)
        loss = torch.empty(1, dtype=F32, device=x.device)
        bs, ld = E._bl(out16)
        ol = None if orig_len is None else orig_len.to(device=x.device, dtype=torch.int64).contiguous()
        N.call("of_mse_fwd", out16.data_ptr(), ld, bs, x.data_ptr(), noise.data_ptr(), float(ta), float(tb), E._p(ol), B, Cc, n,
               accum.data_ptr(), loss.data_ptr())
        fctx.saved = (unet, ctx, xf, out16, x, noise, ta, tb, ol, accum, params)
        return loss[0]

    @staticmethod
    def backward(fctx, gloss):
        unet, ctx, xf, out16, x, noise, ta, tb, ol, accum, params = fctx.saved
        fctx.saved = None
        B, Cc, n = x.shape
        Lp = out16.shape[1]
        bs, ld = E._bl(out16)
        g = gloss.detach().to(F32).reshape(1).contiguous()
        if unet.grad_prescale != 1.0:
            g = g * unet.grad_prescale
        dY16 = torch.empty((B, Lp, 8), dtype=BF16, device=x.device)
        N.call("of_mse_bwd", out16.data_ptr(), ld, bs, x.data_ptr(), noise.data_ptr(), float(ta), float(tb), E._p(ol), B, Cc, n, Lp, 8,
               accum.data_ptr(), g.data_ptr(), dY16.data_ptr())
        grads = unet.backward_from(ctx, xf, dY16, params)
        return (None,) * 12 + tuple(grads)


class BaseOsuFusion(nn.Module):
    def __init__(self, dim_h, dim_h_mult, num_layer_blocks, num_middle_transformers, cross_embed_kernel_sizes, attn_dim_head,
                 attn_heads, attn_kv_heads, attn_context_len, cond_drop_prob) -> None:
        super().__init__()
        self.unet = UNet(TOTAL_DIM, AUDIO_DIM, CONTEXT_DIM, dim_h, dim_h_mult=dim_h_mult, num_layer_blocks=num_layer_blocks,
                         num_middle_transformers=num_middle_transformers, cross_embed_kernel_sizes=cross_embed_kernel_sizes,
                         attn_dim_head=attn_dim_head, attn_heads=attn_heads, attn_kv_heads=attn_kv_heads,
                         attn_context_len=attn_context_len)
        self.cond_drop_prob = cond_drop_prob

    def set_full_bf16(self) -> None:
        """diffusion.py:56-57.  The engine always computes in bf16 with fp32 master weights; kept for API compatibility."""

    def _train_step(self, x, a, t_cond, c, noise, ca, cb, ta, tb, orig_len, cond_mask):
        assert x.shape[-1] == a.shape[-1], "x and a must have the same number of sequence length"
        if cond_mask is None:
            from ..modules import prob_mask_like
            cond_mask = prob_mask_like((x.shape[0],), 1.0 - self.cond_drop_prob, x.device)
        params = list(self.unet.parameters())
        return TrainStepFunction.apply(self.unet, x, a, t_cond, c, cond_mask, noise, ca, cb, ta, tb, orig_len, *params)

    # ------------------------------------------------------------------ sampling
    class _SamplerState:
        pass

    def _sampler_setup(self, a, c, x, cond_scale):
        """Pack inputs once, run the loop-invariant audio encoder once, and duplicate to a 2B CFG batch if needed."""
        unet = self.unet
        b, _, n = a.shape
        dev = a.device
        if x is None:
            x = torch.randn((b, TOTAL_DIM, n), device=dev)
        x = x.contiguous().float()
        Lp = unet.padded_len(n)
        cfg = cond_scale != 1.0
        ctx = Ctx(dev, unet._store, None)
        ctx.attn_variant = unet.attn_variant
        unet._store.begin_forward(False, unet)
        a16 = _pack(a.float(), AUDIO_DIM, Lp, A_PAD_VALUE)
        a_feat = unet.encode_audio(ctx, a16)
        cc = c.float()
        keep = torch.ones(b, dtype=torch.bool, device=dev)
        if cfg:
            a_feat = Act(None, torch.cat([a_feat.bf16, a_feat.bf16], 0))
            cc = torch.cat([cc, cc], 0)
            keep = torch.cat([keep, ~keep], 0)
        x16 = _pack(x, 8, Lp, X_PAD_VALUE)
        s = BaseOsuFusion._SamplerState()
        s.ctx, s.a_feat, s.c, s.keep, s.cfg, s.x, s.x16, s.b, s.n, s.Lp = ctx, a_feat, cc, keep, cfg, x, x16, b, n, Lp
        return s

    def _eval_denoiser(self, s, x16, t_batched):
        """One CFG evaluation = ONE denoiser pass over the 2B batch [cond; null]; returns (cond16, null16|None)."""
        unet = self.unet
        if s.cfg:
            x16 = torch.cat([x16, x16], 0)
            t_batched = torch.cat([t_batched, t_batched], 0)
        out16, _ = unet.denoise(s.ctx, x16, s.a_feat, t_batched, s.c, s.keep)
        if s.cfg:
            return out16[:s.b], out16[s.b:]
        return out16, None

    # ------------------------------------------------------------------ sampling, one CUDA graph per step
    def _graph_key(self, s, cond_scale, tag):
        st = self.unet._store
        vers = sum(p._version for p in self.unet.parameters())
        return (tag, s.b, s.n, bool(s.cfg), float(cond_scale), str(s.x.device), st.param_epoch, vers)

    def _sampler_graphs(self, s, cond_scale, tag, kinds):
        """CUDA graphs of one sampler step each: `[cond; null]` denoiser evaluation (~1.3 k launches) + fused CFG / update kernel,
        with the timestep and the update coefficients read from device buffers, so the same graph is replayed for every step of
        the loop (diffusion.py:71-75, rectified_flow.py:69-79) and for every later `sample()` call of the same shape.
        `kinds`: one (mode, src, dst_x, dst_packed) tuple per graph over the state's static buffers.  Returns None when graphs are
        disabled (OF_SAMPLER_GRAPH=0)."""
        if os.environ.get("OF_SAMPLER_GRAPH", "1") == "0" or not s.x.is_cuda:
            return None
        cache = self.__dict__.setdefault("_sgraphs", {})
        key = self._graph_key(s, cond_scale, tag)
        hit = cache.get(key)
        if hit is None:
            cache.clear()                      # weights / shapes changed: the old graphs reference stale operand buffers
            dev = s.x.device
            st = BaseOsuFusion._SamplerState()
            st.ctx, st.cfg, st.b, st.n, st.Lp = s.ctx, s.cfg, s.b, s.n, s.Lp
            st.a_feat = Act(None, s.a_feat.bf16.clone())
            st.c, st.keep = s.c.clone(), s.keep.clone()
            st.x = torch.empty_like(s.x)
            st.xtmp = torch.empty_like(s.x)
            st.x16 = torch.empty_like(s.x16)
            st.xmid16 = torch.empty_like(s.x16)
            st.t_buf = torch.zeros(s.b, dtype=F32, device=dev)
            st.coef = torch.tensor([0.0, 1.0, 1.0, 0.0], dtype=F32).to(dev)     # identity update during the warm-up evaluation
            st.x.copy_(s.x)
            st.x16.copy_(s.x16)
            st.xmid16.copy_(s.x16)

            def body(mode, src, dst_x, dst_packed):
                # a fresh pool of zero-initialised scratch per evaluation: its chunk fills must be nodes of the captured graph
                # (a chunk zeroed before the capture would carry the previous replay's accumulators)
                st.ctx.zpool = E.ZeroPool(dev)
                E.use_pool(st.ctx.zpool)
                cond16, null16 = self._eval_denoiser(st, getattr(st, src), st.t_buf)
                bs, ld = E._bl(cond16)
                N.call("of_sampler_update_dev", st.x.data_ptr(), cond16.data_ptr(), E._p(null16), ld, bs, float(cond_scale), mode,
                       st.coef.data_ptr(), st.b, TOTAL_DIM, st.n, getattr(st, dst_x).data_ptr(), getattr(st, dst_packed).data_ptr(),
                       st.Lp, 8, X_PAD_VALUE)

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for k in kinds:                 # warm-up: populates operand / RoPE / FiLM-plan caches outside the capture
                    body(*k)
            torch.cuda.current_stream().wait_stream(side)
            graphs = []
            for k in kinds:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    body(*k)
                graphs.append(g)
            hit = cache[key] = (st, graphs)
        st, graphs = hit
        st.a_feat.bf16.copy_(s.a_feat.bf16)
        st.c.copy_(s.c)
        st.x.copy_(s.x)
        st.x16.copy_(s.x16)
        return st, graphs

    def _update(self, s, xin, cond16, null16, cond_scale, mode, c_eps, c_div=1.0, c_x0=0.0, c_dir=0.0):
        xout = torch.empty_like(xin)
        packed = torch.empty((s.b, s.Lp, 8), dtype=BF16, device=xin.device)
        bs, ld = E._bl(cond16)
        N.call("of_sampler_update", xin.data_ptr(), cond16.data_ptr(), E._p(null16), ld, bs, float(cond_scale), mode, float(c_eps),
               float(c_div), float(c_x0), float(c_dir), s.b, TOTAL_DIM, s.n, xout.data_ptr(), packed.data_ptr(), s.Lp, 8,
               X_PAD_VALUE)
        return xout, packed
