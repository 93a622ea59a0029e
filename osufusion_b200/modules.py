"""nn.Module tree of the B200 denoiser — a drop-in for the reference's `osu_fusion.modules.unet.UNet` API surface.

Module/attribute names, constructor signatures and state_dict keys/shapes mirror the reference exactly (unet.py:26-513,
residual.py:14-137, attention.py:15-58) so `load_state_dict` of a reference checkpoint works unchanged and adapter
injectors can target `attn.to_q`, `attn.to_kv` (nn.Linear) and `block1.proj`, `block2.proj` (nn.Conv1d).  The leaf
nn.Conv1d / nn.Linear / nn.GroupNorm / nn.LayerNorm modules are PARAMETER CONTAINERS only: compute never goes through
their torch forward — it runs in libosufusion_sm100.so via engine.py, and raises if that library is unavailable.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

from . import _native as N
from . import engine as E
from . import ops_raw as R
from .engine import BF16, F32, Act, Ctx, ParamStore, Tape

TOTAL_DIM, AUDIO_DIM, CONTEXT_DIM = 6, 96, 5   # encode.py:24-26, scripts/dataset_creator.py:22-25
X_PAD_VALUE, A_PAD_VALUE = -1.0, -23.0          # unet.py:479-480


def prob_mask_like(shape, prob: float, device) -> torch.Tensor:
    """utils.py:15-21 (same RNG consumption: uniform_ is drawn only for 0 < prob < 1)."""
    if prob == 0.0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    if prob == 1.0:
        return torch.ones(shape, device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).uniform_(0.0, 1.0) < prob


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is executed by the CUDA engine (osufusion_b200.engine), not by torch")


class SinusoidalPositionEmbedding(_Container):
    def __init__(self, dim: int, theta: int = 10000) -> None:
        super().__init__()
        self.dim, self.theta = dim, theta


class CrossEmbedLayer(_Container):
    def __init__(self, dim: int, dim_out: int, kernel_sizes: Sequence[int]) -> None:
        super().__init__()
        ks = sorted(kernel_sizes)
        widths = [int(dim / (2 ** i)) for i in range(1, len(ks))]
        widths.append(dim_out - sum(widths))
        self.convs = nn.ModuleList([nn.Conv1d(dim, w, k, padding=k // 2) for k, w in zip(ks, widths)])


class Upsample(_Container):
    def __init__(self, dim_in: int, dim_out: int) -> None:
        super().__init__()
        self.conv = nn.Conv1d(dim_in, dim_out, 3, padding=1)


class Downsample(_Container):
    def __init__(self, dim_in: int, dim_out: int) -> None:
        super().__init__()
        self.conv = nn.Conv1d(dim_in, dim_out, 3, stride=2, padding=0)


class Parallel(_Container):
    def __init__(self, *fns: nn.Module) -> None:
        super().__init__()
        self.fns = nn.ModuleList(fns)


class RotaryPositionEmbedding(_Container):
    def __init__(self, dim: int, theta: int = 10000, scale_base: int = 4096) -> None:
        super().__init__()
        self.scale_base = scale_base
        inv_freq = 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)


class Attend(_Container):
    pass


class Attention(_Container):
    def __init__(self, dim_in: int, dim_head: int, heads: int, kv_heads: int, context_len: int = 4096) -> None:
        super().__init__()
        self.heads, self.kv_heads, self.dim_head = heads, kv_heads, dim_head
        self.norm = nn.LayerNorm(dim_in)
        self.to_q = nn.Linear(dim_in, dim_head * heads, bias=False)
        self.to_kv = nn.Linear(dim_in, dim_head * kv_heads * 2, bias=False)
        self.rotary_emb = RotaryPositionEmbedding(dim_head, scale_base=context_len)
        self.attn = Attend()
        self.to_out = nn.Linear(dim_head * heads, dim_in)


class FeedForward(nn.Sequential):
    def __init__(self, dim: int, dim_mult: int = 2) -> None:
        super().__init__(nn.Linear(dim, dim * dim_mult), nn.SiLU(), nn.Linear(dim * dim_mult, dim))


class TransformerBlock(_Container):
    def __init__(self, dim: int, ff_mult: int = 2, attn_dim_head: int = 64, attn_heads: int = 16, attn_kv_heads: int = 1,
                 attn_context_len: int = 4096) -> None:
        super().__init__()
        self.attn = Attention(dim, attn_dim_head, attn_heads, attn_kv_heads, attn_context_len)
        self.ff = FeedForward(dim, ff_mult)


class GlobalContext(_Container):
    def __init__(self, dim_in: int, dim_out: int, reduction: int = 2, dim_min: int = 8) -> None:
        super().__init__()
        self.to_k = nn.Conv1d(dim_in, 1, 1)
        inner = max(dim_min, dim_out // reduction)
        self.layers = nn.Sequential(nn.Conv1d(dim_in, inner, 1), nn.SiLU(), nn.Conv1d(inner, dim_out, 1), nn.Sigmoid())


class Block(_Container):
    def __init__(self, dim_in: int, dim_out: int, norm: bool = True) -> None:
        super().__init__()
        self.proj = nn.Conv1d(dim_in, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(1, dim_out) if norm else nn.Identity()
        self.activation = nn.SiLU()


class ResidualBlock(_Container):
    def __init__(self, dim_in: int, dim_out: int, dim_time: Optional[int] = None, dim_cond: Optional[int] = None) -> None:
        super().__init__()
        self.mlp = (nn.Sequential(nn.SiLU(), nn.Linear(int(dim_time) + int(dim_cond), dim_out * 2))
                    if (dim_time or dim_cond) else None)
        self.block1 = Block(dim_in, dim_out)
        self.block2 = Block(dim_out, dim_out)
        self.res_conv = nn.Conv1d(dim_in, dim_out, 1) if dim_in != dim_out else nn.Identity()
        self.se = GlobalContext(dim_out, dim_out)


class UNetBlock(_Container):
    def __init__(self, dim_in: int, dim_out: int, dim_time, dim_cond, layer_idx: int, num_layers: int, num_blocks: int,
                 down_block: bool, attn_dim_head: int, attn_heads: int, attn_kv_heads: int, attn_context_len: int) -> None:
        super().__init__()
        self.init_resnet = ResidualBlock(dim_in if down_block else dim_in + dim_out, dim_in, dim_time, dim_cond)
        self.resnets = nn.ModuleList([ResidualBlock(dim_in, dim_in, dim_time, dim_cond) for _ in range(num_blocks)])
        self.transformers = nn.ModuleList([
            TransformerBlock(dim_in, attn_dim_head=attn_dim_head, attn_heads=attn_heads, attn_kv_heads=attn_kv_heads,
                             attn_context_len=attn_context_len) for _ in range(num_blocks)])
        if layer_idx >= num_layers - 1:
            self.sampler = Parallel(nn.Conv1d(dim_in, dim_out, 3, padding=1), nn.Conv1d(dim_in, dim_out, 1))
            self.sampler_kind = "parallel"
        elif down_block:
            self.sampler = Downsample(dim_in, dim_out)
            self.sampler_kind = "down"
        else:
            self.sampler = Upsample(dim_in, dim_out)
            self.sampler_kind = "up"
        self.gradient_checkpointing = False


def _level_dims(dim_h: int, mult: Sequence[int]):
    dims = (dim_h, *[dim_h * m for m in mult])
    return list(zip(dims[:-1], dims[1:]))


class AudioEncoder(_Container):
    def __init__(self, dim_in: int, dim_h: int, dim_h_mult=(1, 2, 3, 4), num_layer_blocks=(3, 3, 3, 3),
                 cross_embed_kernel_sizes=(3, 7, 15), attn_dim_head: int = 64, attn_heads: int = 16, attn_kv_heads: int = 1,
                 attn_context_len: int = 4096) -> None:
        super().__init__()
        self.init_conv = CrossEmbedLayer(dim_in, dim_h, cross_embed_kernel_sizes)
        io = _level_dims(dim_h, dim_h_mult)
        self.layers = nn.ModuleList([
            UNetBlock(i_, o_, None, None, i, len(io), num_layer_blocks[i], True, attn_dim_head, attn_heads, attn_kv_heads,
                      attn_context_len // (2 ** i)) for i, (i_, o_) in enumerate(io)])


def _pack(x: torch.Tensor, Cp: int, Lp: int, pad: float, noise=None, ca=None, cb=None) -> torch.Tensor:
    """(B, C, N) fp32 channel-first -> (B, Lp, Cp) bf16 channels-last (optionally ca*x + cb*noise), right-padded."""
    B, Cc, n = x.shape
    x = x.contiguous().float()
    out = torch.empty((B, Lp, Cp), dtype=BF16, device=x.device)
    N.call("of_pack_input", x.data_ptr(), E._p(noise), E._p(ca), E._p(cb), B, Cc, n, out.data_ptr(), Lp, Cp, pad)
    return out


class UNet(nn.Module):
    """Drop-in for osu_fusion.modules.unet.UNet (unet.py:321-513); `forward(x, a, t, c, cond_drop_prob)`."""

    def __init__(self, dim_in_x: int, dim_in_a: int, dim_in_c: int, dim_h: int, dim_h_mult=(1, 2, 3, 4),
                 num_layer_blocks=(3, 3, 3, 3), num_middle_transformers: int = 3, cross_embed_kernel_sizes=(3, 7, 15),
                 attn_dim_head: int = 64, attn_heads: int = 16, attn_kv_heads: int = 1, attn_context_len: int = 4096) -> None:
        super().__init__()
        self.dim_in_x, self.dim_in_a, self.dim_in_c = dim_in_x, dim_in_a, dim_in_c
        self.dim_h, self.dim_emb, self.attn_context_len = dim_h, dim_h * 4, attn_context_len
        Em = self.dim_emb
        self.init_x = CrossEmbedLayer(dim_in_x, dim_h, cross_embed_kernel_sizes)
        self.audio_encoder = AudioEncoder(dim_in_a, dim_h, dim_h_mult=dim_h_mult, num_layer_blocks=num_layer_blocks,
                                          cross_embed_kernel_sizes=cross_embed_kernel_sizes, attn_dim_head=attn_dim_head,
                                          attn_heads=attn_heads, attn_kv_heads=attn_kv_heads)
        self.final_resnet = ResidualBlock(dim_h * 2, dim_h, Em, Em)
        self.final_conv = nn.Conv1d(dim_h, dim_in_x, 1)
        nn.init.zeros_(self.final_conv.weight)   # unet.py:18-23,354
        nn.init.zeros_(self.final_conv.bias)
        self.time_mlp = nn.Sequential(SinusoidalPositionEmbedding(Em), nn.Linear(Em, Em), nn.SiLU(), nn.Linear(Em, Em))
        self.cond_mlp = nn.Sequential(nn.Linear(dim_in_c, Em), nn.SiLU(), nn.Linear(Em, Em))
        self.null_cond = nn.Parameter(torch.randn(Em))
        io = _level_dims(dim_h, dim_h_mult)
        n = len(io)
        kw = dict(attn_dim_head=attn_dim_head, attn_heads=attn_heads, attn_kv_heads=attn_kv_heads)
        self.down_layers = nn.ModuleList([
            UNetBlock(i_, o_, Em, Em, i, n, num_layer_blocks[i], True, attn_context_len=attn_context_len // (2 ** i), **kw)
            for i, (i_, o_) in enumerate(io)])
        top = io[-1][1]
        self.middle_resnet1 = ResidualBlock(top * 2, top, Em, Em)
        self.middle_transformer = nn.ModuleList([
            TransformerBlock(top, attn_context_len=attn_context_len // (2 ** (n - 1)), **kw)
            for _ in range(num_middle_transformers)])
        self.middle_resnet2 = ResidualBlock(top, top, Em, Em)
        rio, rblocks = list(reversed(io)), list(reversed(num_layer_blocks))
        self.up_layers = nn.ModuleList([
            UNetBlock(hi, lo, Em, Em, i, n, rblocks[i], False, attn_context_len=attn_context_len // (2 ** (n - i - 1)), **kw)
            for i, (lo, hi) in enumerate(rio)])
        self._store = ParamStore()
        self.attn_variant = 0
        self.grad_sync = None    # set by osufusion_b200.ddp.GradAllReducer: called after every backward tape op
        self.grad_finish = None  # ... and once at the end of backward (launch remaining buckets, join the comm stream)
        self.grad_prescale = 1.0 # set by ddp.GradAllReducer: loss-gradient scale 1/world when buckets are reduced with SUM
        self._side_stream = None # second stream of backward (weight / bias gradients), created on first use

    # ------------------------------------------------------------------ reference API
    def set_gradient_checkpointing(self, value: bool) -> None:
        """unet.py:452-456.  Accepted for API compatibility; activations fit comfortably in 180 GB so nothing is recomputed."""
        for _, m in self.named_modules():
            if hasattr(m, "gradient_checkpointing"):
                m.gradient_checkpointing = value

    def forward_with_cond_scale(self, *args, cond_scale: float = 1.0, **kwargs) -> torch.Tensor:
        cond = self(*args, **kwargs)
        if cond_scale == 1.0:
            return cond
        null = self(*args, **kwargs, cond_drop_prob=1.0)
        return null + (cond - null) * cond_scale

    def forward(self, x, a, t, c, cond_drop_prob: float = 0.0, cond_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("osufusion_b200.UNet runs only on CUDA (sm_100a); there is no CPU path")
        if cond_mask is None:
            cond_mask = prob_mask_like((x.shape[0],), 1.0 - cond_drop_prob, x.device)
        params = list(self.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return UNetFunction.apply(self, x, a, t, c, cond_mask, *params)
        out16, _ = self.run(None, x, a, t, c, cond_mask)
        return self.unpack(out16, x.shape[-1])

    # ------------------------------------------------------------------ engine entry points
    def padded_len(self, n: int) -> int:
        depth = len(self.down_layers)
        return n + ((-n) % (2 ** depth))

    def unpack(self, out16: torch.Tensor, n: int) -> torch.Tensor:
        B = out16.shape[0]
        y = torch.empty((B, self.dim_in_x, n), dtype=F32, device=out16.device)
        bs, ld = E._bl(out16)
        N.call("of_unpack_output", out16.data_ptr(), ld, bs, B, self.dim_in_x, n, y.data_ptr())
        return y

    def encode_audio(self, ctx: Ctx, a16: torch.Tensor) -> Act:
        """AudioEncoder.forward (unet.py:314-318) on the packed (B, Lp, 96) bf16 spectrogram."""
        h = E.cross_embed(ctx, self.audio_encoder.init_conv, a16)
        for layer in self.audio_encoder.layers:
            h, _ = E.unet_block(ctx, layer, h)
        return h

    def conditioning(self, ctx: Ctx, t: torch.Tensor, c: torch.Tensor, keep: torch.Tensor) -> None:
        """time_mlp / cond_mlp / null_cond select / SiLU(cat(t, c)) (unet.py:484-493, residual.py:104-105)."""
        st, dev = ctx.store, ctx.device
        B = t.shape[0]
        Em = self.dim_emb
        tf = t.to(F32).contiguous()
        temb = E.empty((B, Em), F32, dev)
        N.call("of_time_embed", tf.data_ptr(), B, Em, float(self.time_mlp[0].theta), temb.data_ptr())
        l1, l3 = self.time_mlp[1], self.time_mlp[3]
        t1, t1pre = E.linear_small_fwd(temb, l1.weight, l1.bias, act=1, want_pre=True)
        tvec, _ = E.linear_small_fwd(t1, l3.weight, l3.bias)
        c0, c2 = self.cond_mlp[0], self.cond_mlp[2]
        cin = c.to(F32).contiguous()
        c1, c1pre = E.linear_small_fwd(cin, c0.weight, c0.bias, act=1, want_pre=True)
        cvec, _ = E.linear_small_fwd(c1, c2.weight, c2.bias)
        csel = torch.where(keep[:, None], cvec, self.null_cond.detach()[None, :].to(F32))
        cat = torch.cat([tvec, csel], dim=1).contiguous()
        emb_act = E.empty((B, 2 * Em), F32, dev)
        N.call("of_silu_small", cat.data_ptr(), None, emb_act.data_ptr(), cat.numel())
        ctx.emb_act = emb_act
        # every FiLM head of the denoiser in one launch (their common input is emb_act; bf16-rounded as under autocast)
        training = ctx.tape is not None
        if training:
            st.ensure_arena(self)
        plan = st.film_plan(self, B, training)
        emb_r = None
        if plan is not None:
            emb_r = emb_act.to(BF16).to(F32)
            ss_all = E.empty((B * plan["rows"],), F32, dev)
            N.call("of_film_fwd", plan["groups"].data_ptr(), plan["num_groups"], plan["rows"], emb_r.data_ptr(), B, plan["K"],
                   ss_all.data_ptr())
            ctx.film = {k: ss_all[o:o + B * n].view(B, n) for k, (o, n) in plan["slices"].items()}
        if training:
            ctx.d_emb_act = E.zeros((B, 2 * Em), F32, dev)
            dss_all = None
            if plan is not None:
                dss_all = torch.zeros((B * plan["rows"],), dtype=F32, device=dev)
                ctx.film_dss = {k: dss_all[o:o + B * n].view(B, n) for k, (o, n) in plan["slices"].items()}

            ctx.film_bwd = None
            if plan is not None:
                def film_bwd(heads):
                    """Weight / bias gradients of the given FiLM heads (one grouped launch over their chunk range; consecutive in the
                    plan) and their contribution to d emb_act; called right after the backward of the unit that owns them."""
                    c0 = min(plan["chunk_range"][id(h)][0] for h in heads)
                    c1 = max(plan["chunk_range"][id(h)][1] for h in heads)
                    N.call("of_film_bwd", plan["groups"].data_ptr(), plan["chunks"].data_ptr() + 8 * c0, c1 - c0,
                           dss_all.data_ptr(), emb_r.data_ptr(), B, plan["K"], ctx.d_emb_act.data_ptr())
                    for h in heads:
                        for p_ in h.mlp[1].parameters():
                            if p_.requires_grad:
                                st.touch(p_)
                ctx.film_bwd = film_bwd

            def backward():
                if plan is not None and self.grad_sync is None:
                    ctx.film_bwd(plan["heads"])
                dcat = E.empty((B, 2 * Em), F32, dev)
                N.call("of_silu_small", cat.data_ptr(), ctx.d_emb_act.data_ptr(), dcat.data_ptr(), cat.numel())
                dt = dcat[:, :Em]
                dcs = dcat[:, Em:]
                if self.null_cond.requires_grad:
                    st.set_grad(self.null_cond, (dcs * (~keep)[:, None]).sum(0))
                dcv = (dcs * keep[:, None]).contiguous()
                dc1 = E.zeros((B, Em), F32, dev)
                E.linear_small_bwd_param(st, dcv, None, 0, c1, c2.weight, c2.bias, dc1)
                E.linear_small_bwd_param(st, dc1, c1pre, 1, cin, c0.weight, c0.bias, None)
                dt1 = E.zeros((B, Em), F32, dev)
                E.linear_small_bwd_param(st, dt, None, 0, t1, l3.weight, l3.bias, dt1)
                E.linear_small_bwd_param(st, dt1, t1pre, 1, temb, l1.weight, l1.bias, None)
            ctx.tape.push(backward)

    def _film_unit(self, ctx: Ctx, modules) -> None:
        """Tape marker placed BEFORE a unit's forward ops: in backward it runs right after the unit and produces the weight
        gradients of the unit's FiLM heads, so they (and their all-reduce bucket) complete progressively instead of at the end."""
        if ctx.tape is None or ctx.film_bwd is None or self.grad_sync is None:
            return         # single GPU: one grouped launch over ALL heads at the end of backward (conditioning) is cheaper
        heads = [m for u in modules for m in u.modules() if type(m).__name__ == "ResidualBlock" and m.mlp is not None]
        if heads:
            ctx.tape.push(lambda: ctx.film_bwd(heads))

    def denoise(self, ctx: Ctx, x16: torch.Tensor, a_feat, t, c, keep) -> tuple:
        """Everything of UNet.forward after the input packing and the audio encoder.  Returns (out16 (B, Lp, 8), final Act).
        `a_feat` is the encoded audio (Act) or a callable producing it: the training step evaluates the audio encoder AFTER the
        down path (it depends on neither x nor t), so that its backward runs right after `middle_resnet1` and the cheap first
        down block — not the 26 % audio encoder — is the un-overlappable tail of the gradient all-reduce."""
        st, dev = ctx.store, ctx.device
        self.conditioning(ctx, t, c, keep)
        # the audio encoder and the down path are independent until `middle_resnet1`: the training step runs the encoder on the
        # second stream (a parallel branch of the captured graph), so each branch fills the other's under-filled launches
        fork = None
        if callable(a_feat) and ctx.tape is not None and E.FWD_SIDE and dev.type == "cuda":
            fork = torch.cuda.Event()
            fork.record()
        x = E.cross_embed(ctx, self.init_x, x16)
        r = x
        skips = []
        for layer in self.down_layers:
            self._film_unit(ctx, [layer])
            x, s = E.unet_block(ctx, layer, x)
            skips.append(s)
        if callable(a_feat) and fork is not None:
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=dev)
            side = self._side_stream
            side.wait_event(fork)
            main_pool, main_rope = ctx.zpool, ctx.rope_cache
            side_pool = E.ZeroPool(dev)
            try:
                with torch.cuda.stream(side):
                    # scratch chunks are zero-filled on the stream that creates them and the RoPE tables are generated lazily: the
                    # branch gets its own so that it never reads what the main stream produced after the fork
                    ctx.zpool, ctx.rope_cache = side_pool, {}
                    E.use_pool(side_pool)
                    a_feat = a_feat()
            finally:
                ctx.zpool, ctx.rope_cache = main_pool, main_rope
                E.use_pool(main_pool)
            torch.cuda.current_stream().wait_stream(side)
        elif callable(a_feat):
            a_feat = a_feat()
        self._film_unit(ctx, [self.middle_resnet1, self.middle_resnet2])
        x = E.concat(ctx, x, a_feat)
        x = E.residual_block(ctx, self.middle_resnet1, x)
        for tr in self.middle_transformer:
            x = E.transformer_block(ctx, tr, x)
        x = E.residual_block(ctx, self.middle_resnet2, x)
        for layer in self.up_layers:
            self._film_unit(ctx, [layer])
            x = E.concat(ctx, x, skips.pop())
            x, _ = E.unet_block(ctx, layer, x)
        self._film_unit(ctx, [self.final_resnet])
        x = E.concat(ctx, x, r)
        x = E.residual_block(ctx, self.final_resnet, x)
        B, Lp, Cc = x.bf16.shape
        wf = st.padded_rows_w(self.final_conv.weight, 8)
        bf = torch.zeros(8, dtype=F32, device=dev)
        bf[:self.dim_in_x] = self.final_conv.bias.detach()
        out16 = E.empty((B, Lp, 8), BF16, dev)
        R.gemm_fwd(x.bf16, wf.view(1, 8, Cc), N_out=8, K=Cc, bias=bf, out_bf16=out16)
        return out16, x

    def final_backward(self, ctx: Ctx, xf: Act, dY16: torch.Tensor) -> None:
        st, dev = ctx.store, ctx.device
        fc = self.final_conv
        Cc = xf.bf16.shape[2]
        if fc.weight.requires_grad:
            tmp = E.zeros((1, 8, Cc), F32, dev)
            R.gemm_wgrad(dY16, xf.bf16, tmp, M=8, N_out=Cc)
            st.set_grad(fc.weight, tmp[0, :self.dim_in_x].reshape(fc.weight.shape).clone())
            db = E.zeros((8,), F32, dev)
            E.colsum(dY16, db)
            st.set_grad(fc.bias, db[:self.dim_in_x].clone())
        wf = st.padded_rows_w(fc.weight, 8)
        E._dgrad_into(xf, dY16, wf.view(1, 8, Cc), N_out=Cc, K=8)

    def run(self, tape: Optional[Tape], x, a, t, c, keep, *, noise=None, ca=None, cb=None, refresh: Optional[bool] = None):
        """Pack inputs, run the audio encoder and the denoiser.  Returns (out16, (ctx, final Act))."""
        n = x.shape[-1]
        Lp = self.padded_len(n)
        ctx = Ctx(x.device, self._store, tape)
        ctx.attn_variant = self.attn_variant
        self._store.begin_forward(refresh if refresh is not None else tape is not None, self)
        x16 = _pack(x, 8, Lp, X_PAD_VALUE, noise, ca, cb)
        a16 = _pack(a, self.dim_in_a, Lp, A_PAD_VALUE)
        out16, xf = self.denoise(ctx, x16, lambda: self.encode_audio(ctx, a16), t, c, keep)
        self._store.end_forward()
        return out16, (ctx, xf)

    def backward_from(self, ctx: Ctx, xf: Act, dY16: torch.Tensor, params):
        st = ctx.store
        E.use_pool(ctx.zpool)
        st.begin_backward(self, film_overwritten=ctx.film_dss is not None)
        # weight / bias gradients go to a second stream (engine.SideLane): parallel branches of the captured step
        st.side = E.SideLane(ctx.device) if (E.WGRAD_SIDE and ctx.device.type == "cuda") else None
        if st.side is not None and self._side_stream is not None:
            st.side.stream = self._side_stream        # one stream per model, not one per step
        elif st.side is not None:
            self._side_stream = st.side.stream
        try:
            if self.grad_sync is not None:
                self.grad_sync(len(ctx.tape.ops) + 1)
            self.final_backward(ctx, xf, dY16)
            ctx.tape.run_backward(self.grad_sync, st.side)
            st.lora_finish()
            if self.grad_finish is not None:
                self.grad_finish()
        finally:
            if st.side is not None:
                st.side.join()
            st.side = None
        return st.take_grads(params)


class UNetFunction(torch.autograd.Function):
    """The whole denoiser as ONE autograd node: forward records the engine tape, backward replays it in reverse."""

    @staticmethod
    def forward(fctx, unet: UNet, x, a, t, c, keep, *params):
        tape = Tape()
        out16, (ctx, xf) = unet.run(tape, x, a, t, c, keep)
        fctx.unet, fctx.ctx, fctx.xf, fctx.params = unet, ctx, xf, params
        fctx.n, fctx.Lp = x.shape[-1], out16.shape[1]
        return unet.unpack(out16, x.shape[-1])

    @staticmethod
    def backward(fctx, dy):
        unet = fctx.unet
        if unet.grad_prescale != 1.0:
            dy = dy * unet.grad_prescale
        dY16 = _pack(dy, 8, fctx.Lp, 0.0)
        grads = unet.backward_from(fctx.ctx, fctx.xf, dY16, fctx.params)
        fctx.ctx = fctx.xf = None
        return (None, None, None, None, None, None, *grads)
