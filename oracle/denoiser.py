"""ORACLE (test infrastructure, not product code): CPU/PyTorch restatement of the OsuFusion denoiser.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (osufusion_b200/) never does.

What is restated (reference file:line, relative to /root/reference):
  * UNet, AudioEncoder, UNetBlock, TransformerBlock, Attention, FeedForward, CrossEmbedLayer, Upsample,
    Downsample, Parallel, SinusoidalPositionEmbedding, zero_init     osu_fusion/modules/unet.py:18-513
  * ResidualBlock, Block, GlobalContext                               osu_fusion/modules/residual.py:14-37,62-137
  * RotaryPositionEmbedding, Attend                                   osu_fusion/modules/attention.py:15-101
  * prob_mask_like, rotate_half, apply_rotary_pos_emb                 osu_fusion/modules/utils.py:15-32

Parity pin: tests/test_oracle_vs_reference.py imports the real reference modules from /root/reference (when
present, i.e. in the build container) and checks this restatement against them with identical weights;
oracle/make_golden.py stores reference outputs as fixtures under tests/golden/ for boxes without the reference.
The reference itself ships no tests or golden vectors.

Module/attribute names mirror the reference so that state_dict keys are identical (1239 keys at dim_h=512).
The implementation style is intentionally independent: plain functional torch, no einops.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn


# --------------------------------------------------------------------------------------------- helpers
def cfg_keep_mask(batch: int, keep_prob: float, device) -> torch.Tensor:
    """utils.py:15-21 — True = keep the conditioning.  RNG is consumed only for 0 < keep_prob < 1."""
    if keep_prob == 0.0:
        return torch.zeros((batch,), device=device, dtype=torch.bool)
    if keep_prob == 1.0:
        return torch.ones((batch,), device=device, dtype=torch.bool)
    return torch.zeros((batch,), device=device).uniform_(0.0, 1.0) < keep_prob


def _half_rotate(x: torch.Tensor) -> torch.Tensor:
    """utils.py:25-27 — (x1, x2) -> (-x2, x1) on the last dim (half-split convention)."""
    d = x.shape[-1] // 2
    return torch.cat((-x[..., d:], x[..., :d]), dim=-1)


def rope_apply(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """utils.py:30-32."""
    return (x * cos) + (_half_rotate(x) * sin)


def rope_tables(seq_len: int, dim: int, scale_base: int, dtype, device, theta: float = 10000.0):
    """attention.py:24-49 — tables are built IN THE DTYPE OF q (bf16 under CUDA autocast), with position
    interpolation t *= scale_base / seq_len.  Returns cos, sin of shape (seq_len, dim)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, dim, 2, device=device).float() / dim))
    t = torch.arange(seq_len, dtype=dtype, device=device)
    t *= scale_base / seq_len
    freqs = torch.einsum("i,j->ij", t, inv_freq.to(dtype))
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


# --------------------------------------------------------------------------------------------- leaf modules
class SinusoidalPositionEmbedding(nn.Module):  # unet.py:26-39
    def __init__(self, dim: int, theta: int = 10000) -> None:
        super().__init__()
        self.dim, self.theta = dim, theta

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        half = self.dim // 2
        step = math.log(self.theta) / (half - 1)
        freqs = torch.exp(torch.arange(half, device=t.device) * -step)
        ang = t[:, None] * freqs[None, :]
        return torch.cat([ang.sin(), ang.cos()], dim=-1)


class CrossEmbedLayer(nn.Module):  # unet.py:42-58
    def __init__(self, dim: int, dim_out: int, kernel_sizes: Sequence[int]) -> None:
        super().__init__()
        ks = sorted(kernel_sizes)
        widths = [int(dim / (2 ** i)) for i in range(1, len(ks))]
        widths.append(dim_out - sum(widths))
        self.convs = nn.ModuleList([nn.Conv1d(dim, w, k, padding=k // 2) for k, w in zip(ks, widths)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.cat([c(x) for c in self.convs], dim=1)


class Upsample(nn.Module):  # unet.py:61-74
    def __init__(self, dim_in: int, dim_out: int) -> None:
        super().__init__()
        self.conv = nn.Conv1d(dim_in, dim_out, 3, padding=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Downsample(nn.Module):  # unet.py:77-92
    def __init__(self, dim_in: int, dim_out: int) -> None:
        super().__init__()
        self.conv = nn.Conv1d(dim_in, dim_out, 3, stride=2, padding=0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.conv(F.pad(x, (0, 1), mode="reflect"))


class Parallel(nn.Module):  # unet.py:95-101
    def __init__(self, *fns: nn.Module) -> None:
        super().__init__()
        self.fns = nn.ModuleList(fns)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self.fns[0](x)
        for f in self.fns[1:]:
            out = out + f(x)
        return out


class RotaryPositionEmbedding(nn.Module):  # attention.py:15-58
    def __init__(self, dim: int, theta: int = 10000, scale_base: int = 4096) -> None:
        super().__init__()
        self.dim, self.theta, self.scale_base = dim, theta, scale_base
        inv_freq = 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self._key = None
        self._tables = None

    def tables(self, q: torch.Tensor):
        key = (q.shape[-2], q.device, q.dtype)
        if key != self._key:
            # the reference decorates this with cuda.amp.autocast(dtype=float32), which merely DISABLES autocast
            with torch.autocast(device_type=q.device.type, enabled=False):
                t = torch.arange(q.shape[-2], dtype=q.dtype, device=q.device)
                t *= self.scale_base / q.shape[-2]
                freqs = torch.einsum("i,j->ij", t, self.inv_freq.to(q.dtype))
                emb = torch.cat([freqs, freqs], dim=-1)
                self._tables = (emb.cos()[None, None], emb.sin()[None, None])
            self._key = key
        return self._tables

    def forward(self, q: torch.Tensor, k: torch.Tensor):
        cos, sin = self.tables(q)
        with torch.autocast(device_type=q.device.type, enabled=False):
            return rope_apply(q, cos, sin), rope_apply(k, cos, sin)


class Attend(nn.Module):  # attention.py:61-101 — q,k,v are ALWAYS cast to bf16 around SDPA
    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
        dtype = v.dtype
        q, k, v = (t.to(torch.bfloat16).contiguous() for t in (q, k, v))
        return F.scaled_dot_product_attention(q, k, v).to(dtype)


class Attention(nn.Module):  # unet.py:104-146
    def __init__(self, dim_in: int, dim_head: int, heads: int, kv_heads: int, context_len: int = 4096) -> None:
        super().__init__()
        self.heads, self.kv_heads, self.dim_head = heads, kv_heads, dim_head
        self.norm = nn.LayerNorm(dim_in)
        self.to_q = nn.Linear(dim_in, dim_head * heads, bias=False)
        self.to_kv = nn.Linear(dim_in, dim_head * kv_heads * 2, bias=False)
        self.rotary_emb = RotaryPositionEmbedding(dim_head, scale_base=context_len)
        self.attn = Attend()
        self.to_out = nn.Linear(dim_head * heads, dim_in)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # x: (B, L, C)
        b, n, _ = x.shape
        x = self.norm(x)  # NB: the residual below is taken from the NORMED x (unet.py:127,141)
        q = self.to_q(x).view(b, n, self.heads, self.dim_head).transpose(1, 2)
        k, v = self.to_kv(x).chunk(2, dim=-1)
        k = k.reshape(b, n, self.kv_heads, self.dim_head).transpose(1, 2)
        v = v.reshape(b, n, self.kv_heads, self.dim_head).transpose(1, 2)
        rep = self.heads // self.kv_heads
        # einops "b h n d -> b (r h) n d": the repeat index is the OUTER factor
        k = k.repeat(1, rep, 1, 1)
        v = v.repeat(1, rep, 1, 1)
        q, k = self.rotary_emb(q, k)
        o = self.attn(q, k, v)
        o = o.transpose(1, 2).reshape(b, n, self.heads * self.dim_head)
        return x + self.to_out(o)


class FeedForward(nn.Sequential):  # unet.py:149-156
    def __init__(self, dim: int, dim_mult: int = 2) -> None:
        super().__init__(nn.Linear(dim, dim * dim_mult), nn.SiLU(), nn.Linear(dim * dim_mult, dim))


class TransformerBlock(nn.Module):  # unet.py:159-183
    def __init__(self, dim: int, ff_mult: int = 2, attn_dim_head: int = 64, attn_heads: int = 16,
                 attn_kv_heads: int = 1, attn_context_len: int = 4096) -> None:
        super().__init__()
        self.attn = Attention(dim, attn_dim_head, attn_heads, attn_kv_heads, attn_context_len)
        self.ff = FeedForward(dim, ff_mult)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # (B, C, L)
        y = self.attn(x.transpose(1, 2))
        y = self.ff(y) + y
        return y.transpose(1, 2)


class GlobalContext(nn.Module):  # residual.py:14-37
    def __init__(self, dim_in: int, dim_out: int, reduction: int = 2, dim_min: int = 8) -> None:
        super().__init__()
        self.to_k = nn.Conv1d(dim_in, 1, 1)
        inner = max(dim_min, dim_out // reduction)
        self.layers = nn.Sequential(nn.Conv1d(dim_in, inner, 1), nn.SiLU(), nn.Conv1d(inner, dim_out, 1), nn.Sigmoid())

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # (B, C, L) -> (B, C, 1)
        attn = self.to_k(x).softmax(dim=-1)
        pooled = torch.einsum("bid,bjd->bij", x, attn)
        return self.layers(pooled)


class Block(nn.Module):  # residual.py:62-88
    def __init__(self, dim_in: int, dim_out: int, norm: bool = True) -> None:
        super().__init__()
        self.proj = nn.Conv1d(dim_in, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(1, dim_out) if norm else nn.Identity()
        self.activation = nn.SiLU()

    def forward(self, x: torch.Tensor, scale_shift=None) -> torch.Tensor:
        x = self.norm(self.proj(x))
        if scale_shift is not None:
            scale, shift = scale_shift
            x = x * (scale + 1) + shift
        return self.activation(x)


class ResidualBlock(nn.Module):  # residual.py:91-137
    def __init__(self, dim_in: int, dim_out: int, dim_time: Optional[int] = None, dim_cond: Optional[int] = None) -> None:
        super().__init__()
        self.mlp = (
            nn.Sequential(nn.SiLU(), nn.Linear(int(dim_time) + int(dim_cond), dim_out * 2))
            if (dim_time or dim_cond) else None
        )
        self.block1 = Block(dim_in, dim_out)
        self.block2 = Block(dim_out, dim_out)
        self.res_conv = nn.Conv1d(dim_in, dim_out, 1) if dim_in != dim_out else nn.Identity()
        self.se = GlobalContext(dim_out, dim_out)

    def forward(self, x: torch.Tensor, t: Optional[torch.Tensor] = None, c: Optional[torch.Tensor] = None) -> torch.Tensor:
        scale_shift = None
        if self.mlp is not None:
            emb = self.mlp(torch.cat([e for e in (t, c) if e is not None], dim=-1))
            scale_shift = emb[:, :, None].chunk(2, dim=1)
        h = self.block2(self.block1(x, scale_shift))
        h = h * self.se(h)
        return h + self.res_conv(x)


class UNetBlock(nn.Module):  # unet.py:186-263
    def __init__(self, dim_in: int, dim_out: int, dim_time, dim_cond, layer_idx: int, num_layers: int, num_blocks: int,
                 down_block: bool, attn_dim_head: int, attn_heads: int, attn_kv_heads: int, attn_context_len: int) -> None:
        super().__init__()
        self.init_resnet = ResidualBlock(dim_in if down_block else dim_in + dim_out, dim_in, dim_time, dim_cond)
        self.resnets = nn.ModuleList([ResidualBlock(dim_in, dim_in, dim_time, dim_cond) for _ in range(num_blocks)])
        self.transformers = nn.ModuleList([
            TransformerBlock(dim_in, attn_dim_head=attn_dim_head, attn_heads=attn_heads, attn_kv_heads=attn_kv_heads,
                             attn_context_len=attn_context_len) for _ in range(num_blocks)])
        last = layer_idx >= num_layers - 1
        if last:
            self.sampler = Parallel(nn.Conv1d(dim_in, dim_out, 3, padding=1), nn.Conv1d(dim_in, dim_out, 1))
        else:
            self.sampler = Downsample(dim_in, dim_out) if down_block else Upsample(dim_in, dim_out)
        self.gradient_checkpointing = False

    def _body(self, x, t=None, c=None):
        x = self.init_resnet(x, t, c)
        for res, tr in zip(self.resnets, self.transformers):
            x = tr(res(x, t, c))
        return self.sampler(x), x

    def forward(self, x, t=None, c=None):
        if self.training and self.gradient_checkpointing:
            return torch.utils.checkpoint.checkpoint(self._body, x, t, c, use_reentrant=True)
        return self._body(x, t, c)


def _level_dims(dim_h: int, mult: Sequence[int]) -> List[Tuple[int, int]]:
    dims = (dim_h, *[dim_h * m for m in mult])
    return list(zip(dims[:-1], dims[1:]))


class AudioEncoder(nn.Module):  # unet.py:266-318
    def __init__(self, dim_in: int, dim_h: int, dim_h_mult=(1, 2, 3, 4), num_layer_blocks=(3, 3, 3, 3),
                 cross_embed_kernel_sizes=(3, 7, 15), attn_dim_head: int = 64, attn_heads: int = 16,
                 attn_kv_heads: int = 1, attn_context_len: int = 4096) -> None:
        super().__init__()
        self.init_conv = CrossEmbedLayer(dim_in, dim_h, cross_embed_kernel_sizes)
        io = _level_dims(dim_h, dim_h_mult)
        self.layers = nn.ModuleList([
            UNetBlock(i_, o_, None, None, i, len(io), num_layer_blocks[i], True, attn_dim_head, attn_heads, attn_kv_heads,
                      attn_context_len // (2 ** i)) for i, (i_, o_) in enumerate(io)])

    def forward(self, a: torch.Tensor) -> torch.Tensor:
        a = self.init_conv(a)
        for layer in self.layers:
            a, _ = layer(a)
        return a


class UNet(nn.Module):  # unet.py:321-513
    def __init__(self, dim_in_x: int, dim_in_a: int, dim_in_c: int, dim_h: int, dim_h_mult=(1, 2, 3, 4),
                 num_layer_blocks=(3, 3, 3, 3), num_middle_transformers: int = 3, cross_embed_kernel_sizes=(3, 7, 15),
                 attn_dim_head: int = 64, attn_heads: int = 16, attn_kv_heads: int = 1, attn_context_len: int = 4096) -> None:
        super().__init__()
        self.dim_h, self.dim_emb, self.attn_context_len = dim_h, dim_h * 4, attn_context_len
        E = self.dim_emb
        self.init_x = CrossEmbedLayer(dim_in_x, dim_h, cross_embed_kernel_sizes)
        # NB: the reference does not forward attn_context_len to the audio encoder (unet.py:343-352) -> default 4096
        self.audio_encoder = AudioEncoder(dim_in_a, dim_h, dim_h_mult=dim_h_mult, num_layer_blocks=num_layer_blocks,
                                          cross_embed_kernel_sizes=cross_embed_kernel_sizes, attn_dim_head=attn_dim_head,
                                          attn_heads=attn_heads, attn_kv_heads=attn_kv_heads)
        self.final_resnet = ResidualBlock(dim_h * 2, dim_h, E, E)
        self.final_conv = nn.Conv1d(dim_h, dim_in_x, 1)
        nn.init.zeros_(self.final_conv.weight)  # unet.py:18-23,354
        nn.init.zeros_(self.final_conv.bias)
        self.time_mlp = nn.Sequential(SinusoidalPositionEmbedding(E), nn.Linear(E, E), nn.SiLU(), nn.Linear(E, E))
        self.cond_mlp = nn.Sequential(nn.Linear(dim_in_c, E), nn.SiLU(), nn.Linear(E, E))
        self.null_cond = nn.Parameter(torch.randn(E))

        io = _level_dims(dim_h, dim_h_mult)
        n = len(io)
        kw = dict(attn_dim_head=attn_dim_head, attn_heads=attn_heads, attn_kv_heads=attn_kv_heads)
        self.down_layers = nn.ModuleList([
            UNetBlock(i_, o_, E, E, i, n, num_layer_blocks[i], True, attn_context_len=attn_context_len // (2 ** i), **kw)
            for i, (i_, o_) in enumerate(io)])
        top = io[-1][1]
        self.middle_resnet1 = ResidualBlock(top * 2, top, E, E)
        self.middle_transformer = nn.ModuleList([
            TransformerBlock(top, attn_context_len=attn_context_len // (2 ** (n - 1)), **kw)
            for _ in range(num_middle_transformers)])
        self.middle_resnet2 = ResidualBlock(top, top, E, E)
        rio = list(reversed(io))
        rblocks = list(reversed(num_layer_blocks))
        self.up_layers = nn.ModuleList([
            UNetBlock(hi, lo, E, E, i, n, rblocks[i], False, attn_context_len=attn_context_len // (2 ** (n - i - 1)), **kw)
            for i, (lo, hi) in enumerate(rio)])

    def set_gradient_checkpointing(self, value: bool) -> None:  # unet.py:452-456
        for _, m in self.named_modules():
            if hasattr(m, "gradient_checkpointing"):
                m.gradient_checkpointing = value

    def forward_with_cond_scale(self, *args, cond_scale: float = 1.0, **kwargs) -> torch.Tensor:  # unet.py:458-465
        cond = self(*args, **kwargs)
        if cond_scale == 1.0:
            return cond
        null = self(*args, **kwargs, cond_drop_prob=1.0)
        return null + (cond - null) * cond_scale

    def forward(self, x, a, t, c, cond_drop_prob: float = 0.0, cond_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """unet.py:467-513.  `cond_mask` (B,) bool, True = keep conditioning, is an oracle-only hook that injects the
        CFG dropout mask instead of drawing it (SURVEY §8c parity protocol)."""
        n = x.shape[-1]
        depth = len(self.down_layers)
        pad = (-n) % (2 ** depth)
        x = F.pad(x, (0, pad), value=-1.0)
        a = F.pad(a, (0, pad), value=-23.0)
        x = self.init_x(x)
        a = self.audio_encoder(a)
        t = self.time_mlp(t)
        r = x.clone()
        if cond_mask is None:
            cond_mask = cfg_keep_mask(x.shape[0], 1.0 - cond_drop_prob, x.device)
        c = torch.where(cond_mask[:, None], self.cond_mlp(c), self.null_cond[None, :].expand(x.shape[0], -1))
        skips = []
        for layer in self.down_layers:
            x, s = layer(x, t, c)
            skips.append(s)
        x = self.middle_resnet1(torch.cat([x, a], dim=1), t, c)
        for tr in self.middle_transformer:
            x = tr(x)
        x = self.middle_resnet2(x, t, c)
        for layer in self.up_layers:
            x, _ = layer(torch.cat([x, skips.pop()], dim=1), t, c)
        x = self.final_resnet(torch.cat([x, r], dim=1), t, c)
        return self.final_conv(x)[:, :, :n]
