"""ORACLE (test infrastructure, not product code): CPU/PyTorch restatement of the reference's two transformer backbones
(SURVEY.md §8f row 3).  Only tests/ may import this module; the product (osufusion_b200/) never does.

What is restated (reference file:line, relative to /root/reference):
  * DiT, DiTBlock, DiTAttention, FinalLayer, MultiHeadRMSNorm, modulate      osu_fusion/modules/dit.py:13-292
  * MMDiT, MMDiTBlock, JointAttention, PatchEmbedding, FinalLayer            osu_fusion/modules/mmdit.py:13-389
  * Attend (q, k, v cast to bf16 around SDPA, result cast to v's dtype)      osu_fusion/modules/attention.py:61-101

Parity pin: tests/test_backbones_oracle.py imports the real reference modules (build container only) and checks this
restatement against them with identical weights; oracle/make_golden_backbones.py stores reference outputs under
tests/golden/backbones_ref.pt for boxes without /root/reference.

Attribute names mirror the reference so that state_dict keys are identical.  One extension (as in oracle/denoiser.py):
`forward(..., cond_mask=...)` lets a test inject the classifier-free-guidance keep mask instead of drawing it.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from .denoiser import Attend, CrossEmbedLayer, SinusoidalPositionEmbedding, cfg_keep_mask


def ada_modulate(h: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """dit.py:13-15 / mmdit.py:13-15: per-sample affine of the normed activations, h (B, L, C), shift/scale (B, C)."""
    return h * (1 + scale[:, None, :]) + shift[:, None, :]


def _mlp(d_in: int, d_hidden: int, d_out: int, bias: bool = True) -> nn.Sequential:
    return nn.Sequential(nn.Linear(d_in, d_hidden, bias=bias), nn.SiLU(), nn.Linear(d_hidden, d_out, bias=bias))


class FeedForward(nn.Sequential):  # dit.py:52-59, mmdit.py:34-41 (default multiplier 4)
    def __init__(self, dim: int, dim_mult: int = 4) -> None:
        super().__init__(nn.Linear(dim, dim * dim_mult), nn.SiLU(), nn.Linear(dim * dim_mult, dim))


class MultiHeadRMSNorm(nn.Module):  # dit.py:62-69, mmdit.py:55-62: unit-normalise each head vector, learned gain, * sqrt(dim)
    def __init__(self, dim: int, heads: int) -> None:
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(heads, 1, dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # x: (B, heads, L, dim)
        return F.normalize(x, dim=-1) * self.gamma * self.scale


def _heads(t: torch.Tensor, h: int) -> torch.Tensor:
    b, n, w = t.shape
    return t.view(b, n, h, w // h).transpose(1, 2)


def _merge(t: torch.Tensor) -> torch.Tensor:
    b, h, n, d = t.shape
    return t.transpose(1, 2).reshape(b, n, h * d)


class _AdaHead(nn.Module):
    """`norm` (no affine, eps 1e-6) + `modulation` = SiLU -> Linear(dim, 2 dim) + `linear` (dit.py:72-86, mmdit.py:219-233)."""

    def __init__(self, dim_h: int, dim_out: int) -> None:
        super().__init__()
        self.norm = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim_h, dim_h * 2, bias=True))
        self.linear = nn.Linear(dim_h, dim_out)

    def forward(self, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        shift, scale = self.modulation(c).chunk(2, dim=1)
        return self.linear(ada_modulate(self.norm(x), shift, scale))


# ------------------------------------------------------------------------------------------------------ DiT
class DiTAttention(nn.Module):  # dit.py:89-118 — fused qkv projection, per-head qk RMS norm, no output projection
    def __init__(self, dim: int, heads: int, dim_head: int, qk_norm: bool = True, context_len: int = 4096) -> None:
        super().__init__()
        self.heads = heads
        self.to_qkv = nn.Linear(dim, dim_head * heads * 3, bias=False)
        self.q_norm = MultiHeadRMSNorm(dim_head, heads=heads) if qk_norm else nn.Identity()
        self.k_norm = MultiHeadRMSNorm(dim_head, heads=heads) if qk_norm else nn.Identity()
        self.attn = Attend()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        q, k, v = (_heads(t, self.heads) for t in self.to_qkv(x).chunk(3, dim=-1))
        return _merge(self.attn(self.q_norm(q), self.k_norm(k), v))


class DiTBlock(nn.Module):  # dit.py:121-159 — adaLN-Zero block
    def __init__(self, dim_h: int, dim_h_mult: int = 4, attn_heads: int = 8, attn_dim_head: int = 64, attn_qk_norm: bool = True,
                 attn_context_len: int = 4096) -> None:
        super().__init__()
        self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim_h, dim_h * 6, bias=True))
        self.norm1 = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.attn = DiTAttention(dim_h, heads=attn_heads, dim_head=attn_dim_head, qk_norm=attn_qk_norm, context_len=attn_context_len)
        self.norm2 = nn.LayerNorm(dim_h, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(dim_h, dim_h_mult)
        self.gradient_checkpointing = False

    def forward(self, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        sh1, sc1, g1, sh2, sc2, g2 = self.modulation(c).chunk(6, dim=1)
        x = x + g1[:, None, :] * self.attn(ada_modulate(self.norm1(x), sh1, sc1))
        return x + g2[:, None, :] * self.ff(ada_modulate(self.norm2(x), sh2, sc2))


class DiT(nn.Module):  # dit.py:162-292
    def __init__(self, dim_in_x: int, dim_in_a: int, dim_in_c: int, dim_h: int, dim_h_mult: int = 4, depth: int = 12,
                 cross_embed_kernel_sizes: Sequence[int] = (3, 7, 15), attn_heads: int = 8, attn_dim_head: int = 64,
                 attn_qk_norm: bool = True, attn_context_len: int = 4096) -> None:
        super().__init__()
        self.dim_in_x = dim_in_x
        self.preprocess = CrossEmbedLayer(dim_in_x + dim_in_a, dim_h, cross_embed_kernel_sizes)
        self.postprocess = nn.Conv1d(dim_h, dim_in_x, 1, bias=False)
        self.mlp_time = nn.Sequential(SinusoidalPositionEmbedding(dim_h), nn.Linear(dim_h, dim_h, bias=False), nn.SiLU(),
                                      nn.Linear(dim_h, dim_h, bias=False))
        self.mlp_cond = _mlp(dim_in_c, dim_h, dim_h)
        self.null_cond = nn.Parameter(torch.randn(dim_h))
        self.feature_extractor_a = nn.Linear(dim_in_a * 2, dim_h)
        self.mlp_audio = _mlp(dim_h, dim_h, dim_h)
        self.blocks = nn.ModuleList([DiTBlock(dim_h, dim_h_mult, attn_heads, attn_dim_head, attn_qk_norm, attn_context_len)
                                     for _ in range(depth)])
        self.final = _AdaHead(dim_h, dim_h)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        """dit.py:223-250 — xavier-uniform + zero bias everywhere, N(0, 0.02) embedders, zero adaLN heads / postprocess."""
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

    def forward_with_cond_scale(self, *args, cond_scale: float = 1.0, **kwargs) -> torch.Tensor:  # dit.py:258-265
        cond = self(*args, **kwargs)
        if cond_scale == 1.0:
            return cond
        null = self(*args, **kwargs, cond_drop_prob=1.0)
        return null + (cond - null) * cond_scale

    def forward(self, x, a, t, c, cond_drop_prob: float = 0.0, cond_mask: Optional[torch.Tensor] = None) -> torch.Tensor:  # dit.py:267-292
        n = x.shape[-1]
        h = self.preprocess(torch.cat([x, a], dim=1)).transpose(1, 2)
        stats = self.feature_extractor_a(torch.cat([a.mean(dim=-1), a.std(dim=-1)], dim=1))
        keep = cond_mask if cond_mask is not None else cfg_keep_mask(h.shape[0], 1.0 - cond_drop_prob, h.device)
        cvec = torch.where(keep[:, None], self.mlp_cond(c), self.null_cond[None, :].expand(h.shape[0], -1))
        cvec = cvec + self.mlp_time(t) + self.mlp_audio(stats)
        for blk in self.blocks:
            h = blk(h, cvec)
        h = self.final(h, cvec).transpose(1, 2)
        return self.postprocess(h[:, :, :n])


# ------------------------------------------------------------------------------------------------------ MMDiT
class PatchEmbedding(nn.Module):  # mmdit.py:44-52 — non-overlapping patches of `patch_size` frames
    def __init__(self, dim_in: int, dim_emb: int, patch_size: int) -> None:
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv1d(dim_in, dim_emb, patch_size, stride=patch_size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] % self.patch_size == 0, "Input sequence length must be divisible by the patch size"
        return self.proj(x).transpose(1, 2)


class JointAttention(nn.Module):  # mmdit.py:65-130 — one softmax over the concatenated [audio ; beatmap] token sequence
    def __init__(self, dim: int, dim_head: int, heads: int, kv_heads: int, qk_norm: bool = True, context_len: int = 4096) -> None:
        super().__init__()
        self.heads, self.kv_heads, self.qk_norm = heads, kv_heads, qk_norm
        for s in ("x", "a"):
            setattr(self, f"to_q_{s}", nn.Linear(dim, dim_head * heads, bias=False))
            setattr(self, f"to_k_{s}", nn.Linear(dim, dim_head * kv_heads, bias=False))
            setattr(self, f"to_v_{s}", nn.Linear(dim, dim_head * kv_heads, bias=False))
            setattr(self, f"q_{s}_norm", MultiHeadRMSNorm(dim_head, heads) if qk_norm else nn.Identity())
            setattr(self, f"k_{s}_norm", MultiHeadRMSNorm(dim_head, kv_heads) if qk_norm else nn.Identity())
        self.attn = Attend()

    def _stream(self, s: str, h: torch.Tensor):
        q = getattr(self, f"q_{s}_norm")(_heads(getattr(self, f"to_q_{s}")(h), self.heads))
        k = getattr(self, f"k_{s}_norm")(_heads(getattr(self, f"to_k_{s}")(h), self.kv_heads))
        v = _heads(getattr(self, f"to_v_{s}")(h), self.kv_heads)
        rep = self.heads // self.kv_heads
        # grouped-query expansion "b h n d -> b (r h) n d": query head j reads kv head j % kv_heads  (mmdit.py:116-117)
        return q, k.repeat(1, rep, 1, 1), v.repeat(1, rep, 1, 1)

    def forward(self, x: torch.Tensor, a: torch.Tensor):
        qx, kx, vx = self._stream("x", x)
        qa, ka, va = self._stream("a", a)
        la = a.shape[1]
        out = self.attn(torch.cat([qa, qx], dim=2), torch.cat([ka, kx], dim=2), torch.cat([va, vx], dim=2))
        return _merge(out[:, :, la:]), _merge(out[:, :, :la])


class MMDiTBlock(nn.Module):  # mmdit.py:133-216
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
        self.attn = JointAttention(dim_h, attn_dim_head, attn_heads, attn_kv_heads, qk_norm=attn_qk_norm, context_len=attn_context_len)
        self.gradient_checkpointing = False

    def forward(self, x: torch.Tensor, a: torch.Tensor, c: torch.Tensor):
        shx1, scx1, gx1, shx2, scx2, gx2 = self.modulation_x(c).chunk(6, dim=1)
        sha1, sca1, ga1, sha2, sca2, ga2 = self.modulation_a(c).chunk(6, dim=1)
        ox, oa = self.attn(ada_modulate(self.norm1_x(x), shx1, scx1), ada_modulate(self.norm1_a(a), sha1, sca1))
        x = x + gx1[:, None, :] * self.attn_out_x(ox)
        a = a + ga1[:, None, :] * self.attn_out_a(oa)
        x = x + gx2[:, None, :] * self.mlp_x(ada_modulate(self.norm2_x(x), shx2, scx2))
        a = a + ga2[:, None, :] * self.mlp_a(ada_modulate(self.norm2_a(a), sha2, sca2))
        return x, a


class MMDiT(nn.Module):  # mmdit.py:236-389
    def __init__(self, dim_in_x: int, dim_in_a: int, dim_in_c: int, dim_h: int, dim_h_mult: int = 4, patch_size: int = 4,
                 depth: int = 12, attn_dim_head: int = 64, attn_heads: int = 8, attn_kv_heads: int = 2, attn_qk_norm: bool = True,
                 attn_context_len: int = 4096) -> None:
        super().__init__()
        self.dim_h, self.dim_in_x, self.patch_size = dim_h, dim_in_x, patch_size
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
        self.final_layer = _AdaHead(dim_h, patch_size * dim_h)
        self.out = nn.Conv1d(dim_h, dim_in_x, 1)
        self.reset_parameters()

    def reset_parameters(self) -> None:
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

    forward_with_cond_scale = DiT.forward_with_cond_scale  # mmdit.py:335-342

    def forward(self, x, a, t, c, cond_drop_prob: float = 0.0, cond_mask: Optional[torch.Tensor] = None) -> torch.Tensor:  # mmdit.py:344-389
        stats = self.feature_extractor_a(torch.cat([a.mean(dim=-1), a.std(dim=-1)], dim=1))
        n = x.shape[-1]
        pad = (-n) % self.patch_size
        hx = self.emb_x(F.pad(x, (0, pad), value=-1.0))
        ha = self.emb_a(F.pad(a, (0, pad), value=-23.0))
        keep = cond_mask if cond_mask is not None else cfg_keep_mask(hx.shape[0], 1.0 - cond_drop_prob, hx.device)
        cvec = torch.where(keep[:, None], self.mlp_cond(c), self.null_cond[None, :].expand(hx.shape[0], -1))
        cvec = cvec + self.mlp_time(t) + self.mlp_a(stats)
        for blk in self.blocks:
            hx, ha = blk(hx, ha, cvec)
        y = self.final_layer(hx, cvec)                       # (B, n/p, p * dim_h)
        b, m, _ = y.shape
        y = y.reshape(b, m * self.patch_size, self.dim_h)    # "b n (p d) -> b d (n p)": patch index is the slow half
        return self.out(y.transpose(1, 2))[:, :, :n]
