"""ORACLE (test infrastructure): the two model wrappers of the reference, restated.

  * DDIM diffusion wrapper      osu_fusion/models/diffusion.py:15-111
  * rectified-flow wrapper      osu_fusion/models/rectified_flow.py:15-111
Constants TOTAL_DIM=6 (osu_fusion/library/osu/data/encode.py:24-26), AUDIO_DIM=96, CONTEXT_DIM=5
(osu_fusion/scripts/dataset_creator.py:22-25).

The reference modules themselves cannot be imported here (they need diffusers / torchdiffeq / librosa / bezier), so
these wrappers are pinned only through the UNet they wrap (tests/test_oracle_vs_reference.py) plus the schedule
identities in tests/test_schedules.py: PARITY UNPINNED for the wrapper arithmetic itself.

Oracle-only hooks (keyword-only, default None) inject the random draws so that two implementations can be compared
on identical noise / timesteps / CFG mask: `noise=`, `timesteps=`, `cond_mask=`.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .denoiser import UNet
from .schedules import DDIMSchedule, odeint_midpoint

TOTAL_DIM, AUDIO_DIM, CONTEXT_DIM = 6, 96, 5


def _masked_mse(pred: torch.Tensor, target: torch.Tensor, orig_len: Optional[torch.Tensor]) -> torch.Tensor:
    """diffusion.py:101-111 / rectified_flow.py:101-111."""
    loss = F.mse_loss(pred, target, reduction="none")
    if orig_len is None:
        return loss.mean()
    b, d, n = loss.shape
    pos = torch.arange(n, device=loss.device)[None, :]
    mask = (pos < orig_len.to(loss.device)[:, None]).to(torch.float32)[:, None, :].expand(b, d, n)
    return (loss * mask).sum() / mask.sum()


class _Base(nn.Module):
    def __init__(self, dim_h, dim_h_mult, num_layer_blocks, num_middle_transformers, cross_embed_kernel_sizes,
                 attn_dim_head, attn_heads, attn_kv_heads, attn_context_len, cond_drop_prob) -> None:
        super().__init__()
        self.unet = UNet(TOTAL_DIM, AUDIO_DIM, CONTEXT_DIM, dim_h, dim_h_mult=dim_h_mult, num_layer_blocks=num_layer_blocks,
                         num_middle_transformers=num_middle_transformers, cross_embed_kernel_sizes=cross_embed_kernel_sizes,
                         attn_dim_head=attn_dim_head, attn_heads=attn_heads, attn_kv_heads=attn_kv_heads,
                         attn_context_len=attn_context_len)
        self.cond_drop_prob = cond_drop_prob

    def set_full_bf16(self) -> None:
        self.unet = self.unet.bfloat16()


class DiffusionOsuFusion(_Base):
    def __init__(self, dim_h: int, dim_h_mult=(1, 2, 3, 4), num_layer_blocks=(3, 3, 3, 3), num_middle_transformers: int = 3,
                 cross_embed_kernel_sizes=(3, 7, 15), attn_dim_head: int = 64, attn_heads: int = 16, attn_kv_heads: int = 1,
                 attn_context_len: int = 4096, cond_drop_prob: float = 0.5, train_timesteps: int = 1000,
                 sampling_timesteps: int = 35) -> None:
        super().__init__(dim_h, dim_h_mult, num_layer_blocks, num_middle_transformers, cross_embed_kernel_sizes,
                         attn_dim_head, attn_heads, attn_kv_heads, attn_context_len, cond_drop_prob)
        self.scheduler = DDIMSchedule(train_timesteps)
        self.train_timesteps = train_timesteps
        self.sampling_timesteps = sampling_timesteps

    @torch.inference_mode()
    def sample(self, a, c, x=None, cond_scale: float = 7.0):  # diffusion.py:59-77
        b, _, n = a.shape
        if x is None:
            x = torch.randn((b, TOTAL_DIM, n), device=a.device)
        self.scheduler.set_timesteps(self.sampling_timesteps)
        for t in self.scheduler.timesteps:
            tb = t.expand(b).long().to(a.device)
            pred = self.unet.forward_with_cond_scale(x, a, tb, c, cond_scale=cond_scale)
            x = self.scheduler.step(pred, t, x)
        return x

    def forward(self, x, a, c, orig_len=None, *, noise=None, timesteps=None, cond_mask=None):  # diffusion.py:79-111
        assert x.shape[-1] == a.shape[-1], "x and a must have the same number of sequence length"
        if noise is None:
            noise = torch.randn_like(x)
        if timesteps is None:
            timesteps = torch.randint(0, self.train_timesteps, (x.shape[0],), dtype=torch.int64, device=x.device)
        x_noisy = self.scheduler.add_noise(x, noise, timesteps)
        pred = self.unet(x_noisy, a, timesteps, c, cond_drop_prob=self.cond_drop_prob, cond_mask=cond_mask)
        return _masked_mse(pred, noise, orig_len)


def cosmap(t: torch.Tensor) -> torch.Tensor:  # rectified_flow.py:15-16
    return 1.0 - (1.0 / (torch.tan(torch.pi / 2 * t) + 1))


class RectifiedFlowOsuFusion(_Base):
    def __init__(self, dim_h: int, dim_h_mult=(1, 2, 3, 4), num_layer_blocks=(3, 3, 3, 3), num_middle_transformers: int = 3,
                 cross_embed_kernel_sizes=(3, 7, 15), attn_dim_head: int = 64, attn_heads: int = 16, attn_kv_heads: int = 1,
                 attn_context_len: int = 4096, cond_drop_prob: float = 0.5, sampling_timesteps: int = 16) -> None:
        super().__init__(dim_h, dim_h_mult, num_layer_blocks, num_middle_transformers, cross_embed_kernel_sizes,
                         attn_dim_head, attn_heads, attn_kv_heads, attn_context_len, cond_drop_prob)
        self.sample_timesteps = sampling_timesteps

    @torch.inference_mode()
    def sample(self, a, c, x=None, cond_scale: float = 2.0):  # rectified_flow.py:57-79
        b, _, n = a.shape
        if x is None:
            x = torch.randn((b, TOTAL_DIM, n), device=a.device)
        times = torch.linspace(0.0, 1.0, self.sample_timesteps, device=a.device)

        def ode_fn(t, y):
            return self.unet.forward_with_cond_scale(y, a, t.expand(b), c, cond_scale=cond_scale)

        return odeint_midpoint(ode_fn, x, times)[-1]

    def forward(self, x, a, c, orig_len=None, *, noise=None, timesteps=None, cond_mask=None):  # rectified_flow.py:81-111
        assert x.shape[-1] == a.shape[-1], "x and a must have the same number of sequence length"
        if noise is None:
            noise = torch.randn_like(x)
        times = torch.rand(x.shape[0], device=x.device) if timesteps is None else timesteps
        t = cosmap(times[:, None, None])
        x_noisy = t * x + (1 - t) * noise
        flow = x - noise
        pred = self.unet(x_noisy, a, times, c, cond_drop_prob=self.cond_drop_prob, cond_mask=cond_mask)
        return _masked_mse(pred, flow, orig_len)
