"""Drop-in for osu_fusion/models/rectified_flow.py; fixed-grid midpoint integration restated from torchdiffeq 0.2.4
(call site rectified_flow.py:78)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._base import BaseOsuFusion


def cosmap(t: torch.Tensor) -> torch.Tensor:  # rectified_flow.py:15-16
    return 1.0 - (1.0 / (torch.tan(torch.pi / 2 * t) + 1))


class OsuFusion(BaseOsuFusion):
    def __init__(self, dim_h: int, dim_h_mult: Tuple[int] = (1, 2, 3, 4), num_layer_blocks: Tuple[int] = (3, 3, 3, 3),
                 num_middle_transformers: int = 3, cross_embed_kernel_sizes: Tuple[int] = (3, 7, 15), attn_dim_head: int = 64,
                 attn_heads: int = 16, attn_kv_heads: int = 1, attn_context_len: int = 4096, cond_drop_prob: float = 0.5,
                 sampling_timesteps: int = 16) -> None:
        super().__init__(dim_h, dim_h_mult, num_layer_blocks, num_middle_transformers, cross_embed_kernel_sizes, attn_dim_head,
                         attn_heads, attn_kv_heads, attn_context_len, cond_drop_prob)
        self.sample_timesteps = sampling_timesteps

    @torch.inference_mode()
    def sample(self, a: torch.Tensor, c: torch.Tensor, x: Optional[torch.Tensor] = None, cond_scale: float = 2.0) -> torch.Tensor:
        """rectified_flow.py:57-79: midpoint over linspace(0, 1, sample_timesteps): two CFG evaluations per interval."""
        s = self._sampler_setup(a, c, x, cond_scale)
        times = torch.linspace(0.0, 1.0, self.sample_timesteps)
        g = self._sampler_graphs(s, cond_scale, "midpoint", [(1, "x16", "xtmp", "xmid16"), (1, "xmid16", "x", "x16")])
        if g is not None:
            st, (g_half, g_full) = g
            t0s, t1s = times[:-1], times[1:]
            dts = t1s - t0s
            zero = torch.zeros_like(dts)
            one = torch.ones_like(dts)
            # rows: [t, c_eps, c_div, c_x0, c_dir] for the half step (y_mid = y + f(t0, y) dt/2) and the full step (y += dt f(t0 + dt/2, y_mid))
            half = torch.stack([t0s, 0.5 * dts, one, zero, zero], 1).to(a.device)
            full = torch.stack([t0s + 0.5 * dts, dts, one, zero, zero], 1).to(a.device)
            for i in range(dts.numel()):
                for row, graph in ((half[i], g_half), (full[i], g_full)):
                    st.t_buf.copy_(row[0].expand(st.b))
                    st.coef.copy_(row[1:])
                    graph.replay()
            return st.x.clone()
        y, y16 = s.x, s.x16
        for t0, t1 in zip(times[:-1].tolist(), times[1:].tolist()):
            dt = t1 - t0
            tb = torch.full((s.b,), t0, dtype=torch.float32, device=a.device)
            cond16, null16 = self._eval_denoiser(s, y16, tb)
            _, ymid16 = self._update(s, y, cond16, null16, cond_scale, 1, 0.5 * dt)
            tb = torch.full((s.b,), t0 + 0.5 * dt, dtype=torch.float32, device=a.device)
            cond16, null16 = self._eval_denoiser(s, ymid16, tb)
            y, y16 = self._update(s, y, cond16, null16, cond_scale, 1, dt)
        return y

    def forward(self, x: torch.Tensor, a: torch.Tensor, c: torch.Tensor, orig_len: Optional[torch.Tensor] = None, *,
                noise: Optional[torch.Tensor] = None, timesteps: Optional[torch.Tensor] = None,
                cond_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """rectified_flow.py:81-111: x_t = t*x + (1-t)*noise with t = cosmap(times); the UNet is conditioned on `times`."""
        assert x.shape[-1] == a.shape[-1], "x and a must have the same number of sequence length"
        if noise is None:
            noise = torch.randn_like(x)
        times = torch.rand(x.shape[0], device=x.device) if timesteps is None else timesteps
        tm = cosmap(times.float())
        return self._train_step(x, a, times, c, noise, tm.contiguous(), (1 - tm).contiguous(), 1.0, -1.0, orig_len, cond_mask)
