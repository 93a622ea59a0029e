"""Drop-in for osu_fusion/models/diffusion.py: `OsuFusion(dim_h, ...)`, `model(x, a, c, orig_len) -> loss`,
`model.sample(a, c, x, cond_scale)`; DDIM schedule restated from diffusers 0.29.2 (call sites diffusion.py:48-51,71-75,96)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._base import BaseOsuFusion


class DDIMScheduler:
    """The subset of diffusers.DDIMScheduler the reference uses (linear betas 1e-4..0.02, leading spacing, eta=0,
    epsilon prediction, clip_sample=True, set_alpha_to_one=True)."""

    class _Cfg:
        pass

    def __init__(self, num_train_timesteps: int = 1000, beta_schedule: str = "linear") -> None:
        assert beta_schedule == "linear"
        self.config = DDIMScheduler._Cfg()
        self.config.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(1e-4, 0.02, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)
        self._dev_cache = {}

    def alphas_cumprod_on(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._dev_cache:
            self._dev_cache[key] = self.alphas_cumprod.to(device=device, dtype=torch.float32)
        return self._dev_cache[key]

    def set_timesteps(self, n: int) -> None:
        self.num_inference_steps = n
        ratio = self.config.num_train_timesteps // n
        self.timesteps = (torch.arange(0, n) * ratio).round().flip(0).to(torch.int64)

    def step_coeffs(self, t: int):
        t_prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[t_prev]) if t_prev >= 0 else 1.0
        return (1 - a_t) ** 0.5, a_t ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5


class OsuFusion(BaseOsuFusion):
    def __init__(self, dim_h: int, dim_h_mult: Tuple[int] = (1, 2, 3, 4), num_layer_blocks: Tuple[int] = (3, 3, 3, 3),
                 num_middle_transformers: int = 3, cross_embed_kernel_sizes: Tuple[int] = (3, 7, 15), attn_dim_head: int = 64,
                 attn_heads: int = 16, attn_kv_heads: int = 1, attn_context_len: int = 4096, cond_drop_prob: float = 0.5,
                 train_timesteps: int = 1000, sampling_timesteps: int = 35) -> None:
        super().__init__(dim_h, dim_h_mult, num_layer_blocks, num_middle_transformers, cross_embed_kernel_sizes, attn_dim_head,
                         attn_heads, attn_kv_heads, attn_context_len, cond_drop_prob)
        self.scheduler = DDIMScheduler(num_train_timesteps=train_timesteps, beta_schedule="linear")
        self.train_timesteps = train_timesteps
        self.sampling_timesteps = sampling_timesteps

    @torch.inference_mode()
    def sample(self, a: torch.Tensor, c: torch.Tensor, x: Optional[torch.Tensor] = None, cond_scale: float = 7.0) -> torch.Tensor:
        """diffusion.py:59-77.  Audio encoder evaluated once (loop-invariant), cond+null batched, CFG+DDIM update fused."""
        s = self._sampler_setup(a, c, x, cond_scale)
        self.scheduler.set_timesteps(self.sampling_timesteps)
        steps = self.scheduler.timesteps.tolist()
        g = self._sampler_graphs(s, cond_scale, "ddim", [(0, "x16", "x", "x16")])
        if g is not None:
            st, (graph,) = g
            tt = torch.tensor(steps, dtype=torch.float32).to(a.device)
            coefs = torch.tensor([self.scheduler.step_coeffs(t) for t in steps], dtype=torch.float32).to(a.device)
            for i in range(len(steps)):
                st.t_buf.copy_(tt[i].expand(st.b))          # timestep and DDIM coefficients are device data of the captured step
                st.coef.copy_(coefs[i])
                graph.replay()
            return st.x.clone()
        xcur, x16 = s.x, s.x16
        for t in steps:
            tb = torch.full((s.b,), t, dtype=torch.int64, device=a.device)
            cond16, null16 = self._eval_denoiser(s, x16, tb)
            c_eps, c_div, c_x0, c_dir = self.scheduler.step_coeffs(t)
            xcur, x16 = self._update(s, xcur, cond16, null16, cond_scale, 0, c_eps, c_div, c_x0, c_dir)
        return xcur

    def forward(self, x: torch.Tensor, a: torch.Tensor, c: torch.Tensor, orig_len: Optional[torch.Tensor] = None, *,
                noise: Optional[torch.Tensor] = None, timesteps: Optional[torch.Tensor] = None,
                cond_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """diffusion.py:79-111 (same RNG draw order: randn_like -> randint -> CFG uniform_)."""
        assert x.shape[-1] == a.shape[-1], "x and a must have the same number of sequence length"
        if noise is None:
            noise = torch.randn_like(x)
        if timesteps is None:
            timesteps = torch.randint(0, self.scheduler.config.num_train_timesteps, (x.shape[0],), dtype=torch.int64, device=x.device)
        ac = self.scheduler.alphas_cumprod_on(x.device)[timesteps]
        ca, cb = (ac ** 0.5).contiguous(), ((1 - ac) ** 0.5).contiguous()
        return self._train_step(x, a, timesteps, c, noise, ca, cb, 0.0, 1.0, orig_len, cond_mask)
