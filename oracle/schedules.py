"""ORACLE (test infrastructure): restatement of the third-party arithmetic the reference calls on its hot path.

The sources are NOT vendored in /root/reference and not installed in this image (no network):
  * diffusers==0.29.2  DDIMScheduler.__init__/add_noise/set_timesteps/step — call sites
    osu_fusion/models/diffusion.py:48-51,71,72,75,96 (requirements.txt:3)
  * torchdiffeq==0.2.4 odeint(..., method="midpoint") — call site osu_fusion/models/rectified_flow.py:78
    (requirements.txt:13)
Their published algorithms are restated below from those pinned versions.  PARITY UNPINNED against the real
packages (no golden vectors exist anywhere in the reference); the self-consistency identities of SURVEY.md §8c
are tested in tests/test_schedules.py.
"""
from __future__ import annotations

from typing import Callable, List

import torch


class DDIMSchedule:
    """DDIMScheduler(num_train_timesteps=T, beta_schedule="linear") with diffusers 0.29.2 defaults:
    beta_start=1e-4, beta_end=0.02, clip_sample=True (range 1.0), set_alpha_to_one=True, steps_offset=0,
    prediction_type="epsilon", timestep_spacing="leading"; step() is called with eta=0."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02) -> None:
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def add_noise(self, x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        ac = self.alphas_cumprod.to(device=x0.device, dtype=x0.dtype)
        sa = ac[t] ** 0.5
        sb = (1 - ac[t]) ** 0.5
        while sa.dim() < x0.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * x0 + sb * noise

    def set_timesteps(self, n: int) -> None:
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = (torch.arange(0, n) * ratio).round().flip(0).to(torch.int64)

    def step(self, eps: torch.Tensor, t: int, x: torch.Tensor) -> torch.Tensor:
        t = int(t)
        t_prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t_prev] if t_prev >= 0 else self.final_alpha_cumprod
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        x0 = x0.clamp(-1.0, 1.0)
        # eta = 0, use_clipped_model_output=False: eps is NOT recomputed from the clamped x0
        return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps


def odeint_midpoint(f: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], y0: torch.Tensor, times: torch.Tensor) -> List[torch.Tensor]:
    """torchdiffeq fixed-grid midpoint: the grid is `times` itself; two evaluations of f per interval."""
    ys = [y0]
    y = y0
    for t0, t1 in zip(times[:-1], times[1:]):
        dt = t1 - t0
        half = 0.5 * dt
        y_mid = y + f(t0, y) * half
        y = y + dt * f(t0 + half, y_mid)
        ys.append(y)
    return ys
