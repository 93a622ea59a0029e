"""Self-consistency identities for the restated third-party arithmetic: parity is UNPINNED against
diffusers / torchdiffeq themselves (absent from the image)."""
import torch

from oracle.schedules import DDIMSchedule, odeint_midpoint


def test_ddim_timesteps_leading():
    s = DDIMSchedule(1000)
    s.set_timesteps(35)
    assert s.timesteps[0] == 952 and s.timesteps[1] == 924 and s.timesteps[-1] == 0 and len(s.timesteps) == 35


def test_ddim_last_step_returns_clamped_x0():
    s = DDIMSchedule(1000)
    s.set_timesteps(35)
    x = torch.randn(2, 6, 16) * 3
    eps = torch.randn(2, 6, 16)
    a0 = s.alphas_cumprod[0]
    out = s.step(eps, 0, x)
    x0 = ((x - (1 - a0) ** 0.5 * eps) / a0 ** 0.5).clamp(-1, 1)
    assert torch.allclose(out, x0)


def test_add_noise_small_t_is_near_identity():
    s = DDIMSchedule(1000)
    x = torch.randn(3, 6, 8)
    out = s.add_noise(x, torch.zeros_like(x), torch.zeros(3, dtype=torch.long))
    assert torch.allclose(out, x * (1 - 1e-4) ** 0.5)


def test_add_noise_then_step_roundtrip():
    s = DDIMSchedule(1000)
    s.set_timesteps(1000)
    x0 = torch.rand(2, 6, 8) * 2 - 1
    eps = torch.randn_like(x0)
    t = 500
    xt = s.add_noise(x0, eps, torch.full((2,), t))
    prev = s.step(eps, t, xt)
    expect = s.add_noise(x0, eps, torch.full((2,), t - 1))
    assert torch.allclose(prev, expect, atol=1e-5)


def test_midpoint_exact_for_constant_and_linear_field():
    times = torch.linspace(0, 1, 16)
    y0 = torch.randn(4)
    v = torch.randn(4)
    out = odeint_midpoint(lambda t, y: v, y0, times)[-1]
    assert torch.allclose(out, y0 + v, atol=1e-6)
    calls = []
    out = odeint_midpoint(lambda t, y: (calls.append(float(t)), 2 * t * torch.ones_like(y))[1], y0, times)[-1]
    assert torch.allclose(out, y0 + 1.0, atol=1e-6)  # midpoint integrates linear-in-t fields exactly
    assert len(calls) == 30  # 15 intervals x 2 evaluations (rectified_flow.py:70 tqdm total)
