"""ORACLE (test infrastructure): restatement of the reference's LoRA/DoRA forward for Conv1d (osu_fusion/modules/lora_layers.py:
59-92,292-328) and of peft 0.12.0's DoRA nn.Linear path.  Plain torch, autograd-differentiable, used only by tests.

Pinning: `dora_conv1d` is PINNED against the reference's own `LoraConv1d` / `DoraConv1dLayer` code, executed through the peft
bookkeeping stub of tests/peft_stub.py (tests/test_lora_reference_pin.py: live reference in the build container, golden vectors
tests/golden/lora_conv1d_ref.pt everywhere: outputs <= 1e-6, gradients <= 1e-5).  `dora_linear` restates peft's own
`DoraLinearLayer.forward`, whose source is absent from the image: PARITY UNPINNED (formula from SURVEY.md Appendix B; it is the
same expression as the pinned Conv1d variant with `linear` in place of `conv1d`)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dora_conv1d(x, W, bias, A, B, mag, scaling, padding):
    """result = base(x) + (s - 1) * conv(x, W) + s * B(A(x)) * scaling,  s = mag / ||W + scaling*BA|| (detached)."""
    base = F.conv1d(x, W, bias, padding=padding)
    lora_weight = (B.flatten(1) @ A.flatten(1)).reshape(W.shape)
    norm = (W + scaling * lora_weight.detach()).norm(p=2, dim=(1, 2), keepdim=True).transpose(1, 0).detach()
    s = mag / norm
    return base + (s - 1) * F.conv1d(x, W, None, padding=padding) + s * F.conv1d(F.conv1d(x, A, None, padding=padding), B) * scaling


def dora_linear(x, W, bias, A, B, mag, scaling):
    base = F.linear(x, W, bias)
    norm = (W + scaling * (B @ A).detach()).norm(p=2, dim=1).detach()
    s = mag / norm
    return base + (s - 1) * F.linear(x, W) + s * F.linear(F.linear(x, A), B) * scaling
