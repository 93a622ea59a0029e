"""Fused optimizer step for the B200 denoiser.

Replaces the tail of the reference's training loop (trainer.py:230-236, 302-309):

    accelerator.clip_grad_norm_(model.parameters(), 1.0)      # 1239 per-tensor norms + one host sync
    optimizer.step()                                          # torch.optim.AdamW(lr=1e-5), per-tensor (foreach) kernels
    scheduler.step()                                          # diffusers get_cosine_schedule_with_warmup

with two launches over the engine's flat gradient arena: `of_grad_sumsq` (global gradient norm, left on the device) and
`of_adamw_step` (every parameter tensor in one launch; the clip factor is read from device memory, so nothing synchronises).
Moments live in two arenas with the gradient arena's layout.  `FusedAdamW` is a `torch.optim.Optimizer`, so LR schedulers
(`torch.optim.lr_scheduler.LambdaLR`, what diffusers' cosine schedule is) drive `param_groups[0]["lr"]` as usual.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _native as N


def cosine_schedule_with_warmup(optimizer, num_warmup_steps: int, num_training_steps: int, num_cycles: float = 0.5):
    """diffusers.optimization.get_cosine_schedule_with_warmup (call site trainer.py:232-236), restated."""
    def lr_lambda(step: int) -> float:
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        progress = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda)


class FusedAdamW(torch.optim.Optimizer):
    """AdamW over `model.unet`'s trainable parameters with the gradient-norm clip fused in (torch.optim.AdamW defaults)."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = 1.0) -> None:
        unet = model.unet if hasattr(model, "unet") else model
        self.unet = unet
        store = unet._store
        store.ensure_arena(unet)
        self.store = store
        params = list(store.arena_params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self._step = 0
        self._plan_key = None
        self._build()

    def _build(self) -> None:
        st = self.store
        dev = st.arena.device
        lib = N.lib()
        rows, cta = [], 0
        for p, (s0, _) in zip(st.arena_params, st.arena_offsets):
            assert p.is_contiguous() and p.dtype == torch.float32
            rows.append(N.OptTensor(p.data_ptr(), s0, p.numel(), cta, 0))
            cta += lib.of_opt_tensor_ctas(p.numel())
        arr = (N.OptTensor * len(rows))(*rows)
        self._table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self._num, self._ctas = len(rows), cta
        if getattr(self, "exp_avg", None) is None or self.exp_avg.numel() != st.arena.numel():
            self.exp_avg = torch.zeros_like(st.arena)
            self.exp_avg_sq = torch.zeros_like(st.arena)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._plan_key = (st.arena.data_ptr(), tuple(p.data_ptr() for p in st.arena_params))

    @property
    def grad_norm(self) -> torch.Tensor:
        """Global gradient L2 norm of the last step (0-dim tensor on the device; reading it is the only sync)."""
        return self._sumsq.sqrt().float()[0]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        st = self.store
        if self._plan_key != (st.arena.data_ptr(), tuple(p.data_ptr() for p in st.arena_params)):
            self._build()
        # gradients must be the engine's arena views (they are after a backward pass of the engine); anything else is copied in
        for p in st.arena_params:
            v = st.arena_views[id(p)]
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
        g = self.param_groups[0]
        self._step += 1
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        if clip:
            N.call("of_grad_sumsq", st.arena.data_ptr(), st.arena.numel(), self._sumsq.data_ptr())
        N.call("of_adamw_step", self._table.data_ptr(), self._num, self._ctas, st.arena.data_ptr(), self.exp_avg.data_ptr(),
               self.exp_avg_sq.data_ptr(), self._sumsq.data_ptr() if clip else None, float(self.max_grad_norm or 0.0), float(g["lr"]),
               float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self._step)
        st.param_epoch += 1     # parameters changed through raw pointers (torch's _version did not move): invalidate operand caches
        return loss

    def state_dict(self):
        d = super().state_dict()
        d["fused"] = {"step": self._step, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}
        return d

    def load_state_dict(self, sd):
        fused = sd.get("fused")
        super().load_state_dict({k: v for k, v in sd.items() if k != "fused"})
        if fused is not None:
            self._step = int(fused["step"])
            self.exp_avg.copy_(fused["exp_avg"])
            self.exp_avg_sq.copy_(fused["exp_avg_sq"])
