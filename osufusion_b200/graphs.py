"""CUDA-graph capture of the launch-bound inner loops (a whole training micro-step; a whole sampling step).

The denoiser issues ~2.7k small kernel launches per fwd+bwd; replaying them as one CUDA graph removes the host from
the critical path (north_star "CUDA streams and graphs instead of a tracing compiler").
"""
from __future__ import annotations

from typing import Optional

import torch


class GraphedTrainStep:
    """Captures `zero_grad(set_to_none) -> loss = model(x, a, c[, orig_len]) -> loss.backward()` into one CUDA graph.

    After each `__call__`, `p.grad` of every trainable parameter holds the fresh gradient (static graph-pool tensors)
    and the returned 0-dim loss tensor is the static loss buffer.  Random draws (noise, timesteps, CFG mask) come from
    torch's graph-safe CUDA generator, exactly like eager mode.
    """

    def __init__(self, model: torch.nn.Module, x: torch.Tensor, a: torch.Tensor, c: torch.Tensor,
                 orig_len: Optional[torch.Tensor] = None, warmup: int = 2, post_backward=None) -> None:
        self.model = model
        self.x, self.a, self.c = x.clone(), a.clone(), c.clone()
        self.orig_len = None if orig_len is None else orig_len.to(x.device).clone()
        self.post_backward = post_backward

        def step():
            model.zero_grad(set_to_none=True)
            args = (self.x, self.a, self.c) if self.orig_len is None else (self.x, self.a, self.c, self.orig_len)
            loss = model(*args)
            loss.backward()
            if self.post_backward is not None:
                self.post_backward()
            return loss

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        model.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = step()

    def __call__(self, x=None, a=None, c=None, orig_len=None) -> torch.Tensor:
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if a is not None:
            self.a.copy_(a, non_blocking=True)
        if c is not None:
            self.c.copy_(c, non_blocking=True)
        if orig_len is not None and self.orig_len is not None:
            self.orig_len.copy_(orig_len, non_blocking=True)
        self.graph.replay()
        return self.loss


class GraphedCallable:
    """Captures an arbitrary launch-only closure (e.g. `zero_grad -> loss = f(backbone(x, a, t, c)) -> loss.backward()` of the DiT /
    MMDiT backbones) into one CUDA graph.  The closure must read its inputs from tensors that stay alive and are updated in
    place between replays, and must not synchronise with the host.  `__call__` replays and returns what the closure returned."""

    def __init__(self, fn, warmup: int = 2) -> None:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def __call__(self):
        self.graph.replay()
        return self.result
