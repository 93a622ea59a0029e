"""Input side of the hot path: the reference's batch collation and a pinned, one-batch-ahead host-to-device
prefetcher, so that batch assembly and the H2D copy never sit between two training micro-steps.

`collate_fn` keeps the contract of trainer.py:74-95: variable-length `(x (6, n_i), a (96, n_i), c (5,))` items are right-padded to
the longest item with -1.0 (beatmap) / -23.0 (log-mel silence) and stacked, and the original lengths are returned for the masked
loss (diffusion.py:101-111).  The padding values are the ones the denoiser itself uses for its own alignment padding (unet.py:475-480).
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

X_PAD_VALUE, A_PAD_VALUE = -1.0, -23.0

Batch = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]


def collate_fn(batch: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], pin: bool = False) -> Batch:
    """trainer.py:74-95.  Writes straight into the stacked (optionally pinned) output buffers: one pass, no per-item F.pad copies."""
    n_max = max(x.shape[1] for x, _, _ in batch)
    B = len(batch)
    x0, a0, c0 = batch[0]
    out_x = torch.full((B, x0.shape[0], n_max), X_PAD_VALUE, dtype=x0.dtype, pin_memory=pin)
    out_a = torch.full((B, a0.shape[0], n_max), A_PAD_VALUE, dtype=a0.dtype, pin_memory=pin)
    out_c = torch.empty((B, *c0.shape), dtype=c0.dtype, pin_memory=pin)
    orig_len = torch.empty(B, dtype=torch.int64, pin_memory=pin)
    for i, (x, a, c) in enumerate(batch):
        n = x.shape[1]
        assert a.shape[1] == n, "x and a must have the same number of sequence length"
        out_x[i, :, :n] = x
        out_a[i, :, :n] = a
        out_c[i] = c
        orig_len[i] = n
    return out_x, out_a, out_c, orig_len


class DevicePrefetcher:
    """Iterates `(x, a, c[, orig_len])` host batches and yields them on `device`, copying batch i+1 on a side stream while batch i
    is being consumed (the reference relies on DataLoader workers + a synchronous `.to(device)` inside `accelerator.prepare`)."""

    def __init__(self, batches: Iterable, device: torch.device) -> None:
        self.batches, self.device = batches, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def _stage(self, batch):
        if self.stream is None:
            return tuple(t.to(self.device) for t in batch)
        with torch.cuda.stream(self.stream):
            staged = tuple((t if t.is_pinned() else t.pin_memory()).to(self.device, non_blocking=True) for t in batch)
        return staged

    def __iter__(self) -> Iterator:
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        for batch in it:
            cur = nxt
            if self.stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self.stream)
                for t in cur:
                    t.record_stream(torch.cuda.current_stream(self.device))
            nxt = self._stage(batch)
            yield cur
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            for t in nxt:
                t.record_stream(torch.cuda.current_stream(self.device))
        yield nxt
