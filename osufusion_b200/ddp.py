"""Data-parallel training: bucketed NCCL gradient all-reduce overlapped with the engine's backward pass.

Replaces what `accelerator.prepare(model)` gives the reference (torch DistributedDataParallel over NCCL, 25 MB buckets,
trainer.py:211-220,264-269,301).  Differences that matter on an NVSwitch box:
  * gradients live in ONE flat fp32 arena laid out in backward-completion order, so a bucket is a contiguous slice: the
    all-reduce runs in place on the arena (no flatten / unflatten copies) and `p.grad` are views of it;
  * buckets are large (default 256 MB: NVLink-5/NVSwitch collectives are latency- not link-bound) and are launched on a
    side stream from inside the backward tape as soon as the last layer writing into them has run;
  * the whole step (backward + collectives) is CUDA-graph capturable;
  * `reserve_sms` SMs are left to the NCCL kernels: the engine's persistent GEMM sizes its grid for the remaining SMs
    (of_set_sm_limit), otherwise every GEMM launched while a bucket is in flight runs as two waves (measured at 2 GPUs:
    67.9 -> 65.7 ms per step with 16 reserved SMs and NCCL_MAX_CTAS=16).
The path has exactly one exchange step: sum of gradients; everything else is batch-sharded.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


from .engine import backward_param_order  # noqa: E402,F401  (re-exported: the arena order is defined by the engine)


class GradAllReducer:
    """Owns the gradient arena of `model.unet` and all-reduces (averages) it bucket by bucket during backward."""

    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 256 << 20, group=None, overlap: bool = True,
                 reserve_sms: int = 0) -> None:
        unet = model.unet if hasattr(model, "unet") else model
        self.unet, self.group, self.overlap = unet, group, overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        store = unet._store
        store.ensure_arena(unet)                 # the engine owns the gradient arena; buckets are contiguous slices of it
        self.store = store
        params = store.arena_params
        dev = params[0].device
        self.buckets = []      # (start, end, [param ids])
        b_start, b_ids = 0, []
        end = 0
        for p, (s0, e0) in zip(params, store.arena_offsets):
            b_ids.append(id(p))
            end = e0
            if (end - b_start) * 4 >= bucket_bytes:
                self.buckets.append((b_start, end, b_ids))
                b_start, b_ids = end, []
        if b_ids:
            self.buckets.append((b_start, end, b_ids))
        self.bucket_of = {pid: bi for bi, (_, _, ids) in enumerate(self.buckets) for pid in ids}
        self.ready_at = None           # bucket index -> tape op index after which it is complete (learned in step 1)
        self.comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._pending: List[int] = []
        store.on_backward_begin = self._begin
        store.on_touch = self._touch
        unet.grad_sync = self._after_op
        unet.grad_finish = self.finish
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        # leave `reserve_sms` SMs to the NCCL kernels: the persistent GEMM grid shrinks accordingly (see of_set_sm_limit)
        self.reserve_sms = reserve_sms
        if reserve_sms > 0 and dev.type == "cuda" and self.world > 1:
            from . import _native as N
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            N.lib().of_set_sm_limit(max(2, (sms - reserve_sms) // 2 * 2))
        self._touch_log = {}
        self._op = None
        self._launched = set()

    # ---- hooks called by the engine
    @property
    def arena(self) -> torch.Tensor:
        return self.store.arena

    @property
    def views(self):
        return self.store.arena_views

    def _begin(self) -> None:
        self._launched = set()
        self._touch_log = {}

    def _touch(self, pid: int) -> None:
        self._touch_log[pid] = self._op

    def _after_op(self, i: int) -> None:
        """Called by Tape.run_backward after tape op i (ops run from last to first)."""
        self._op = i - 1
        if self.ready_at is None or not self.overlap or self.world == 1:
            return
        for bi, ready in enumerate(self.ready_at):
            if ready is not None and ready >= i and bi not in self._launched:
                self._launch(bi)

    def _reduce(self, t: torch.Tensor) -> None:
        if self._avg:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:                          # gloo (CPU tests) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def _launch(self, bi: int) -> None:
        self._launched.add(bi)
        s, e, _ = self.buckets[bi]
        if self.comm is None:
            self._reduce(self.arena[s:e])
            return
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        with torch.cuda.stream(self.comm):
            self._reduce(self.arena[s:e])

    def finish(self) -> None:
        """End of backward (still inside the autograd node): launch what is left, then join the comm stream."""
        if self.world > 1:
            for bi in range(len(self.buckets)):
                if bi not in self._launched:
                    self._launch(bi)
            if self.comm is not None:
                torch.cuda.current_stream().wait_stream(self.comm)
        if self.ready_at is None and self._touch_log:
            ready = []
            for _, _, ids in self.buckets:
                idx = [self._touch_log[p] for p in ids if p in self._touch_log and self._touch_log[p] is not None]
                ready.append(min(idx) if idx else 0)
            self.ready_at = ready

    # kept for API symmetry with bench.py's post_backward hook: all work already happened inside backward
    def all_reduce(self) -> None:
        return None
