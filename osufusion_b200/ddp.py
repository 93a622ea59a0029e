"""Data-parallel training: bucketed NCCL gradient all-reduce overlapped with the engine's backward pass.

Replaces what `accelerator.prepare(model)` gives the reference (torch DistributedDataParallel over NCCL, 25 MB buckets,
trainer.py:211-220,264-269,301).  Differences that matter on an NVSwitch box:
  * gradients live in ONE flat fp32 arena laid out in backward-completion order, so a bucket is a contiguous slice: the
    all-reduce runs in place on the arena (no flatten / unflatten copies) and `p.grad` are views of it;
  * buckets are large (default 256 MB: NVLink-5/NVSwitch collectives are latency- not link-bound) and are launched on a
    side stream from inside the backward tape as soon as the last layer writing into them has run;
  * the whole step (backward + collectives) is CUDA-graph capturable;
  * `reserve_sms` SMs are left to the NCCL kernels while buckets are in flight (from the first bucket launch of a backward pass
    to its end): the engine's persistent GEMM sizes its grid for the remaining SMs (of_set_sm_limit), otherwise every GEMM
    launched while a bucket is in flight runs as two waves (measured at 2 GPUs: 67.9 -> 65.7 ms per step with 16 reserved SMs
    and NCCL_MAX_CTAS=16); the forward pass keeps every SM;
  * the mean is taken by PRE-DIVISION for power-of-two worlds: the loss gradient is scaled by 1/world at the start of backward (exact:
    an exponent shift) and buckets are reduced with SUM, the only form NCCL's in-switch NVLS reduction supports — with AVG the tuner
    falls back to 16-channel rings (8 x B200: 61.3 -> 55.1 ms per step); the arena can live in NCCL-registered memory
    (`registered_arena`, ncclMemAlloc + register_mem_pool) so the collective runs zero-copy on the user buffer;
  * the engine's second stream (weight gradients, engine.SideLane) is a dependency of every bucket launch, like the main stream;
  * buckets taper: 256 MB while plenty of backward is left to hide them, 32 MB over the last 192 MB of the arena, and the arena
    order puts the audio encoder BEFORE the down path (engine.backward_param_plan), so the exposed tail is one small bucket.
The path has exactly one exchange step: sum of gradients; everything else is batch-sharded.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


from .engine import backward_param_order  # noqa: E402,F401  (re-exported: the arena order is defined by the engine)


class GradAllReducer:
    """Owns the gradient arena of `model.unet` and all-reduces (averages) it bucket by bucket during backward.

    Construction mirrors what DDP construction gives the reference (`accelerator.prepare`, trainer.py:264-269): parameters and
    buffers are broadcast from rank 0, so replicas start identical whatever each rank's seed / checkpoint state was."""

    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 256 << 20, group=None, overlap: bool = True,
                 reserve_sms: int = 0, tail_bucket_bytes: int = 32 << 20, tail_bytes: int = 192 << 20,
                 broadcast_params: bool = True, registered_arena: Optional[bool] = None, prescale: Optional[bool] = None) -> None:
        if registered_arena is None:
            registered_arena = os.environ.get("OF_DDP_REGISTERED_ARENA", "1") != "0"
        unet = model.unet if hasattr(model, "unet") else model
        self.unet, self.group, self.overlap = unet, group, overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_bytes, self.tail_bucket_bytes, self.tail_bytes = bucket_bytes, tail_bucket_bytes, tail_bytes
        store = unet._store
        self.store = store
        if broadcast_params and self.world > 1:
            self.broadcast_parameters(model)
        self.registered = False
        if self.world > 1 and registered_arena and dist.get_backend(group) == "nccl":
            self._use_registered_arena(store)
        store.ensure_arena(unet)                 # the engine owns the gradient arena; buckets are contiguous slices of it
        dev = store.arena_params[0].device
        self._build_buckets()
        self.comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._pending: List[int] = []
        store.on_backward_begin = self._begin
        store.on_touch = self._touch
        unet.grad_sync = self._after_op
        unet.grad_finish = self.finish
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        # Mean by pre-division: NCCL's in-switch reduction (NVLS, what makes an 8-GPU NVSwitch all-reduce cheap) exists for SUM but not
        # for AVG (a pre-multiplied sum): with AVG the tuner falls back to 16-channel rings that move 1.75x the bytes.  For a power-of-two
        # world the engine scales the loss gradient by 1/world at the start of backward instead — exact in bf16 / fp32 (an exponent
        # shift commutes with every rounding step of the linear backward pass) — and the buckets are all-reduced with SUM.
        pow2 = self.world > 1 and (self.world & (self.world - 1)) == 0
        if prescale is None:             # automatic: NCCL + power-of-two world; `prescale=True` forces it on any backend (CPU tests)
            prescale = self._avg and os.environ.get("OF_DDP_PRESCALE", "1") != "0"
        self.prescale = bool(prescale and pow2)
        unet.grad_prescale = 1.0 / self.world if self.prescale else 1.0
        # leave `reserve_sms` SMs to the NCCL kernels WHILE buckets are in flight: the persistent GEMM grid shrinks from the first
        # bucket launch of a backward pass to its end (of_set_sm_limit is read at launch time, i.e. baked into a captured graph);
        # forward and the part of backward before the first bucket keep every SM.
        self.reserve_sms = reserve_sms if (dev.type == "cuda" and self.world > 1) else 0
        self._sms = torch.cuda.get_device_properties(dev).multi_processor_count if dev.type == "cuda" else 0
        self._reserved = False
        self.dry_run = False
        self._touch_log = {}
        self._op = None
        self._launched = set()
        self._comm_used = False

    # ---- NCCL user-buffer registration
    def _use_registered_arena(self, store) -> None:
        """Allocate the gradient arena from NCCL's own allocator (ncclMemAlloc) and register it with the communicator: in-place
        all-reduces of registered buffers take NCCL's zero-copy NVLS path on an NVSwitch box (the switch reduces, the SMs only
        issue multimem loads / stores), which needs far fewer CTAs for the same bandwidth than the copy-through-FIFO path.
        Falls back to a plain arena when the backend cannot do it."""
        try:
            pg = self.group if self.group is not None else dist.distributed_c10d._get_default_group()
            dev = next(self.unet.parameters()).device
            backend = pg._get_backend(dev)
            pool = torch.cuda.MemPool(backend.mem_allocator)

            def alloc(n, d):
                try:
                    with torch.cuda.use_mem_pool(pool):
                        t = torch.zeros(n, dtype=torch.float32, device=d)
                    backend.register_mem_pool(pool)
                    return t
                except Exception as e:  # noqa: BLE001
                    self.registered = False
                    print(f"[osufusion_b200.ddp] NCCL-registered arena failed ({type(e).__name__}: {str(e)[:200]}); plain arena", flush=True)
                    return torch.zeros(n, dtype=torch.float32, device=d)
            store.arena_alloc = alloc
            store.arena_key = None           # force a rebuild of the arena from the registered pool
            self._pool, self.registered = pool, True
        except Exception as e:  # noqa: BLE001
            store.arena_alloc = None
            self.registered = False
            print(f"[osufusion_b200.ddp] NCCL-registered gradient arena unavailable ({type(e).__name__}: {str(e)[:200]}); plain arena", flush=True)

    # ---- replica consistency
    def broadcast_parameters(self, model: torch.nn.Module, src: int = 0) -> None:
        """Every parameter and buffer takes rank `src`'s value (what DDP's constructor does); operand caches are invalidated."""
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)
        self.store.param_epoch += 1

    # ---- bucket table
    def _build_buckets(self) -> None:
        """Contiguous slices of the arena in backward-completion order; large buckets first (NVSwitch collectives are latency-
        bound), small ones over the last `tail_bytes` so that what cannot hide under the remaining backward is short."""
        store = self.store
        params, offs = store.arena_params, store.arena_offsets
        total = offs[-1][1] if offs else 0
        tail_start = max(0, total - self.tail_bytes // 4)
        self.buckets = []      # (start, end, [param ids])
        b_start, b_ids, end = 0, [], 0
        for p, (s0, e0) in zip(params, offs):
            b_ids.append(id(p))
            end = e0
            limit = min(self.bucket_bytes, self.tail_bucket_bytes) if s0 >= tail_start else self.bucket_bytes
            if (end - b_start) * 4 >= limit:
                self.buckets.append((b_start, end, b_ids))
                b_start, b_ids = end, []
        if b_ids:
            self.buckets.append((b_start, end, b_ids))
        self.bucket_of = {pid: bi for bi, (_, _, ids) in enumerate(self.buckets) for pid in ids}
        self.ready_at = None           # bucket index -> tape op index after which it is complete (learned in one backward pass)
        self._plan_key = (store.arena_key, store.arena.data_ptr() if store.arena is not None else 0)
        self._tape_len = None

    # ---- hooks called by the engine
    @property
    def arena(self) -> torch.Tensor:
        return self.store.arena

    @property
    def views(self):
        return self.store.arena_views

    def _set_reserved(self, on: bool) -> None:
        if self.reserve_sms <= 0 or on == self._reserved:
            return
        from . import _native as N
        N.lib().of_set_sm_limit(max(2, (self._sms - self.reserve_sms) // 2 * 2) if on else 0)
        self._reserved = on

    def _begin(self) -> None:
        store = self.store
        # the arena is rebuilt whenever the trainable set changes (adapter injection, merge_and_unload, requires_grad_ toggles):
        # old offsets / learned readiness would then address the wrong slices
        if self._plan_key != (store.arena_key, store.arena.data_ptr() if store.arena is not None else 0):
            self._build_buckets()
        self._launched = set()
        self._touch_log = {}

    def _touch(self, pid: int) -> None:
        self._touch_log[pid] = self._op

    def _after_op(self, i: int) -> None:
        """Called by the engine before the first tape op (i = len(tape) + 1) and by Tape.run_backward after tape op i (ops run from
        last to first)."""
        if self._op is None or i > self._op + 1:           # first call of this backward pass: i = tape length + 1
            if self._tape_len is not None and self._tape_len != i:
                self.ready_at = None                       # the tape changed shape: learned readiness is stale
            self._tape_len = i
        self._op = i - 1
        if self.ready_at is None or not self.overlap or self.world == 1:
            return
        for bi, ready in enumerate(self.ready_at):
            if ready is not None and ready >= i and bi not in self._launched:
                self._launch(bi)

    def _reduce(self, t: torch.Tensor) -> None:
        if self.prescale:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        elif self._avg:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:                          # gloo (CPU tests) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def _launch(self, bi: int) -> None:
        self._launched.add(bi)
        s, e, _ = self.buckets[bi]
        if self.dry_run:                   # bench.py: same launch schedule and SM reservation, no collective (isolates its cost)
            self._set_reserved(True)
            return
        if self.comm is None:
            self._reduce(self.arena[s:e])
            return
        self._set_reserved(True)
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        side = self.store.side
        if side is not None and side.active:     # weight gradients of the bucket may still be running on the engine's second stream
            self.comm.wait_stream(side.stream)
        self._comm_used = True
        with torch.cuda.stream(self.comm):
            self._reduce(self.arena[s:e])

    def finish(self) -> None:
        """End of backward (still inside the autograd node): launch what is left, then join the comm stream."""
        if self.world > 1:
            for bi in range(len(self.buckets)):
                if bi not in self._launched:
                    self._launch(bi)
            if self.comm is not None and self._comm_used:      # (a dry run launches nothing there: no edge to an uncaptured stream)
                torch.cuda.current_stream().wait_stream(self.comm)
        self._comm_used = False
        self._set_reserved(False)
        if self.ready_at is None and self._touch_log:
            ready = []
            for _, _, ids in self.buckets:
                idx = [self._touch_log[p] for p in ids if p in self._touch_log and self._touch_log[p] is not None]
                ready.append(min(idx) if idx else 0)
            self.ready_at = ready
        self._op = None

    # kept for API symmetry with bench.py's post_backward hook: all work already happened inside backward
    def all_reduce(self) -> None:
        return None
