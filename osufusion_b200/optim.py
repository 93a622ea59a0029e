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
    """AdamW over `model.unet`'s trainable parameters with the gradient-norm clip fused in (torch.optim.AdamW defaults).

    `param_groups[0]["params"]` lists the trainable parameters in `model.parameters()` order — the order
    `torch.optim.AdamW(model.parameters())` of the reference (trainer.py:230) uses — and `state_dict()` emits torch's own
    per-parameter format (`state[i] = {step, exp_avg, exp_avg_sq}`), so optimizer checkpoints move between the reference and
    this repo in both directions.  Internally the moments live in two flat arenas with the gradient arena's layout."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = 1.0, emit_operands: bool = True) -> None:
        unet = model.unet if hasattr(model, "unet") else model
        self.emit_operands = emit_operands
        self.unet = unet
        store = unet._store
        store.ensure_arena(unet)
        self.store = store
        in_arena = {id(p) for p in store.arena_params}
        params = [p for p in model.parameters() if id(p) in in_arena]
        assert len(params) == len(in_arena), "every trainable parameter of the model must belong to model.unet"
        self._n_model_params = sum(1 for _ in model.parameters())
        self._model_index = [i for i, p in enumerate(model.parameters()) if id(p) in in_arena]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self._step = 0
        self._plan_key = None
        self._tables = {}
        self._build()

    def _table(self, skip: frozenset):
        """Device table of the tensors to update (parameters whose gradient is None are skipped, as torch.optim.AdamW does)."""
        hit = self._tables.get(skip)
        if hit is not None:
            return hit
        st = self.store
        lib = N.lib()
        targets = st.operand_targets(self.unet) if self.emit_operands else {}
        rows, cta, emitted = [], 0, 0
        for p, (s0, _) in zip(st.arena_params, st.arena_offsets):
            if id(p) in skip:
                continue
            assert p.is_contiguous() and p.dtype == torch.float32
            Cout, Cin, k = st.packed.get(id(p), (0, 0, 1))
            dst = targets.get(id(p), 0)
            if k == 1 and p.dim() == 3 and p.shape[2] > 1:
                dst = 0          # a k > 1 conv whose gradient is NOT packed: its operand layout differs from the parameter's
            emitted += 1 if dst else 0
            rows.append(N.OptTensor(p.data_ptr(), s0, p.numel(), dst, cta, Cout, Cin, k))
            n = lib.of_opt_tensor_ctas2(p.numel(), Cout, Cin, k)
            assert n > 0
            cta += n
        if not rows:
            hit = (None, 0, 0, False)
        else:
            arr = (N.OptTensor * len(rows))(*rows)
            # operands are complete only if every weight of the pack plan was refreshed by this table
            complete = bool(targets) and emitted == len(targets)
            hit = (torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(st.arena.device), len(rows), cta, complete)
        self._tables[skip] = hit
        return hit

    def _build(self) -> None:
        st = self.store
        dev = st.arena.device
        self._tables = {}
        if getattr(self, "exp_avg", None) is None or self.exp_avg.numel() != st.arena.numel():
            self.exp_avg = torch.zeros_like(st.arena)
            self.exp_avg_sq = torch.zeros_like(st.arena)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._plan_key = (st.arena.data_ptr(), tuple(p.data_ptr() for p in st.arena_params))
        # the operand buffer moves when the pack plan is rebuilt (adapter injection): tables are rebuilt with it
        self._pack_buf = st.pack_plan["buf"].data_ptr() if st.pack_plan is not None else 0

    @property
    def grad_norm(self) -> torch.Tensor:
        """Global gradient L2 norm of the last step (0-dim tensor on the device; reading it is the only sync)."""
        return self._sumsq.sqrt().float()[0]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        st = self.store
        buf = st.pack_plan["buf"].data_ptr() if st.pack_plan is not None else 0
        if self._plan_key != (st.arena.data_ptr(), tuple(p.data_ptr() for p in st.arena_params)) or buf != self._pack_buf:
            self._build()
        # gradients must be the engine's arena views (they are after a backward pass of the engine); anything else is copied in
        skip = []
        for p in st.arena_params:
            v = st.arena_views[id(p)]
            if p.grad is None:
                v.zero_()                # keeps the global norm right; the tensor itself is left untouched (torch skips grad None)
                skip.append(id(p))
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
        table, num, ctas, complete = self._table(frozenset(skip))
        g = self.param_groups[0]
        self._step += 1
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        if clip:
            N.call("of_grad_sumsq", st.arena.data_ptr(), st.arena.numel(), self._sumsq.data_ptr())
        if table is not None:
            N.call("of_adamw_step", table.data_ptr(), num, ctas, st.arena.data_ptr(), self.exp_avg.data_ptr(),
                   self.exp_avg_sq.data_ptr(), self._sumsq.data_ptr() if clip else None, float(self.max_grad_norm or 0.0), float(g["lr"]),
                   float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self._step)
        st.param_epoch += 1     # parameters changed through raw pointers (torch's _version did not move): invalidate operand caches
        if complete:
            st.mark_operands_current()      # ... except the grouped bf16 GEMM operands, which the kernel has just rewritten
        return loss

    # ---- checkpoints: torch.optim.AdamW's own state layout
    def _moment_views(self, p):
        (s0, _) = self.store.arena_offsets[self._arena_index[id(p)]]
        if id(p) in self.store.packed:      # moments share the gradient's GEMM layout [k][Cout][Cin]
            Cout, Cin, k = self.store.packed[id(p)]
            return (self.exp_avg[s0:s0 + p.numel()].view(k, Cout, Cin).permute(1, 2, 0),
                    self.exp_avg_sq[s0:s0 + p.numel()].view(k, Cout, Cin).permute(1, 2, 0))
        return self.exp_avg[s0:s0 + p.numel()].view(p.shape), self.exp_avg_sq[s0:s0 + p.numel()].view(p.shape)

    @property
    def _arena_index(self):
        return {id(p): i for i, p in enumerate(self.store.arena_params)}

    def state_dict(self):
        """torch format: `state[i] = {"step", "exp_avg", "exp_avg_sq"}` with i indexing the trainable parameters in
        `model.parameters()` order (empty before the first step, like torch)."""
        self.state.clear()
        if self._step > 0:
            for p in self.param_groups[0]["params"]:
                m, v = self._moment_views(p)
                self.state[p] = {"step": torch.tensor(float(self._step)), "exp_avg": m, "exp_avg_sq": v}
        d = super().state_dict()
        self.state.clear()
        return d

    def load_state_dict(self, sd):
        """Accepts this class's own dict and a reference `torch.optim.AdamW(model.parameters())` dict — including one saved over ALL
        model parameters with some of them frozen (trainer_peft.py: state only for the adapter tensors)."""
        sd = {k: v for k, v in sd.items() if k != "fused"}          # round-1 private key: ignored
        groups = sd["param_groups"]
        n_here = len(self.param_groups[0]["params"])
        ids = [i for g in groups for i in g["params"]]
        if len(groups) == 1 and len(ids) == self._n_model_params and n_here != self._n_model_params:
            keep = [ids[i] for i in self._model_index]                # indices of our trainable tensors in the reference's numbering
            remap = {old: new for new, old in enumerate(keep)}
            sd = {"state": {remap[k]: v for k, v in sd["state"].items() if k in remap},
                  "param_groups": [dict(groups[0], params=list(range(n_here)))]}
        super().load_state_dict(sd)                                  # raises ValueError on a genuine size mismatch
        steps = []
        for p in self.param_groups[0]["params"]:
            stt = self.state.get(p)
            m, v = self._moment_views(p)
            if stt:
                m.copy_(stt["exp_avg"])
                v.copy_(stt["exp_avg_sq"])
                steps.append(int(float(stt["step"])))
            else:
                m.zero_()
                v.zero_()
        self._step = max(steps) if steps else 0
        self.state.clear()
