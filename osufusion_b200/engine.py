"""Execution engine of the B200 denoiser: a hand-rolled forward/backward "tape" over the C-ABI kernels.

torch is used for device memory (torch.empty/zeros), the CUDA stream, parameters and tiny (B, dim_emb) glue; every
FLOP and every full-tensor pass of the denoiser runs in libosufusion_sm100.so.  torch.autograd sees the whole UNet as
ONE node (modules.UNetFunction): the backward pass is the reversed tape below, which lets gradients be accumulated by
GEMM epilogues / fp32 atomics in place and lets the data-parallel wrapper launch NCCL buckets while backward still runs.

Numerics = the reference under CUDA bf16 autocast (mode "M1"): bf16 GEMM/attention operands with
fp32 accumulation, fp32 residual stream, fp32 norms; rounding points of the reference are reproduced where cheap.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional

import torch

from . import _native as N
from . import ops_raw as R

BF16, F32, F64 = torch.bfloat16, torch.float32, torch.float64
# GlobalContext forward: one-pass online-softmax pooling (of_rb_logit_pool) vs logits / softmax / pooling as three launches
GCTX_FUSED = os.environ.get("OF_GCTX_FUSED", "1") != "0"
# LoRA / DoRA backward through the rank-r activations (no full weight gradient of the frozen base) vs full wgrad + projection
LORA_RANK_R = os.environ.get("OF_LORA_RANK_R", "1") != "0"
# effective-weight merge W + scaling*B*A as a K = r tensor-core GEMM (+ a row-wise norm/scale/pack kernel) vs the CUDA-core kernel
LORA_MERGE_TC = os.environ.get("OF_LORA_MERGE_TC", "1") != "0"
# adapter-side operands of ALL adapted layers from one grouped launch per step, rank-r glue folded into the merge kernel, one grouped
# finishing launch at the end of backward, lora_A conv gradients accumulated in the GEMM layout (no per-layer cast / pack / prep /
# unpack / finish launches)
LORA_GROUPED = os.environ.get("OF_LORA_GROUPED", "1") != "0"
# training forward with adapters: every layer's effective-weight merge is issued at the start of the step on the second stream, in the
# order the forward pass will need them (learned from the previous forward); the main stream waits per layer
LORA_MERGE_AHEAD = os.environ.get("OF_LORA_MERGE_AHEAD", "1") != "0"
# conv-weight gradients / moments in the GEMM layout [k][Cout][Cin] (no unpack pass, no per-layer scratch fill)
PACKED_ARENA = os.environ.get("OF_PACKED_ARENA", "1") != "0"
# weight / bias gradients of backward on a second stream (parallel branches of the captured graph), off the dgrad critical path
WGRAD_SIDE = os.environ.get("OF_WGRAD_SIDE", "1") != "0"
# the dgrad GEMM that completes an activation's gradient also emits its bf16 copy when the producer's backward needs one as a GEMM
# operand (no separate fp32 -> bf16 cast pass at the start of TransformerBlock / sampler / CrossEmbed backward)
GRAD16 = os.environ.get("OF_GRAD16", "1") != "0"
# training forward: the audio encoder (independent of x and t) on the second stream, concurrent with the down path
FWD_SIDE = os.environ.get("OF_FWD_SIDE", "1") != "0"


def _p(t):
    return None if t is None else t.data_ptr()


def _bl(t: torch.Tensor):
    assert t.dim() == 3 and t.stride(2) == 1, (t.shape, t.stride())
    return t.stride(0), t.stride(1)


class Act:
    """Channels-last activation (B, L, C): fp32 residual-stream copy and/or bf16 GEMM-operand copy, plus its fp32 grad."""
    __slots__ = ("f32", "bf16", "grad", "grad16", "want16")

    def __init__(self, f32: Optional[torch.Tensor] = None, bf16: Optional[torch.Tensor] = None) -> None:
        self.f32, self.bf16, self.grad = f32, bf16, None
        # grad16: bf16 copy of the COMPLETE gradient, emitted by the dgrad GEMM that makes the last contribution (ResidualBlock's conv1)
        # when the producer of this activation asked for it (want16): its backward then needs no fp32 -> bf16 cast pass
        self.grad16, self.want16 = None, False

    def take_grad16(self, shape, device) -> torch.Tensor:
        """(gradient as a bf16 GEMM operand, fp32 gradient); clears both.  Uses the epilogue-emitted copy when there is one."""
        g, g16 = self.grad, self.grad16
        self.grad = self.grad16 = None
        if g16 is None:
            g16 = empty(shape, BF16, device)
            cast_copy(g, g16)
        return g16, g

    @property
    def shape(self):
        return (self.f32 if self.f32 is not None else self.bf16).shape

    def add_grad(self, g: torch.Tensor) -> None:
        """Accumulate an fp32 (B, L, C) gradient view; the first contribution is adopted without a copy."""
        self.grad16 = None
        if self.grad is None:
            self.grad = g
        else:
            cast_copy(g, self.grad, accumulate=True)


class Tape:
    def __init__(self) -> None:
        self.ops: List[Callable[[], None]] = []

    def push(self, fn: Callable[[], None]) -> None:
        self.ops.append(fn)

    def run_backward(self, after_op: Optional[Callable[[int], None]] = None, side: Optional["SideLane"] = None) -> None:
        n = len(self.ops)
        for i in range(n - 1, -1, -1):
            self.ops[i]()
            self.ops[i] = None
            if side is not None:
                side.rotate()
            if after_op is not None:
                after_op(i)
        self.ops.clear()


class SideLane:
    """Second CUDA stream for the part of backward that nothing downstream waits for: weight- and bias-gradient kernels.

    The critical path of backward is the chain dY -> dgrad GEMM -> norm / activation backward -> next dgrad; the weight-gradient
    GEMMs (half of backward's FLOPs) only feed the gradient arena.  Issued in stream order they sit between the chain's kernels,
    so every under-filled launch of the chain (65-86 % SM fill at the deep levels, one-CTA-per-sample reductions, the tails of
    33 us persistent GEMMs) leaves SMs idle.  On a second stream — a parallel branch once the step is captured as a CUDA graph —
    their CTAs are scheduled into exactly those holes.

    Memory safety without `record_stream` (which would defer every free to the end of a capture): tensors a side kernel reads are
    kept alive here and released only after the main stream has waited for the side work of that generation; a generation = one
    tape op (one block's backward), and the main stream may run at most `LAG` generations ahead."""
    LAG = int(os.environ.get("OF_SIDE_LAG", "2"))

    def __init__(self, device) -> None:
        self.stream = torch.cuda.Stream(device=device)
        self.gens: list = []
        self.cur: list = []
        self.active = False

    def run(self, fn: Callable[[], None], *keep) -> None:
        main = torch.cuda.current_stream()
        self.stream.wait_stream(main)         # everything launched so far (the producers of dY, the saved activations) is visible
        with torch.cuda.stream(self.stream):
            fn()
        self.cur.append(keep)
        self.active = True

    def rotate(self) -> None:
        """End of a tape op: close the current generation, make the main stream wait for the one LAG ops back and drop its tensors."""
        ev = None
        if self.cur:
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.gens.append((ev, self.cur))
        self.cur = []
        while len(self.gens) > self.LAG:
            old, _keep = self.gens.pop(0)
            if old is not None:
                torch.cuda.current_stream().wait_event(old)

    def join(self) -> None:
        if self.active:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.gens, self.cur, self.active = [], [], False


# ------------------------------------------------------------------------------------------------ raw helpers
def empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


class ZeroPool:
    """Bump allocator over freshly zeroed chunks: the hundreds of tiny zero-initialised scratch tensors of one step (GroupNorm
    statistics, per-channel accumulators, gate-MLP gradients) cost one fill per 4 MB chunk instead of one launch each.  Views
    keep their chunk alive, so lifetimes are ordinary tensor lifetimes."""
    CHUNK = 4 << 20

    def __init__(self, device) -> None:
        self.device, self.buf, self.off = device, None, 0

    def take(self, shape, dtype) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        aligned = (nbytes + 255) // 256 * 256
        if aligned > self.CHUNK // 4:
            return torch.zeros(shape, dtype=dtype, device=self.device)
        if self.buf is None or self.off + aligned > self.CHUNK:
            self.buf = torch.zeros(self.CHUNK, dtype=torch.uint8, device=self.device)
            self.off = 0
        v = self.buf[self.off:self.off + nbytes].view(dtype).view(shape)
        self.off += aligned
        return v


_POOL: Optional[ZeroPool] = None     # pool of the forward/backward pass currently being recorded (set by Ctx / backward_from)


def use_pool(pool: Optional[ZeroPool]) -> None:
    global _POOL
    _POOL = pool


def zeros(shape, dtype, device):
    if _POOL is not None and _POOL.device == device:
        return _POOL.take(tuple(shape), dtype)
    return torch.zeros(shape, dtype=dtype, device=device)


def cast_copy(src: torch.Tensor, dst: torch.Tensor, accumulate: bool = False) -> None:
    """dst (fp32|bf16 (B,L,C) view) = / += src (fp32|bf16 (B,L,C) view)."""
    B, L, Cc = src.shape
    s_bs, s_ld = _bl(src)
    d_bs, d_ld = _bl(dst)
    N.call("of_cast_copy", _p(src) if src.dtype == F32 else None, _p(src) if src.dtype == BF16 else None, s_ld, s_bs,
           B, L, Cc, _p(dst) if dst.dtype == F32 else None, _p(dst) if dst.dtype == BF16 else None, d_ld, d_bs,
           int(accumulate))


def linear_small_fwd(x, W, bias, act=0, want_pre=False):
    """x (M,K) fp32 contiguous, W (N,K[,1]) fp32 parameter -> y (M,N) fp32 (bf16-rounded values)."""
    M, K = x.shape
    Nn = W.shape[0]
    y = empty((M, Nn), F32, x.device)
    pre = empty((M, Nn), F32, x.device) if want_pre else None
    N.call("of_linear_small_fwd", _p(x), x.stride(0), M, Nn, K, _p(W), K, _p(bias), act, 1, _p(y), Nn, _p(pre))
    return y, pre


def linear_small_bwd(dy, pre, act, x, W, dW, dbias, dx, accumulate=True):
    M, K = x.shape
    Nn = W.shape[0]
    N.call("of_linear_small_bwd", _p(dy), dy.stride(0), _p(pre), act, _p(x), x.stride(0), M, Nn, K, _p(W), K, 1, _p(dW),
           _p(dbias), _p(dx), dx.stride(0) if dx is not None else 0, int(accumulate))


def linear_small_bwd_param(st, dy, pre, act, x, w, b, dx):
    """Backward of a small Linear / 1x1-conv whose parameters are `w` (viewed as (N, K)) and `b`: gradients go straight into the
    arena; the first contribution of a step overwrites (the arena holds zeros), later ones accumulate."""
    fresh = id(w) not in st.touched and (b is None or id(b) not in st.touched)
    linear_small_bwd(dy, pre, act, x, w.view(w.shape[0], -1), st.grad_opt(w), st.grad_opt(b), dx, accumulate=not fresh)


def colsum(dy16: torch.Tensor, out: torch.Tensor) -> None:
    """out[n] += sum over all rows of the (B, L, N) bf16 view (rows must be uniformly strided: bs == L*ld)."""
    B, L, Nn = dy16.shape
    bs, ld = _bl(dy16)
    if B > 1 and bs != L * ld:
        for b in range(B):
            N.call("of_colsum_bf16", _p(dy16[b]), ld, L, Nn, _p(out))
    else:
        N.call("of_colsum_bf16", _p(dy16), ld, B * L, Nn, _p(out))


def _base(m):
    """The nn.Conv1d / nn.Linear holding the (possibly frozen) base weight: adapters wrap it as `.base_layer` (peft naming)."""
    return getattr(m, "base_layer", m)


def _adapter(m):
    return m if hasattr(m, "base_layer") else None


def _lora_fast(ad) -> bool:
    """Adapted layer handled by the grouped LoRA / DoRA path (ParamStore.refresh_lora): tensor-core merge + rank-r backward."""
    W = ad.base_layer.weight
    Cin = W.shape[1]
    k = W.shape[2] if W.dim() == 3 else 1
    return (LORA_GROUPED and LORA_RANK_R and LORA_MERGE_TC and ad.r % 8 == 0 and Cin % 8 == 0 and Cin * k <= 8192 and k <= 4
            and W.is_contiguous() and ad.lora_A["default"].weight.requires_grad)


# ------------------------------------------------------------------------------------------------ parameter staging
def backward_param_order(unet) -> List[torch.nn.Parameter]:
    """Parameters in the order their gradients become final during the engine's backward pass (see backward_param_plan)."""
    return backward_param_plan(unet)[0]


def packed_conv_params(unet):
    """{id(weight): (Cout, Cin, k)} of the k > 1 Conv1d weights whose gradient comes from the implicit-GEMM weight-gradient kernel
    (`_wgrad_conv`): un-adapted `Block.proj`, `Upsample.conv`, `Parallel.fns[0]` with Cin a multiple of 8.  Their slice of the
    gradient arena (and of the optimizer's moment arenas) uses the GEMM layout [k][Cout][Cin], so the split-K accumulation lands in
    place and no unpack pass exists; the fp32 master weight keeps torch's (Cout, Cin, k) layout."""
    out = {}

    def add(conv):
        ad = _adapter(conv)
        if ad is not None:
            A = ad.lora_A["default"].weight      # (r, Cin, k): its rank-r weight gradient lands in [k][r][Cin] as well
            if _lora_fast(ad) and A.dim() == 3 and A.shape[2] > 1:
                out[id(A)] = tuple(A.shape)
            return
        w = conv.weight
        Cout, Cin, k = w.shape
        if 1 < k <= 4 and Cin % 8 == 0 and w.requires_grad:
            out[id(w)] = (Cout, Cin, k)

    for m in unet.modules():
        kind = type(m).__name__
        if kind == "ResidualBlock":
            add(m.block1.proj)
            add(m.block2.proj)
        elif kind == "Upsample":
            add(m.conv)
        elif kind == "Parallel":
            add(m.fns[0])
    return out


def film_units(unet):
    """Top-level units of the denoiser in FORWARD order, each a list of modules whose FiLM heads (`ResidualBlock.mlp`) get their
    weight gradients from ONE grouped `of_film_bwd` launch as soon as the unit's backward has run (modules.UNet.denoise)."""
    units = [[layer] for layer in unet.down_layers]
    units.append([unet.middle_resnet1, *unet.middle_transformer, unet.middle_resnet2])
    units += [[layer] for layer in unet.up_layers]
    units.append([unet.final_resnet])
    return units


def backward_param_plan(unet):
    """(order, film_ranges): parameters in the order their gradients become final during the engine's backward pass.
    The FiLM heads (`ResidualBlock.mlp`) of one unit (a UNetBlock, the middle, the final resnet) get their gradients from one
    grouped kernel right after that unit's backward, so they sit together at the END of the unit's parameters; `film_ranges`
    lists their (start, end) positions in `order`.  The audio encoder's backward runs right after `middle_resnet1` (its forward
    is recorded after the down path, modules.UNet.run), so the tail of backward is the cheap first down block."""
    order: List[torch.nn.Parameter] = []
    seen = set()
    film_ranges = []

    def add(module, film):
        for name, p in module.named_parameters():
            if id(p) in seen:
                continue
            seen.add(id(p))
            (film if name.startswith("mlp.") or ".mlp." in name else order).append(p)

    def add_block(blk, film):
        add(blk.sampler, film)
        for res, tr in reversed(list(zip(blk.resnets, blk.transformers))):
            add(tr, film)
            add(res, film)
        add(blk.init_resnet, film)

    def close(film):
        if film:
            film_ranges.append((len(order), len(order) + len(film)))
            order.extend(film)

    add(unet.final_conv, [])
    film = []
    add(unet.final_resnet, film)
    close(film)
    for blk in reversed(unet.up_layers):
        film = []
        add_block(blk, film)
        close(film)
    film = []
    add(unet.middle_resnet2, film)
    for tr in reversed(unet.middle_transformer):
        add(tr, film)
    add(unet.middle_resnet1, film)
    close(film)
    for blk in reversed(unet.audio_encoder.layers):
        film = []
        add_block(blk, film)
        close(film)
    add(unet.audio_encoder.init_conv, [])
    for blk in reversed(unet.down_layers):
        film = []
        add_block(blk, film)
        close(film)
    add(unet.init_x, [])
    for p in list(unet.time_mlp.parameters()) + list(unet.cond_mlp.parameters()) + [unet.null_cond]:
        if id(p) not in seen:
            seen.add(id(p))
            order.append(p)
    for p in unet.parameters():          # anything not covered above
        if id(p) not in seen:
            seen.add(id(p))
            order.append(p)
    return order, film_ranges


class ParamStore:
    """bf16 GEMM-operand copies of the fp32 master parameters and the fp32 gradient arena the kernels accumulate into.

    `refresh=True` (training): operand copies are rebuilt on every forward (the reference's autocast also re-casts every
    weight each iteration) by ONE grouped launch.  `refresh=False` (sampling): copies are cached per parameter version.

    Gradients of all trainable parameters live in ONE flat fp32 arena laid out in backward-completion order: one memset
    per step instead of one fill per tensor, `p.grad` are views of it, and the data-parallel wrapper all-reduces contiguous
    slices in place (osufusion_b200/ddp.py).
    """
    ALIGN = 32   # floats: every gradient view starts 128-byte aligned (kernels use 16-byte vector atomics)

    def __init__(self) -> None:
        self.cache = {}
        self.refresh = True
        self.epoch = 0
        self.arena = None
        self.arena_views = {}
        self.arena_params = []       # trainable parameters in arena order
        self.arena_offsets = []      # (start, end) in floats, aligned
        self.arena_key = None
        self.film_floats = None      # [(start, end)] float ranges of the FiLM-head blocks (written, never accumulated -> not zeroed)
        self.packed = {}             # id(p) -> (Cout, Cin, k): conv weights whose gradient lives in the GEMM layout [k][Cout][Cin]
        self.arena_packed_views = {} # id(p) -> (k, Cout, Cin) contiguous view of the same arena slice
        self.touched = set()
        self.param_epoch = 0         # bumped by optimizers that update parameters through raw pointers (osufusion_b200/optim.py)
        self.pack_plan = None
        self.lora_plan = None
        self._merge_log = None        # merges of the forward pass being recorded, in first-use order: [(key, thunk)]
        self._merge_order = None      # ... of the previous forward pass (what merge-ahead replays)
        self._merge_events = {}
        self._merge_seen = set()
        self._merge_stream = None
        self._lora_ads = None         # (parameter count, [adapted layers on the grouped path]): the module tree is walked once
        self._film_heads = None
        self._n_adapters = None
        self.film_plans = {}
        self.arena_alloc = None      # optional allocator of the gradient arena (ddp: NCCL-registered memory), (numel, device) -> zeros
        self.side: Optional[SideLane] = None     # second stream of the current backward pass (weight / bias gradients), or None
        # set by ddp.GradAllReducer
        self.on_backward_begin = None
        self.on_touch = None

    # ---- gradient arena
    def ensure_arena(self, unet) -> None:
        custom = getattr(unet, "backward_param_order", None)
        if custom is not None:       # other backbones (osufusion_b200/backbones.py) supply their own completion order; no FiLM block
            order, franges = custom(), []
        else:
            order, franges = backward_param_plan(unet)
        film_ids = {id(p) for fs, fe in franges for p in order[fs:fe]}
        params = [p for p in order if p.requires_grad]
        key = tuple(id(p) for p in params) + (str(params[0].device) if params else "",)
        if key == self.arena_key:
            return
        if not params:
            self.arena, self.arena_views, self.arena_params, self.arena_offsets = None, {}, [], []
            self.film_floats, self.arena_key, self.film_plans = None, key, {}
            return
        A = self.ALIGN
        total = sum((p.numel() + A - 1) // A * A for p in params)
        dev = params[0].device
        self.arena = (self.arena_alloc or (lambda n, d: torch.zeros(n, dtype=F32, device=d)))(max(total, 1), dev)
        self.arena_views, self.arena_params, self.arena_offsets = {}, params, []
        self.packed = packed_conv_params(unet) if PACKED_ARENA and custom is None else {}
        self.arena_packed_views = {}
        off = 0
        ranges = []
        for p in params:
            n = (p.numel() + A - 1) // A * A
            if id(p) in self.packed:
                # gradient kept in the weight-gradient GEMM's own layout [k][Cout][Cin]; `p.grad` is the (Cout, Cin, k) permuted view
                Cout, Cin, k = self.packed[id(p)]
                pv = self.arena[off:off + p.numel()].view(k, Cout, Cin)
                self.arena_packed_views[id(p)] = pv
                self.arena_views[id(p)] = pv.permute(1, 2, 0)
            else:
                self.arena_views[id(p)] = self.arena[off:off + p.numel()].view(p.shape)
            self.arena_offsets.append((off, off + n))
            if id(p) in film_ids:
                if ranges and ranges[-1][1] == off:
                    ranges[-1][1] = off + n
                else:
                    ranges.append([off, off + n])
            off += n
        self.film_floats = [tuple(r) for r in ranges] or None
        self.arena_key = key
        self.film_plans = {}

    def begin_backward(self, unet, film_overwritten: bool) -> None:
        """Start of a backward pass: detach any `p.grad` still aliasing the arena (gradient accumulation without
        zero_grad(set_to_none=True)), zero the arena, forget which gradients have been produced."""
        self.ensure_arena(unet)
        for p in self.arena_params:
            g = p.grad
            if g is not None and g.data_ptr() == self.arena_views[id(p)].data_ptr():
                p.grad = g.clone()
        if self.arena is None:
            pass
        elif film_overwritten and self.film_floats is not None:
            pos = 0
            for f0, f1 in self.film_floats:      # the FiLM-head gradients are plain stores: zero only what lies between them
                if f0 > pos:
                    self.arena[pos:f0].zero_()
                pos = f1
            if pos < self.arena.numel():
                self.arena[pos:].zero_()
        else:
            self.arena.zero_()
        self.touched = set()
        if self.lora_plan is not None:
            self.lora_plan["acc32"].zero_()      # dB_raw / d magnitude accumulators of every adapted layer: one memset
            self.lora_plan["pending"] = False
        if self.on_backward_begin is not None:
            self.on_backward_begin()

    def begin_forward(self, refresh: bool, unet=None) -> None:
        self.refresh = refresh
        self.epoch += 1
        if unet is not None:
            self.refresh_operands(unet)
            self._merge_ahead()

    def _merge_ahead(self) -> None:
        """LoRA / DoRA training forward: issue every effective-weight merge now, on the second stream, in first-use order."""
        order, self._merge_order = self._merge_order, None
        self._merge_events = {}
        self._merge_log = [] if (self.lora_plan is not None and self.refresh) else None
        self._merge_seen = set()
        if self._merge_log is None or not order or not LORA_MERGE_AHEAD or self.lora_plan["buf16"].device.type != "cuda":
            return
        if order[0] != self.lora_plan["ptrs"]:
            return
        if self._merge_stream is None:
            self._merge_stream = torch.cuda.Stream(device=self.lora_plan["buf16"].device)
        side, main = self._merge_stream, torch.cuda.current_stream()
        side.wait_stream(main)
        pool = _POOL
        use_pool(None)           # scratch chunks are zero-filled on the stream that creates them: none may be born on the side stream
        try:
            with torch.cuda.stream(side):
                for key, thunk in order[1]:
                    thunk()
                    ev = torch.cuda.Event()
                    ev.record(side)
                    self._merge_events[key] = ev
        finally:
            use_pool(pool)

    def _merged(self, key, thunk, val):
        """Bookkeeping of one adapted operand at its first use in a forward pass: log it for the next step's merge-ahead and make
        the main stream wait for the side-stream merge that produced it."""
        if self._merge_log is not None and key not in self._merge_seen:
            self._merge_seen.add(key)
            self._merge_log.append((key, thunk))
        ev = self._merge_events.pop(key, None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        return val

    def end_forward(self) -> None:
        if self._merge_log is not None and self.lora_plan is not None:
            self._merge_order = (self.lora_plan["ptrs"], self._merge_log)
        self._merge_log = None
        for ev in self._merge_events.values():       # merged but unused this pass (should not happen): still join the side stream
            torch.cuda.current_stream().wait_event(ev)
        self._merge_events = {}

    # ---- grouped operand packing: every plain Conv1d / Linear weight of the denoiser in one launch
    def _build_pack_plan(self, unet):
        entries = []     # (cache key, params, [(param, Cout, Cin, k, cin_pad, element offset)], shape)
        total = 0

        def conv_entry(w):
            nonlocal total
            Cout, Cin, k = w.shape
            cp = (Cin + 7) // 8 * 8
            if k > 4:
                return
            entries.append((("conv", id(w)), (w,), [(w, Cout, Cin, k, cp, total)], (k, Cout, cp), total))
            total += (k * Cout * cp + 127) // 128 * 128

        def lin_entry(*ws):
            nonlocal total
            K = ws[0].shape[1]
            segs, r = [], 0
            for w in ws:
                segs.append((w, w.shape[0], K, 1, K, total + r * K))
                r += w.shape[0]
            entries.append((("lin",) + tuple(id(w) for w in ws), tuple(ws), segs, (r, K), total))
            total += (r * K + 127) // 128 * 128

        for m in unet.modules():
            kind = type(m).__name__
            if kind == "ResidualBlock":
                for blk in (m.block1, m.block2):
                    if _adapter(blk.proj) is None:
                        conv_entry(blk.proj.weight)
                if not isinstance(m.res_conv, torch.nn.Identity):
                    lin_entry(m.res_conv.weight)
            elif kind == "TransformerBlock":
                at = m.attn
                if _adapter(at.to_q) is None and _adapter(at.to_kv) is None:
                    lin_entry(at.to_q.weight, at.to_kv.weight)
                lin_entry(at.to_out.weight)
                lin_entry(m.ff[0].weight)
                lin_entry(m.ff[2].weight)
            elif kind == "Upsample":
                conv_entry(m.conv.weight)
            elif kind == "Parallel":
                conv_entry(m.fns[0].weight)
                lin_entry(m.fns[1].weight)
            elif kind == "DiTBlock":          # osufusion_b200/backbones.py
                lin_entry(m.attn.to_qkv.weight)
                lin_entry(m.ff[0].weight)
                lin_entry(m.ff[2].weight)
            elif kind == "MMDiTBlock":
                for s_ in ("x", "a"):
                    at = m.attn
                    lin_entry(getattr(at, f"to_q_{s_}").weight, getattr(at, f"to_k_{s_}").weight, getattr(at, f"to_v_{s_}").weight)
                    lin_entry(getattr(m, f"attn_out_{s_}").weight)
                    ff = getattr(m, f"mlp_{s_}")
                    lin_entry(ff[0].weight)
                    lin_entry(ff[2].weight)
            elif kind == "FinalLayer":
                lin_entry(m.linear.weight)
        if not entries:
            return None
        dev = entries[0][1][0].device
        buf = torch.empty(total, dtype=BF16, device=dev)
        lib = N.lib()
        segs = []
        segs_host = [seg for _, _, seglist, _, _ in entries for seg in seglist]
        cta = 0
        for _, _, seglist, _, _ in entries:
            for (w, Cout, Cin, k, cp, off) in seglist:
                n = lib.of_pack_seg_ctas(Cout, Cin, k, cp)
                assert n > 0
                segs.append(N.PackSeg(w.data_ptr(), buf.data_ptr() + 2 * off, Cout, Cin, k, cp, cta, 0))
                cta += n
        arr = (N.PackSeg * len(segs))(*segs)
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        views = []
        for key, params, _, shape, off0 in entries:
            n = 1
            for d in shape:
                n *= d
            views.append((key, params, buf[off0:off0 + n].view(shape)))
        ptrs = tuple(p.data_ptr() for _, ps, _ in views for p in ps)
        return {"table": table, "num_segs": len(segs), "ctas": cta, "views": views, "ptrs": ptrs, "buf": buf, "vers": None,
                "segs_host": segs_host}

    def refresh_operands(self, unet) -> None:
        plan = self.pack_plan
        n_params = sum(1 for _ in unet.parameters())          # adapter injection / merge_and_unload changes the parameter count
        if self._n_adapters is None or self._n_adapters[0] != n_params:
            self._n_adapters = (n_params, sum(1 for m in unet.modules() if hasattr(m, "base_layer")))
        n_adapters = self._n_adapters[1]
        if plan is not None:
            ptrs = tuple(p.data_ptr() for _, ps, _ in plan["views"] for p in ps)
            if ptrs != plan["ptrs"] or plan.get("n_adapters") != n_adapters:
                plan = None
        if plan is None:
            plan = self.pack_plan = self._build_pack_plan(unet)
            if plan is None:
                return
            plan["n_adapters"] = n_adapters
        vers = tuple(p._version for _, ps, _ in plan["views"] for p in ps) + (self.param_epoch,)
        if vers != plan["vers"]:
            # weights changed since the operand buffer was written (foreign optimizer, load_state_dict, ...).  The fused optimizer
            # (optim.FusedAdamW) emits the bf16 operands itself while it updates the fp32 masters and marks them current
            # (mark_operands_current), so a training loop built on it never comes through here.
            N.call("of_pack_weights", plan["table"].data_ptr(), plan["num_segs"], plan["ctas"])
            plan["vers"] = vers
        for key, params, view in plan["views"]:
            self.cache[key] = (tuple((p.data_ptr(), p._version) for p in params) + (self.param_epoch,), view, self.epoch)
        self.refresh_lora(unet, n_adapters)

    # ---- grouped LoRA / DoRA operands
    def refresh_lora(self, unet, n_adapters: int) -> None:
        """bf16 operands of every adapted layer's adapter side — A as the merge GEMM's operand (r, Cin*k), A in the conv layout
        [k][r][Cin] for u = conv(x, A), scaling * B — written by ONE grouped launch (of_pack_weights) whenever an adapter tensor
        changed; plus persistent per-layer buffers for what the merge kernel leaves for backward (n2, rowscale, (scaling s B)^T) and
        the accumulators backward fills (dB_raw, d magnitude)."""
        if n_adapters == 0 or not LORA_GROUPED:
            self.lora_plan = None
            return
        n_params = self._n_adapters[0] if self._n_adapters else -1
        if self._lora_ads is None or self._lora_ads[0] != n_params:
            self._lora_ads = (n_params, [m for m in unet.modules() if hasattr(m, "base_layer") and _lora_fast(m)])
        ads = self._lora_ads[1]
        if not ads:
            self.lora_plan = None
            return
        plan = self.lora_plan
        ptrs = tuple(t.data_ptr() for ad in ads for t in (ad.lora_A["default"].weight, ad.lora_B["default"].weight))
        if plan is None or plan["ptrs"] != ptrs:
            plan = self.lora_plan = self._build_lora_plan(ads, ptrs)
        vers = tuple(t._version for ad in ads for t in (ad.lora_A["default"].weight, ad.lora_B["default"].weight)) + (self.param_epoch,)
        if vers != plan["vers"]:
            N.call("of_pack_weights", plan["table"].data_ptr(), plan["num_segs"], plan["ctas"])
            plan["vers"] = vers

    def _build_lora_plan(self, ads, ptrs):
        lib = N.lib()
        dev = ads[0].base_layer.weight.device
        al = lambda n: (n + 127) // 128 * 128
        n16 = n32 = nacc = 0
        lay = []
        for ad in ads:
            W = ad.base_layer.weight
            Cout, Cin = W.shape[0], W.shape[1]
            k = W.shape[2] if W.dim() == 3 else 1
            r, E = ad.r, Cin * k
            o = {"A16": n16}
            n16 += al(r * E)
            if k > 1:
                o["Apk"] = n16
                n16 += al(r * E)
            o["B16s"] = n16
            n16 += al(Cout * r)
            o["Bst"] = n16
            n16 += al(Cout * r)
            o["n2"], o["rowscale"] = n32, n32 + al(Cout)
            n32 += 2 * al(Cout)
            o["dBraw"], o["dm"] = nacc, nacc + al(Cout * r)
            nacc += al(Cout * r) + al(Cout)
            lay.append((ad, Cout, Cin, k, r, E, o))
        buf16 = torch.zeros(n16, dtype=BF16, device=dev)
        buf32 = torch.zeros(n32, dtype=F32, device=dev)
        acc32 = torch.zeros(nacc, dtype=F32, device=dev)
        segs, cta, entries = [], 0, {}

        def seg(src, dst_off, Cout, Cin, k, cp, scale):
            nonlocal cta
            n = lib.of_pack_seg_ctas(Cout, Cin, k, cp)
            assert n > 0
            segs.append(N.PackSeg(src.data_ptr(), buf16.data_ptr() + 2 * dst_off, Cout, Cin, k, cp, cta, scale))
            cta += n

        for ad, Cout, Cin, k, r, E, o in lay:
            A, Bm = ad.lora_A["default"].weight, ad.lora_B["default"].weight
            seg(A, o["A16"], r, E, 1, E, 0.0)
            if k > 1:
                seg(A, o["Apk"], r, Cin, k, Cin, 0.0)
            seg(Bm, o["B16s"], Cout, r, 1, r, float(ad.scaling))
            A16 = buf16[o["A16"]:o["A16"] + r * E].view(r, E)
            entries[id(ad)] = {
                "A16": A16, "Apk": buf16[o["Apk"]:o["Apk"] + r * E].view(k, r, Cin) if k > 1 else A16.view(1, r, Cin),
                "B16s": buf16[o["B16s"]:o["B16s"] + Cout * r].view(Cout, r), "Bst": buf16[o["Bst"]:o["Bst"] + Cout * r].view(r, Cout),
                "n2": buf32[o["n2"]:o["n2"] + Cout], "rowscale": buf32[o["rowscale"]:o["rowscale"] + Cout],
                "dBraw": acc32[o["dBraw"]:o["dBraw"] + Cout * r].view(1, Cout, r), "dm": acc32[o["dm"]:o["dm"] + Cout],
            }
        arr = (N.PackSeg * len(segs))(*segs)
        return {"ptrs": ptrs, "vers": None, "table": torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev),
                "num_segs": len(segs), "ctas": cta, "buf16": buf16, "buf32": buf32, "acc32": acc32, "entries": entries, "ads": ads,
                "finish": None, "pending": False}

    def lora_entry(self, ad):
        plan = self.lora_plan
        return None if plan is None else plan["entries"].get(id(ad))

    def lora_finish(self) -> None:
        """End of backward: gB += rowscale * dB_raw and g_mag += dm / mag for EVERY adapted layer in one launch."""
        plan = self.lora_plan
        if plan is None or not plan["pending"]:
            return
        plan["pending"] = False
        fin = plan["finish"]
        key = self.arena.data_ptr() if self.arena is not None else 0
        if fin is None or fin["key"] != key:
            segs, cta = [], 0
            for ad in plan["ads"]:
                e = plan["entries"][id(ad)]
                Bm, mag = ad.lora_B["default"].weight, ad.magnitude()
                if not Bm.requires_grad:
                    continue
                Cout, r = Bm.shape[0], ad.r
                has_mag = mag is not None and mag.requires_grad
                segs.append(N.LoraFinishSeg(e["dBraw"].data_ptr(), e["rowscale"].data_ptr(), self.arena_views[id(Bm)].data_ptr(),
                                            e["dm"].data_ptr() if has_mag else 0, mag.data_ptr() if has_mag else 0,
                                            self.arena_views[id(mag)].data_ptr() if has_mag else 0, Cout, r, cta, 0))
                cta += (Cout * r + 255) // 256
            arr = (N.LoraFinishSeg * len(segs))(*segs)
            fin = plan["finish"] = {"key": key, "num": len(segs), "ctas": cta,
                                    "table": torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.arena.device)}
        for ad in plan["ads"]:
            Bm, mag = ad.lora_B["default"].weight, ad.magnitude()
            if Bm.requires_grad:
                self.touch(Bm)
            if mag is not None and mag.requires_grad:
                self.touch(mag)
        if fin["num"]:
            _off_path(self, lambda: N.call("of_lora_finish_all", fin["table"].data_ptr(), fin["num"], fin["ctas"]))

    def operand_targets(self, unet):
        """{id(param): (device pointer of its bf16 operand copy, element pitch check)} for every weight of the grouped pack plan — what
        the fused optimizer writes while updating the master weights.  Conv operands are only offered when unpadded (cin_pad == Cin)."""
        if self.pack_plan is None:
            self.refresh_operands(unet)
        plan = self.pack_plan
        out = {}
        if plan is None:
            return out
        for (w, Cout, Cin, k, cp, off) in plan["segs_host"]:
            if cp == Cin:
                out[id(w)] = plan["buf"].data_ptr() + 2 * off
        return out

    def mark_operands_current(self) -> None:
        """Called by the fused optimizer after a step in which it refreshed EVERY operand of the pack plan."""
        plan = self.pack_plan
        if plan is not None:
            plan["vers"] = tuple(p._version for _, ps, _ in plan["views"] for p in ps) + (self.param_epoch,)

    # ---- grouped FiLM heads
    def film_plan(self, unet, B: int, with_grads: bool):
        """Descriptor tables of every `ResidualBlock.mlp[1]` head for batch size B (device-resident, built once)."""
        key = (B, with_grads)
        plan = self.film_plans.get(key)
        heads = self._film_heads
        if heads is None:        # the module tree is static (adapter injection never touches the FiLM heads)
            heads = self._film_heads = [m for m in unet.modules() if type(m).__name__ == "ResidualBlock" and m.mlp is not None]
        if not heads:
            return None
        ptrs = tuple(h.mlp[1].weight.data_ptr() for h in heads) + (self.arena.data_ptr() if (with_grads and self.arena is not None) else 0,)
        if plan is not None and plan["ptrs"] == ptrs:
            return plan
        dev = heads[0].mlp[1].weight.device
        K = heads[0].mlp[1].weight.shape[1]
        groups, chunks, slices, chunk_range = [], [], {}, {}
        row = 0
        ch = N.lib().of_film_chunk_rows()
        for gi, h in enumerate(heads):
            c0 = len(chunks) // 2
            lin = h.mlp[1]
            Nn = lin.weight.shape[0]
            assert lin.weight.shape[1] == K and Nn % 4 == 0
            dW = db = 0
            if with_grads:
                if lin.weight.requires_grad:
                    dW = self.arena_views[id(lin.weight)].data_ptr()
                if lin.bias is not None and lin.bias.requires_grad:
                    db = self.arena_views[id(lin.bias)].data_ptr()
            groups.append(N.FilmGroup(lin.weight.data_ptr(), lin.bias.data_ptr() if lin.bias is not None else 0, dW, db,
                                      B * row, Nn, row))
            for n0 in range(0, Nn, ch):
                chunks += [gi, n0]
            slices[id(h)] = (B * row, Nn)
            chunk_range[id(h)] = (c0, len(chunks) // 2)
            row += Nn
        arr = (N.FilmGroup * len(groups))(*groups)
        plan = {
            "ptrs": ptrs, "heads": heads, "K": K, "rows": row, "num_groups": len(groups), "slices": slices, "chunk_range": chunk_range,
            "groups": torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev),
            "chunks": torch.tensor(chunks, dtype=torch.int32).to(dev), "num_chunks": len(chunks) // 2,
        }
        self.film_plans[key] = plan
        return plan

    def _cached(self, key, params, build):
        ver = tuple((p.data_ptr(), p._version) for p in params) + (self.param_epoch,)
        hit = self.cache.get(key)
        if hit is not None and hit[0] == ver and (not self.refresh or hit[2] == self.epoch):
            return hit[1]
        val = build()
        self.cache[key] = (ver, val, self.epoch)
        return val

    # ---- operand builders
    def linear_w(self, *ws: torch.nn.Parameter) -> torch.Tensor:
        """[sum N_i][K] bf16 = row-concatenation of Linear / 1x1-conv weights."""
        def build():
            K = ws[0].shape[1]
            out = empty((sum(w.shape[0] for w in ws), K), BF16, ws[0].device)
            r = 0
            for w in ws:
                N.call("of_cast_f32_bf16", _p(w), out[r:].data_ptr(), w.numel())
                r += w.shape[0]
            return out
        return self._cached(("lin",) + tuple(id(w) for w in ws), ws, build)

    def conv_w(self, w: torch.nn.Parameter) -> torch.Tensor:
        """(Cout, Cin, k) fp32 -> [k][Cout][Cin_pad] bf16."""
        def build():
            Cout, Cin, k = w.shape
            cp = (Cin + 7) // 8 * 8
            out = empty((k, Cout, cp), BF16, w.device)
            N.call("of_pack_conv_weight", _p(w), Cout, Cin, k, _p(out), cp, 0, k)
            return out
        return self._cached(("conv", id(w)), (w,), build)

    def down_w(self, w: torch.nn.Parameter) -> torch.Tensor:
        """stride-2 conv as a 2-tap conv over the (B, L/2, 2C) view: [2][Cout][2Cin] = [W0|W1], [W2|0]."""
        def build():
            Cout, Cin, _ = w.shape
            out = zeros((2, Cout, 2 * Cin), F32, w.device)
            out[0, :, :Cin] = w[:, :, 0]
            out[0, :, Cin:] = w[:, :, 1]
            out[1, :, :Cin] = w[:, :, 2]
            return out.to(BF16)
        return self._cached(("down", id(w)), (w,), build)

    def cross_w(self, convs) -> torch.Tensor:
        """CrossEmbedLayer: all branches zero-padded to the widest kernel -> [kmax][dim_out][Cin_pad]."""
        ws = tuple(c.weight for c in convs)

        def build():
            kmax = max(w.shape[2] for w in ws)
            Cin = ws[0].shape[1]
            cp = (Cin + 7) // 8 * 8
            out = zeros((kmax, sum(w.shape[0] for w in ws), cp), F32, ws[0].device)
            r = 0
            for w in ws:
                k = w.shape[2]
                o = kmax // 2 - k // 2
                out[o:o + k, r:r + w.shape[0], :Cin] = w.detach().permute(2, 0, 1)
                r += w.shape[0]
            return out.to(BF16)
        return self._cached(("cross",) + tuple(id(w) for w in ws), ws, build)

    def padded_rows_w(self, w: torch.nn.Parameter, rows: int) -> torch.Tensor:
        """(N, K[,1]) -> [rows][K] bf16 with zero rows appended (final_conv: N=6 -> 8)."""
        def build():
            out = zeros((rows, w.shape[1]), BF16, w.device)
            out[:w.shape[0]] = w.detach().reshape(w.shape[0], w.shape[1]).to(BF16)
            return out
        return self._cached(("padrows", id(w), rows), (w,), build)

    # ---- LoRA / DoRA: effective weights W_eff = s * (W + scaling * B A) merged straight into the GEMM operand layout
    def _dora_merge_into(self, ad, out_rows: torch.Tensor, cin_pad: int, tap_stride: int):
        base = ad.base_layer
        W = base.weight
        Cout, Cin = W.shape[0], W.shape[1]
        k = W.shape[2] if W.dim() == 3 else 1
        A, Bm, mag = ad.lora_A["default"].weight, ad.lora_B["default"].weight, ad.magnitude()
        E = Cin * k
        ent = self.lora_entry(ad)
        if ent is not None and cin_pad == Cin:
            # grouped path: A / scaling*B operands come from the step's one grouped launch; the norm / scale / pack kernel also leaves
            # rowscale = scaling * s and (rowscale (.) B)^T for the rank-r backward
            V = empty((1, Cout, E), F32, W.device)
            R.gemm_fwd(ent["B16s"].view(1, Cout, ad.r), ent["A16"].view(1, ad.r, E), N_out=E, K=ad.r, b_mn_major=True,
                       aux_f32=W.detach().view(1, Cout, E), out_f32=V)
            N.call("of_dora_scale_pack_prep", _p(V), _p(mag), Cout, Cin, k, ent["n2"].data_ptr(), out_rows.data_ptr(), cin_pad, tap_stride,
                   _p(Bm), float(ad.scaling), ad.r, ent["Bst"].data_ptr(), ent["rowscale"].data_ptr())
            return ent["n2"]
        n2 = empty((Cout,), F32, W.device)
        if LORA_MERGE_TC and ad.r % 8 == 0 and E % 8 == 0 and cin_pad == Cin and E <= 8192 and W.is_contiguous():
            # V = W + scaling * B A on the tensor cores (K = r GEMM with the fp32 weight as residual), then norm + scale + pack per row
            r = ad.r
            A16 = empty((r, E), BF16, W.device)
            N.call("of_cast_f32_bf16", _p(A), _p(A16), r * E)
            B16 = empty((Cout, r), BF16, W.device)
            N.call("of_scale_cast_f32_bf16", _p(Bm), float(ad.scaling), _p(B16), Cout * r)
            V = empty((1, Cout, E), F32, W.device)
            R.gemm_fwd(B16.view(1, Cout, r), A16.view(1, r, E), N_out=E, K=r, b_mn_major=True, aux_f32=W.detach().view(1, Cout, E),
                       out_f32=V)
            N.call("of_dora_scale_pack", _p(V), _p(mag), Cout, Cin, k, _p(n2), out_rows.data_ptr(), cin_pad, tap_stride)
            return n2
        N.call("of_dora_merge", _p(W), _p(A), _p(Bm), _p(mag), float(ad.scaling), Cout, Cin, k, ad.r, _p(n2), out_rows.data_ptr(),
               cin_pad, tap_stride, None)
        return n2

    def conv_w_mod(self, conv) -> torch.Tensor:
        """GEMM operand of a (possibly adapted) Conv1d: [k][Cout][Cin_pad] bf16."""
        ad = _adapter(conv)
        if ad is None:
            return self.conv_w(conv.weight)
        W = ad.base_layer.weight
        ps = (W, ad.lora_A["default"].weight, ad.lora_B["default"].weight) + ((ad.magnitude(),) if ad.use_dora else ())

        def build():
            Cout, Cin, k = W.shape
            cp = (Cin + 7) // 8 * 8
            out = zeros((k, Cout, cp), BF16, W.device) if cp != Cin else empty((k, Cout, cp), BF16, W.device)
            n2 = self._dora_merge_into(ad, out, cp, Cout * cp)
            return out, n2
        key = ("dconv", id(W))
        return self._merged(key, lambda: self.conv_w_mod(conv), self._cached(key, ps, build)[0])

    def dora_n2(self, mod):
        W = mod.base_layer.weight
        hit = self.cache.get(("dconv", id(W))) or self.cache.get(("dlin", id(W)))
        return hit[1][1]

    def linear_w_mods(self, *mods) -> torch.Tensor:
        """[sum N_i][K] bf16 row-concatenation of (possibly adapted) Linear weights."""
        if all(_adapter(m) is None for m in mods):
            return self.linear_w(*[m.weight for m in mods])
        ps = []
        for m in mods:
            ad = _adapter(m)
            ps += [_base(m).weight] + ([ad.lora_A["default"].weight, ad.lora_B["default"].weight] if ad is not None else [])
            if ad is not None and ad.use_dora:
                ps.append(ad.magnitude())

        def build():
            K = _base(mods[0]).weight.shape[1]
            out = empty((sum(_base(m).weight.shape[0] for m in mods), K), BF16, _base(mods[0]).weight.device)
            r = 0
            for m in mods:
                w = _base(m).weight
                ad = _adapter(m)
                if ad is None:
                    N.call("of_cast_f32_bf16", _p(w), out[r:].data_ptr(), w.numel())
                else:
                    n2 = self._dora_merge_into(ad, out[r:], K, w.shape[0] * K)
                    self.cache[("dlin", id(w))] = (None, (None, n2), self.epoch)
                r += w.shape[0]
            return out
        key = ("linm",) + tuple(id(_base(m).weight) for m in mods)
        return self._merged(key, lambda: self.linear_w_mods(*mods), self._cached(key, tuple(ps), build))

    # ---- gradient buffers: zero-initialised fp32 views of the arena in the reference's parameter layout
    def touch(self, p: torch.nn.Parameter) -> bool:
        """Mark p's gradient as produced; returns True the first time (the buffer still holds zeros)."""
        first = id(p) not in self.touched
        self.touched.add(id(p))
        if self.on_touch is not None:
            self.on_touch(id(p))
        return first

    def grad_opt(self, p):
        """Gradient buffer of p, or None when p is absent / frozen (kernels skip NULL outputs)."""
        if p is None or not p.requires_grad:
            return None
        return self.grad(p)

    def grad(self, p: torch.nn.Parameter) -> torch.Tensor:
        if not p.requires_grad:      # frozen (LoRA fine-tuning): kernels with mandatory outputs write into throw-away scratch
            return zeros(p.shape, F32, p.device)
        self.touch(p)
        return self.arena_views[id(p)]

    def set_grad(self, p: torch.nn.Parameter, g: torch.Tensor) -> None:
        if not p.requires_grad:
            return
        first = self.touch(p)
        v = self.arena_views[id(p)]
        if first:
            v.copy_(g.view(v.shape))
        else:
            v.add_(g.view(v.shape))

    def take_grads(self, params):
        """Fresh view objects (so autograd's AccumulateGrad adopts them without a copy) of every gradient produced."""
        out = []
        for p in params:
            if p.requires_grad and id(p) in self.touched:
                v = self.arena_views[id(p)]
                if id(p) in self.packed:
                    # permuted view of the GEMM-layout slice: autograd's AccumulateGrad would copy it into the parameter's own
                    # layout (its "gradient layout contract"), so `.grad` is assigned here and autograd is told nothing
                    if p.grad is None:
                        p.grad = v.view(v.shape)
                    elif p.grad.data_ptr() != v.data_ptr():
                        p.grad.add_(v)
                    out.append(None)
                else:
                    out.append(v.view(v.shape))
            else:
                out.append(None)
        return out


# ------------------------------------------------------------------------------------------------ GEMM-shaped layers
def _dora_backward(store: ParamStore, ad, dW_packed: torch.Tensor, cin_pad: int, tap_stride: int) -> None:
    base = ad.base_layer
    W = base.weight
    Cout, Cin = W.shape[0], W.shape[1]
    k = W.shape[2] if W.dim() == 3 else 1
    A, Bm, mag = ad.lora_A["default"].weight, ad.lora_B["default"].weight, ad.magnitude()
    n2 = store.dora_n2(ad) if mag is not None else None
    N.call("of_dora_grad", _p(W), _p(A), _p(Bm), _p(mag), float(ad.scaling), Cout, Cin, k, ad.r, _p(n2), _p(dW_packed), cin_pad,
           tap_stride, store.grad(A).data_ptr(), store.grad(Bm).data_ptr(), _p(store.grad(mag)) if mag is not None else None)


def _coldot(dy16: torch.Tensor, y16: torch.Tensor, bias, out: torch.Tensor) -> None:
    """out[n] += sum over all rows of dy[., n] * (y[., n] - bias[n]) on (B, L, N) bf16 views."""
    B, L, Nn = dy16.shape
    d_bs, d_ld = _bl(dy16)
    y_bs, y_ld = _bl(y16)
    if B > 1 and (d_bs != L * d_ld or y_bs != L * y_ld):
        for b in range(B):
            N.call("of_coldot_bf16", _p(dy16[b]), d_ld, _p(y16[b]), y_ld, L, Nn, _p(bias), _p(out))
    else:
        N.call("of_coldot_bf16", _p(dy16), d_ld, _p(y16), y_ld, B * L, Nn, _p(bias), _p(out))


def _adapter_backward_rank_r(store: ParamStore, ad, dy16, x16, y16, taps, shift0):
    """Adapter gradients WITHOUT the full weight gradient: every term goes through the rank-r activations, on the tensor cores
    (the frozen base weight needs no gradient; the reference gets the same structure from autograd, lora_layers.py:80-90):
        u  = conv(x, A)                                  (B, L, r)        -- what `lora_A(x)` is in the reference
        dB = scaling * s (.) (dy^T u)                    (Cout, r)
        e  = dy @ (scaling * s (.) B)                    (B, L, r)
        dA = e^T * x   (a rank-r conv weight gradient)   (r, Cin, k)
        d mag = sum_l dy (y - b) / mag                   (Cout)           -- y = s (.) conv(x, V) + b is the saved layer output
    s = mag / ||W + scaling B A|| (detached; 1 for plain LoRA)."""
    base = ad.base_layer
    W = base.weight
    Cout, Cin = W.shape[0], W.shape[1]
    k = W.shape[2] if W.dim() == 3 else 1
    r, sc, dev = ad.r, float(ad.scaling), dy16.device
    A, Bm, mag = ad.lora_A["default"].weight, ad.lora_B["default"].weight, ad.magnitude()
    cp = (Cin + 7) // 8 * 8
    Bsz, L = x16.shape[0], x16.shape[1]
    ent = store.lora_entry(ad)
    if ent is not None:
        # grouped path: operands were prepared by the forward merge, accumulators are persistent (zeroed once per backward), lora_A's
        # gradient accumulates in place, the finishing arithmetic runs once for all layers (ParamStore.lora_finish); nothing here is
        # on the critical path of backward, so the four GEMMs go to the second stream
        u = empty((Bsz, L, r), BF16, dev)
        e = empty((Bsz, L, r), BF16, dev)
        store.touch(A)
        gA = store.arena_packed_views[id(A)] if k > 1 else store.arena_views[id(A)].view(1, r, Cin)
        step = 1 if k > 1 else 0

        def rank_r():
            R.gemm_fwd(x16, ent["Apk"], N_out=r, K=Cin, taps=k, shift0=shift0, shift_step=step, out_bf16=u)
            R.gemm_wgrad(dy16, u, ent["dBraw"], M=Cout, N_out=r)
            R.gemm_fwd(dy16, ent["Bst"].view(1, r, Cout), N_out=r, K=Cout, out_bf16=e)
            R.gemm_wgrad(e, x16, gA, M=r, N_out=Cin, taps=taps, shift0=shift0, shift_step=step)
            if mag is not None:
                _coldot(dy16, y16, base.bias, ent["dm"])
        _off_path(store, rank_r, dy16, x16, y16, u, e)
        store.lora_plan["pending"] = True
        return
    n2 = store.dora_n2(ad) if mag is not None else None
    # scaled B^T operand and the per-row factor scaling * s in one tiny launch
    Bst = empty((r, Cout), BF16, dev)
    rowscale = empty((Cout,), F32, dev)
    N.call("of_dora_rankr_prep", _p(Bm), _p(mag), _p(n2), sc, Cout, r, _p(Bst), _p(rowscale))
    # u = conv(x, A): A packed like any conv weight
    Apk = (zeros if cp != Cin else empty)((k, r, cp), BF16, dev)
    N.call("of_pack_conv_weight", _p(A), r, Cin, k, _p(Apk), cp, 0, k)
    u = empty((Bsz, L, r), BF16, dev)
    R.gemm_fwd(x16, Apk, N_out=r, K=cp, taps=k, shift0=shift0, shift_step=1 if k > 1 else 0, out_bf16=u)
    # dB_raw = dy^T u
    dBraw = zeros((1, Cout, r), F32, dev)
    R.gemm_wgrad(dy16, u, dBraw, M=Cout, N_out=r)
    # e = dy @ (scaling * s (.) B): B operand [N = r][K = Cout], K-major
    e = empty((Bsz, L, r), BF16, dev)
    R.gemm_fwd(dy16, Bst.view(1, r, Cout), N_out=r, K=Cout, out_bf16=e)
    # dA = e^T * x
    tmp = zeros((k, r, cp), F32, dev)
    R.gemm_wgrad(e, x16, tmp, M=r, N_out=cp, taps=taps, shift0=shift0, shift_step=1 if k > 1 else 0)
    acc = 0 if store.touch(A) else 1
    N.call("of_unpack_conv_wgrad", _p(tmp), r, Cin, k, cp, 0, _p(store.grad(A)), acc, 0)
    # d magnitude from the activations, then dB / d mag accumulated into the arena
    dm = None
    if mag is not None:
        dm = zeros((Cout,), F32, dev)
        _coldot(dy16, y16, base.bias, dm)
    N.call("of_dora_rankr_finish", _p(dBraw), _p(rowscale), _p(store.grad(Bm)), _p(dm), _p(mag),
           _p(store.grad(mag)) if mag is not None else None, Cout, r)


def _wgrad_conv_mod(store: ParamStore, conv, dy16, x16, taps, shift0, y16=None):
    ad = _adapter(conv)
    if ad is None:
        return _wgrad_conv(store, conv.weight, dy16, x16, taps, shift0)
    if LORA_RANK_R and (y16 is not None or not ad.use_dora) and ad.r % 8 == 0:
        return _adapter_backward_rank_r(store, ad, dy16, x16, y16, taps, shift0)
    Cout, Cin, k = ad.base_layer.weight.shape
    cp = (Cin + 7) // 8 * 8
    tmp = zeros((k, Cout, cp), F32, dy16.device)
    R.gemm_wgrad(dy16, x16, tmp, M=Cout, N_out=cp, taps=taps, shift0=shift0, shift_step=1)
    _dora_backward(store, ad, tmp, cp, Cout * cp)


def _wgrad_linear_mod(store: ParamStore, lin, dy16, x16, y16=None):
    ad = _adapter(lin)
    if ad is None:
        return _wgrad_linear(store, lin.weight, dy16, x16)
    if LORA_RANK_R and (y16 is not None or not ad.use_dora) and ad.r % 8 == 0:
        return _adapter_backward_rank_r(store, ad, dy16, x16, y16, 1, 0)
    Nn, K = ad.base_layer.weight.shape
    tmp = zeros((1, Nn, K), F32, dy16.device)
    R.gemm_wgrad(dy16, x16, tmp, M=Nn, N_out=K)
    _dora_backward(store, ad, tmp, K, Nn * K)


def _wgrad_conv(store: ParamStore, w: torch.nn.Parameter, dy16, x16, taps, shift0):
    """Conv1d weight gradient: packed [k][Cout][Cin_pad] fp32 accumulation, then back to the (Cout, Cin, k) layout."""
    if not w.requires_grad:
        return
    Cout, Cin, k = w.shape
    cp = (Cin + 7) // 8 * 8
    pv = store.arena_packed_views.get(id(w))
    if pv is not None:       # the arena slice IS the packed accumulator (zeroed at the start of backward)
        store.touch(w)
        _off_path(store, lambda: R.gemm_wgrad(dy16, x16, pv, M=Cout, N_out=Cin, taps=taps, shift0=shift0, shift_step=1), dy16, x16)
        return
    # a fresh allocation (not a persistent scratch): the caching allocator hands back the block the previous layer just released,
    # so fill -> split-K atomics -> unpack stay L2-resident (measured: a persistent 3 GB scratch made the unpack 6x slower)
    tmp = zeros((k, Cout, cp), F32, w.device)
    R.gemm_wgrad(dy16, x16, tmp, M=Cout, N_out=cp, taps=taps, shift0=shift0, shift_step=1)
    acc = 0 if store.touch(w) else 1
    g = store.grad(w)
    N.call("of_unpack_conv_wgrad", _p(tmp), Cout, Cin, k, cp, 0, _p(g), acc, 0)


def _wgrad_linear(store: ParamStore, w: torch.nn.Parameter, dy16, x16, row0: int = 0, rows: Optional[int] = None):
    """Linear / 1x1-conv weight gradient accumulated atomically straight into the (N, K) fp32 grad buffer."""
    if not w.requires_grad:
        return
    g = store.grad(w)
    Nn = w.shape[0] if rows is None else rows
    K = w.shape[1]
    _off_path(store, lambda: R.gemm_wgrad(dy16, x16, g.view(1, w.shape[0], K)[:, row0:row0 + Nn], M=Nn, N_out=K), dy16, x16)


def _bias_grad(store: ParamStore, b: Optional[torch.nn.Parameter], dy16) -> None:
    if b is not None and b.requires_grad:
        g = store.grad(b)
        _off_path(store, lambda: colsum(dy16, g), dy16)


def _off_path(store: ParamStore, fn: Callable[[], None], *keep) -> None:
    """Run a weight- / bias-gradient launch off the critical path (SideLane) when the backward pass has a second stream."""
    if _DEBUG_SKIP_OFFPATH:      # measurement only (wrong gradients): how long is the critical path alone?
        global _debug_warned
        if not _debug_warned:
            import sys
            print("[osufusion_b200] OF_DEBUG_SKIP_OFFPATH=1: weight / bias gradients are NOT computed (timing aid only)", file=sys.stderr)
            _debug_warned = True
        return
    if store.side is not None:
        store.side.run(fn, *keep)
    else:
        fn()


_DEBUG_SKIP_OFFPATH = os.environ.get("OF_DEBUG_SKIP_OFFPATH", "0") == "1"
_debug_warned = False


def _dgrad_into(x: Act, dy16, wpack, *, N_out, K, taps=1, shift0=0, shift_step=0, b_ld=None, want_bf16=False,
                want_f32=True):
    """x.grad (+)= dY * W^T.  The GEMM epilogue adds the already-accumulated gradient (aux) and writes in place."""
    B, L = dy16.shape[0], dy16.shape[1]
    prev = x.grad
    x.grad16 = None
    out = None
    if want_f32:
        out = prev if prev is not None else empty((B, L, N_out), F32, dy16.device)
    o16 = empty((B, L, N_out), BF16, dy16.device) if want_bf16 else None
    R.gemm_fwd(dy16, wpack, N_out=N_out, K=K, taps=taps, shift0=shift0, shift_step=shift_step, b_mn_major=True,
               b_ld=b_ld, aux_f32=prev, out_f32=out, out_bf16=o16)
    x.grad = out
    return o16


class Ctx:
    """Per-forward context: device, tape (None = inference), parameter store, conditioning vector."""

    def __init__(self, device, store: ParamStore, tape: Optional[Tape]) -> None:
        self.device, self.store, self.tape = device, store, tape
        self.emb_act = None      # (B, 2*dim_emb) fp32 = SiLU(cat(t, c))  (shared input of every FiLM head)
        self.d_emb_act = None    # its gradient accumulator
        self.rope_cache = {}
        self.attn_variant = 0
        self.zpool = ZeroPool(device)
        use_pool(self.zpool)
        self.film = None         # id(ResidualBlock) -> (B, 2C) fp32 view of the grouped FiLM output
        self.film_dss = None     # ... and of its gradient accumulator
        self.film_bwd = None     # callable(heads): grouped FiLM weight-gradient launch for a unit's heads (modules.UNet.conditioning)


def conv3(ctx: Ctx, x16, conv, *, stats=None, out=None):
    """nn.Conv1d(k=3, padding=1) as implicit GEMM; returns bf16 (B, L, Cout)."""
    B, L, _ = x16.shape
    base = _base(conv)
    Cout, Cin, k = base.weight.shape
    w = ctx.store.conv_w_mod(conv)
    y = out if out is not None else empty((B, L, Cout), BF16, ctx.device)
    R.gemm_fwd(x16, w, N_out=Cout, K=w.shape[2], taps=k, shift0=-(k // 2), shift_step=1, bias=base.bias, out_bf16=y, stats=stats)
    return y


def conv3_bwd(ctx: Ctx, conv, x: Act, x16, dy16, need_dx=True, want_bf16=False, want_f32=True, y16=None):
    Cout, Cin, k = _base(conv).weight.shape
    _wgrad_conv_mod(ctx.store, conv, dy16, x16, taps=k, shift0=-(k // 2), y16=y16)
    if need_dx:
        w = ctx.store.conv_w_mod(conv)
        return _dgrad_into(x, dy16, w, N_out=Cin, K=Cout, taps=k, shift0=k // 2, shift_step=-1, b_ld=w.shape[2],
                           want_bf16=want_bf16, want_f32=want_f32)
    return None


# ------------------------------------------------------------------------------------------------ ResidualBlock
def _rb_args(B, L, Cc, y, stats, norm, ss):
    a = N.RbArgs()
    a.B, a.L, a.C, a.eps = B, L, Cc, norm.eps
    a.y = y.data_ptr()
    a.y_bs, a.y_ld = _bl(y)
    a.stats = stats.data_ptr()
    a.gamma, a.beta = norm.weight.data_ptr(), norm.bias.data_ptr()
    a.ss = _p(ss)
    return a


def residual_block(ctx: Ctx, m, x: Act) -> Act:
    """ResidualBlock.forward (reference residual.py:118-137)."""
    st, dev = ctx.store, ctx.device
    x16 = x.bf16
    B, L, Cin = x16.shape
    Cout = _base(m.block1.proj).weight.shape[0]
    ss = None
    if m.mlp is not None:
        ss = ctx.film[id(m)]     # computed for all heads at once by UNet.conditioning (of_film_fwd)
    stats1 = zeros((B, 2), F64, dev)
    y1 = conv3(ctx, x16, m.block1.proj, stats=stats1)
    a1 = _rb_args(B, L, Cout, y1, stats1, m.block1.norm, ss)
    h1 = empty((B, L, Cout), BF16, dev)
    a1.out_bf16 = h1.data_ptr()
    a1.out_bf16_bs, a1.out_bf16_ld = _bl(h1)
    N.call("of_rb_apply_fwd", C.byref(a1))
    stats2 = zeros((B, 2), F64, dev)
    y2 = conv3(ctx, h1, m.block2.proj, stats=stats2)
    # GlobalContext: logits -> softmax over L -> pooled -> 2-layer gate MLP
    se = m.se
    wk = se.to_k.weight.view(-1)
    a2 = _rb_args(B, L, Cout, y2, stats2, m.block2.norm, None)
    p = empty((B, L), F32, dev)
    a2.mode = 0
    a2.vec, a2.vec_bs, a2.vec_bias = wk.data_ptr(), 0, se.to_k.bias.data_ptr()
    a2.out_rows = p.data_ptr()
    if GCTX_FUSED:
        pooled = empty((B, Cout), F32, dev)
        nparts = N.lib().of_rb_pool_parts(C.byref(a2))
        part = empty((B, nparts, Cout + 2), F32, dev)
        N.call("of_rb_logit_pool", C.byref(a2), part.data_ptr(), pooled.data_ptr())   # logits -> softmax -> pooled in one pass over y2
        a2.p = p.data_ptr()
    else:
        N.call("of_rb_rowdot", C.byref(a2))
        N.call("of_softmax_rows", _p(p), B, L)
        pooled = zeros((B, Cout), F32, dev)
        a2.p, a2.acc_bc = p.data_ptr(), pooled.data_ptr()
        N.call("of_rb_pool", C.byref(a2))
    Wa, Wb = se.layers[0].weight, se.layers[2].weight
    g1, g1pre = linear_small_fwd(pooled, Wa.view(Wa.shape[0], -1), se.layers[0].bias, act=1, want_pre=True)
    gate, gatepre = linear_small_fwd(g1, Wb.view(Wb.shape[0], -1), se.layers[2].bias, act=2, want_pre=True)
    # residual branch
    has_res = not isinstance(m.res_conv, torch.nn.Identity)
    out32 = empty((B, L, Cout), F32, dev)
    out16 = empty((B, L, Cout), BF16, dev)
    a2.gate = gate.data_ptr()
    if has_res:
        wres = st.linear_w(m.res_conv.weight)
        r16 = empty((B, L, Cout), BF16, dev)
        R.gemm_fwd(x16, wres, N_out=Cout, K=Cin, bias=m.res_conv.bias, out_bf16=r16)
        a2.res_bf16 = r16.data_ptr()
        a2.res_bf16_bs, a2.res_bf16_ld = _bl(r16)
    elif x.f32 is not None:
        a2.res_f32 = x.f32.data_ptr()
        a2.res_f32_bs, a2.res_f32_ld = _bl(x.f32)
    else:
        a2.res_bf16 = x16.data_ptr()
        a2.res_bf16_bs, a2.res_bf16_ld = _bl(x16)
    a2.out_f32 = out32.data_ptr()
    a2.out_f32_bs, a2.out_f32_ld = _bl(out32)
    a2.out_bf16 = out16.data_ptr()
    a2.out_bf16_bs, a2.out_bf16_ld = _bl(out16)
    N.call("of_rb_gate_fwd", C.byref(a2))
    out = Act(out32, out16)

    if ctx.tape is not None:
        def backward():
            dout = out.grad
            out.grad = None
            d_bs, d_ld = _bl(dout)
            b2 = _rb_args(B, L, Cout, y2, stats2, m.block2.norm, None)
            b2.dout_f32 = dout.data_ptr()
            b2.dout_f32_bs, b2.dout_f32_ld = d_bs, d_ld
            dgate = zeros((B, Cout), F32, dev)
            b2.acc_bc = dgate.data_ptr()
            N.call("of_rb_gate_bwd_reduce", C.byref(b2))
            # gate MLP backward
            dg1 = zeros((B, Wa.shape[0]), F32, dev)
            linear_small_bwd_param(st, dgate, gatepre, 2, g1, Wb, se.layers[2].bias, dg1)
            dpooled = zeros((B, Cout), F32, dev)
            linear_small_bwd_param(st, dg1, g1pre, 1, pooled, Wa, se.layers[0].bias, dpooled)
            # d logits
            da = empty((B, L), F32, dev)
            b2.mode = 1
            b2.vec, b2.vec_bs = dpooled.data_ptr(), Cout
            b2.p, b2.pooled, b2.out_rows = p.data_ptr(), pooled.data_ptr(), da.data_ptr()
            N.call("of_rb_rowdot", C.byref(b2))
            N.call("of_softmax_bwd_rows", _p(p), _p(da), B, L)
            # GroupNorm-2 backward, pass 1
            dxh = empty((B, L, Cout), BF16, dev)
            dstats = zeros((B, 2), F64, dev)
            b2.mode = 0
            b2.gate, b2.dpooled, b2.da, b2.wk = gate.data_ptr(), dpooled.data_ptr(), da.data_ptr(), wk.data_ptr()
            b2.dstats = dstats.data_ptr()
            b2.dgamma, b2.dbeta = st.grad(m.block2.norm.weight).data_ptr(), st.grad(m.block2.norm.bias).data_ptr()
            b2.dwk, b2.dbk = st.grad(se.to_k.weight).data_ptr(), st.grad(se.to_k.bias).data_ptr()
            b2.dxhat_bf16 = dxh.data_ptr()
            b2.dxhat_bs, b2.dxhat_ld = _bl(dxh)
            dout16 = None
            if has_res:
                dout16 = empty((B, L, Cout), BF16, dev)
                b2.dout_bf16 = dout16.data_ptr()
                b2.dout_bf16_bs, b2.dout_bf16_ld = _bl(dout16)
            N.call("of_rb_bwd_pass1", C.byref(b2))
            dy2 = empty((B, L, Cout), BF16, dev)
            b2.dy_bf16 = dy2.data_ptr()
            b2.dy_bs, b2.dy_ld = _bl(dy2)
            b2.dbias = _p(st.grad_opt(_base(m.block2.proj).bias))
            N.call("of_rb_bwd_apply", C.byref(b2))
            # conv2 backward
            h1a = Act(None, h1)
            dh1 = conv3_bwd(ctx, m.block2.proj, h1a, h1, dy2, want_bf16=True, want_f32=False, y16=y2)
            # GroupNorm-1 (+FiLM) backward
            b1 = _rb_args(B, L, Cout, y1, stats1, m.block1.norm, ss)
            b1.mode = 1
            b1.dh_bf16 = dh1.data_ptr()
            b1.dh_bs, b1.dh_ld = _bl(dh1)
            dxh1 = dxh  # reuse
            dstats1 = zeros((B, 2), F64, dev)
            b1.dstats = dstats1.data_ptr()
            b1.dgamma, b1.dbeta = st.grad(m.block1.norm.weight).data_ptr(), st.grad(m.block1.norm.bias).data_ptr()
            if ss is not None:
                b1.dss = ctx.film_dss[id(m)].data_ptr()   # consumed by of_film_bwd in UNet.conditioning's backward
            b1.dxhat_bf16 = dxh1.data_ptr()
            b1.dxhat_bs, b1.dxhat_ld = _bl(dxh1)
            N.call("of_rb_bwd_pass1", C.byref(b1))
            dy1 = dy2 if st.side is None else empty((B, L, Cout), BF16, dev)   # conv2's weight gradient may still be reading dy2
            b1.dy_bf16 = dy1.data_ptr()
            b1.dy_bs, b1.dy_ld = _bl(dy1)
            b1.dbias = _p(st.grad_opt(_base(m.block1.proj).bias))
            N.call("of_rb_bwd_apply", C.byref(b1))
            # input gradient: identity residual (adopt dout) or res_conv dgrad, then conv1 dgrad accumulated in place
            if has_res:
                _wgrad_linear(st, m.res_conv.weight, dout16, x16)
                _bias_grad(st, m.res_conv.bias, dout16)
                _dgrad_into(x, dout16, st.linear_w(m.res_conv.weight), N_out=Cin, K=Cout)
            else:
                x.add_grad(dout)
            g16 = conv3_bwd(ctx, m.block1.proj, x, x16, dy1, y16=y1, want_bf16=x.want16)
            x.grad16 = g16          # conv1's dgrad is the last contribution to x's gradient (every other consumer of x comes later in
            #                         forward order, i.e. earlier in backward)
        ctx.tape.push(backward)
    return out


# ------------------------------------------------------------------------------------------------ TransformerBlock
def rope_tables(ctx: Ctx, L: int, D: int, scale_base: int, f32: bool = False):
    """cos/sin (L, D) generated exactly like attention.py:33-49: in q's dtype — bf16 under autocast, fp32 when a DoRA-adapted
    to_q has promoted q to fp32 (peft: `(mag/norm - 1) * F.linear(x, W)` is fp32 x bf16 -> fp32)."""
    dt = F32 if f32 else BF16
    key = (L, D, scale_base, f32)
    hit = ctx.rope_cache.get(key)
    if hit is None:
        inv_freq = 1.0 / (10000 ** (torch.arange(0, D, 2, device=ctx.device).float() / D))
        t = torch.arange(L, dtype=dt, device=ctx.device)
        t *= scale_base / L
        freqs = torch.einsum("i,j->ij", t, inv_freq.to(dt))
        emb = torch.cat([freqs, freqs], dim=-1)
        hit = (emb.cos().contiguous(), emb.sin().contiguous())
        ctx.rope_cache[key] = hit
    return hit


def transformer_block(ctx: Ctx, m, x: Act) -> Act:
    """TransformerBlock.forward (reference unet.py:179-183) = Attention.forward_body (unet.py:125-141) + FeedForward."""
    st, dev = ctx.store, ctx.device
    at = m.attn
    B, L, Cc = x.f32.shape
    H, KVH, D = at.heads, at.kv_heads, at.dim_head
    HD, KD = H * D, KVH * D
    rows = B * L
    x32 = x.f32
    assert x32.stride(0) == L * x32.stride(1)
    xn32 = empty((B, L, Cc), F32, dev)
    xn16 = empty((B, L, Cc), BF16, dev)
    mr = empty((rows, 2), F32, dev)
    N.call("of_layernorm_fwd", _p(x32), x32.stride(1), rows, Cc, _p(at.norm.weight), _p(at.norm.bias), at.norm.eps,
           _p(xn32), _p(xn16), Cc, _p(mr))
    wqkv = st.linear_w_mods(at.to_q, at.to_kv)
    qkv = empty((B, L, HD + 2 * KD), BF16, dev)
    R.gemm_fwd(xn16, wqkv, N_out=HD + 2 * KD, K=Cc, out_bf16=qkv)
    ad_q = _adapter(at.to_q)
    rope_f32 = ad_q is not None and ad_q.use_dora
    ad_kv = _adapter(at.to_kv)
    need_pre = ctx.tape is not None and LORA_RANK_R and any(a_ is not None and a_.use_dora for a_ in (ad_q, ad_kv))
    qkv_pre = qkv.clone() if need_pre else None    # the DoRA magnitude gradient needs the projection output before RoPE rotates it
    cosT, sinT = rope_tables(ctx, L, D, at.rotary_emb.scale_base, rope_f32)
    q_bs, q_ld = _bl(qkv)
    N.call("of_rope_fwd", _p(qkv), q_ld, q_bs, B, L, H, KVH, D, _p(cosT), _p(sinT), int(rope_f32))
    q, k, v = qkv[:, :, :HD], qkv[:, :, HD:HD + KD], qkv[:, :, HD + KD:]
    o16 = empty((B, L, HD), BF16, dev)
    lse = empty((B, H, L), F32, dev)
    R.attn_fwd(q, k, v, o16, lse, H=H, KVH=KVH, D=D, variant=ctx.attn_variant)
    wout = st.linear_w(at.to_out.weight)
    x2_32 = empty((B, L, Cc), F32, dev)
    x2_16 = empty((B, L, Cc), BF16, dev)
    R.gemm_fwd(o16, wout, N_out=Cc, K=HD, bias=at.to_out.bias, aux_f32=xn32, out_f32=x2_32, out_bf16=x2_16)
    ff1, ff2 = m.ff[0], m.ff[2]
    Ci = ff1.weight.shape[0]
    w1, w2 = st.linear_w(ff1.weight), st.linear_w(ff2.weight)
    u16 = empty((B, L, Ci), BF16, dev) if ctx.tape is not None else None
    s16 = empty((B, L, Ci), BF16, dev)
    R.gemm_fwd(x2_16, w1, N_out=Ci, K=Cc, bias=ff1.bias, act=R.ACT_SILU, pre_bf16=u16, out_bf16=s16)
    out32 = empty((B, L, Cc), F32, dev)
    out16 = empty((B, L, Cc), BF16, dev)
    R.gemm_fwd(s16, w2, N_out=Cc, K=Ci, bias=ff2.bias, aux_f32=x2_32, out_f32=out32, out_bf16=out16)
    out = Act(out32, out16)
    out.want16 = GRAD16

    if ctx.tape is not None:
        def backward():
            dout16, dout = out.take_grad16((B, L, Cc), dev)
            # FeedForward
            dU = empty((B, L, Ci), BF16, dev)
            R.gemm_fwd(dout16, w2, N_out=Ci, K=Cc, b_mn_major=True, aux_bf16=u16, aux_is_dsilu=True, out_bf16=dU)
            _wgrad_linear(st, ff2.weight, dout16, s16)
            _bias_grad(st, ff2.bias, dout16)
            x2 = Act(None, None)
            x2.grad = dout  # residual `ff(x) + x`
            dx2_16 = _dgrad_into(x2, dU, w1, N_out=Cc, K=Ci, want_bf16=True)
            _wgrad_linear(st, ff1.weight, dU, x2_16)
            _bias_grad(st, ff1.bias, dU)
            # to_out
            dO = empty((B, L, HD), BF16, dev)
            R.gemm_fwd(dx2_16, wout, N_out=HD, K=Cc, b_mn_major=True, out_bf16=dO)
            _wgrad_linear(st, at.to_out.weight, dx2_16, o16)
            _bias_grad(st, at.to_out.bias, dx2_16)
            # attention core
            delta = empty((B, H, L), F32, dev)
            dq = empty((B, L, HD), F32, dev)         # zero-filled by the delta pre-pass of of_attn_bwd
            dkv = empty((B, L, 2 * KD), F32, dev)
            R.attn_bwd(q, k, v, o16, lse, dO, delta, dq, dkv[:, :, :KD], dkv[:, :, KD:], H=H, KVH=KVH, D=D, zero_grads=True)
            dqkv = empty((B, L, HD + 2 * KD), BF16, dev)
            dq_bs, dq_ld = _bl(dq)
            dkv_bs, dkv_ld = _bl(dkv)
            o_bs, o_ld = _bl(dqkv)
            N.call("of_rope_bwd", _p(dq), dq_ld, dq_bs, dkv[:, :, :KD].data_ptr(), dkv[:, :, KD:].data_ptr(), dkv_ld, dkv_bs,
                   _p(dqkv), o_ld, o_bs, B, L, H, KVH, D, _p(cosT), _p(sinT), int(rope_f32))
            # q/kv projections: d(xn) = residual grad (x2.grad) + dqkv W
            _dgrad_into(x2, dqkv, wqkv, N_out=Cc, K=HD + 2 * KD)
            _wgrad_linear_mod(st, at.to_q, dqkv[:, :, :HD], xn16, y16=None if qkv_pre is None else qkv_pre[:, :, :HD])
            _wgrad_linear_mod(st, at.to_kv, dqkv[:, :, HD:], xn16, y16=None if qkv_pre is None else qkv_pre[:, :, HD:])
            # LayerNorm
            dxn = x2.grad
            assert dxn.stride(0) == L * dxn.stride(1)
            dx = empty((B, L, Cc), F32, dev)
            N.call("of_layernorm_bwd", _p(dxn), dxn.stride(1), _p(x32), x32.stride(1), rows, Cc, _p(at.norm.weight), _p(mr),
                   _p(dx), None, Cc, st.grad(at.norm.weight).data_ptr(), st.grad(at.norm.bias).data_ptr())
            x.add_grad(dx)
        ctx.tape.push(backward)
    return out


# ------------------------------------------------------------------------------------------------ samplers
def downsample(ctx: Ctx, m, x: Act) -> Act:
    """Downsample (unet.py:77-92): reflect-pad right by 1, conv k=3 stride 2 — as a 2-tap conv over the (B, L/2, 2C) view
    plus a one-row fix-up for the reflected sample."""
    st, dev = ctx.store, ctx.device
    conv = m.conv
    x16 = x.bf16
    B, L, Cin = x16.shape
    assert L % 2 == 0 and x16.is_contiguous()
    Cout = conv.weight.shape[0]
    Lh = L // 2
    w = st.down_w(conv.weight)                       # [2][Cout][2Cin]
    xv = x16.view(B, Lh, 2 * Cin)
    y = empty((B, Lh, Cout), BF16, dev)
    R.gemm_fwd(xv, w, N_out=Cout, K=2 * Cin, taps=2, shift0=0, shift_step=1, bias=conv.bias, out_bf16=y)
    # reflect: x_pad[L] = x[L-2]  ->  y[Lh-1] += W2 x[L-2]
    xrow = x16[:, L - 2:L - 1, :]
    ylast = y[:, Lh - 1:Lh, :]
    R.gemm_fwd(xrow, w[1], N_out=Cout, K=Cin, b_ld=2 * Cin, aux_bf16=ylast, out_bf16=ylast)
    out = Act(None, y)
    out.want16 = GRAD16

    if ctx.tape is not None:
        def backward():
            dy16, dy = out.take_grad16((B, Lh, Cout), dev)
            _bias_grad(st, conv.bias, dy16)
            if conv.weight.requires_grad:
                tmp = zeros((2, Cout, 2 * Cin), F32, dev)
                R.gemm_wgrad(dy16, xv, tmp, M=Cout, N_out=2 * Cin, taps=2, shift0=0, shift_step=1)
                R.gemm_wgrad(dy16[:, Lh - 1:Lh, :], xrow, tmp[1:2, :, :Cin], M=Cout, N_out=Cin)
                g = torch.stack([tmp[0, :, :Cin], tmp[0, :, Cin:], tmp[1, :, :Cin]], dim=2)
                st.set_grad(conv.weight, g)
            # dgrad over the (B, Lh, 2Cin) view, then the reflected row
            x.grad16 = None
            if x.grad is None:
                x.grad = empty((B, L, Cin), F32, dev)
                prev = None
            else:
                if not x.grad.is_contiguous():
                    t = empty((B, L, Cin), F32, dev)
                    cast_copy(x.grad, t)
                    x.grad = t
                prev = x.grad.view(B, Lh, 2 * Cin)
            gv = x.grad.view(B, Lh, 2 * Cin)
            R.gemm_fwd(dy16, w, N_out=2 * Cin, K=Cout, taps=2, shift0=0, shift_step=-1, b_mn_major=True, b_ld=2 * Cin,
                       aux_f32=prev, out_f32=gv)
            grow = x.grad[:, L - 2:L - 1, :]
            R.gemm_fwd(dy16[:, Lh - 1:Lh, :], w[1], N_out=Cin, K=Cout, b_mn_major=True, b_ld=2 * Cin, aux_f32=grow, out_f32=grow)
        ctx.tape.push(backward)
    return out


def upsample(ctx: Ctx, m, x: Act) -> Act:
    """Upsample (unet.py:61-74): nearest x2 then conv k=3."""
    st, dev = ctx.store, ctx.device
    conv = m.conv
    x16 = x.bf16
    B, L, Cin = x16.shape
    up = empty((B, 2 * L, Cin), BF16, dev)
    xb, xl = _bl(x16)
    N.call("of_upsample2x_fwd", _p(x16), xl, xb, B, L, Cin, _p(up), Cin, 2 * L * Cin)
    y = conv3(ctx, up, conv)
    out = Act(None, y)
    out.want16 = GRAD16

    if ctx.tape is not None:
        def backward():
            dy16, dy = out.take_grad16(y.shape, dev)
            _bias_grad(st, conv.bias, dy16)
            upa = Act(None, up)
            conv3_bwd(ctx, conv, upa, up, dy16)
            d = upa.grad
            dx = empty((B, L, Cin), F32, dev)
            N.call("of_upsample2x_bwd", _p(d), Cin, 2 * L * Cin, B, L, Cin, _p(dx), None, Cin, L * Cin)
            x.add_grad(dx)
        ctx.tape.push(backward)
    return out


def parallel_sampler(ctx: Ctx, m, x: Act) -> Act:
    """Parallel(conv3, conv1) (unet.py:95-101,225-236): bf16 sum of the two bf16 conv outputs."""
    st, dev = ctx.store, ctx.device
    c3, c1 = m.fns[0], m.fns[1]
    x16 = x.bf16
    B, L, Cin = x16.shape
    Cout = c3.weight.shape[0]
    y3 = conv3(ctx, x16, c3)
    w1 = st.linear_w(c1.weight)
    y = empty((B, L, Cout), BF16, dev)
    R.gemm_fwd(x16, w1, N_out=Cout, K=Cin, bias=c1.bias, aux_bf16=y3, out_bf16=y)
    out = Act(None, y)
    out.want16 = GRAD16

    if ctx.tape is not None:
        def backward():
            dy16, dy = out.take_grad16(y.shape, dev)
            _bias_grad(st, c3.bias, dy16)
            _bias_grad(st, c1.bias, dy16)
            _wgrad_linear(st, c1.weight, dy16, x16)
            _dgrad_into(x, dy16, w1, N_out=Cin, K=Cout)
            conv3_bwd(ctx, c3, x, x16, dy16)
        ctx.tape.push(backward)
    return out


def cross_embed(ctx: Ctx, m, x16: torch.Tensor) -> Act:
    """CrossEmbedLayer (unet.py:42-58): the three branches as one zero-padded k=15 implicit GEMM."""
    st, dev = ctx.store, ctx.device
    B, L, Cp = x16.shape
    w = st.cross_w(m.convs)
    kmax, Cout = w.shape[0], w.shape[1]
    bias = torch.cat([c.bias.detach() for c in m.convs])
    y = empty((B, L, Cout), BF16, dev)
    R.gemm_fwd(x16, w, N_out=Cout, K=Cp, taps=kmax, shift0=-(kmax // 2), shift_step=1, bias=bias, out_bf16=y)
    out = Act(None, y)
    out.want16 = GRAD16

    if ctx.tape is not None:
        def backward():
            if out.grad is None:
                return
            dy16, dy = out.take_grad16(y.shape, dev)
            tmp = zeros((kmax, Cout, Cp), F32, dev)
            R.gemm_wgrad(dy16, x16, tmp, M=Cout, N_out=Cp, taps=kmax, shift0=-(kmax // 2), shift_step=1)
            db = zeros((Cout,), F32, dev)
            colsum(dy16, db)
            r = 0
            for c in m.convs:
                co, ci, k = c.weight.shape
                o = kmax // 2 - k // 2
                if c.weight.requires_grad:
                    st.set_grad(c.weight, tmp[o:o + k, r:r + co, :ci].permute(1, 2, 0).contiguous())
                    st.set_grad(c.bias, db[r:r + co].clone())
                r += co
        ctx.tape.push(backward)
    return out


def concat(ctx: Ctx, a: Act, b: Act) -> Act:
    """torch.cat([a, b], dim=channels) on the bf16 operand copies (both consumers of a concat are GEMMs)."""
    dev = ctx.device
    B, L, Ca = a.bf16.shape
    Cb = b.bf16.shape[2]
    wide = empty((B, L, Ca + Cb), BF16, dev)
    cast_copy(a.bf16, wide[:, :, :Ca])
    cast_copy(b.bf16, wide[:, :, Ca:])
    out = Act(None, wide)

    if ctx.tape is not None:
        def backward():
            g = out.grad
            out.grad = None
            a.add_grad(g[:, :, :Ca])
            b.add_grad(g[:, :, Ca:])
        ctx.tape.push(backward)
    return out


def unet_block(ctx: Ctx, m, x: Act):
    """UNetBlock.forward_body (unet.py:240-252): returns (sampler(x), x)."""
    x = residual_block(ctx, m.init_resnet, x)
    for res, tr in zip(m.resnets, m.transformers):
        x = residual_block(ctx, res, x)
        x = transformer_block(ctx, tr, x)
    kind = m.sampler_kind
    if kind == "down":
        y = downsample(ctx, m.sampler, x)
    elif kind == "up":
        y = upsample(ctx, m.sampler, x)
    else:
        y = parallel_sampler(ctx, m.sampler, x)
    return y, x
