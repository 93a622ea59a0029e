#!/usr/bin/env python
"""bench.py — denoiser fwd+bwd samples/s on N B200 (BASELINE.json configs[1..2]): one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size L|S] [--batch B] [--frames N]
    torchrun --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU, NCCL)

A "step" = one training micro-step of the reference trainer (trainer.py:295-301): noise + timestep draw, add_noise,
denoiser forward (cond_drop_prob 0.5), masked-MSE loss, full backward to all 1239 parameter gradients (+ gradient
all-reduce when N > 1).  The optimizer is excluded (SURVEY.md §8d metric 1).  Default workload: CFG-L (dim_h=512,
1.28 B params), per-GPU batch 4, 4096 frames, bf16 compute / fp32 master weights — the trainer defaults.

--impl reference: the reference's own CPU implementation of the same step (the oracle port, proven bit-identical to
/root/reference on CPU), fp32, all host threads, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SIZES = {"L": 512, "S": 128}
# algorithmic forward FLOPs per sample at N=4096: 3x for fwd+bwd
FWD_TFLOP = {"L": 2.500, "S": 0.974}


ATTN_TFLOP_4096 = 0.8246   # 4*H*D*sum_layers L_l^2 at N=4096 (16 heads x 64, 39 attention layers): quadratic in N, the rest is linear


def fwd_tflop(size: str, frames: int) -> float:
    """Algorithmic forward TFLOP of one sample of `frames` frames (SURVEY.md §8d figures at 4096, split into the linear and the
    quadratic (attention) part)."""
    r = frames / 4096.0
    return (FWD_TFLOP[size] - ATTN_TFLOP_4096) * r + ATTN_TFLOP_4096 * r * r


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1394.6), d.get("hbm_gbs", 6551.0), "measured"
    return 1400.0, 6650.0, "fallback"


def gemm_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step's most frequent GEMM shape (the level-0 3-tap conv,
    B4 L4096 N512 K512), from the committed `ncu --set full` capture of this round (profiles/r02_attn_gemm_ncu_full.json, written by
    tools/ncu_rep_summary.py; round-1 file as a fallback); None when absent."""
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for name in ("r02_attn_gemm_ncu_full.json", "r01_gemm_ncu_full.json"):
        p = ROOT / "profiles" / name
        if not p.exists():
            continue
        launches = [l for l in json.loads(p.read_text())["launches"] if "gemm_kernel" in l.get("kernel", "gemm_kernel")]
        if not launches:
            continue
        d = launches[0]
        t = d["dram_read"] * mult.get(d["dram_read_unit"], 1.0) + d["dram_write"] * mult.get(d["dram_write_unit"], 1.0)
        dur = d.get("duration_us", d.get("duration"))
        return t, (f"bytes per launch of the level-0 3-tap conv GEMM B4 L4096 N512 K512 ({dur:.1f} us under ncu, {name}); algorithmic 35.1 MB = "
                   "16.8 MB activations + 1.5 MB weights read, 16.8 MB written (the write stays in the 126 MB L2 during the capture)")
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int) -> None:
        self.index, self.rows, self.proc = index, [], None

    def start(self) -> None:
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def synth_batch(batch: int, frames: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 6, frames, generator=g)
    a = torch.randn(batch, 96, frames, generator=g)
    c = torch.randn(batch, 5, generator=g)
    return x, a, c


# ------------------------------------------------------------------------------------------------ CPU reference / baseline
def cpu_reference_step_time(size: str, frames: int, steps: int, warmup: int):
    """Oracle (port of the reference) fwd+bwd on the host cores, fp32 (mode M2), batch 1."""
    from oracle.models import DiffusionOsuFusion as OracleModel

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = OracleModel(SIZES[size])
    torch.nn.init.normal_(model.unet.final_conv.weight, std=0.02)
    x, a, c = synth_batch(1, frames, 1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        loss = model(x, a, c)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), cores, float(loss)


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = args.ref_frames
    sec, cores, loss = cpu_reference_step_time(args.size, frames, args.steps, args.warmup)
    # one bounded step = batch 1 x `frames` frames; expressed in the workload's unit (samples of args.frames frames) by algorithmic FLOPs
    equiv = fwd_tflop(args.size, frames) / fwd_tflop(args.size, args.frames)
    value = equiv / sec
    sample = (f"batch 1 x {frames} frames per step = {equiv:.3f} workload samples by algorithmic FLOPs (workload: batch {args.batch} x "
              f"{args.frames} frames per GPU), fp32, oracle port of the reference, {sec:.2f} s per step")
    line = {
        "impl": "reference", "metric": "denoiser fwd+bwd samples/s", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload / config keys as our arm (run_ours); `sample` says which bounded part of it one timed step covers
        "config": {"workload": f"CFG-{args.size} dim_h={SIZES[args.size]} denoiser train micro-step (fwd+bwd, all grads), cond_drop_prob=0.5",
                   "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus, "frames": args.frames,
                   "parallelism": f"dp{args.gpus}", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.ref_validate_frames and args.ref_validate_frames != frames:
        # ONE step at the workload's own length (batch 1): checks the FLOP-scaled bounded sample above (attention is 6 % of the FLOPs at
        # 1024 frames and 33 % at 4096, and CPU bf16 SDPA is slower per FLOP than the CPU convolutions)
        vsec, _, _ = cpu_reference_step_time(args.size, args.ref_validate_frames, 1, 0)
        vequiv = fwd_tflop(args.size, args.ref_validate_frames) / fwd_tflop(args.size, args.frames)
        line["validation"] = {"frames": args.ref_validate_frames, "batch": 1, "sec_per_step": vsec, "value": vequiv / vsec,
                              "unit": "samples/s", "note": "single cold step (no warm-up) at the workload's sequence length; "
                                                           "the headline value of this line is the bounded-sample figure"}
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(size: str, dx, da, dc, steps: int = 3):
    """"The reference on the same box" (SURVEY.md §2.3 / §8d): the oracle port of the reference's module tree under
    torch.autocast(cuda, bf16) — cuDNN convs, cuBLAS linears, SDPA flash attention, ATen norms, torch autograd — on the SAME GPU,
    SAME batch, in the same run: eager (what trainer.py runs) and captured as one CUDA graph (host launch overhead excluded).
    A reported baseline; never on the product path."""
    from oracle.models import DiffusionOsuFusion as OracleModel
    dev = dx.device
    torch.manual_seed(0)
    ora = OracleModel(SIZES[size]).to(dev)
    torch.nn.init.normal_(ora.unet.final_conv.weight, std=0.02)
    ora.scheduler.alphas_cumprod = ora.scheduler.alphas_cumprod.to(dev)      # no host->device copy inside the captured step

    def step():
        ora.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = ora(dx, da, dc)
        loss.backward()
        return loss

    def timed(fn, k):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    out = {"what": "oracle port of the reference under torch.autocast(bf16): torch eager + cuDNN/cuBLAS/SDPA on this GPU, same batch"}
    for _ in range(2):
        step()
    out["eager_ms_per_step"] = timed(step, steps)
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        ora.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            step()
        g.replay()
        out["cuda_graph_ms_per_step"] = timed(g.replay, steps)
        del g
    except Exception as e:  # noqa: BLE001
        out["cuda_graph_ms_per_step"] = None
        out["cuda_graph_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    del ora
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args) -> None:
    import torch.distributed as dist

    from osufusion_b200 import _native as NN
    from osufusion_b200 import ops_raw as R
    from osufusion_b200.graphs import GraphedTrainStep
    from osufusion_b200.models import DiffusionOsuFusion

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if args.reserve_sms > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.reserve_sms))   # the collective gets exactly the SMs the GEMM grid leaves free
        dist.init_process_group("nccl", device_id=dev)
    B, n, size = args.batch, args.frames, args.size
    torch.manual_seed(0)
    model = DiffusionOsuFusion(SIZES[size]).to(dev)
    torch.nn.init.normal_(model.unet.final_conv.weight, std=0.02)
    adapted = None
    if args.lora:      # BASELINE.json configs[4]: trainer_peft.py LoRA/DoRA fine-tuning step, base weights frozen
        from osufusion_b200 import lora
        adapted = lora.inject_adapters(model, r=32, lora_alpha=32, use_dora=True)
        model.to(dev)
    sync = None
    if world > 1:
        from osufusion_b200.ddp import GradAllReducer
        sync = GradAllReducer(model, reserve_sms=args.reserve_sms)
    x, a, c = synth_batch(B, n, 1234 + rank)
    hx, ha, hc = x.pin_memory(), a.pin_memory(), c.pin_memory()
    dx, da, dc = hx.to(dev), ha.to(dev), hc.to(dev)

    NN.lib().of_reset_launch_count()
    step = GraphedTrainStep(model, dx, da, dc, warmup=2, post_backward=(sync.all_reduce if sync else None))
    launches_per_step = NN.lib().of_launch_count() // 3   # 2 warm-up steps + 1 captured step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if os.environ.get("OF_PROFILE_STEP") == "1":     # ncu --profile-from-start off: exactly one graph replay is profiled
        torch.cuda.cudart().cudaProfilerStart()
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = B * world / (ms * 1e-3)

    # ---- end-to-end: host buffers in pinned memory -> H2D -> step -> D2H of the loss, every step
    hloss = torch.empty(1).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(hx, ha, hc)                       # copies into the graph's static inputs, then replays
        hloss.copy_(step.loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t)
    e2e_value = B * world / (e2e_ms * 1e-3)
    h2d = hx.numel() * 4 + ha.numel() * 4 + hc.numel() * 4

    # ---- roofline of the dominant kernel family: one instrumented eager step with CUDA events around every launch
    # (every rank runs it so that the gradient all-reduces stay matched; rank 0 reports)
    roof = None
    del loss
    model.zero_grad(set_to_none=True)
    # ... on ONE stream: with the second-stream branches on, launches of the two streams share the SMs and each one's event-to-event
    # time would include its neighbour's work
    from osufusion_b200 import engine as EN
    saved_flags = (EN.WGRAD_SIDE, EN.FWD_SIDE, EN.LORA_MERGE_AHEAD)
    EN.WGRAD_SIDE = EN.FWD_SIDE = EN.LORA_MERGE_AHEAD = False
    NN.PROFILE = []
    l2 = model(dx, da, dc)
    l2.backward()
    torch.cuda.synchronize()
    EN.WGRAD_SIDE, EN.FWD_SIDE, EN.LORA_MERGE_AHEAD = saved_flags
    if rank == 0:
        tf_peak, hbm_peak, how = peaks()
        agg = {}
        for name, flops, ev0, ev1, _tag in NN.PROFILE:
            d = agg.setdefault(name, [0.0, 0.0, 0])
            d[0] += ev0.elapsed_time(ev1)
            d[1] += flops
            d[2] += 1
        top = max(agg.items(), key=lambda kv: kv[1][0])
        name, (tms, fl, cnt) = top
        ach = fl / (tms * 1e-3) / 1e12
        traffic, traffic_note = gemm_traffic()
        roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "traffic": traffic, "traffic_note": traffic_note, "launches": cnt, "ms_per_step": tms,
                "peak_source": how + " (bf16_tflops_sustained)",
                "families_ms": {k: round(v[0], 3) for k, v in agg.items()},
                "families_tflops": {k: round(v[1] / (v[0] * 1e-3) / 1e12, 1) for k, v in agg.items() if v[0] > 0}}

    NN.PROFILE = None
    del l2
    # ---- optimizer tail (excluded from the metric, reported beside it): fused grad-norm clip + AdamW over the gradient arena
    opt_ms = None
    if not args.no_optimizer:
        from osufusion_b200.optim import FusedAdamW
        step()
        opt = FusedAdamW(model, lr=1e-5)
        for _ in range(2):
            opt.step()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(5):
            opt.step()
        o1.record()
        torch.cuda.synchronize()
        opt_ms = o0.elapsed_time(o1) / 5
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, cores, _ = cpu_reference_step_time(size, args.ref_frames, 4, 1)      # ~10 s of CPU work at CFG-L, first (cold) step untimed
        equiv = fwd_tflop(size, args.ref_frames) / fwd_tflop(size, n)
        cpu = {"value": equiv / sec, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"mean of 4 steps (after 1 warm-up) of batch 1 x {args.ref_frames} frames = {equiv:.3f} workload samples each by "
                         f"algorithmic FLOPs, fp32 oracle port on the host CPU ({sec:.1f} s per step)"}

    # ---- cost of the gradient exchange (N > 1): the same captured step with the collectives removed (a) keeping the launch schedule
    # and the SM reservation, (b) without either = the single-GPU step in this process
    comm = None
    if world > 1 and sync is not None and not args.no_comm_breakdown:
        def time_variant(dry_run, world_one):
            sync.dry_run = dry_run
            saved = sync.world
            if world_one:
                sync.world = 1
            g2 = GraphedTrainStep(model, dx, da, dc, warmup=2)
            for _ in range(3):
                g2()
            barrier()
            t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0_.record()
            for _ in range(args.steps):
                g2()
            t1_.record()
            barrier()
            sync.dry_run, sync.world = False, saved
            t = torch.tensor([t0_.elapsed_time(t1_) / args.steps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            del g2
            return float(t)
        try:
            ms_resv = time_variant(True, False)
            ms_single = time_variant(False, True)
        except Exception as e:  # noqa: BLE001  (the breakdown is extra information: never lose the bench line over it)
            ms_resv = ms_single = float("nan")
            sync.dry_run = False
            sync.world = world
            print(f"comm breakdown failed: {type(e).__name__}: {str(e)[:300]}", file=sys.stderr, flush=True)
        comm = {"step_ms": ms, "no_collective_same_schedule_and_reservation_ms": ms_resv, "no_collective_no_reservation_ms": ms_single,
                "exposed_comm_ms": ms - ms_resv, "sm_reservation_ms": ms_resv - ms_single, "total_comm_cost_ms": ms - ms_single,
                "buckets": len(sync.buckets), "bucket_mb": [round((e - s_) * 4 / 2 ** 20, 1) for s_, e, _ in sync.buckets],
                "gradient_bytes": int(sync.arena.numel() * 4)}
    eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager_baseline:
        eager = gpu_eager_baseline(size, dx, da, dc)
        eager["speedup_vs_eager"] = eager["eager_ms_per_step"] / ms
        if eager.get("cuda_graph_ms_per_step"):
            eager["speedup_vs_eager_cuda_graph"] = eager["cuda_graph_ms_per_step"] / ms

    if rank == 0:
        tf_peak, _, _ = peaks()
        step_tflop = 3 * fwd_tflop(size, n) * B
        line = {
            "metric": "denoiser fwd+bwd samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (f"CFG-{size} dim_h={SIZES[size]} denoiser train micro-step (fwd+bwd, all grads), cond_drop_prob=0.5"
                                    if adapted is None else
                                    f"CFG-{size} dim_h={SIZES[size]} LoRA/DoRA fine-tuning micro-step (r=32, alpha=32, {len(adapted)} adapted "
                                    f"modules, base frozen), cond_drop_prob=0.5"),
                       "per_gpu_batch": B, "global_batch": B * world, "frames": n, "parallelism": f"dp{world}",
                       "reserve_sms_for_nccl": (args.reserve_sms if world > 1 else 0),
                       "l2": "working set (2.6 GB bf16 weights + activations) >> 126 MB L2; no explicit flush",
                       "cuda_graph": True,
                       "weight_operands": "bf16 GEMM operands are written by the fused optimizer kernel together with the fp32 master update "
                                          "(optimizer_ms_per_step, outside the metric like the reference's optimizer.step()); the micro-step "
                                          "re-packs them only when weights changed by other means"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step * args.steps),
            "model_tflops_per_gpu": step_tflop / (ms * 1e-3), "mfu_vs_measured_peak": step_tflop / (ms * 1e-3) / tf_peak,
            "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "comm": comm, "loss": float(step.loss.detach()),
            "optimizer_ms_per_step": opt_ms,
        }
        print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world: int) -> None:
    """Multi-rank exit: tearing the NCCL communicator down while CUDA graphs that captured its collectives are still alive can
    block for minutes, so ranks synchronise, flush and leave without destroy_process_group()."""
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------------ sampling (metric 2)
def run_sampling(args) -> None:
    """End-to-end sampling beatmap-frames/s (BASELINE.json configs[3]): complete DDIM loop (35 steps, CFG -> 70 denoiser
    evaluations), batch-sharded across ranks with no collective."""
    import torch.distributed as dist

    from osufusion_b200 import _native as NN
    from osufusion_b200.models import DiffusionOsuFusion

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if args.reserve_sms > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.reserve_sms))   # the collective gets exactly the SMs the GEMM grid leaves free
        dist.init_process_group("nccl", device_id=dev)
    B, n, size = args.batch, args.frames, args.size
    torch.manual_seed(0)
    model = DiffusionOsuFusion(SIZES[size]).to(dev).eval()
    torch.nn.init.normal_(model.unet.final_conv.weight, std=0.02)
    x, a, c = synth_batch(B, n, 1234 + rank)
    ha, hc, hx = a.pin_memory(), c.pin_memory(), x.pin_memory()
    hout = torch.empty(B, 6, n).pin_memory()

    def one_song():
        da, dc, dx = ha.to(dev, non_blocking=True), hc.to(dev, non_blocking=True), hx.to(dev, non_blocking=True)
        y = model.sample(da, dc, dx, cond_scale=args.cond_scale)
        hout.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(1, min(args.warmup, 1))):
        one_song()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    NN.lib().of_reset_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        one_song()
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms, wall_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall_ms = float(t[0]), float(t[1])
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        tf_peak, _, how = peaks()
        fwd = {("S", 32768): 53.97, ("L", 32768): 66.18, ("S", 4096): 0.974, ("L", 4096): 2.5, ("S", 65536): 213.5}.get((size, n))
        evals = model.sampling_timesteps * (2 if args.cond_scale != 1.0 else 1)
        line = {
            "metric": "sampling beatmap-frames/s", "value": B * world * n / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"CFG-{size} dim_h={SIZES[size]} full DDIM sampling loop, {model.sampling_timesteps} steps, "
                                   f"cond_scale={args.cond_scale} ({evals} denoiser evaluations), audio encoder cached",
                       "songs_per_gpu": B, "frames": n, "parallelism": f"replicas x{world} (no collective)",
                       "l2": "per-evaluation working set >> 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": B * world * n / (wall_ms * 1e-3), "unit": "frames/s", "ms_per_step": wall_ms,
                    "h2d_bytes_per_step": (ha.numel() + hc.numel() + hx.numel()) * 4, "d2h_bytes_per_step": hout.numel() * 4},
            "gpu_launches": int(NN.lib().of_launch_count()),
        }
        if fwd is not None:
            tfl = evals * fwd * B / (ms * 1e-3)
            line["algorithmic_tflops_per_gpu"] = tfl
            line["roofline"] = {"kernel": "attn_fwd_kernel + gemm_kernel (whole sampler)", "bound": "tensor", "achieved": tfl,
                                "peak": tf_peak, "unit": "TFLOP/s", "frac": tfl / tf_peak, "traffic": None,
                                "note": f"algorithmic FLOPs = {evals} x F_fwd (reference's count; the cached audio encoder removes work)",
                                "peak_source": how}
        print(json.dumps(line), flush=True)
    _finish(world)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="L", choices=["L", "S"])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--ref-frames", type=int, default=1024, help="frames of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true", help="skip timing the oracle under torch eager bf16 on the same GPU")
    ap.add_argument("--no-comm-breakdown", action="store_true", help="N>1: skip the two extra captured steps that isolate the all-reduce cost")
    ap.add_argument("--ref-validate-frames", type=int, default=4096,
                    help="--impl reference: also time ONE step at this length (batch 1) to validate the FLOP-scaled bounded sample; 0 = off")
    ap.add_argument("--reserve-sms", type=int, default=16,
                    help="N>1: SMs left to the overlapped NCCL all-reduce (NCCL_MAX_CTAS and the persistent GEMM grid; measured at 2 GPUs: "
                         "0 -> 67.9 ms, 8 -> 70.0, 16 -> 65.7, 32 -> 70.2)")
    ap.add_argument("--lora", action="store_true", help="BASELINE.json configs[4]: LoRA/DoRA fine-tuning step (base frozen)")
    ap.add_argument("--no-optimizer", action="store_true",
                    help="skip timing the fused clip + AdamW step (reported separately as optimizer_ms_per_step; it also emits the bf16 GEMM "
                         "operands of the next forward pass, which is why the micro-step itself has no weight re-cast pass)")
    ap.add_argument("--mode", default="train", choices=["train", "sample"], help="train: fwd+bwd samples/s; sample: frames/s")
    ap.add_argument("--cond-scale", type=float, default=2.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "sample":
        run_sampling(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
