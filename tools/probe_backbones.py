"""Times one training micro-step (forward + backward, all gradients) of the DiT / MMDiT backbones on the engine and, beside it, the
oracle under torch-eager bf16 autocast on the same GPU (cuBLAS / SDPA: what the reference would run).  CUDA events, max of nothing
(1 GPU), `--steps` timed iterations after `--warmup`.  Prints one JSON line per backbone.

    python tools/probe_backbones.py [--dim 512] [--depth 12] [--batch 4] [--frames 4096] [--steps 10] [--warmup 3] [--no-eager]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def flops_fwd(kind, B, n, C, depth, mult=4, H=8, D=64, KVH=2, p=4):
    """Algorithmic forward FLOPs (2 * MACs; attention 4 * H * L^2 * D), embeddings and output convs included."""
    if kind == "dit":
        L = n
        blk = 2 * L * C * 3 * C + 4 * H * L * L * D + 2 * 2 * L * C * mult * C
        emb = 2 * L * 102 * (51 * 3 + 25 * 7 + (C - 76) * 15) + 2 * L * C * C + 2 * L * C * 6
        return B * (depth * blk + emb)
    m = -(-n // p)
    W = (H + 2 * KVH) * D
    per_stream = 2 * m * C * W + 2 * m * C * C + 2 * 2 * m * C * mult * C
    blk = 2 * per_stream + 4 * H * (2 * m) * (2 * m) * D
    emb = 2 * m * p * 6 * C + 2 * m * p * 96 * C + 2 * m * C * p * C + 2 * m * p * C * 6
    return B * (depth * blk + emb)


def time_steps(step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--only", choices=["dit", "mmdit"], default=None)
    args = ap.parse_args()
    from oracle.backbones import DiT as ODiT, MMDiT as OMMDiT
    from oracle.synth import synth_inputs, synth_state_dict
    from osufusion_b200 import _native as N
    from osufusion_b200.backbones import DiT, MMDiT
    peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())
    peak = float(peaks.get("bf16_tflops_sustained", 1394.6))
    dev = "cuda"
    x, a, c, t, noise, mask = (v.to(dev) for v in synth_inputs(args.batch, args.frames, 1234))
    for kind, new_cls, ora_cls in (("dit", DiT, ODiT), ("mmdit", MMDiT, OMMDiT)):
        if args.only is not None and kind != args.only:
            continue
        cfg = dict(dim_h=args.dim, depth=args.depth)
        ora = ora_cls(6, 96, 5, **cfg)
        ora.load_state_dict(synth_state_dict(ora))
        new = new_cls(6, 96, 5, **cfg)
        new.load_state_dict(ora.state_dict())
        new = new.to(dev)

        def step_new():
            new.zero_grad(set_to_none=True)
            torch.nn.functional.mse_loss(new(x, a, t, c, cond_mask=mask), noise).backward()

        N.lib().of_reset_launch_count()
        step_new()
        torch.cuda.synchronize()
        launches = N.lib().of_launch_count()
        ms_eager = time_steps(step_new, args.steps, args.warmup)
        from osufusion_b200.graphs import GraphedCallable
        graphed = GraphedCallable(step_new)
        ms = time_steps(graphed, args.steps, args.warmup)
        N.PROFILE = []                       # one instrumented eager step: CUDA events around every entry point
        step_new()
        torch.cuda.synchronize()
        fam = {}
        for name, _, e0, e1, _ in N.PROFILE:
            fam[name] = fam.get(name, 0.0) + e0.elapsed_time(e1)
        N.PROFILE = None
        fl = 3.0 * flops_fwd(kind, args.batch, args.frames, args.dim, args.depth)
        line = {"backbone": kind, "metric": "fwd+bwd samples/s", "value": args.batch / ms * 1e3, "ms_per_step": ms,
                "config": {"dim_h": args.dim, "depth": args.depth, "batch": args.batch, "frames": args.frames, "heads": "8x64",
                           "cuda_graph": True},
                "algorithmic_tflops": fl / ms / 1e9, "frac_of_measured_bf16_peak": fl / ms / 1e9 / peak, "peak_tflops": peak,
                "gpu_launches_per_step": launches, "batched_adaln_gate": os.environ.get("OF_BACKBONE_BATCHED", "1"),
                "headnorm_variant": os.environ.get("OF_HEADNORM_VARIANT", "0 (auto: 2)"),
                "families_ms_one_eager_step": {k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}, "ms_per_step_eager_launches": ms_eager, "params_m": sum(p.numel() for p in new.parameters()) / 1e6}
        if not args.no_eager:
            ora = ora.to(dev)

            def step_ref():
                ora.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    y = ora(x, a, t, c, cond_mask=mask)
                torch.nn.functional.mse_loss(y.float(), noise).backward()
            line["torch_eager_bf16_autocast_ms"] = time_steps(step_ref, max(3, args.steps // 2), 2)
            del ora
        print(json.dumps(line), flush=True)
        del new
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
