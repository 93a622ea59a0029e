#!/usr/bin/env python
"""Minimal training loop on synthetic data that mirrors the reference trainer's micro-step (trainer.py:290-309) with this repo's
drop-in pieces: model, (optional) CUDA-graph step, data-parallel gradient all-reduce, fused clip + AdamW, cosine schedule and
reference-layout checkpoints.

    python examples/train_synthetic.py --steps 20                       # 1 GPU
    torchrun --nproc-per-node 8 examples/train_synthetic.py --steps 20  # 8 GPUs, NCCL all-reduce overlapped with backward
    python examples/train_synthetic.py --lora --steps 20                # trainer_peft.py: DoRA adapters, base frozen
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-dim", type=int, default=128)       # trainer.py --model-dim (512 = CFG-L)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lr", type=float, default=1e-5)
    ap.add_argument("--lora", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the micro-step as one CUDA graph")
    ap.add_argument("--out", type=str, default="")
    args = ap.parse_args()

    import torch.distributed as dist
    from osufusion_b200 import checkpoint as ck
    from osufusion_b200.graphs import GraphedTrainStep
    from osufusion_b200.models import DiffusionOsuFusion
    from osufusion_b200.optim import FusedAdamW, cosine_schedule_with_warmup

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = DiffusionOsuFusion(args.model_dim).to(dev)
    torch.nn.init.normal_(model.unet.final_conv.weight, std=0.02)      # the reference zero-initialises it: gradients would be zero
    if args.lora:
        from osufusion_b200 import lora
        lora.inject_adapters(model, r=32, lora_alpha=32, use_dora=True)
        model.to(dev)
    if world > 1:
        from osufusion_b200.ddp import GradAllReducer
        GradAllReducer(model, reserve_sms=16)
    opt = FusedAdamW(model, lr=args.lr, max_grad_norm=1.0)
    sched = cosine_schedule_with_warmup(opt, 5, args.steps)

    g = torch.Generator().manual_seed(1234 + rank)

    def batch():
        return (torch.randn(args.batch, 6, args.frames, generator=g).pin_memory().to(dev, non_blocking=True),
                torch.randn(args.batch, 96, args.frames, generator=g).pin_memory().to(dev, non_blocking=True),
                torch.randn(args.batch, 5, generator=g).pin_memory().to(dev, non_blocking=True))

    step_fn = GraphedTrainStep(model, *batch()) if args.graph else None
    t0 = time.perf_counter()
    for it in range(args.steps):
        x, a, c = batch()
        if step_fn is not None:
            loss = step_fn(x, a, c)
        else:
            model.zero_grad(set_to_none=True)
            loss = model(x, a, c)
            loss.backward()
        opt.step()
        sched.step()
        if rank == 0 and (it % 5 == 0 or it == args.steps - 1):
            print(f"step {it:4d}  loss {float(loss):.4f}  grad-norm {float(opt.grad_norm):.3f}  lr {sched.get_last_lr()[0]:.2e}", flush=True)
    torch.cuda.synchronize()
    if rank == 0:
        dt = time.perf_counter() - t0
        print(f"{args.steps} steps in {dt:.2f} s  ({args.steps * args.batch * world / dt:.1f} samples/s incl. optimizer and host batches)")
        if args.out:
            if args.lora:
                print("saved", ck.save_peft_checkpoint(model, opt, sched, args.steps - 1, Path(args.out)))
            else:
                ck.save_model_sd(model, Path(args.out))
                print("saved", ck.save_checkpoint(model, opt, sched, args.steps - 1, Path(args.out)))
    if world > 1:
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
