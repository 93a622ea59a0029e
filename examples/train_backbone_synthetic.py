#!/usr/bin/env python
"""Minimal training loop of the DiT / MMDiT backbones (osu_fusion/modules/dit.py, mmdit.py) on synthetic data with this repo's
drop-in pieces: backbone, noise-prediction loss on DDIM-noised inputs, one CUDA graph per micro-step, (optional) data-parallel
gradient all-reduce overlapped with backward, fused clip + AdamW.  The reference does not wire these backbones into its trainers;
this mirrors what trainer.py does with the U-Net (trainer.py:290-309).  Status: the captured micro-step is what tools/probe_backbones.py
times on the B200 and the optimizer pairing is covered by tests/test_backbones_host_cpu.py, but this script as a whole has not been run
on a B200 yet (round-1 GPU budget).

    python examples/train_backbone_synthetic.py --backbone mmdit --steps 20
    torchrun --nproc-per-node 8 examples/train_backbone_synthetic.py --backbone dit --steps 20
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
    ap.add_argument("--backbone", choices=["dit", "mmdit"], default="mmdit")
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lr", type=float, default=1e-4)
    args = ap.parse_args()

    import torch.distributed as dist
    from osufusion_b200.backbones import DiT, MMDiT
    from osufusion_b200.graphs import GraphedCallable
    from osufusion_b200.models.diffusion import DDIMScheduler
    from osufusion_b200.optim import FusedAdamW, cosine_schedule_with_warmup

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = (DiT if args.backbone == "dit" else MMDiT)(6, 96, 5, dim_h=args.dim, depth=args.depth).to(dev)
    # the reference zero-initialises the adaLN heads and the output conv (dit.py:238-250): the output conv needs a non-zero start
    # for any gradient to flow
    out_conv = net.postprocess if args.backbone == "dit" else net.out
    torch.nn.init.normal_(out_conv.weight, std=0.02)
    if world > 1:
        from osufusion_b200.ddp import GradAllReducer
        GradAllReducer(net, reserve_sms=16)
    opt = FusedAdamW(net, lr=args.lr, max_grad_norm=1.0)
    sched = cosine_schedule_with_warmup(opt, 5, args.steps)
    alphas = DDIMScheduler(1000).alphas_cumprod_on(dev)

    g = torch.Generator().manual_seed(1234 + rank)
    B, n = args.batch, args.frames
    x = torch.randn(B, 6, n, generator=g).to(dev)
    a = torch.randn(B, 96, n, generator=g).to(dev)
    c = torch.randn(B, 5, generator=g).to(dev)
    # static buffers the captured step reads; refreshed in place every iteration
    noise, xt = torch.empty_like(x), torch.empty_like(x)
    t = torch.zeros(B, dtype=torch.int64, device=dev)
    keep = torch.ones(B, dtype=torch.bool, device=dev)

    def micro_step():
        net.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(net(xt, a, t, c, cond_mask=keep), noise)
        loss.backward()
        return loss

    step = None
    t0 = time.perf_counter()
    for it in range(args.steps):
        noise.normal_()
        t.random_(0, 1000)
        keep.copy_(torch.rand(B, device=dev) < 0.5)                  # cond_drop_prob = 0.5
        ac = alphas[t].view(B, 1, 1)
        xt.copy_(ac.sqrt() * x + (1 - ac).sqrt() * noise)             # DDIMScheduler.add_noise (diffusion.py:96)
        if step is None:
            step = GraphedCallable(micro_step)                       # warm-up + capture with the current buffers
        loss = step()
        opt.step()
        sched.step()
        if rank == 0 and (it % 5 == 0 or it == args.steps - 1):
            print(f"step {it:4d}  loss {loss.item():.4f}  grad-norm {float(opt.grad_norm):.3f}  lr {sched.get_last_lr()[0]:.2e}")
    torch.cuda.synchronize()
    if rank == 0:
        dt = (time.perf_counter() - t0) / args.steps
        print(f"{args.backbone}: {B * world / dt:.1f} samples/s over {world} GPU(s) ({dt * 1e3:.1f} ms per step incl. optimizer)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
