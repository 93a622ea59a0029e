"""Timing of individual bandwidth kernels at the level-0 shape (CUDA events, 20 launches)."""
import sys
import torch
sys.path.insert(0, ".")
from osufusion_b200 import engine as E
dev = "cuda"
B, L = 4, 4096
for C in (512, 1024):
    dy = torch.randn(B, L, C, device=dev).bfloat16()
    db = torch.zeros(C, device=dev)
    for _ in range(3):
        E.colsum(dy, db)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        E.colsum(dy, db)
    e1.record()
    torch.cuda.synchronize()
    ref = dy.float().sum((0, 1)) * 23
    err = ((db - ref).abs().max() / ref.abs().max()).item()
    print(f"colsum B{B} L{L} C{C}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us  err {err:.2e}")
