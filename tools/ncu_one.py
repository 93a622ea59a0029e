"""Tiny driver for ncu --set full captures: one attention forward/backward and a few GEMM shapes."""
import sys
import torch
sys.path.insert(0, ".")
from osufusion_b200 import ops_raw as R
dev = "cuda"
torch.manual_seed(0)
B, L, H, D = 4, 4096, 16, 64
qkv = torch.randn(B, L, (H + 2) * D, device=dev).bfloat16()
q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:(H + 1) * D], qkv[:, :, (H + 1) * D:]
out = torch.zeros(B, L, H * D, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, L, device=dev)
for _ in range(2):
    R.attn_fwd(q, k, v, out, lse, H=H, KVH=1, D=D)
dout = torch.randn(B, L, H * D, device=dev).bfloat16()
delta = torch.zeros(B, H, L, device=dev)
dq = torch.zeros(B, L, H * D, device=dev)
dkv = torch.zeros(B, L, 2 * D, device=dev)
for _ in range(2):
    R.attn_bwd(q, k, v, out, lse, dout, delta, dq, dkv[:, :, :D], dkv[:, :, D:], H=H, KVH=1, D=D)
for (Bb, Ll, N, K, T) in [(4, 4096, 512, 512, 3), (4, 4096, 1024, 512, 1), (4, 512, 2048, 2048, 3)]:
    x = torch.randn(Bb, Ll, K, device=dev).bfloat16()
    w = (torch.randn(T, N, K, device=dev) / (K * T) ** 0.5).bfloat16()
    o = torch.empty(Bb, Ll, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        R.gemm_fwd(x, w, N_out=N, K=K, taps=T, shift0=-(T // 2), shift_step=1, out_bf16=o)
torch.cuda.synchronize()
print("ok")
