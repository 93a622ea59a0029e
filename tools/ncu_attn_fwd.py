"""ncu driver: of_attn_fwd at the level-0 shape (B4 H16 KVH1 L4096 D64), default variant."""
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
for _ in range(3):
    R.attn_fwd(q, k, v, out, lse, H=H, KVH=1, D=D)
torch.cuda.synchronize()
print("ok")
