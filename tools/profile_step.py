"""Per-kernel-family and per-GEMM-shape time breakdown of one eager training micro-step (CUDA events around each launch)."""
import sys
from collections import defaultdict

import torch

sys.path.insert(0, ".")
from osufusion_b200 import _native as NN  # noqa: E402
from osufusion_b200.models import DiffusionOsuFusion  # noqa: E402

size, B, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = "cuda"
torch.manual_seed(0)
model = DiffusionOsuFusion({"L": 512, "S": 128}[size]).to(dev)
torch.nn.init.normal_(model.unet.final_conv.weight, std=0.02)
if len(sys.argv) > 4 and sys.argv[4] == "lora":
    from osufusion_b200 import lora
    lora.inject_adapters(model, r=32, lora_alpha=32, use_dora=True)
    model.to(dev)
x, a, c = torch.randn(B, 6, n, device=dev), torch.randn(B, 96, n, device=dev), torch.randn(B, 5, device=dev)
for _ in range(2):
    model.zero_grad(set_to_none=True)
    model(x, a, c).backward()
torch.cuda.synchronize()
model.zero_grad(set_to_none=True)
NN.PROFILE = []
model(x, a, c).backward()
torch.cuda.synchronize()
fam = defaultdict(lambda: [0.0, 0.0, 0])
shp = defaultdict(lambda: [0.0, 0.0, 0])
for name, flops, e0, e1, tag in NN.PROFILE:
    ms = e0.elapsed_time(e1)
    for d, k in ((fam, name), (shp, (name, tag))):
        d[k][0] += ms
        d[k][1] += flops
        d[k][2] += 1
tot = sum(v[0] for v in fam.values())
print(f"total instrumented {tot:.1f} ms")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]:8.2f} ms {100 * v[0] / tot:5.1f}%  n={v[2]:4d}  {v[1] / max(v[0], 1e-9) / 1e9:7.1f} TF/s  {k}")
print("--- GEMM shapes")
for k, v in sorted(((k, v) for k, v in shp.items() if k[1]), key=lambda kv: -kv[1][0])[:45]:
    print(f"{v[0]:8.2f} ms n={v[2]:3d} {v[0] / v[2] * 1e3:8.1f} us/call {v[1] / max(v[0], 1e-9) / 1e9:7.1f} TF/s  {k[1]}")
