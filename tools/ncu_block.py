"""Driver for `ncu --set full` captures of the bandwidth-bound kernels: a one-level CFG-L-width denoiser (every block at
B x L x 512, the level-0 shape of the headline workload) doing one eager fwd+bwd.

    ncu --set full --clock-control none -k regex:'rb_|linear_small|layernorm|colsum|pack_|cast_|rope|softmax' \
        -o gpurun_out/prof_hbm python tools/ncu_block.py 4 4096
"""
import sys

import torch

sys.path.insert(0, ".")
from osufusion_b200.modules import UNet  # noqa: E402

B, n = int(sys.argv[1]), int(sys.argv[2])
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 512
dev = "cuda"
torch.manual_seed(0)
net = UNet(6, 96, 5, dim, dim_h_mult=(1,), num_layer_blocks=(1,), num_middle_transformers=1).to(dev)
torch.nn.init.normal_(net.final_conv.weight, std=0.02)
x, a, c = torch.randn(B, 6, n, device=dev), torch.randn(B, 96, n, device=dev), torch.randn(B, 5, device=dev)
t = torch.randint(0, 1000, (B,), device=dev)
y = net(x, a, t, c, cond_drop_prob=0.5)
y.square().mean().backward()
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
