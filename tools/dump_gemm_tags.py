"""One instrumented eager training micro-step (CFG-L B4 N4096): every entry-point launch in issue order with its family, shape tag,
algorithmic FLOPs and CUDA-event time -> JSON list on stdout.  Joined offline with an ncu launch list (tools/gemm_shapes.py)."""
import json
import os
import sys
os.environ.setdefault("OF_WGRAD_SIDE", "0")
sys.path.insert(0, ".")
import torch
from osufusion_b200 import _native as NN
from osufusion_b200.models import DiffusionOsuFusion
from bench import synth_batch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
lora = "--lora" in sys.argv
model = DiffusionOsuFusion(512).to(dev)
torch.nn.init.normal_(model.unet.final_conv.weight, std=0.02)
if lora:
    from osufusion_b200 import lora as L
    L.inject_adapters(model, r=32, lora_alpha=32, use_dora=True)
    model.to(dev)
x, a, c = (t.to(dev) for t in synth_batch(4, 4096, 1234))
for _ in range(2):
    model.zero_grad(set_to_none=True)
    model(x, a, c).backward()
torch.cuda.synchronize()
model.zero_grad(set_to_none=True)
NN.PROFILE = []
model(x, a, c).backward()
torch.cuda.synchronize()
out = [{"family": f, "tag": tag, "flops": fl, "ms": e0.elapsed_time(e1)} for f, fl, e0, e1, tag in NN.PROFILE]
print(json.dumps(out))
