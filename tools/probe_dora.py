"""Timing / ncu driver for the DoRA merge + gradient-projection kernels on one conv-sized module."""
import sys
import torch
sys.path.insert(0, ".")
from osufusion_b200 import _native as N
dev = "cuda"
Cout, Cin, k, r = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (1024, 1024, 3, 32)))
torch.manual_seed(0)
W = torch.randn(Cout, Cin, k, device=dev) / (Cin * k) ** 0.5
A = torch.randn(r, Cin, k, device=dev) / (Cin * k) ** 0.5
Bm = 0.05 * torch.randn(Cout, r, device=dev)
mag = torch.rand(Cout, device=dev) + 0.5
cp = (Cin + 7) // 8 * 8
packed = torch.zeros(k, Cout, cp, device=dev, dtype=torch.bfloat16)
n2 = torch.empty(Cout, device=dev)
dWp = torch.randn(k, Cout, cp, device=dev)
dA, dB, dm = torch.zeros_like(A), torch.zeros_like(Bm), torch.zeros(Cout, device=dev)


def merge():
    N.call("of_dora_merge", W.data_ptr(), A.data_ptr(), Bm.data_ptr(), mag.data_ptr(), 1.0, Cout, Cin, k, r, n2.data_ptr(),
           packed.data_ptr(), cp, Cout * cp, None)


def grad():
    N.call("of_dora_grad", W.data_ptr(), A.data_ptr(), Bm.data_ptr(), mag.data_ptr(), 1.0, Cout, Cin, k, r, n2.data_ptr(),
           dWp.data_ptr(), cp, Cout * cp, dA.data_ptr(), dB.data_ptr(), dm.data_ptr())


for name, fn in (("merge", merge), ("grad", grad)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"dora {name} {Cout}x{Cin}x{k} r{r}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
