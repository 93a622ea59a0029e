"""A/B probe of the of_attn_fwd schedule variants (see attn_fwd.cu: alternation of the two softmax groups, polynomial exp2 share)."""
import sys

import torch

sys.path.insert(0, ".")
from osufusion_b200 import ops_raw as R  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def ref_attn(q, k, v, H, D):
    B, L, _ = q.shape
    qh = q.float().view(B, L, H, D).transpose(1, 2)
    s = (qh @ k.float().view(B, 1, L, D).transpose(-1, -2)) / D ** 0.5
    o = s.softmax(-1) @ v.float().view(B, 1, L, D)
    return o.transpose(1, 2).reshape(B, L, H * D), torch.logsumexp(s, -1) * 1.4426950408889634


def run(B, L, H, D, variant, check):
    qkv = torch.randn(B, L, (H + 2) * D, device=dev).bfloat16()
    q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:(H + 1) * D], qkv[:, :, (H + 1) * D:]
    out = torch.zeros(B, L, H * D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=dev)
    for _ in range(3):
        R.attn_fwd(q, k, v, out, lse, H=H, KVH=1, D=D, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        R.attn_fwd(q, k, v, out, lse, H=H, KVH=1, D=D, variant=variant)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    msg = f"variant {variant:2d} B{B} L{L}: {ms * 1e3:8.1f} us {4.0 * B * H * L * L * D / ms / 1e9:7.1f} TFLOP/s"
    if check:
        o_ref, lse_ref = ref_attn(q, k, v, H, D)
        msg += f"  out rel {rel(out, o_ref):.2e} lse rel {rel(lse, lse_ref):.2e}"
    print(msg, flush=True)


if __name__ == "__main__":
    variants = [int(a) for a in sys.argv[1:]] or [2, 10, 11, 12, 13, 14, 21, 22]
    for v in variants:
        run(2, 1024, 16, 64, v, True)
    for v in variants:
        run(4, 4096, 16, 64, v, False)
    for v in variants[:1] + variants[1:4]:
        run(1, 32768, 16, 64, v, False)
