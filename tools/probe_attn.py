"""GPU probe for of_attn_fwd / of_attn_bwd vs torch SDPA (fp32 math on the same bf16 inputs)."""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from osufusion_b200 import ops_raw as R  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def ref_attn(q, k, v, H, KVH, D):
    B, L, _ = q.shape
    qh = q.float().view(B, L, H, D).transpose(1, 2)
    kh = k.float().view(B, L, KVH, D).transpose(1, 2).repeat(1, H // KVH, 1, 1)
    vh = v.float().view(B, L, KVH, D).transpose(1, 2).repeat(1, H // KVH, 1, 1)
    s = (qh @ kh.transpose(-1, -2)) / D ** 0.5
    p = s.softmax(-1)
    o = p @ vh
    lse2 = torch.logsumexp(s, -1) * 1.4426950408889634
    return o.transpose(1, 2).reshape(B, L, H * D), lse2


def case_fwd(B, L, H, KVH, D, variant, qscale=1.0):
    qkv = (torch.randn(B, L, (H + 2 * KVH) * D, device=dev) * qscale).bfloat16()
    q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:(H + KVH) * D], qkv[:, :, (H + KVH) * D:]
    out = torch.zeros(B, L, H * D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=dev)
    R.attn_fwd(q, k, v, out, lse, H=H, KVH=KVH, D=D, variant=variant)
    torch.cuda.synchronize()
    o_ref, lse_ref = ref_attn(q, k, v, H, KVH, D)
    print(f"attn fwd v{variant} B{B} L{L} H{H} KVH{KVH} D{D} qs{qscale}: out rel={rel(out, o_ref):.2e} lse rel={rel(lse, lse_ref):.2e}", flush=True)


def perf_fwd(B, L, H, D, variant):
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
    fl = 4.0 * B * H * L * L * D
    print(f"perf attn fwd v{variant} B{B} L{L}: {ms * 1e3:.1f} us {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    # torch SDPA for comparison
    qh = q.reshape(B, L, H, D).transpose(1, 2).contiguous()
    kh = k.reshape(B, L, 1, D).transpose(1, 2).expand(B, H, L, D).contiguous()
    vh = v.reshape(B, L, 1, D).transpose(1, 2).expand(B, H, L, D).contiguous()
    for _ in range(3):
        F.scaled_dot_product_attention(qh, kh, vh)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        F.scaled_dot_product_attention(qh, kh, vh)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"     torch sdpa B{B} L{L}: {ms * 1e3:.1f} us {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


def _bwd_setup(B, L, H, KVH, D):
    qkv = torch.randn(B, L, (H + 2 * KVH) * D, device=dev).bfloat16()
    q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:(H + KVH) * D], qkv[:, :, (H + KVH) * D:]
    out = torch.zeros(B, L, H * D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=dev)
    R.attn_fwd(q, k, v, out, lse, H=H, KVH=KVH, D=D)
    dout = torch.randn(B, L, H * D, device=dev).bfloat16()
    delta = torch.zeros(B, H, L, device=dev)
    dq = torch.zeros(B, L, H * D, device=dev)
    dkv = torch.zeros(B, L, 2 * KVH * D, device=dev)
    return q, k, v, out, lse, dout, delta, dq, dkv


def case_bwd(B, L, H, KVH, D):
    q, k, v, out, lse, dout, delta, dq, dkv = _bwd_setup(B, L, H, KVH, D)
    R.attn_bwd(q, k, v, out, lse, dout, delta, dq, dkv[:, :, :KVH * D], dkv[:, :, KVH * D:], H=H, KVH=KVH, D=D)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = ref_attn(qf, kf, vf, H, KVH, D)
    o_ref.backward(dout.float())
    print(f"attn bwd B{B} L{L} H{H} KVH{KVH} D{D}: dq rel={rel(dq, qf.grad):.2e} dk rel={rel(dkv[:, :, :KVH * D], kf.grad):.2e} "
          f"dv rel={rel(dkv[:, :, KVH * D:], vf.grad):.2e}", flush=True)


def perf_bwd(B, L, H, D):
    q, k, v, out, lse, dout, delta, dq, dkv = _bwd_setup(B, L, H, 1, D)
    f = lambda: R.attn_bwd(q, k, v, out, lse, dout, delta, dq, dkv[:, :, :D], dkv[:, :, D:], H=H, KVH=1, D=D)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 10.0 * B * H * L * L * D
    print(f"perf attn bwd B{B} L{L}: {ms * 1e3:.1f} us {fl / ms / 1e9:.1f} TFLOP/s (5 MMAs)", flush=True)
    qh = q.reshape(B, L, H, D).transpose(1, 2).contiguous().requires_grad_(True)
    kh = k.reshape(B, L, 1, D).transpose(1, 2).expand(B, H, L, D).contiguous().requires_grad_(True)
    vh = v.reshape(B, L, 1, D).transpose(1, 2).expand(B, H, L, D).contiguous().requires_grad_(True)
    o = F.scaled_dot_product_attention(qh, kh, vh)
    g = torch.randn_like(o)
    for _ in range(3):
        torch.autograd.grad(o, (qh, kh, vh), g, retain_graph=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        torch.autograd.grad(o, (qh, kh, vh), g, retain_graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"     torch sdpa bwd B{B} L{L}: {ms * 1e3:.1f} us {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    grp = sys.argv[1]
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if grp == "fwd":
        case_fwd(1, 128, 1, 1, 64, variant)
        case_fwd(1, 256, 2, 1, 64, variant)
        case_fwd(2, 1024, 16, 1, 64, variant)
        case_fwd(2, 200, 2, 1, 16, variant)
        case_fwd(1, 72, 4, 2, 32, variant)
        case_fwd(1, 512, 4, 1, 64, variant, qscale=6.0)
    if grp == "perf":
        perf_fwd(4, 4096, 16, 64, variant)
        perf_fwd(4, 1024, 16, 64, variant)
        perf_fwd(1, 32768, 16, 64, variant)
    if grp == "bwd":
        case_bwd(1, 128, 1, 1, 64)
        case_bwd(1, 256, 2, 1, 64)
        case_bwd(2, 1024, 16, 1, 64)
        case_bwd(2, 200, 2, 1, 16)
        case_bwd(1, 72, 4, 2, 32)
    if grp == "perfbwd":
        perf_bwd(4, 4096, 16, 64)
        perf_bwd(4, 1024, 16, 64)
    print("done", grp, variant, flush=True)
