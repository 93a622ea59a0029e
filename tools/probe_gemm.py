"""GPU probe for of_gemm: every operand-major / tap / epilogue variant against a torch fp32 reference.
Usage: python tools/probe_gemm.py [group]   (group in fwd, bmn, wgrad, epi, perf, all)"""
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


def conv_ref(x, w, shift0):
    # x (B,L,C) bf16; w [T][N][K]; out[b,l,n] = sum_t sum_k x[b,l+shift0+t,k] w[t][n][k]
    T = w.shape[0]
    B, L, Cc = x.shape
    out = torch.zeros(B, L, w.shape[1], device=x.device, dtype=torch.float32)
    xf = x.float()
    for t in range(T):
        s = shift0 + t
        xs = torch.zeros_like(xf)
        lo, hi = max(0, -s), min(L, L - s)
        if hi > lo:
            xs[:, lo:hi] = xf[:, lo + s:hi + s]
        out += xs @ w[t].float().t()
    return out


def case_fwd(B, L, N, K, T, block_n=0):
    x = torch.randn(B, L, K, device=dev).bfloat16()
    w = (torch.randn(T, N, K, device=dev) / (K * T) ** 0.5).bfloat16()
    out = torch.empty(B, L, N, device=dev, dtype=torch.bfloat16)
    o32 = torch.empty(B, L, N, device=dev, dtype=torch.float32)
    R.gemm_fwd(x, w, N_out=N, K=K, taps=T, shift0=-(T // 2), shift_step=1, out_bf16=out, out_f32=o32, block_n=block_n)
    torch.cuda.synchronize()
    ref = conv_ref(x, w, -(T // 2))
    print(f"fwd  B{B} L{L} N{N} K{K} T{T} bn{block_n}: rel32={rel(o32, ref):.2e} rel16={rel(out, ref):.2e}", flush=True)


def case_bmn(B, L, N, K, T):
    # dgrad-like: B stored [T][K][N]
    x = torch.randn(B, L, K, device=dev).bfloat16()
    wt = (torch.randn(T, K, N, device=dev) / (K * T) ** 0.5).bfloat16()
    o32 = torch.empty(B, L, N, device=dev, dtype=torch.float32)
    R.gemm_fwd(x, wt, N_out=N, K=K, taps=T, shift0=(T // 2), shift_step=-1, b_mn_major=True, out_f32=o32)
    torch.cuda.synchronize()
    # reference: out[l] = sum_t x[l + T//2 - t] @ wt[t]
    xf = x.float()
    ref = torch.zeros_like(o32)
    for t in range(T):
        s = (T // 2) - t
        xs = torch.zeros_like(xf)
        lo, hi = max(0, -s), min(L, L - s)
        if hi > lo:
            xs[:, lo:hi] = xf[:, lo + s:hi + s]
        ref += xs @ wt[t].float()
    print(f"bmn  B{B} L{L} N{N} K{K} T{T}: rel32={rel(o32, ref):.2e}", flush=True)


def case_wgrad(B, L, M, N, T, split_k=0):
    dy = torch.randn(B, L, M, device=dev).bfloat16()
    x = torch.randn(B, L, N, device=dev).bfloat16()
    out = torch.zeros(T, M, N, device=dev, dtype=torch.float32)
    R.gemm_wgrad(dy, x, out, M=M, N_out=N, taps=T, shift0=-(T // 2), shift_step=1, split_k=split_k)
    torch.cuda.synchronize()
    ref = torch.zeros_like(out)
    xf = x.float()
    for t in range(T):
        s = -(T // 2) + t
        xs = torch.zeros_like(xf)
        lo, hi = max(0, -s), min(L, L - s)
        if hi > lo:
            xs[:, lo:hi] = xf[:, lo + s:hi + s]
        ref[t] = torch.einsum("blm,bln->mn", dy.float(), xs)
    print(f"wgrd B{B} L{L} M{M} N{N} T{T} sk{split_k}: rel32={rel(out, ref):.2e}", flush=True)


def case_epi():
    B, L, N, K = 2, 300, 136, 72
    x = torch.randn(B, L, K, device=dev).bfloat16()
    w = (torch.randn(1, N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    aux32 = torch.randn(B, L, N, device=dev)
    aux16 = torch.randn(B, L, N, device=dev).bfloat16()
    base = x.float() @ w[0].float().t() + bias
    # (c) bias + aux32 -> f32 + bf16
    o32 = torch.empty(B, L, N, device=dev)
    o16 = torch.empty(B, L, N, device=dev, dtype=torch.bfloat16)
    R.gemm_fwd(x, w, N_out=N, K=K, bias=bias, aux_f32=aux32, out_f32=o32, out_bf16=o16)
    print(f"epi bias+aux32: {rel(o32, base + aux32):.2e} {rel(o16, base + aux32):.2e}", flush=True)
    # (b) silu with pre-activation output + stats
    pre = torch.empty_like(o16)
    stats = torch.zeros(B, 2, device=dev, dtype=torch.float64)
    R.gemm_fwd(x, w, N_out=N, K=K, bias=bias, act=R.ACT_SILU, pre_bf16=pre, out_bf16=o16, stats=stats)
    ref = F.silu(base)
    rb = ref.bfloat16().double()
    sref = torch.stack([rb.sum((1, 2)), (rb * rb).sum((1, 2))], 1)
    print(f"epi silu: pre {rel(pre, base):.2e} out {rel(o16, ref):.2e} stats {rel(stats, sref):.2e}", flush=True)
    # (d) dsilu multiply
    R.gemm_fwd(x, w, N_out=N, K=K, aux_bf16=aux16, aux_is_dsilu=True, out_bf16=o16)
    a = aux16.float()
    s = torch.sigmoid(a)
    ref = (base - bias) * (s * (1 + a * (1 - s)))
    print(f"epi dsilu: {rel(o16, ref):.2e}", flush=True)
    # aux bf16 add, strided output (channel slice of a wider buffer)
    wide = torch.zeros(B, L, N + 64, device=dev, dtype=torch.bfloat16)
    R.gemm_fwd(x, w, N_out=N, K=K, aux_bf16=aux16, out_bf16=wide[:, :, 64:])
    print(f"epi aux16 + strided out: {rel(wide[:, :, 64:], base - bias + aux16.float()):.2e} "
          f"untouched={wide[:, :, :64].abs().max().item()}", flush=True)


def perf():
    for (B, L, N, K, T) in [(4, 4096, 512, 512, 3), (4, 4096, 1024, 512, 1), (4, 1024, 1024, 1024, 3), (4, 512, 2048, 2048, 3)]:
        x = torch.randn(B, L, K, device=dev).bfloat16()
        w = (torch.randn(T, N, K, device=dev) / (K * T) ** 0.5).bfloat16()
        out = torch.empty(B, L, N, device=dev, dtype=torch.bfloat16)
        for bn in (128, 256):
            for _ in range(3):
                R.gemm_fwd(x, w, N_out=N, K=K, taps=T, shift0=-(T // 2), shift_step=1, out_bf16=out, block_n=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                R.gemm_fwd(x, w, N_out=N, K=K, taps=T, shift0=-(T // 2), shift_step=1, out_bf16=out, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            fl = 2.0 * B * L * N * K * T
            print(f"perf B{B} L{L} N{N} K{K} T{T} bn{bn}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        # wgrad
        dy = torch.randn(B, L, N, device=dev).bfloat16()
        dw = torch.zeros(T, N, K, device=dev)
        for _ in range(3):
            R.gemm_wgrad(dy, x, dw, M=N, N_out=K, taps=T, shift0=-(T // 2), shift_step=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            R.gemm_wgrad(dy, x, dw, M=N, N_out=K, taps=T, shift0=-(T // 2), shift_step=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"perf wgrad M{N} N{K} T{T}: {ms * 1e3:.1f} us  {2.0 * B * L * N * K * T / ms / 1e9:.1f} TFLOP/s", flush=True)


def perfepi():
    """Epilogue-variant timing on the transformer-block shapes (which fused option costs what)."""
    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3
    for (B, L, N, K) in [(4, 4096, 512, 1024), (4, 4096, 1024, 512)]:
        x = torch.randn(B, L, K, device=dev).bfloat16()
        w = (torch.randn(1, N, K, device=dev) / K ** 0.5).bfloat16()
        wt = w.transpose(1, 2).contiguous()
        bias = torch.randn(N, device=dev)
        aux32 = torch.randn(B, L, N, device=dev)
        aux16 = torch.randn(B, L, N, device=dev).bfloat16()
        o16 = torch.empty(B, L, N, device=dev, dtype=torch.bfloat16)
        pre = torch.empty_like(o16)
        o32 = torch.empty(B, L, N, device=dev)
        fl = 2.0 * B * L * N * K
        variants = {
            "o16": dict(out_bf16=o16), "bias+o16": dict(bias=bias, out_bf16=o16), "o32": dict(out_f32=o32),
            "o32+o16": dict(out_f32=o32, out_bf16=o16), "aux32+o32": dict(aux_f32=aux32, out_f32=o32),
            "aux32+o32 inplace": dict(aux_f32=o32, out_f32=o32),
            "bias+aux32+o32+o16": dict(bias=bias, aux_f32=aux32, out_f32=o32, out_bf16=o16),
            "bias+silu+pre+o16": dict(bias=bias, act=R.ACT_SILU, pre_bf16=pre, out_bf16=o16),
            "aux16dsilu+o16": dict(aux_bf16=aux16, aux_is_dsilu=True, out_bf16=o16),
        }
        for name, kw in variants.items():
            us = timeit(lambda: R.gemm_fwd(x, w, N_out=N, K=K, **kw))
            us2 = timeit(lambda: R.gemm_fwd(x, wt, N_out=N, K=K, b_mn_major=True, **kw))
            print(f"perfepi B{B} L{L} N{N} K{K} {name:22s}: {us:6.1f} us {fl / us / 1e6:7.1f} TF/s | B mn-major {us2:6.1f} us", flush=True)


if __name__ == "__main__":
    grp = sys.argv[1] if len(sys.argv) > 1 else "all"
    t0 = time.time()
    if grp in ("fwd", "all"):
        case_fwd(1, 128, 64, 64, 1)
        case_fwd(1, 128, 256, 128, 1)
        case_fwd(2, 256, 512, 512, 1)
        case_fwd(2, 200, 96, 96, 3)
        case_fwd(3, 1000, 520, 264, 3)
        case_fwd(1, 16, 8, 8, 15)
        case_fwd(2, 4096, 512, 512, 3, block_n=128)
        case_fwd(2, 4096, 512, 512, 3, block_n=256)
    if grp in ("bmn", "all"):
        case_bmn(1, 128, 64, 64, 1)
        case_bmn(2, 200, 136, 96, 3)
        case_bmn(2, 1024, 512, 1024, 3)
    if grp in ("wgrad", "all"):
        case_wgrad(1, 64, 128, 64, 1, split_k=1)
        case_wgrad(1, 256, 128, 128, 1, split_k=1)
        case_wgrad(2, 200, 96, 136, 3)
        case_wgrad(4, 1024, 512, 512, 3)
    if grp in ("epi", "all"):
        case_epi()
    if grp in ("perf",):
        perf()
    if grp in ("perfepi",):
        perfepi()
    print(f"done {grp} in {time.time() - t0:.1f}s", flush=True)
