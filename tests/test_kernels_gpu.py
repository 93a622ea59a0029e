"""GPU parity of the individual kernel families through the C-ABI, against torch fp32 math on the same inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def shifted(xf, s):
    L = xf.shape[1]
    out = torch.zeros_like(xf)
    lo, hi = max(0, -s), min(L, L - s)
    if hi > lo:
        out[:, lo:hi] = xf[:, lo + s:hi + s]
    return out


@pytest.mark.parametrize("B,L,N,K,T", [(1, 128, 64, 64, 1), (2, 200, 96, 96, 3), (3, 1000, 520, 264, 3), (1, 16, 8, 8, 15),
                                       (2, 4096, 512, 512, 3)])
def test_gemm_conv_forward(B, L, N, K, T):
    from osufusion_b200 import ops_raw as R
    torch.manual_seed(0)
    x = torch.randn(B, L, K, device=dev).bfloat16()
    w = (torch.randn(T, N, K, device=dev) / (K * T) ** 0.5).bfloat16()
    o32 = torch.empty(B, L, N, device=dev)
    o16 = torch.empty(B, L, N, device=dev, dtype=torch.bfloat16)
    R.gemm_fwd(x, w, N_out=N, K=K, taps=T, shift0=-(T // 2), shift_step=1, out_f32=o32, out_bf16=o16)
    ref = sum(shifted(x.float(), -(T // 2) + t) @ w[t].float().t() for t in range(T))
    assert rel(o32, ref) < 1e-4          # fp32 path tolerance (north_star: <= 1e-4)
    assert rel(o16, ref) < 1e-2          # bf16 output rounding


@pytest.mark.parametrize("B,L,N,K,T", [(1, 128, 64, 64, 1), (2, 200, 136, 96, 3), (2, 1024, 512, 1024, 3)])
def test_gemm_dgrad_mn_major(B, L, N, K, T):
    from osufusion_b200 import ops_raw as R
    torch.manual_seed(1)
    x = torch.randn(B, L, K, device=dev).bfloat16()
    wt = (torch.randn(T, K, N, device=dev) / (K * T) ** 0.5).bfloat16()
    o32 = torch.empty(B, L, N, device=dev)
    R.gemm_fwd(x, wt, N_out=N, K=K, taps=T, shift0=T // 2, shift_step=-1, b_mn_major=True, out_f32=o32)
    ref = sum(shifted(x.float(), T // 2 - t) @ wt[t].float() for t in range(T))
    assert rel(o32, ref) < 1e-4


@pytest.mark.parametrize("B,L,M,N,T", [(1, 64, 128, 64, 1), (2, 200, 96, 136, 3), (4, 1024, 512, 512, 3)])
def test_gemm_wgrad_splitk(B, L, M, N, T):
    from osufusion_b200 import ops_raw as R
    torch.manual_seed(2)
    dy = torch.randn(B, L, M, device=dev).bfloat16()
    x = torch.randn(B, L, N, device=dev).bfloat16()
    out = torch.zeros(T, M, N, device=dev)
    R.gemm_wgrad(dy, x, out, M=M, N_out=N, taps=T, shift0=-(T // 2), shift_step=1)
    ref = torch.stack([torch.einsum("blm,bln->mn", dy.float(), shifted(x.float(), -(T // 2) + t)) for t in range(T)])
    assert rel(out, ref) < 1e-4


def test_gemm_epilogues():
    from osufusion_b200 import ops_raw as R
    torch.manual_seed(3)
    B, L, N, K = 2, 300, 136, 72
    x = torch.randn(B, L, K, device=dev).bfloat16()
    w = (torch.randn(1, N, K, device=dev) / K ** 0.5).bfloat16()
    bias, aux32 = torch.randn(N, device=dev), torch.randn(B, L, N, device=dev)
    aux16 = torch.randn(B, L, N, device=dev).bfloat16()
    base = x.float() @ w[0].float().t() + bias
    o32 = torch.empty(B, L, N, device=dev)
    o16 = torch.empty(B, L, N, device=dev, dtype=torch.bfloat16)
    R.gemm_fwd(x, w, N_out=N, K=K, bias=bias, aux_f32=aux32, out_f32=o32, out_bf16=o16)
    assert rel(o32, base + aux32) < 1e-4
    pre = torch.empty_like(o16)
    stats = torch.zeros(B, 2, device=dev, dtype=torch.float64)
    R.gemm_fwd(x, w, N_out=N, K=K, bias=bias, act=R.ACT_SILU, pre_bf16=pre, out_bf16=o16, stats=stats)
    ref = torch.nn.functional.silu(base)
    rb = ref.bfloat16().double()
    assert rel(pre, base) < 1e-2 and rel(o16, ref) < 1e-2
    assert rel(stats, torch.stack([rb.sum((1, 2)), (rb * rb).sum((1, 2))], 1)) < 1e-3
    R.gemm_fwd(x, w, N_out=N, K=K, aux_bf16=aux16, aux_is_dsilu=True, out_bf16=o16)
    a = aux16.float()
    s = torch.sigmoid(a)
    assert rel(o16, (base - bias) * (s * (1 + a * (1 - s)))) < 1e-2
    wide = torch.zeros(B, L, N + 64, device=dev, dtype=torch.bfloat16)
    R.gemm_fwd(x, w, N_out=N, K=K, aux_bf16=aux16, out_bf16=wide[:, :, 64:])
    assert rel(wide[:, :, 64:], base - bias + aux16.float()) < 1e-2 and wide[:, :, :64].abs().max() == 0


def ref_attn(q, k, v, H, KVH, D):
    B, L, _ = q.shape
    qh = q.float().view(B, L, H, D).transpose(1, 2)
    kh = k.float().view(B, L, KVH, D).transpose(1, 2).repeat(1, H // KVH, 1, 1)
    vh = v.float().view(B, L, KVH, D).transpose(1, 2).repeat(1, H // KVH, 1, 1)
    s = (qh @ kh.transpose(-1, -2)) / D ** 0.5
    o = s.softmax(-1) @ vh
    return o.transpose(1, 2).reshape(B, L, H * D), torch.logsumexp(s, -1) * 1.4426950408889634


@pytest.mark.parametrize("variant", [0, 2, 10, 14, 22, 30, 32, 34])   # default / round-1 schedule / alternation / + polynomial exp2 / 2 CTAs per SM
@pytest.mark.parametrize("B,L,H,KVH,D,qs", [(1, 128, 1, 1, 64, 1.0), (2, 1024, 16, 1, 64, 1.0), (2, 200, 2, 1, 16, 1.0), (1, 72, 4, 2, 32, 1.0),
                                            (1, 512, 4, 1, 64, 6.0)])
def test_attention_forward(B, L, H, KVH, D, qs, variant):
    from osufusion_b200 import ops_raw as R
    torch.manual_seed(4)
    qkv = (torch.randn(B, L, (H + 2 * KVH) * D, device=dev) * qs).bfloat16()
    q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:(H + KVH) * D], qkv[:, :, (H + KVH) * D:]
    out = torch.zeros(B, L, H * D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=dev)
    R.attn_fwd(q, k, v, out, lse, H=H, KVH=KVH, D=D, variant=variant)
    o_ref, lse_ref = ref_attn(q, k, v, H, KVH, D)
    assert rel(out, o_ref) < 1e-2      # bf16 P and output (reference SDPA is bf16 as well)
    assert rel(lse, lse_ref) < 1e-4


@pytest.mark.parametrize("zero_grads", [False, True])     # True: the delta pre-pass zero-fills dq / dk / dv (buffers start as garbage)
@pytest.mark.parametrize("B,L,H,KVH,D", [(1, 128, 1, 1, 64), (2, 1024, 16, 1, 64), (2, 200, 2, 1, 16), (1, 72, 4, 2, 32)])
def test_attention_backward(B, L, H, KVH, D, zero_grads):
    from osufusion_b200 import ops_raw as R
    torch.manual_seed(5)
    qkv = torch.randn(B, L, (H + 2 * KVH) * D, device=dev).bfloat16()
    q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:(H + KVH) * D], qkv[:, :, (H + KVH) * D:]
    out = torch.zeros(B, L, H * D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=dev)
    R.attn_fwd(q, k, v, out, lse, H=H, KVH=KVH, D=D)
    dout = torch.randn(B, L, H * D, device=dev).bfloat16()
    delta = torch.zeros(B, H, L, device=dev)
    fill = 7.5 if zero_grads else 0.0
    dq = torch.full((B, L, H * D), fill, device=dev)
    dkv = torch.full((B, L, 2 * KVH * D), fill, device=dev)
    R.attn_bwd(q, k, v, out, lse, dout, delta, dq, dkv[:, :, :KVH * D], dkv[:, :, KVH * D:], H=H, KVH=KVH, D=D, zero_grads=zero_grads)
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = ref_attn(qf, kf, vf, H, KVH, D)
    o_ref.backward(dout.float())
    assert rel(dq, qf.grad) < 1e-2 and rel(dkv[:, :, :KVH * D], kf.grad) < 1e-2 and rel(dkv[:, :, KVH * D:], vf.grad) < 1e-2


def test_layernorm_and_small_linear():
    from osufusion_b200 import _native as N
    from osufusion_b200 import engine as E
    torch.manual_seed(6)
    rows, C = 300, 520
    x = torch.randn(rows, C, device=dev) * 2 + 0.3
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    o32 = torch.empty(rows, C, device=dev)
    o16 = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    mr = torch.empty(rows, 2, device=dev)
    N.call("of_layernorm_fwd", x.data_ptr(), C, rows, C, g.data_ptr(), b.data_ptr(), 1e-5, o32.data_ptr(), o16.data_ptr(), C, mr.data_ptr())
    ref = torch.nn.functional.layer_norm(x, (C,), g, b)
    assert rel(o32, ref) < 1e-4
    dy = torch.randn(rows, C, device=dev)
    dx = torch.empty(rows, C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    N.call("of_layernorm_bwd", dy.data_ptr(), C, x.data_ptr(), C, rows, C, g.data_ptr(), mr.data_ptr(), dx.data_ptr(), None, C,
           dg.data_ptr(), db.data_ptr())
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (C,), gr, br).backward(dy)
    assert rel(dx, xr.grad) < 1e-4 and rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    # small-M linear (bf16-rounded operands, fp32 accumulate) forward + backward
    M, K, Nn = 4, 520, 96
    xs = torch.randn(M, K, device=dev)
    W = torch.randn(Nn, K, device=dev) / K ** 0.5
    bias = torch.randn(Nn, device=dev)
    y, pre = E.linear_small_fwd(xs, W, bias, act=1, want_pre=True)
    xb, Wb = xs.bfloat16().float().requires_grad_(True), W.bfloat16().float().requires_grad_(True)
    pr = (xb @ Wb.t() + bias).bfloat16().float()
    assert rel(pre, pr) < 5e-3 and rel(y, torch.nn.functional.silu(pr)) < 5e-3
    dyl = torch.randn(M, Nn, device=dev)
    dW, dbias, dxs = torch.zeros_like(W), torch.zeros(Nn, device=dev), torch.zeros(M, K, device=dev)
    E.linear_small_bwd(dyl, pre, 1, xs, W, dW, dbias, dxs)
    pre_r = (xb @ Wb.t() + bias)
    torch.nn.functional.silu(pre_r).backward(dyl)
    assert rel(dW, Wb.grad) < 5e-3 and rel(dxs, xb.grad) < 5e-3


@pytest.mark.parametrize("M,K,Ns", [(4, 512, (192, 64, 256)), (2, 768, (384,)), (9, 256, (128, 132))])
def test_grouped_film_heads(M, K, Ns):
    """of_film_fwd / of_film_bwd (all FiLM heads in one launch) vs torch on bf16-rounded operands."""
    from osufusion_b200 import _native as N
    torch.manual_seed(3)
    lib = N.lib()
    Ws = [torch.randn(n, K, device=dev) / K ** 0.5 for n in Ns]
    bs = [torch.randn(n, device=dev) for n in Ns]
    dWs = [torch.full((n, K), 7.0, device=dev) for n in Ns]       # must be overwritten, not accumulated
    dbs = [torch.full((n,), 7.0, device=dev) for n in Ns]
    x = torch.randn(M, K, device=dev).bfloat16().float()
    groups, chunks, row = [], [], 0
    ch = lib.of_film_chunk_rows()
    for gi, n in enumerate(Ns):
        groups.append(N.FilmGroup(Ws[gi].data_ptr(), bs[gi].data_ptr(), dWs[gi].data_ptr(), dbs[gi].data_ptr(), M * row, n, row))
        for n0 in range(0, n, ch):
            chunks += [gi, n0]
        row += n
    gt = torch.frombuffer(bytearray(bytes((N.FilmGroup * len(groups))(*groups))), dtype=torch.uint8).to(dev)
    ct = torch.tensor(chunks, dtype=torch.int32, device=dev)
    out = torch.empty(M * row, device=dev)
    N.call("of_film_fwd", gt.data_ptr(), len(groups), row, x.data_ptr(), M, K, out.data_ptr())
    dss = torch.randn(M * row, device=dev)
    demb = torch.zeros(M, K, device=dev)
    N.call("of_film_bwd", gt.data_ptr(), ct.data_ptr(), len(chunks) // 2, dss.data_ptr(), x.data_ptr(), M, K, demb.data_ptr())
    off, demb_ref = 0, torch.zeros(M, K, device=dev)
    for gi, n in enumerate(Ns):
        w16 = Ws[gi].bfloat16().float()
        ref = (x @ w16.t() + bs[gi]).bfloat16().float()
        assert rel(out[off:off + M * n].view(M, n), ref) < 1e-2 and rel(out[off:off + M * n].view(M, n), x @ w16.t() + bs[gi]) < 1e-2
        d = dss[off:off + M * n].view(M, n)
        assert rel(dWs[gi], d.t() @ x) < 1e-4
        assert rel(dbs[gi], d.sum(0)) < 1e-4
        demb_ref += d @ w16
        off += M * n
    assert rel(demb, demb_ref) < 1e-4


def test_grouped_weight_pack():
    """of_pack_weights: conv (Cout,Cin,k) -> [k][Cout][cin_pad] bf16 and plain casts, many tensors in one launch."""
    from osufusion_b200 import _native as N
    torch.manual_seed(4)
    lib = N.lib()
    shapes = [(96, 96, 3), (520, 264, 1), (64, 100, 3), (1024, 2048, 1), (7, 13, 1), (130, 2051, 3)]
    ws = [torch.randn(s, device=dev) for s in shapes]
    segs, outs, cta = [], [], 0
    for w in ws:
        Cout, Cin, k = w.shape
        cp = Cin if k == 1 else (Cin + 7) // 8 * 8
        o = torch.full((k, Cout, cp), 9.0, device=dev, dtype=torch.bfloat16)
        outs.append(o)
        segs.append(N.PackSeg(w.data_ptr(), o.data_ptr(), Cout, Cin, k, cp, cta, 0))
        cta += lib.of_pack_seg_ctas(Cout, Cin, k, cp)
    table = torch.frombuffer(bytearray(bytes((N.PackSeg * len(segs))(*segs))), dtype=torch.uint8).to(dev)
    N.call("of_pack_weights", table.data_ptr(), len(segs), cta)
    for w, o in zip(ws, outs):
        Cout, Cin, k = w.shape
        ref = torch.zeros_like(o)
        ref[:, :, :Cin] = w.permute(2, 0, 1).bfloat16()
        assert torch.equal(o, ref), w.shape
