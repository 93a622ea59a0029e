"""GPU parity of the fused LoRA/DoRA path (one GEMM on the effective weight + projected gradients) against the oracle's
restatement of the reference's three-term forward (lora_layers.py:59-92) under autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def nrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def _build(use_dora, cfg=None, r=8, alpha=16):
    from oracle.denoiser import UNet as OracleUNet
    from oracle.dora import dora_conv1d, dora_linear
    from oracle.synth import TINY
    from osufusion_b200 import lora
    from osufusion_b200.modules import UNet
    torch.manual_seed(0)
    cfg = TINY if cfg is None else cfg
    ora = OracleUNet(6, 96, 5, **cfg)
    torch.nn.init.normal_(ora.final_conv.weight, std=0.02)
    new = UNet(6, 96, 5, **cfg)
    new.load_state_dict(ora.state_dict())
    ora, new = ora.to(dev), new.to(dev)
    names = lora.inject_adapters(new, r=r, lora_alpha=alpha, use_dora=use_dora)
    for p in ora.parameters():
        p.requires_grad_(False)
    g = torch.Generator(device="cpu").manual_seed(1)
    leaves = {}
    for name in names:
        ad = new.get_submodule(name)
        with torch.no_grad():
            ad.lora_B["default"].weight.copy_(0.05 * torch.randn(ad.lora_B["default"].weight.shape, generator=g))
            if use_dora:
                ad.magnitude().mul_(1 + 0.1 * torch.randn(ad.magnitude().shape, generator=g).to(dev))
        A = ad.lora_A["default"].weight.detach().clone().requires_grad_(True)
        Bm = ad.lora_B["default"].weight.detach().clone().requires_grad_(True)
        mag = ad.magnitude().detach().clone().requires_grad_(True) if use_dora else None
        leaves[name] = (A, Bm, mag)
        om = ora.get_submodule(name)
        sc = ad.scaling
        if isinstance(om, torch.nn.Conv1d):
            def fwd(x, om=om, A=A, Bm=Bm, mag=mag, sc=sc):
                if mag is None:
                    return torch.nn.functional.conv1d(x, om.weight, om.bias, padding=1) + torch.nn.functional.conv1d(
                        torch.nn.functional.conv1d(x, A, None, padding=1), Bm) * sc
                return dora_conv1d(x, om.weight, om.bias, A, Bm, mag, sc, padding=om.padding[0])
        else:
            def fwd(x, om=om, A=A, Bm=Bm, mag=mag, sc=sc):
                if mag is None:
                    return torch.nn.functional.linear(x, om.weight, om.bias) + torch.nn.functional.linear(
                        torch.nn.functional.linear(x, A), Bm) * sc
                return dora_linear(x, om.weight, om.bias, A, Bm, mag, sc)
        om.forward = fwd
    return ora, new, names, leaves


def test_cfg_l_r32_dora_forward_backward():
    """The benchmarked adapter configuration (BASELINE.json configs[4], trainer_peft.py:236-244): CFG-L dim_h=512, r=32, alpha=32,
    DoRA on all 180 adapted modules (102 Conv1d + 78 Linear), base frozen; 1024 frames."""
    from oracle.synth import LARGE
    _run_whole_model(True, LARGE, 32, 32, 1, 1024)
    torch.cuda.empty_cache()


@pytest.mark.parametrize("use_dora", [True, False])
def test_lora_dora_forward_backward(use_dora):
    _run_whole_model(use_dora, None, 8, 16, 2, 120)


def _run_whole_model(use_dora, cfg, r, alpha, B, n):
    from oracle.synth import synth_inputs
    ora, new, names, leaves = _build(use_dora, cfg, r, alpha)
    if cfg is not None:
        assert len(names) == 180
    x, a, c, t, noise, keep = (v.to(dev) for v in synth_inputs(B, n, 5))

    def run_oracle(autocast):
        for A, Bm, mag in leaves.values():
            for v in (A, Bm, mag):
                if v is not None:
                    v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = ora(x, a, t, c, cond_mask=keep)
        torch.nn.functional.mse_loss(y.float(), noise).backward()
        return y.detach().float(), {n: tuple(None if v is None else v.grad.detach().clone() for v in leaves[n]) for n in names}

    y_ref, g_ref = run_oracle(True)
    y_tru, g_tru = run_oracle(False)
    new.zero_grad(set_to_none=True)
    y_new = new(x, a, t, c, cond_mask=keep)
    torch.nn.functional.mse_loss(y_new, noise).backward()
    assert nrel(y_new, y_tru) <= max(1e-2, 2 * nrel(y_ref, y_tru))
    assert all(p.grad is None for n, p in new.named_parameters() if "lora_" not in n)      # base weights stay frozen
    bad = []
    for n in names:
        ad = new.get_submodule(n)
        mine = (ad.lora_A["default"].weight.grad, ad.lora_B["default"].weight.grad, ad.magnitude().grad if use_dora else None)
        for which, gm, gr, gt in zip("ABm", mine, g_ref[n], g_tru[n]):
            if gt is None:
                continue
            e_new, e_ref = nrel(gm.view(gt.shape), gt), nrel(gr, gt)
            # Under DoRA the reference's type promotion keeps conv/linear outputs in fp32 (fp32 scale x bf16), so its own error
            # drops below the plain bf16-autocast level (~4e-2 on this config, see the use_dora=False arm); the engine keeps bf16
            # GEMM outputs, i.e. plain-autocast precision: allow that floor.
            if e_new > max(1e-2, 3 * e_ref, 6e-2 if use_dora else 0.0):
                bad.append((n, which, e_new, e_ref))
    assert not bad, bad[:6]


@pytest.mark.parametrize("Cout,Cin,k,r", [(96, 96, 3, 8), (40, 24, 1, 4), (512, 512, 3, 32)])
def test_dora_merge_and_grad_kernels(Cout, Cin, k, r):
    from osufusion_b200 import _native as N
    torch.manual_seed(0)
    W = torch.randn(Cout, Cin, k, device=dev) / (Cin * k) ** 0.5
    A = torch.randn(r, Cin, k, device=dev) / (Cin * k) ** 0.5
    Bm = 0.05 * torch.randn(Cout, r, device=dev)
    sc = 2.0
    v = W + sc * (Bm @ A.flatten(1)).reshape(W.shape)
    n = v.flatten(1).norm(dim=1)
    mag = n * (1 + 0.1 * torch.randn(Cout, device=dev))
    s = mag / n
    cp = (Cin + 7) // 8 * 8
    packed = torch.zeros(k, Cout, cp, device=dev, dtype=torch.bfloat16)
    n2 = torch.empty(Cout, device=dev)
    s_out = torch.empty(Cout, device=dev)
    N.call("of_dora_merge", W.data_ptr(), A.data_ptr(), Bm.data_ptr(), mag.data_ptr(), sc, Cout, Cin, k, r, n2.data_ptr(),
           packed.data_ptr(), cp, Cout * cp, s_out.data_ptr())
    ref = (s[:, None, None] * v).permute(2, 0, 1)
    assert nrel(packed[:, :, :Cin], ref) < 1e-2 and nrel(n2, n * n) < 1e-4 and nrel(s_out, s) < 1e-4
    dWp = torch.randn(k, Cout, cp, device=dev)
    dA, dB, dm = torch.zeros_like(A), torch.zeros_like(Bm), torch.zeros(Cout, device=dev)
    N.call("of_dora_grad", W.data_ptr(), A.data_ptr(), Bm.data_ptr(), mag.data_ptr(), sc, Cout, Cin, k, r, n2.data_ptr(),
           dWp.data_ptr(), cp, Cout * cp, dA.data_ptr(), dB.data_ptr(), dm.data_ptr())
    dW = dWp[:, :, :Cin].permute(1, 2, 0)                # (Cout, Cin, k)
    G = sc * s[:, None, None] * dW
    assert nrel(dB, G.flatten(1) @ A.flatten(1).t()) < 1e-4
    assert nrel(dA, (Bm.t() @ G.flatten(1)).reshape(A.shape)) < 1e-4
    assert nrel(dm, (dW * v).flatten(1).sum(1) / n) < 1e-4


# ------------------------------------------------------------------------------------------------ the DEFAULT path's kernels
@pytest.mark.parametrize("name", ["conv3_dora", "conv1_dora", "conv3_dora_r32"])
def test_effective_weight_matches_reference_merged_weight_golden(name):
    """W_eff = s (.) (W + scaling B A) built by the engine (tensor-core merge GEMM + of_dora_scale_pack, and the CUDA-core
    of_dora_merge) equals the weight the REFERENCE's own `LoraConv1d.merge` produces (lora_layers.py:197-238; golden vectors from
    oracle/make_golden_lora.py) up to the bf16 rounding of the GEMM operand."""
    from pathlib import Path

    from osufusion_b200 import _native as N
    from osufusion_b200 import ops_raw as R
    g = torch.load(Path(__file__).parent / "golden" / "lora_conv1d_ref.pt", weights_only=False)[name]
    d = g["dims"]
    Cout, Cin, k, r, sc = d["Cout"], d["Cin"], d["k"], d["r"], g["scaling"]
    W, A, Bm, mag = (g[x].to(dev).contiguous() for x in ("W", "A", "B", "mag"))
    E = Cin * k
    ref = g["W_merged"].to(dev).permute(2, 0, 1)                         # [k][Cout][Cin]
    # CUDA-core merge
    packed = torch.zeros(k, Cout, Cin, device=dev, dtype=torch.bfloat16)
    n2 = torch.empty(Cout, device=dev)
    N.call("of_dora_merge", W.data_ptr(), A.data_ptr(), Bm.data_ptr(), mag.data_ptr(), sc, Cout, Cin, k, r, n2.data_ptr(),
           packed.data_ptr(), Cin, Cout * Cin, None)
    assert nrel(packed, ref) < 1e-2
    if r % 8 or E % 8:
        return          # the engine takes the tensor-core path only for r, Cin*k multiples of 8 (engine.ParamStore._dora_merge_into)
    # tensor-core merge: V = W + (scaling B) A as a K = r GEMM, then norm / scale / pack per output channel
    A16 = torch.empty(r, E, device=dev, dtype=torch.bfloat16)
    N.call("of_cast_f32_bf16", A.data_ptr(), A16.data_ptr(), r * E)
    B16 = torch.empty(Cout, r, device=dev, dtype=torch.bfloat16)
    N.call("of_scale_cast_f32_bf16", Bm.data_ptr(), float(sc), B16.data_ptr(), Cout * r)
    assert nrel(B16, sc * Bm.view(Cout, r)) < 1e-2
    V = torch.empty(1, Cout, E, device=dev)
    R.gemm_fwd(B16.view(1, Cout, r), A16.view(1, r, E), N_out=E, K=r, b_mn_major=True, aux_f32=W.view(1, Cout, E), out_f32=V)
    v_ref = W.view(Cout, E) + sc * (Bm.view(Cout, r) @ A.view(r, E))
    assert nrel(V[0], v_ref) < 2e-3                                       # bf16 A, B operands, fp32 accumulate on top of the fp32 W
    packed2 = torch.zeros(k, Cout, Cin, device=dev, dtype=torch.bfloat16)
    n2b = torch.empty(Cout, device=dev)
    N.call("of_dora_scale_pack", V.data_ptr(), mag.data_ptr(), Cout, Cin, k, n2b.data_ptr(), packed2.data_ptr(), Cin, Cout * Cin)
    assert nrel(packed2, ref) < 1e-2 and nrel(n2b, (v_ref * v_ref).sum(1)) < 2e-3 and nrel(n2, (v_ref * v_ref).sum(1)) < 1e-4


@pytest.mark.parametrize("Cout,r,with_mag", [(96, 8, True), (512, 32, True), (40, 4, False)])
def test_rankr_glue_kernels(Cout, r, with_mag):
    """of_dora_rankr_prep / of_dora_rankr_finish (engine._adapter_backward_rank_r) against their definitions in the header."""
    from osufusion_b200 import _native as N
    torch.manual_seed(1)
    Bm = 0.05 * torch.randn(Cout, r, device=dev)
    mag = (1 + 0.1 * torch.randn(Cout, device=dev)).abs() if with_mag else None
    n2 = (1 + 0.2 * torch.rand(Cout, device=dev)) if with_mag else None
    sc = 2.0
    Bst = torch.empty(r, Cout, device=dev, dtype=torch.bfloat16)
    rowscale = torch.empty(Cout, device=dev)
    N.call("of_dora_rankr_prep", Bm.data_ptr(), N.ptr(mag), N.ptr(n2), sc, Cout, r, Bst.data_ptr(), rowscale.data_ptr())
    rs = sc * (mag / n2.sqrt() if with_mag else torch.ones(Cout, device=dev))
    assert nrel(rowscale, rs) < 1e-6 and nrel(Bst, (rs[:, None] * Bm).t()) < 1e-2
    dBraw = torch.randn(Cout, r, device=dev)
    gB = torch.randn(Cout, r, device=dev)
    gB0 = gB.clone()
    dm = torch.randn(Cout, device=dev) if with_mag else None
    gmag = torch.randn(Cout, device=dev) if with_mag else None
    gmag0 = gmag.clone() if with_mag else None
    N.call("of_dora_rankr_finish", dBraw.data_ptr(), rowscale.data_ptr(), gB.data_ptr(), N.ptr(dm), N.ptr(mag), N.ptr(gmag), Cout, r)
    assert nrel(gB, gB0 + rs[:, None] * dBraw) < 1e-6
    if with_mag:
        assert nrel(gmag, gmag0 + dm / mag) < 1e-6


@pytest.mark.parametrize("rows,Nn,ld_pad,bias", [(300, 96, 0, True), (4096, 512, 640, True), (1000, 40, 8, False)])
def test_coldot_bf16_kernel(rows, Nn, ld_pad, bias):
    """of_coldot_bf16: out[n] += sum_rows dy[row, n] * (y[row, n] - bias[n]) on strided bf16 views (the DoRA magnitude gradient
    taken from the saved layer output, lora_layers.py:76-90 with the norm detached)."""
    from osufusion_b200 import _native as N
    torch.manual_seed(2)
    dy = torch.randn(rows, Nn + ld_pad, device=dev).bfloat16()
    y = torch.randn(rows, Nn + ld_pad, device=dev).bfloat16()
    b = torch.randn(Nn, device=dev) if bias else None
    out = torch.randn(Nn, device=dev)
    out0 = out.clone()
    N.call("of_coldot_bf16", dy.data_ptr(), Nn + ld_pad, y.data_ptr(), Nn + ld_pad, rows, Nn, N.ptr(b), out.data_ptr())
    yy = y[:, :Nn].float() - (b if bias else 0.0)
    ref = out0 + (dy[:, :Nn].float() * yy).sum(0)
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 1e-4
