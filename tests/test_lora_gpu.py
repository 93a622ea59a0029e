"""GPU parity of the fused LoRA/DoRA path (one GEMM on the effective weight + projected gradients) against the oracle's
restatement of the reference's three-term forward (lora_layers.py:59-92) under autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def nrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def _build(use_dora):
    from oracle.denoiser import UNet as OracleUNet
    from oracle.dora import dora_conv1d, dora_linear
    from oracle.synth import TINY
    from osufusion_b200 import lora
    from osufusion_b200.modules import UNet
    torch.manual_seed(0)
    ora = OracleUNet(6, 96, 5, **TINY)
    torch.nn.init.normal_(ora.final_conv.weight, std=0.02)
    new = UNet(6, 96, 5, **TINY)
    new.load_state_dict(ora.state_dict())
    ora, new = ora.to(dev), new.to(dev)
    names = lora.inject_adapters(new, r=8, lora_alpha=16, use_dora=use_dora)
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


@pytest.mark.parametrize("use_dora", [True, False])
def test_lora_dora_forward_backward(use_dora):
    from oracle.synth import synth_inputs
    ora, new, names, leaves = _build(use_dora)
    x, a, c, t, noise, keep = (v.to(dev) for v in synth_inputs(2, 120, 5))

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
