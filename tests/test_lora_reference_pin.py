"""Pins the LoRA / DoRA restatements (oracle/dora.py, osufusion_b200/lora.py) to the REFERENCE's own code.

`osu_fusion/modules/lora_layers.py` only fails to import here because of four `peft` imports (:8-11); tests/peft_stub.py supplies
the bookkeeping base classes, everything numerical is the reference's (`DoraConv1dLayer.forward` :59-92, `get_weight_norm` :16-26,
`LoraConv1d.get_delta_weight` :258-290, `merge` / `unmerge` :197-256, `forward` :292-328).  Two arms:
  * golden vectors produced by that code (oracle/make_golden_lora.py -> tests/golden/lora_conv1d_ref.pt) — run anywhere;
  * the live reference from /root/reference on fresh random inputs — in the build container only.
peft's own nn.Linear DoRA layer (attn.to_q / attn.to_kv) has no source in the image and stays "parity unpinned".
"""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from conftest import HAVE_REFERENCE
from oracle.dora import dora_conv1d
from osufusion_b200 import lora

GOLD = Path(__file__).parent / "golden" / "lora_conv1d_ref.pt"


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def oracle_forward(g, x, A, Bm, mag):
    d = g["dims"]
    if d["use_dora"]:
        return dora_conv1d(x, g["W"], g["bias"], A, Bm, mag, g["scaling"], d["k"] // 2)
    return F.conv1d(x, g["W"], g["bias"], padding=d["k"] // 2) + F.conv1d(F.conv1d(x, A, None, padding=d["k"] // 2), Bm) * g["scaling"]


def adapted_layer(g):
    d = g["dims"]
    conv = torch.nn.Conv1d(d["Cin"], d["Cout"], d["k"], padding=d["k"] // 2)
    with torch.no_grad():
        conv.weight.copy_(g["W"])
        conv.bias.copy_(g["bias"])
    ad = lora.AdaptedLayer(conv, d["r"], d["alpha"], d["use_dora"])
    return conv, ad


@pytest.mark.parametrize("name", ["conv3_dora", "conv3_lora", "conv1_dora", "conv3_dora_r32"])
def test_oracle_and_adapter_host_logic_match_reference_golden(name):
    g = torch.load(GOLD, weights_only=False)[name]
    d = g["dims"]
    x = g["x"].clone().requires_grad_(True)
    A, Bm = g["A"].clone().requires_grad_(True), g["B"].clone().requires_grad_(True)
    mag = g["mag"].clone().requires_grad_(True) if d["use_dora"] else None
    y = oracle_forward(g, x, A, Bm, mag)
    assert rel(y.detach(), g["y"]) < 1e-5
    y.backward(g["dy"])
    assert rel(x.grad, g["dx"]) < 1e-5 and rel(A.grad, g["dA"]) < 1e-5 and rel(Bm.grad, g["dB"]) < 1e-5
    if d["use_dora"]:
        assert rel(mag.grad, g["dmag"]) < 1e-5
    # the product's host-side adapter logic: init, delta weight, merge, unmerge
    conv, ad = adapted_layer(g)
    assert ad.scaling == g["scaling"]
    if d["use_dora"]:
        assert ad.magnitude().shape == g["mag_init"].shape and rel(ad.magnitude().detach(), g["mag_init"]) < 1e-6
    with torch.no_grad():
        ad.lora_A["default"].weight.copy_(g["A"])
        ad.lora_B["default"].weight.copy_(g["B"])
        if d["use_dora"]:
            ad.magnitude().copy_(g["mag"])
    assert rel(ad.delta_weight().detach(), g["delta"]) < 1e-5
    ad.merge()
    assert rel(conv.weight.detach(), g["W_merged"]) < 1e-5
    ad.unmerge()
    assert rel(conv.weight.detach(), g["W_unmerged"]) < 1e-5 and rel(conv.weight.detach(), g["W"]) < 1e-5


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("Cin,Cout,k,r,alpha,use_dora", [(12, 20, 3, 4, 8, True), (12, 20, 1, 4, 4, True), (20, 12, 3, 8, 8, False)])
def test_oracle_matches_live_reference_lora_conv1d(Cin, Cout, k, r, alpha, use_dora):
    import peft_stub
    ref = peft_stub.load_reference_lora_layers()
    torch.manual_seed(7)
    conv = torch.nn.Conv1d(Cin, Cout, k, padding=k // 2)
    layer = ref.LoraConv1d(conv, "default", r=r, lora_alpha=alpha, use_dora=use_dora)
    x0 = torch.randn(2, Cin, 40)
    assert torch.equal(layer(x0), conv(x0))                       # B = 0 at init: identity, DoRA scale exactly 1
    with torch.no_grad():
        layer.lora_B["default"].weight.normal_(std=0.05)
        if use_dora:
            layer.lora_magnitude_vector["default"].weight.mul_(1 + 0.1 * torch.randn(1, Cout, 1))
    g = {"dims": dict(k=k, use_dora=use_dora), "W": conv.weight.detach(), "bias": conv.bias.detach(), "scaling": layer.scaling["default"]}
    x = torch.randn(2, Cin, 40, requires_grad=True)
    dy = torch.randn(2, Cout, 40)
    y_ref = layer(x)
    y_ref.backward(dy)
    A, Bm = layer.lora_A["default"].weight, layer.lora_B["default"].weight
    m = layer.lora_magnitude_vector["default"].weight if use_dora else None
    x2 = x.detach().clone().requires_grad_(True)
    A2, B2 = A.detach().clone().requires_grad_(True), Bm.detach().clone().requires_grad_(True)
    m2 = m.detach().clone().requires_grad_(True) if use_dora else None
    y = oracle_forward(g, x2, A2, B2, m2)
    y.backward(dy)
    assert rel(y.detach(), y_ref.detach()) < 1e-6
    assert rel(x2.grad, x.grad) < 1e-5 and rel(A2.grad, A.grad) < 1e-5 and rel(B2.grad, Bm.grad) < 1e-5
    if use_dora:
        assert rel(m2.grad, m.grad) < 1e-5
    # merged weight of the reference == effective weight the engine feeds its single GEMM (s * (W + scaling B A))
    W0 = conv.weight.detach().clone()
    layer.merge()
    delta = (B2.detach().flatten(1) @ A2.detach().flatten(1)).reshape(W0.shape) * g["scaling"]
    if use_dora:
        s = (m2.detach() / (W0 + delta).norm(p=2, dim=(1, 2), keepdim=True).transpose(1, 0)).view(-1, 1, 1)
        assert rel(conv.weight.detach(), s * (W0 + delta)) < 1e-6
    else:
        assert rel(conv.weight.detach(), W0 + delta) < 1e-6
