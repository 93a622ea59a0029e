"""Host-side adapter logic (no GPU): injection targets, peft-compatible names, init identities, merge/unmerge."""
import torch

from oracle.synth import TINY
from osufusion_b200 import lora
from osufusion_b200.modules import UNet


def test_injection_counts_names_and_freezing():
    net = UNet(6, 96, 5, **TINY)
    n_base = len(net.state_dict())
    adapted = lora.inject_adapters(net, r=8, lora_alpha=8, use_dora=True)
    convs = [a for a in adapted if a.endswith("proj")]
    lins = [a for a in adapted if "attn.to_" in a]
    n_res = sum(1 for _, m in net.named_modules() if type(m).__name__ == "ResidualBlock")
    n_tr = sum(1 for _, m in net.named_modules() if type(m).__name__ == "TransformerBlock")
    assert len(convs) == 2 * n_res and len(lins) == 2 * n_tr          # "attn.linear" matches nothing
    sd = net.state_dict()
    assert "down_layers.0.resnets.0.block1.proj.base_layer.weight" in sd
    assert "down_layers.0.resnets.0.block1.proj.lora_A.default.weight" in sd
    assert "down_layers.0.resnets.0.block1.proj.lora_magnitude_vector.default.weight" in sd
    assert sd["down_layers.0.resnets.0.block1.proj.lora_magnitude_vector.default.weight"].shape == (1, 96, 1)
    assert sd["down_layers.0.transformers.0.attn.to_q.lora_magnitude_vector.default.weight"].shape == (32,)
    assert all(p.requires_grad == ("lora_" in n) for n, p in net.named_parameters())
    ad = lora.adapter_state_dict(net)
    assert all(k.startswith("base_model.model.") and ".default" not in k for k in ad)
    assert len(ad) == 3 * len(adapted)
    lora.merge_and_unload(net)
    assert len(net.state_dict()) == n_base


def test_init_identities_and_merge_roundtrip():
    torch.manual_seed(0)
    net = UNet(6, 96, 5, **TINY)
    lora.inject_adapters(net, r=8, lora_alpha=8)
    m = net.down_layers[0].resnets[0].block1.proj
    W0 = m.base_layer.weight.detach().clone()
    assert torch.count_nonzero(m.delta_weight()) == 0                    # B = 0 at init: adapter is the identity
    assert torch.allclose(m.magnitude(), m.weight_norm())                # DoRA scale is exactly 1 at init
    with torch.no_grad():
        m.lora_B["default"].weight.normal_(std=0.05)
        m.magnitude().mul_(1.1)
    m.merge()
    assert not torch.allclose(m.base_layer.weight, W0)
    m.unmerge()
    assert torch.allclose(m.base_layer.weight, W0, atol=1e-6)            # merge then unmerge restores W
