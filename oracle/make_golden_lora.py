"""ORACLE TOOLING (test infrastructure): golden vectors of the reference's OWN LoRA / DoRA Conv1d code.

Runs `/root/reference/osu_fusion/modules/lora_layers.py` (LoraConv1d + DoraConv1dLayer, :15-332) on CPU fp32 through the
peft bookkeeping stub of tests/peft_stub.py and stores inputs, adapter tensors, forward outputs, all gradients, the delta weight
and the merged weight in tests/golden/lora_conv1d_ref.pt, so the pin also holds where /root/reference is absent (GPU box).

    python -m oracle.make_golden_lora
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))

CASES = [  # name, Cin, Cout, k, r, alpha, use_dora, batch, length
    ("conv3_dora", 24, 40, 3, 8, 16, True, 2, 50),
    ("conv3_lora", 24, 40, 3, 8, 8, False, 2, 50),
    ("conv1_dora", 16, 32, 1, 4, 8, True, 3, 20),
    ("conv3_dora_r32", 64, 64, 3, 32, 32, True, 1, 33),
]


def run_case(mod, Cin, Cout, k, r, alpha, use_dora, B, L, seed):
    torch.manual_seed(seed)
    conv = torch.nn.Conv1d(Cin, Cout, k, padding=k // 2)
    layer = mod.LoraConv1d(conv, "default", r=r, lora_alpha=alpha, use_dora=use_dora)
    mag0 = layer.lora_magnitude_vector["default"].weight.detach().clone() if use_dora else None
    with torch.no_grad():
        layer.lora_B["default"].weight.normal_(std=0.05)
        if use_dora:
            layer.lora_magnitude_vector["default"].weight.mul_(1 + 0.1 * torch.randn(1, Cout, 1))
    x = torch.randn(B, Cin, L, requires_grad=True)
    dy = torch.randn(B, Cout, L)
    y = layer(x)
    y.backward(dy)
    A, Bm = layer.lora_A["default"].weight, layer.lora_B["default"].weight
    out = {
        "dims": dict(Cin=Cin, Cout=Cout, k=k, r=r, alpha=alpha, use_dora=use_dora),
        "W": conv.weight.detach().clone(), "bias": conv.bias.detach().clone(),
        "A": A.detach().clone(), "B": Bm.detach().clone(), "scaling": layer.scaling["default"],
        "x": x.detach().clone(), "dy": dy, "y": y.detach().clone(), "dx": x.grad.clone(),
        "dA": A.grad.clone(), "dB": Bm.grad.clone(),
        "delta": layer.get_delta_weight("default").detach().clone(),
        "mag_init": mag0,
    }
    if use_dora:
        m = layer.lora_magnitude_vector["default"].weight
        out["mag"], out["dmag"] = m.detach().clone(), m.grad.clone()
    layer.merge()
    out["W_merged"] = conv.weight.detach().clone()
    layer.unmerge()
    out["W_unmerged"] = conv.weight.detach().clone()
    return out


def main() -> None:
    import peft_stub
    mod = peft_stub.load_reference_lora_layers()
    gold = {name: run_case(mod, *rest, seed=100 + i) for i, (name, *rest) in enumerate(CASES)}
    path = ROOT / "tests" / "golden" / "lora_conv1d_ref.pt"
    torch.save(gold, path)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
