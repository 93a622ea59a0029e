"""LoRA / DoRA adapters for the B200 denoiser — stand-in for `peft.get_peft_model(model, LoraConfig(...))` as used by
trainer_peft.py:236-244 together with the reference's `LoraConv1d` / `DoraConv1dLayer` (lora_layers.py:15-332).

peft is not installed in this image, so the injector below reproduces the parts the reference relies on: suffix matching of
`target_modules`, the wrapper's sub-module / parameter names (`base_layer`, `lora_A.default`, `lora_B.default`,
`lora_magnitude_vector.default.weight`), initialisation (A: kaiming-uniform a=sqrt(5), B: zeros, magnitude = ||W||),
freezing of every non-adapter parameter, adapter-only state dicts with peft's key names, and merge / unmerge.
The forward/backward of adapted layers runs in the CUDA engine as ONE GEMM on the effective weight (csrc/dora.cu).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List

import torch
from torch import nn

DEFAULT_TARGETS = ["attn.to_q", "attn.to_kv", "attn.linear", "block1.proj", "block2.proj"]  # trainer_peft.py:241


class _Magnitude(nn.Module):
    def __init__(self, weight: torch.Tensor) -> None:
        super().__init__()
        self.weight = nn.Parameter(weight, requires_grad=True)


class AdaptedLayer(nn.Module):
    """Wrapper placed where the nn.Conv1d / nn.Linear used to be (same position in the module tree as peft's LoraLayer)."""

    def __init__(self, base_layer: nn.Module, r: int, lora_alpha: int, use_dora: bool) -> None:
        super().__init__()
        self.base_layer = base_layer
        self.r, self.lora_alpha, self.use_dora = r, lora_alpha, use_dora
        self.scaling = lora_alpha / r
        self.is_conv = isinstance(base_layer, nn.Conv1d)
        if self.is_conv:
            k, stride, pad = base_layer.kernel_size[0], base_layer.stride[0], base_layer.padding[0]
            A = nn.Conv1d(base_layer.in_channels, r, k, stride=stride, padding=pad, bias=False)   # lora_layers.py:151-158
            Bm = nn.Conv1d(r, base_layer.out_channels, 1, stride=1, bias=False)                    # lora_layers.py:159
        else:
            A = nn.Linear(base_layer.in_features, r, bias=False)
            Bm = nn.Linear(r, base_layer.out_features, bias=False)
        nn.init.kaiming_uniform_(A.weight, a=math.sqrt(5))     # peft reset_lora_parameters
        nn.init.zeros_(Bm.weight)
        dev, dt = base_layer.weight.device, base_layer.weight.dtype
        self.lora_A = nn.ModuleDict({"default": A.to(device=dev, dtype=dt)})
        self.lora_B = nn.ModuleDict({"default": Bm.to(device=dev, dtype=dt)})
        self.lora_magnitude_vector = nn.ModuleDict()
        if use_dora:
            self.lora_magnitude_vector["default"] = _Magnitude(self.weight_norm().detach().clone())
        self.merged = False
        self._cached_norm = None

    # ---- helpers (host-side, small tensors)
    def delta_weight(self) -> torch.Tensor:
        """scaling * B A in the base weight's shape (lora_layers.py:258-290)."""
        A, Bm = self.lora_A["default"].weight, self.lora_B["default"].weight
        d = (Bm.flatten(1) @ A.flatten(1)).reshape(self.base_layer.weight.shape)
        return d * self.scaling

    def weight_norm(self) -> torch.Tensor:
        """||W + scaling*BA||_2 per output channel; shape (1, Cout, 1) for Conv1d (lora_layers.py:16-26), (Cout,) for Linear."""
        w = self.base_layer.weight + self.delta_weight()
        if self.is_conv:
            return w.norm(p=2, dim=(1, 2), keepdim=True).transpose(1, 0)
        return w.norm(p=2, dim=1)

    def magnitude(self):
        return self.lora_magnitude_vector["default"].weight if self.use_dora else None

    def merge(self) -> None:
        """Fold the adapter into the base weight (lora_layers.py:197-238)."""
        if self.merged:
            return
        with torch.no_grad():
            W = self.base_layer.weight
            delta = self.delta_weight()
            if self.use_dora:
                norm = self.weight_norm().detach()
                self._cached_norm = norm
                factor = (self.magnitude() / norm).view(-1, *([1] * (W.dim() - 1)))
                W.copy_(factor * (W + delta))
            else:
                W.add_(delta)
        self.merged = True

    def unmerge(self) -> None:
        """lora_layers.py:240-256."""
        if not self.merged:
            return
        with torch.no_grad():
            W = self.base_layer.weight
            delta = self.delta_weight()
            if self.use_dora:
                factor = (self.magnitude() / self._cached_norm).view(-1, *([1] * (W.dim() - 1)))
                W.copy_(W / factor - delta)
            else:
                W.sub_(delta)
        self.merged = False

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("AdaptedLayer is executed by the CUDA engine (osufusion_b200.engine), not by torch")


def _matches(name: str, targets: Iterable[str]) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)   # peft's suffix rule


def inject_adapters(model: nn.Module, r: int = 32, lora_alpha: int = 32, use_dora: bool = True,
                    target_modules: Iterable[str] = DEFAULT_TARGETS) -> List[str]:
    """Wrap every nn.Conv1d / nn.Linear whose qualified name matches `target_modules`, freeze everything else.
    Returns the list of adapted module names (102 Conv1d + 78 Linear with the trainer_peft.py defaults)."""
    targets = list(target_modules)
    adapted = []
    for name, module in list(model.named_modules()):
        if not isinstance(module, (nn.Conv1d, nn.Linear)) or not _matches(name, targets):
            continue
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        wrapper = AdaptedLayer(module, r, lora_alpha, use_dora)
        if isinstance(parent, (nn.Sequential, nn.ModuleList)) and leaf.isdigit():
            parent[int(leaf)] = wrapper
        else:
            setattr(parent, leaf, wrapper)
        adapted.append(name)
    for n, p in model.named_parameters():
        p.requires_grad_("lora_" in n)
    return adapted


def adapter_state_dict(model: nn.Module, prefix: str = "base_model.model.") -> Dict[str, torch.Tensor]:
    """Adapter-only tensors with peft's `save_pretrained` key names (adapter name "default" stripped)."""
    out = {}
    for k, v in model.state_dict().items():
        if "lora_" in k:
            out[prefix + k.replace(".default", "")] = v.detach().clone()
    return out


def load_adapter_state_dict(model: nn.Module, sd: Dict[str, torch.Tensor], prefix: str = "base_model.model.") -> None:
    own = model.state_dict()
    for k, v in sd.items():
        kk = k[len(prefix):] if k.startswith(prefix) else k
        for part in ("lora_A", "lora_B", "lora_magnitude_vector"):
            kk = kk.replace(f"{part}.weight", f"{part}.default.weight")
        own[kk].copy_(v)


def merge_and_unload(model: nn.Module) -> nn.Module:
    """peft `merge_and_unload` (trainer_peft.py:161-164): fold adapters into the base weights and restore plain layers, so
    the resulting state_dict has the reference's 1239 keys again."""
    for name, module in list(model.named_modules()):
        if isinstance(module, AdaptedLayer):
            module.merge()
            parent_name, _, leaf = name.rpartition(".")
            parent = model.get_submodule(parent_name) if parent_name else model
            if isinstance(parent, (nn.Sequential, nn.ModuleList)) and leaf.isdigit():
                parent[int(leaf)] = module.base_layer
            else:
                setattr(parent, leaf, module.base_layer)
    return model
