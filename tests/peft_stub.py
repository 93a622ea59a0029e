"""TEST INFRASTRUCTURE: a minimal stand-in for the four `peft` imports of the reference's
`osu_fusion/modules/lora_layers.py:8-11` so that the reference's OWN DoRA code can be imported and executed in this image
(peft 0.12.0 is pinned in the reference's requirements.txt but not installed here, and there is no network).

What is stubbed is only peft's bookkeeping base classes (adapter dictionaries, `reset_lora_parameters`, merge caches) —
restated from peft 0.12.0's published behaviour.  Everything numerical that the pin tests exercise is the REFERENCE's code:
`DoraConv1dLayer.get_weight_norm / update_layer / forward` (lora_layers.py:15-96) and `LoraConv1d.update_layer / dora_init /
merge / unmerge / get_delta_weight / forward` (lora_layers.py:99-332).  peft's own `DoraLinearLayer.forward` (the nn.Linear
path used for attn.to_q / attn.to_kv) is NOT reproduced here and stays "parity unpinned".
"""
from __future__ import annotations

import contextlib
import math
import sys
import types

import torch
from torch import nn


class LoraLayer:
    """peft.tuners.lora.layer.LoraLayer (0.12.0): adapter containers + helpers used by lora_layers.py."""
    adapter_layer_names = ("lora_A", "lora_B", "lora_embedding_A", "lora_embedding_B")
    other_param_names = ("r", "lora_alpha", "scaling", "lora_dropout")

    def __init__(self, base_layer: nn.Module, **kwargs) -> None:
        self.base_layer = base_layer
        self.r = {}
        self.lora_alpha = {}
        self.scaling = {}
        self.lora_dropout = nn.ModuleDict({})
        self.lora_A = nn.ModuleDict({})
        self.lora_B = nn.ModuleDict({})
        self.lora_embedding_A = nn.ParameterDict({})
        self.lora_embedding_B = nn.ParameterDict({})
        self._disable_adapters = False
        self.merged_adapters = []
        self.use_dora = {}
        self.lora_magnitude_vector = nn.ModuleDict()
        self._caches = {}
        self.kwargs = kwargs

    # ---- BaseTunerLayer bits
    def get_base_layer(self) -> nn.Module:
        base = self
        while hasattr(base, "base_layer"):
            base = base.base_layer
        return base

    @property
    def merged(self) -> bool:
        return bool(self.merged_adapters)

    @property
    def disable_adapters(self) -> bool:
        return self._disable_adapters

    @property
    def active_adapters(self):
        a = self._active_adapter
        return [a] if isinstance(a, str) else a

    def set_adapter(self, adapter_names) -> None:
        if isinstance(adapter_names, str):
            adapter_names = [adapter_names]
        for layer_name in self.adapter_layer_names:
            for key, layer in getattr(self, layer_name).items():
                layer.requires_grad_(key in adapter_names)
        self._active_adapter = adapter_names

    def _move_adapter_to_device_of_base_layer(self, adapter_name: str) -> None:
        w = self.get_base_layer().weight
        for layer_name in self.adapter_layer_names + ("lora_magnitude_vector",):
            d = getattr(self, layer_name, None)
            if isinstance(d, (nn.ModuleDict, nn.ParameterDict)) and adapter_name in d:
                d[adapter_name].to(w.device)

    def reset_lora_parameters(self, adapter_name: str, init_lora_weights) -> None:
        if init_lora_weights is False:
            return
        if adapter_name in self.lora_A.keys():
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B[adapter_name].weight)

    def _cache_store(self, key, value) -> None:
        self._caches[key] = value

    def _cache_pop(self, key):
        return self._caches.pop(key)

    def _check_forward_args(self, x, *args, **kwargs) -> None:
        return None


class DoraLinearLayer(nn.Module):
    """peft.tuners.lora.dora.DoraLinearLayer (0.12.0) — constructor only; the reference overrides the rest."""

    def __init__(self, fan_in_fan_out: bool) -> None:
        super().__init__()
        self.fan_in_fan_out = fan_in_fan_out


def check_adapters_to_merge(module, adapter_names=None):
    """peft.tuners.tuners_utils.check_adapters_to_merge: the active adapters that are not merged yet."""
    if adapter_names is None:
        adapter_names = module.active_adapters
    if module.merged:
        adapter_names = [n for n in adapter_names if n not in module.merged_adapters]
    return adapter_names


def dequantize_module_weight(module: nn.Module) -> torch.Tensor:
    return module.weight


@contextlib.contextmanager
def gather_params_ctx(param, modifier_rank: int = 0, fwd_module=None):
    yield


def install() -> None:
    """Register the stub under the module names lora_layers.py imports (no-op when a real peft is importable)."""
    try:
        import peft  # noqa: F401
        return
    except Exception:
        pass

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    mod("peft", __stub__=True)
    mod("peft.tuners")
    mod("peft.tuners.lora", LoraLayer=LoraLayer)
    mod("peft.tuners.lora.dora", DoraLinearLayer=DoraLinearLayer)
    mod("peft.tuners.tuners_utils", check_adapters_to_merge=check_adapters_to_merge)
    mod("peft.utils")
    mod("peft.utils.integrations", dequantize_module_weight=dequantize_module_weight, gather_params_ctx=gather_params_ctx)


def load_reference_lora_layers(reference_root: str = "/root/reference"):
    """Import the reference's lora_layers.py by file path (its package __init__ chain pulls in absent dependencies)."""
    import importlib.util
    install()
    spec = importlib.util.spec_from_file_location("_ref_lora_layers", f"{reference_root}/osu_fusion/modules/lora_layers.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m
