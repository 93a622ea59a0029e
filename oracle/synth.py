"""ORACLE (test infrastructure): deterministic synthetic weights and inputs shared by the reference, the oracle and
the CUDA implementation.  Everything is generated on the CPU from per-tensor seeds, so any module tree with the
reference's state_dict keys/shapes gets bit-identical parameters irrespective of construction order."""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import torch

TINY = dict(dim_h=96, dim_h_mult=(1, 2), num_layer_blocks=(1, 1), num_middle_transformers=1, attn_heads=2,
            attn_dim_head=16)
SMALL = dict(dim_h=128)   # inference_gradio.py:40  (CFG-S)
LARGE = dict(dim_h=512)   # trainer.py:379         (CFG-L)


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
    return g


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = _gen(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith("null_cond"):
        return torch.randn(shape, generator=g)
    if leaf == "bias":
        return 0.05 * torch.randn(shape, generator=g)
    if leaf == "gamma" or (leaf == "weight" and len(shape) == 1):  # norm gains (incl. the DiT / MMDiT per-head RMS-norm gamma)
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    std = 1.0 / (3.0 * max(fan_in, 1)) ** 0.5   # same second moment as torch's default kaiming-uniform(a=sqrt(5)) init
    if key in ("postprocess.weight", "out.weight"):
        std = 0.05  # zero-initialised 6-channel output convs of DiT / MMDiT (dit.py:250, mmdit.py:326): re-randomise likewise
    if "final_conv" in key:
        std = 0.05  # the reference zero-inits final_conv (unet.py:354), which makes parity vacuous: re-randomise
    return std * torch.randn(shape, generator=g)


def synth_state_dict(model: torch.nn.Module, seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(v.shape), seed).to(v.dtype) for k, v in model.state_dict().items()}


def synth_inputs(batch: int, n: int, seed: int = 1234, timesteps: int = 1000):
    """x (B,6,N), a (B,96,N), c (B,5), t (B,) int64, noise (B,6,N), cond_mask (B,) — DummyDataset convention
    (osu_fusion/library/dataset.py:125-131): N(0,1) tensors."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.randn(batch, 6, n, generator=g)
    a = torch.randn(batch, 96, n, generator=g)
    c = torch.randn(batch, 5, generator=g)
    t = torch.randint(0, timesteps, (batch,), generator=g, dtype=torch.int64)
    noise = torch.randn(batch, 6, n, generator=g)
    mask = torch.rand(batch, generator=g) < 0.5
    return x, a, c, t, noise, mask
