"""ORACLE (test infrastructure): generate golden vectors by running the REAL reference (imported from
/root/reference, CPU, fp32 weights + its built-in bf16 SDPA cast = mode M2) on deterministic synthetic weights and
inputs.  Run in the build container only:   python -m oracle.make_golden
Writes tests/golden/unet_tiny_ref.pt (a few hundred KB).  /root/reference is never read at test/bench time.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

from .synth import TINY, synth_inputs, synth_state_dict

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load_reference_unet(**cfg):
    """Import osu_fusion.modules.unet.UNet from the reference and patch the CUDA-less Attend bug
    (attention.py:68-69 returns before setting cuda_config; forward reads it at :87)."""
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import warnings

    warnings.filterwarnings("ignore")
    from osu_fusion.modules.attention import Attend, _config  # type: ignore
    from osu_fusion.modules.unet import UNet  # type: ignore

    net = UNet(6, 96, 5, **cfg)
    for m in net.modules():
        if isinstance(m, Attend):
            m.cuda_config = _config(True, False, False)
    return net


def run_case(net, batch: int, n: int, seed: int, cond_drop_prob: float):
    x, a, c, t, noise, _ = synth_inputs(batch, n, seed)
    net.zero_grad(set_to_none=True)
    y = net(x, a, t, c, cond_drop_prob=cond_drop_prob)
    loss = torch.nn.functional.mse_loss(y, noise)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
    return y.detach(), loss.detach(), grads


def grad_digest(grads):
    """Per-parameter (l2 norm, sum, first 4 entries) — compact but sensitive."""
    out = {}
    for k, g in grads.items():
        f = g.flatten().double()
        out[k] = torch.cat([f.norm()[None], f.sum()[None], f[:4] if f.numel() >= 4 else torch.nn.functional.pad(f, (0, 4 - f.numel()))]).float()
    return out


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(8)
    net = load_reference_unet(**TINY)
    net.load_state_dict(synth_state_dict(net, seed=0))
    net.train()
    cases = {}
    for name, (b, n, seed, p) in {
        "b2_n64_cond": (2, 64, 1234, 0.0),
        "b2_n40_ragged_null": (2, 40, 99, 1.0),      # n not a multiple of 4 -> exercises the -1/-23 padding
        "b1_n128_cond": (1, 128, 7, 0.0),
    }.items():
        y, loss, grads = run_case(net, b, n, seed, p)
        cases[name] = dict(batch=b, n=n, seed=seed, cond_drop_prob=p, y=y, loss=loss, grad_digest=grad_digest(grads))
        print(name, "loss", float(loss), "y absmax", float(y.abs().max()))
    OUT.mkdir(parents=True, exist_ok=True)
    torch.save(dict(config=TINY, weight_seed=0, cases=cases, torch=torch.__version__,
                    note="reference osu_fusion.modules.unet.UNet, CPU, fp32 + bf16 SDPA (M2)"), OUT / "unet_tiny_ref.pt")
    print("wrote", OUT / "unet_tiny_ref.pt")


if __name__ == "__main__":
    main()
