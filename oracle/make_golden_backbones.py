"""ORACLE (test infrastructure): golden vectors of the reference's DiT / MMDiT backbones, produced by
running the REAL reference modules (imported from /root/reference; CPU, fp32 weights + the built-in bf16 SDPA cast) on
deterministic synthetic weights and inputs.  Run in the build container only:   python -m oracle.make_golden_backbones
Writes tests/golden/backbones_ref.pt.  /root/reference is never read at test/bench time.
"""
from __future__ import annotations

import sys
import warnings

import torch

from .make_golden import OUT, REF, grad_digest, run_case
from .synth import synth_state_dict

DIT_TINY = dict(dim_h=128, depth=2, attn_heads=2, attn_dim_head=64)      # DiTAttention has no out-projection: heads * dim_head == dim_h
MMDIT_TINY = dict(dim_h=128, depth=2, patch_size=4, attn_heads=4, attn_kv_heads=2, attn_dim_head=32)
CASES = {"b2_n64_cond": (2, 64, 1234, 0.0), "b2_n50_ragged_null": (2, 50, 99, 1.0)}


def load_reference_backbone(kind: str, **cfg):
    """Import osu_fusion.modules.{dit.DiT, mmdit.MMDiT} and patch the CUDA-less Attend bug (attention.py:68-69,87)."""
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    warnings.filterwarnings("ignore")
    from osu_fusion.modules.attention import Attend, _config  # type: ignore
    if kind == "dit":
        from osu_fusion.modules.dit import DiT as Net  # type: ignore
    else:
        from osu_fusion.modules.mmdit import MMDiT as Net  # type: ignore
    net = Net(6, 96, 5, **cfg)
    for m in net.modules():
        if isinstance(m, Attend):
            m.cuda_config = _config(True, False, False)
    return net


def main() -> None:
    torch.set_num_threads(8)
    blob = {}
    for kind, cfg in (("dit", DIT_TINY), ("mmdit", MMDIT_TINY)):
        net = load_reference_backbone(kind, **cfg)
        net.load_state_dict(synth_state_dict(net, seed=0))
        net.train()
        cases = {}
        for name, (b, n, seed, p) in CASES.items():
            y, loss, grads = run_case(net, b, n, seed, p)
            cases[name] = dict(batch=b, n=n, seed=seed, cond_drop_prob=p, y=y, loss=loss, grad_digest=grad_digest(grads))
            print(kind, name, "loss", float(loss), "y absmax", float(y.abs().max()))
        blob[kind] = dict(config=cfg, weight_seed=0, cases=cases)
    blob["note"] = "reference osu_fusion.modules.dit.DiT / mmdit.MMDiT, CPU, fp32 + bf16 SDPA"
    blob["torch"] = torch.__version__
    OUT.mkdir(parents=True, exist_ok=True)
    torch.save(blob, OUT / "backbones_ref.pt")
    print("wrote", OUT / "backbones_ref.pt")


if __name__ == "__main__":
    main()
