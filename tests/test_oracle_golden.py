"""Oracle vs golden vectors produced by the real reference (oracle/make_golden.py → tests/golden/unet_tiny_ref.pt)."""
from pathlib import Path

import torch

from oracle.denoiser import UNet
from oracle.make_golden import grad_digest, run_case
from oracle.synth import synth_state_dict

GOLD = Path(__file__).parent / "golden" / "unet_tiny_ref.pt"


def test_oracle_reproduces_reference_golden():
    gold = torch.load(GOLD, weights_only=False)
    net = UNet(6, 96, 5, **gold["config"])
    net.load_state_dict(synth_state_dict(net, seed=gold["weight_seed"]))
    net.train()
    for name, case in gold["cases"].items():
        y, loss, grads = run_case(net, case["batch"], case["n"], case["seed"], case["cond_drop_prob"])
        assert (y - case["y"]).abs().max() <= 1e-4 * case["y"].abs().max(), name
        assert abs(float(loss) - float(case["loss"])) <= 1e-4 * abs(float(case["loss"])), name
        dig = grad_digest(grads)
        assert set(dig) == set(case["grad_digest"])
        for k, d in dig.items():
            ref = case["grad_digest"][k]
            assert abs(float(d[0]) - float(ref[0])) <= 1e-3 * float(ref[0]) + 1e-9, (name, k)
