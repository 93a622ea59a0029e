"""Pins the oracle restatement against the REAL reference modules (build container only; the GPU box has no
/root/reference and this file skips there)."""
import warnings

import pytest
import torch

from conftest import HAVE_REFERENCE

pytestmark = pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference not present")


@pytest.fixture(scope="module")
def pair():
    warnings.filterwarnings("ignore")
    from oracle.denoiser import UNet
    from oracle.make_golden import load_reference_unet
    from oracle.synth import TINY, synth_state_dict

    ref = load_reference_unet(**TINY)
    ora = UNet(6, 96, 5, **TINY)
    sd = synth_state_dict(ref)
    ref.load_state_dict(sd)
    ora.load_state_dict(sd)
    return ref, ora


def test_state_dict_keys_identical(pair):
    ref, ora = pair
    rk, ok = ref.state_dict(), ora.state_dict()
    assert list(rk.keys()) == list(ok.keys())
    assert all(rk[k].shape == ok[k].shape for k in rk)


@pytest.mark.parametrize("b,n,seed,p", [(2, 64, 1234, 0.0), (2, 40, 99, 1.0), (1, 37, 5, 0.0)])
def test_forward_backward_match_reference(pair, b, n, seed, p):
    from oracle.make_golden import run_case

    ref, ora = pair
    y1, l1, g1 = run_case(ref, b, n, seed, p)
    y2, l2, g2 = run_case(ora, b, n, seed, p)
    assert y1.abs().max() > 0.1  # final_conv re-randomised: not the vacuous all-zero output
    assert (y1 - y2).abs().max() <= 1e-6 * y1.abs().max()
    assert set(g1) == set(g2)
    for k in g1:
        assert (g1[k] - g2[k]).abs().max() <= 1e-5 * g1[k].abs().max().clamp_min(1e-12), k


def test_cond_scale_matches_reference(pair):
    from oracle.synth import synth_inputs

    ref, ora = pair
    x, a, c, t, _, _ = synth_inputs(2, 48, 3)
    with torch.no_grad():
        r = ref.forward_with_cond_scale(x, a, t, c, cond_scale=2.0)
        o = ora.forward_with_cond_scale(x, a, t, c, cond_scale=2.0)
    assert (r - o).abs().max() <= 1e-6 * r.abs().max()


def test_zero_init_final_conv_like_reference():
    from oracle.denoiser import UNet
    from oracle.synth import TINY, synth_inputs

    torch.manual_seed(0)
    net = UNet(6, 96, 5, **TINY)
    x, a, c, t, _, _ = synth_inputs(1, 32, 1)
    assert net(x, a, t, c).abs().max() == 0  # unet.py:354 zero_init(final_conv)
