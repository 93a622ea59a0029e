"""Host side of the U-Net engine (engine.py / modules.py / models/) on the GPU-less build box: the C-ABI is replaced by
tests/fake_native.py (torch restatement of the header contracts over host pointers), so the complete tape — grouped weight packing,
grouped FiLM heads, ResidualBlock / Transformer / sampler blocks, gradient arena, training-step and sampling wrappers — runs end to
end and is compared with the oracle under the criterion of tests/test_model_parity_gpu.py (oracle under CPU bf16 autocast = the
reference's arithmetic, oracle fp32 = truth).  The CUDA kernels themselves are covered by the `-m gpu` tests."""
import pytest
import torch

import fake_native

GRAD_SLACK = 3.0


def nrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


@pytest.fixture()
def fake_abi(monkeypatch):
    from osufusion_b200 import _native
    monkeypatch.setattr(_native, "call", fake_native.call)
    fake_native.CALLS.clear()
    return fake_native


def _unet_pair(init):
    from oracle.denoiser import UNet as OracleUNet
    from oracle.synth import TINY, synth_state_dict
    from osufusion_b200.modules import UNet
    torch.manual_seed(0)
    ora = OracleUNet(6, 96, 5, **TINY)
    if init == "synth":
        ora.load_state_dict(synth_state_dict(ora))
    else:
        torch.nn.init.normal_(ora.final_conv.weight, std=0.02)
    new = UNet(6, 96, 5, **TINY)
    new.load_state_dict(ora.state_dict())
    return ora, new


def _oracle_fwd_bwd(ora, inputs, keep, autocast):
    x, a, c, t, noise = inputs
    ora.zero_grad(set_to_none=True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        y = ora(x, a, t, c, cond_mask=keep)
    torch.nn.functional.mse_loss(y.float(), noise).backward()
    return y.detach().float(), {k: p.grad.detach().float().clone() for k, p in ora.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("fused_gctx", [True, False])
@pytest.mark.parametrize("init,n,drop", [("synth", 64, False), ("default", 50, True)])
def test_unet_tape_matches_oracle_through_emulated_abi(fake_abi, monkeypatch, init, n, drop, fused_gctx):
    from oracle.synth import synth_inputs
    from osufusion_b200 import engine
    from osufusion_b200.modules import UNetFunction
    monkeypatch.setattr(engine, "GCTX_FUSED", fused_gctx)
    ora, new = _unet_pair(init)
    x, a, c, t, noise, mask = synth_inputs(2, n, 1234)
    keep = mask if drop else torch.ones(2, dtype=torch.bool)
    inputs = (x, a, c, t, noise)
    y_ref, g_ref = _oracle_fwd_bwd(ora, inputs, keep, True)
    y_tru, g_tru = _oracle_fwd_bwd(ora, inputs, keep, False)
    y_new = UNetFunction.apply(new, x, a, t, c, keep, *list(new.parameters()))
    assert y_new.shape == y_tru.shape
    assert nrel(y_new, y_tru) <= max(1e-2, 2 * nrel(y_ref, y_tru))
    torch.nn.functional.mse_loss(y_new, noise).backward()
    g_new = {k: p.grad.detach().float() for k, p in new.named_parameters() if p.grad is not None}
    assert set(g_new) == set(g_tru)
    bad = []
    for k in g_tru:
        if k.endswith("se.to_k.bias"):      # softmax is shift invariant: the true gradient is exactly zero
            continue
        e_new, e_ref = nrel(g_new[k], g_tru[k]), nrel(g_ref[k], g_tru[k])
        if e_new > max(2e-2, GRAD_SLACK * e_ref):
            bad.append((k, e_new, e_ref))
    assert not bad, bad[:6]
    used = set(fake_abi.CALLS)
    assert {"of_pack_weights", "of_film_fwd", "of_film_bwd", "of_rb_apply_fwd", "of_rb_gate_fwd", "of_rb_bwd_pass1", "of_rb_bwd_apply",
            "of_rope_fwd", "of_rope_bwd", "of_attn_fwd", "of_attn_bwd", "of_upsample2x_fwd", "of_upsample2x_bwd",
            "of_unpack_conv_wgrad"} <= used
    assert ("of_rb_logit_pool" in used) == fused_gctx and ("of_rb_pool" in used) == (not fused_gctx)
    # second backward without zero_grad accumulates (p.grad aliases the arena after the first one)
    y2 = UNetFunction.apply(new, x, a, t, c, keep, *list(new.parameters()))
    torch.nn.functional.mse_loss(y2, noise).backward()
    k = "final_resnet.block1.proj.weight"
    assert nrel(dict(new.named_parameters())[k].grad, 2 * g_new[k]) <= 1e-5


@pytest.mark.parametrize("kind", ["diffusion", "rectified_flow"])
def test_train_step_wrapper_matches_oracle(fake_abi, kind):
    """model(x, a, c, orig_len) with injected noise / timesteps / mask (diffusion.py:79-111, rectified_flow.py:81-111)."""
    from oracle.models import DiffusionOsuFusion as OD, RectifiedFlowOsuFusion as OR
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200.models import DiffusionOsuFusion, RectifiedFlowOsuFusion
    OC, NC = (OD, DiffusionOsuFusion) if kind == "diffusion" else (OR, RectifiedFlowOsuFusion)
    torch.manual_seed(0)
    ora = OC(**TINY)
    torch.nn.init.normal_(ora.unet.final_conv.weight, std=0.02)
    new = NC(**TINY)
    new.load_state_dict(ora.state_dict())
    x, a, c, t, noise, mask = synth_inputs(3, 72, 77)
    ts = t if kind == "diffusion" else torch.rand(3) * 0.9 + 0.05
    orig_len = torch.tensor([72, 40, 64])
    with torch.autocast("cpu", dtype=torch.bfloat16):
        l_ref = ora(x, a, c, orig_len, noise=noise, timesteps=ts, cond_mask=mask)
    l_ref.backward()
    l_new = new(x, a, c, orig_len, noise=noise, timesteps=ts, cond_mask=mask)
    l_new.backward()
    assert abs(l_new.item() - l_ref.item()) <= 1e-2 * abs(l_ref.item())
    for k in ("unet.final_resnet.block1.proj.weight", "unet.down_layers.0.resnets.0.block2.proj.weight", "unet.time_mlp.1.weight"):
        gn, gr = dict(new.named_parameters())[k].grad, dict(ora.named_parameters())[k].grad
        assert nrel(gn, gr) < 6e-2, (k, nrel(gn, gr))
    with pytest.raises(AssertionError):
        new(x, a[:, :, :-1], c)
    assert {"of_mse_fwd", "of_mse_bwd"} <= set(fake_abi.CALLS)


@pytest.mark.parametrize("kind,steps", [("diffusion", 1), ("rectified_flow", 2), ("diffusion", 4)])
def test_sampler_wrapper_matches_oracle(fake_abi, kind, steps):
    """CFG-batched, audio-cached sampler with the fused update vs the oracle's plain loop (diffusion.py:59-77, rectified_flow.py:57-79)."""
    from oracle.models import DiffusionOsuFusion as OD, RectifiedFlowOsuFusion as OR
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200.models import DiffusionOsuFusion, RectifiedFlowOsuFusion
    OC, NC = (OD, DiffusionOsuFusion) if kind == "diffusion" else (OR, RectifiedFlowOsuFusion)
    torch.manual_seed(0)
    ora = OC(**TINY, sampling_timesteps=steps).eval()
    torch.nn.init.normal_(ora.unet.final_conv.weight, std=0.02)
    new = NC(**TINY, sampling_timesteps=steps).eval()
    new.load_state_dict(ora.state_dict())
    x, a, c, _, noise, _ = synth_inputs(2, 60, 11)
    for scale in (1.0, 2.0):
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y_ref = ora.sample(a, c, noise.clone(), cond_scale=scale)
        y_new = new.sample(a, c, noise.clone(), cond_scale=scale)
        assert y_new.shape == y_ref.shape == (2, 6, 60)
        assert nrel(y_new, y_ref) < (3e-2 if steps <= 2 else 0.3), (kind, steps, scale, nrel(y_new, y_ref))
        assert (y_new - y_ref).abs().mean() < 3e-2 * y_ref.abs().mean().clamp_min(0.1)
    assert "of_sampler_update" in set(fake_abi.CALLS)
