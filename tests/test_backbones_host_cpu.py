"""Host side of osufusion_b200/backbones.py (DiT / MMDiT on the engine tape) on the GPU-less build box: the C-ABI is replaced by
tests/fake_native.py (a torch restatement of the header contracts over host pointers), so argument order, strides / views, tape
order, gradient routing and the gradient arena are exercised end to end against the oracle.  The CUDA kernels themselves are
covered by tests/test_zz_backbones_gpu.py."""
import pytest
import torch

import fake_native


def nrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


@pytest.fixture()
def fake_abi(monkeypatch):
    from osufusion_b200 import _native
    monkeypatch.setattr(_native, "call", fake_native.call)
    fake_native.CALLS.clear()
    return fake_native


def _pair(kind):
    from oracle.backbones import DiT as ODiT, MMDiT as OMMDiT
    from oracle.make_golden_backbones import DIT_TINY, MMDIT_TINY
    from oracle.synth import synth_state_dict
    from osufusion_b200.backbones import DiT, MMDiT
    new_cls, ora_cls, cfg = (DiT, ODiT, DIT_TINY) if kind == "dit" else (MMDiT, OMMDiT, MMDIT_TINY)
    ora = ora_cls(6, 96, 5, **cfg)
    ora.load_state_dict(synth_state_dict(ora))
    new = new_cls(6, 96, 5, **cfg)
    new.load_state_dict(ora.state_dict())
    return ora, new


@pytest.mark.parametrize("batched", [False, True])
@pytest.mark.parametrize("kind,n", [("dit", 64), ("dit", 56), ("mmdit", 64), ("mmdit", 50)])
def test_engine_tape_matches_oracle_through_emulated_abi(fake_abi, monkeypatch, kind, n, batched):
    from oracle.synth import synth_inputs
    from osufusion_b200 import backbones
    from osufusion_b200.modules import UNetFunction
    monkeypatch.setattr(backbones, "BATCHED", batched)
    monkeypatch.setattr(backbones, "GROUPED_PACK", batched)
    monkeypatch.setattr(backbones, "GROUPED_MOD", batched)
    ora, new = _pair(kind)
    x, a, c, t, noise, mask = synth_inputs(2, n, 1234)
    y_o = ora(x, a, t, c, cond_mask=mask)
    torch.nn.functional.mse_loss(y_o, noise).backward()
    with torch.autocast("cpu", dtype=torch.bfloat16), torch.no_grad():
        e_ref = nrel(ora(x, a, t, c, cond_mask=mask), y_o)
    y_n = UNetFunction.apply(new, x, a, t, c, mask, *list(new.parameters()))
    assert y_n.shape == y_o.shape
    assert nrel(y_n, y_o) <= max(1e-2, 2 * e_ref)
    torch.nn.functional.mse_loss(y_n, noise).backward()
    go = dict(ora.named_parameters())
    for k, p in new.named_parameters():
        assert (p.grad is None) == (go[k].grad is None), k
        if p.grad is not None:
            assert nrel(p.grad, go[k].grad) <= 4e-2, (k, nrel(p.grad, go[k].grad))
    used = set(fake_abi.CALLS)
    assert {"of_gemm", "of_attn_fwd", "of_attn_bwd", "of_headnorm_fwd", "of_headnorm_bwd", "of_gate_residual_fwd", "of_row_mean_std"} <= used
    if batched:
        assert "of_pack_weights" in used and "of_cast_f32_bf16" not in used  # every projection weight goes through the grouped pack
        assert {"of_film_fwd", "of_film_bwd"} <= used                          # every adaLN head in one launch each way
        assert {"of_adaln_fwd", "of_adaln_bwd", "of_gate_bwd"} <= used and not ({"of_layernorm_fwd", "of_coldot_bf16"} & used)
    else:
        assert {"of_layernorm_fwd", "of_layernorm_bwd", "of_gate_mul_bwd", "of_coldot_bf16"} <= used
    # inference path (no tape) gives the same output
    with torch.no_grad():
        out16, _ = new.run(None, x, a, t, c, mask)
        assert nrel(new.unpack(out16, n), y_n) <= 1e-6


@pytest.mark.parametrize("kind", ["dit", "mmdit"])
def test_reference_api_surface(kind):
    from oracle.make_golden_backbones import DIT_TINY, MMDIT_TINY
    from osufusion_b200.backbones import DiT, MMDiT
    net = (DiT if kind == "dit" else MMDiT)(6, 96, 5, **(DIT_TINY if kind == "dit" else MMDIT_TINY))
    net.set_gradient_checkpointing(True)
    assert all(b.gradient_checkpointing for b in net.blocks)
    x = torch.zeros(1, 6, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(x, torch.zeros(1, 96, 16), torch.zeros(1, dtype=torch.long), torch.zeros(1, 5))
    # zero-initialised adaLN heads / output convs like the reference (dit.py:238-250, mmdit.py:314-327)
    out = net.postprocess if kind == "dit" else net.out
    assert out.weight.abs().max() == 0
    assert all(b.modulation[1].weight.abs().max() == 0 for b in net.blocks) if kind == "dit" else \
        all(b.modulation_x[1].weight.abs().max() == 0 and b.modulation_a[1].weight.abs().max() == 0 for b in net.blocks)
    with pytest.raises(ValueError):
        (DiT if kind == "dit" else MMDiT)(6, 96, 5, dim_h=128, depth=1, attn_heads=3, attn_dim_head=32)


@pytest.mark.parametrize("kind", ["dit", "mmdit"])
def test_cfg_batched_pass_equals_two_passes(fake_abi, kind):
    """forward_with_cond_scale at inference = one forward over [cond ; null]; must equal the reference's two separate passes."""
    from oracle.synth import synth_inputs
    ora, new = _pair(kind)
    x, a, c, t, _, _ = synth_inputs(2, 44, 3)
    with torch.no_grad():
        ones, zeros = torch.ones(2, dtype=torch.bool), torch.zeros(2, dtype=torch.bool)
        cond = new.unpack(new.run(None, x, a, t, c, ones)[0], 44)
        null = new.unpack(new.run(None, x, a, t, c, zeros)[0], 44)
        two = null + (cond - null) * 2.0
        one = new._cfg_batched(x, a, t, c, 2.0)
        assert nrel(one, two) <= 1e-5
        ref = ora.forward_with_cond_scale(x, a, t, c, cond_scale=2.0)
        assert nrel(one, ref) <= 3e-2


def test_fused_adamw_drives_a_backbone(fake_abi):
    """FusedAdamW / the gradient arena are backbone-agnostic: two optimizer steps on the MMDiT equal torch.optim.AdamW on the same gradients,
    and the engine sees the updated weights (operand caches are invalidated through `param_epoch`)."""
    import copy

    from oracle.synth import synth_inputs
    from osufusion_b200.modules import UNetFunction
    from osufusion_b200.optim import FusedAdamW
    _, net = _pair("mmdit")
    ref = copy.deepcopy(net)
    opt = FusedAdamW(net, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=1e-2)
    x, a, c, t, noise, mask = synth_inputs(2, 48, 5)
    outs = []
    for _ in range(2):
        net.zero_grad(set_to_none=True)
        y = UNetFunction.apply(net, x, a, t, c, mask, *list(net.parameters()))
        outs.append(y.detach().clone())
        torch.nn.functional.mse_loss(y, noise).backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            q.grad = None if p.grad is None else p.grad.detach().clone()
        torch.nn.utils.clip_grad_norm_([q for q in ref.parameters() if q.grad is not None], 1.0)
        ropt.step()                         # parameters without a gradient are skipped by both (no decay, no moment update)
        opt.step()
        worst = max(((p - q).abs().max() / q.abs().max().clamp_min(1e-12)).item() for p, q in zip(net.parameters(), ref.parameters()))
        assert worst < 1e-5, worst
    assert nrel(outs[1], outs[0]) > 1e-4        # the second forward ran with the updated weights
