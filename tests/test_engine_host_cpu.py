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
            "of_rope_fwd", "of_rope_bwd", "of_attn_fwd", "of_attn_bwd", "of_upsample2x_fwd", "of_upsample2x_bwd"} <= used
    assert "of_unpack_conv_wgrad" not in used      # conv weight gradients accumulate straight into the packed arena slices
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


# ------------------------------------------------------------------------------------------------ LoRA / DoRA and the optimizer tail
def _lora_build(use_dora):
    """Engine UNet with injected adapters + the oracle UNet whose adapted layers evaluate the reference's three-term forward
    (lora_layers.py:59-92) under autograd on cloned leaves — the construction of tests/test_lora_gpu.py, on the CPU."""
    from oracle.denoiser import UNet as OracleUNet
    from oracle.dora import dora_conv1d, dora_linear
    from oracle.synth import TINY
    from osufusion_b200 import lora
    from osufusion_b200.modules import UNet
    F = torch.nn.functional
    torch.manual_seed(0)
    ora = OracleUNet(6, 96, 5, **TINY)
    torch.nn.init.normal_(ora.final_conv.weight, std=0.02)
    new = UNet(6, 96, 5, **TINY)
    new.load_state_dict(ora.state_dict())
    names = lora.inject_adapters(new, r=8, lora_alpha=16, use_dora=use_dora)
    for p in ora.parameters():
        p.requires_grad_(False)
    g = torch.Generator(device="cpu").manual_seed(1)
    leaves = {}
    for name in names:
        ad = new.get_submodule(name)
        with torch.no_grad():
            ad.lora_B["default"].weight.copy_(0.05 * torch.randn(ad.lora_B["default"].weight.shape, generator=g))
            if use_dora:
                ad.magnitude().mul_(1 + 0.1 * torch.randn(ad.magnitude().shape, generator=g))
        A = ad.lora_A["default"].weight.detach().clone().requires_grad_(True)
        Bm = ad.lora_B["default"].weight.detach().clone().requires_grad_(True)
        mag = ad.magnitude().detach().clone().requires_grad_(True) if use_dora else None
        leaves[name] = (A, Bm, mag)
        om = ora.get_submodule(name)
        sc = ad.scaling
        if isinstance(om, torch.nn.Conv1d):
            def fwd(x, om=om, A=A, Bm=Bm, mag=mag, sc=sc):
                if mag is None:
                    return F.conv1d(x, om.weight, om.bias, padding=1) + F.conv1d(F.conv1d(x, A, None, padding=1), Bm) * sc
                return dora_conv1d(x, om.weight, om.bias, A, Bm, mag, sc, padding=om.padding[0])
        else:
            def fwd(x, om=om, A=A, Bm=Bm, mag=mag, sc=sc):
                if mag is None:
                    return F.linear(x, om.weight, om.bias) + F.linear(F.linear(x, A), Bm) * sc
                return dora_linear(x, om.weight, om.bias, A, Bm, mag, sc)
        om.forward = fwd
    return ora, new, names, leaves


@pytest.mark.parametrize("rank_r,merge_tc,grouped", [(True, True, True), (True, True, False), (False, False, False)])
@pytest.mark.parametrize("use_dora", [True, False])
def test_lora_dora_tape_matches_oracle(fake_abi, monkeypatch, use_dora, rank_r, merge_tc, grouped):
    from oracle.synth import synth_inputs
    from osufusion_b200 import engine
    from osufusion_b200.modules import UNetFunction
    monkeypatch.setattr(engine, "LORA_RANK_R", rank_r)
    monkeypatch.setattr(engine, "LORA_MERGE_TC", merge_tc)
    monkeypatch.setattr(engine, "LORA_GROUPED", grouped)
    ora, new, names, leaves = _lora_build(use_dora)
    x, a, c, t, noise, keep = synth_inputs(2, 56, 5)

    def run_oracle(autocast):
        for vs in leaves.values():
            for v in vs:
                if v is not None:
                    v.grad = None
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            y = ora(x, a, t, c, cond_mask=keep)
        torch.nn.functional.mse_loss(y.float(), noise).backward()
        return y.detach().float(), {n: tuple(None if v is None else v.grad.detach().clone() for v in leaves[n]) for n in names}

    y_ref, g_ref = run_oracle(True)
    y_tru, g_tru = run_oracle(False)
    y_new = UNetFunction.apply(new, x, a, t, c, keep, *list(new.parameters()))
    torch.nn.functional.mse_loss(y_new, noise).backward()
    assert nrel(y_new, y_tru) <= max(1e-2, 2 * nrel(y_ref, y_tru))
    assert all(p.grad is None for n, p in new.named_parameters() if "lora_" not in n)      # base weights stay frozen
    bad = []
    for n in names:
        ad = new.get_submodule(n)
        mine = (ad.lora_A["default"].weight.grad, ad.lora_B["default"].weight.grad, ad.magnitude().grad if use_dora else None)
        for which, gm, gr, gt in zip("ABm", mine, g_ref[n], g_tru[n]):
            if gt is None:
                continue
            e_new, e_ref = nrel(gm.view(gt.shape), gt), nrel(gr, gt)
            if e_new > max(2e-2, 3 * e_ref, 6e-2 if use_dora else 0.0):
                bad.append((n, which, e_new, e_ref))
    assert not bad, bad[:6]
    used = set(fake_abi.CALLS)
    if grouped:      # one grouped operand launch, prep folded into the merge kernel, one finishing launch, no unpack pass
        assert {"of_dora_scale_pack_prep", "of_lora_finish_all"} <= used
        assert not ({"of_dora_rankr_prep", "of_dora_rankr_finish", "of_dora_grad", "of_unpack_conv_wgrad", "of_scale_cast_f32_bf16"} & used)
    elif rank_r:
        assert {"of_dora_rankr_prep", "of_dora_rankr_finish"} <= used and "of_dora_grad" not in used
    else:
        assert "of_dora_grad" in used
    if not grouped:
        assert ("of_dora_scale_pack" in used) == merge_tc and ("of_dora_merge" in used) == (not merge_tc)


@pytest.mark.parametrize("max_norm", [1.0, None])
def test_fused_adamw_host_logic_matches_torch(fake_abi, max_norm):
    import copy

    from oracle.synth import TINY, synth_inputs
    from osufusion_b200.modules import UNet, UNetFunction
    from osufusion_b200.optim import FusedAdamW, cosine_schedule_with_warmup
    torch.manual_seed(0)
    net = UNet(6, 96, 5, **TINY)
    torch.nn.init.normal_(net.final_conv.weight, std=0.02)
    ref = copy.deepcopy(net)
    opt = FusedAdamW(net, lr=1e-3, weight_decay=1e-2, max_grad_norm=max_norm)
    sched = cosine_schedule_with_warmup(opt, 2, 10)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=1e-2)
    rsched = cosine_schedule_with_warmup(ropt, 2, 10)
    x, a, c, t, noise, keep = synth_inputs(2, 48, 7)
    y0 = None
    for it in range(3):
        net.zero_grad(set_to_none=True)
        y = UNetFunction.apply(net, x, a, t, c, keep, *list(net.parameters()))
        y0 = y.detach().clone() if y0 is None else y0
        torch.nn.functional.mse_loss(y, noise).backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            q.grad = p.grad.detach().clone()
        if max_norm is not None:
            tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        ropt.step()
        rsched.step()
        opt.step()
        sched.step()
        if max_norm is not None:
            assert abs(float(opt.grad_norm) - float(tn)) <= 1e-4 * float(tn)
        worst = max(((p - q).abs().max() / q.abs().max().clamp_min(1e-12)).item() for p, q in zip(net.parameters(), ref.parameters()))
        assert worst < 1e-5, (it, worst)
    # parameters were updated through raw pointers: the operand caches must have been invalidated (param_epoch) ...
    with torch.no_grad():
        out16, _ = net.run(None, x, a, t, c, keep)
        y3 = net.unpack(out16, 48)
        assert nrel(y3, y0) > 1e-4
    # ... while the grouped bf16 GEMM operands were rewritten by the optimizer kernel itself: ONE of_pack_weights launch in total
    # (the first forward), and the result equals a fresh model that packs the updated weights from scratch
    assert fake_abi.CALLS.count("of_pack_weights") == 1
    fresh = UNet(6, 96, 5, **TINY)
    fresh.load_state_dict(net.state_dict())
    with torch.no_grad():
        out16f, _ = fresh.run(None, x, a, t, c, keep)
    assert nrel(y3, fresh.unpack(out16f, 48)) <= 1e-6
    assert len(net._store.packed) > 0 and all(net._store.arena_views[i].stride() != net._store.arena_views[i].contiguous().stride()
                                               for i in net._store.packed)


@pytest.mark.parametrize("use_dora", [True, False])
def test_adapted_forward_equals_merged_model(fake_abi, use_dora):
    """The engine's on-the-fly effective weight s * (W + scaling * B A) must give the same output as the plain model after
    `merge_and_unload` (peft's merge, lora_layers.py:197-256): the adapter directory and the merged checkpoint describe one model."""
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200 import lora
    from osufusion_b200.modules import UNet
    torch.manual_seed(0)
    net = UNet(6, 96, 5, **TINY)
    torch.nn.init.normal_(net.final_conv.weight, std=0.02)
    names = lora.inject_adapters(net, r=8, lora_alpha=16, use_dora=use_dora)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for name in names:
            ad = net.get_submodule(name)
            ad.lora_B["default"].weight.copy_(0.05 * torch.randn(ad.lora_B["default"].weight.shape, generator=g))
            if use_dora:
                ad.magnitude().mul_(1 + 0.1 * torch.randn(ad.magnitude().shape, generator=g))
    x, a, c, t, _, keep = synth_inputs(2, 40, 9)
    with torch.no_grad():
        y_adapted = net.unpack(net.run(None, x, a, t, c, keep)[0], 40)
        lora.merge_and_unload(net)
        assert not any(hasattr(m, "base_layer") for m in net.modules())
        y_merged = net.unpack(net.run(None, x, a, t, c, keep)[0], 40)
    assert y_adapted.abs().max() > 1e-3
    # both are bf16 evaluations of the same effective weights; under DoRA the adapted to_q keeps q in fp32 through RoPE (the reference's
    # type promotion, lora_layers / peft) while the merged model rotates in bf16, hence bf16-level rather than bit-level agreement
    assert nrel(y_adapted, y_merged) <= 5e-2
    assert (y_adapted - y_merged).abs().mean() <= 3e-2 * y_merged.abs().mean()


@pytest.mark.parametrize("kind", ["diffusion", "rectified_flow"])
def test_training_rng_draw_order_matches_reference(fake_abi, kind):
    """Without injected draws the wrappers must consume the generator like the reference (diffusion.py:88-98: randn_like -> randint ->
    CFG uniform_; rectified_flow.py:89-98: randn_like -> rand -> uniform_): same seed => same noise / timesteps / mask => same loss."""
    from oracle.models import DiffusionOsuFusion as OD, RectifiedFlowOsuFusion as OR
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200.models import DiffusionOsuFusion, RectifiedFlowOsuFusion
    OC, NC = (OD, DiffusionOsuFusion) if kind == "diffusion" else (OR, RectifiedFlowOsuFusion)
    torch.manual_seed(0)
    ora = OC(**TINY)
    torch.nn.init.normal_(ora.unet.final_conv.weight, std=0.02)
    new = NC(**TINY)
    new.load_state_dict(ora.state_dict())
    x, a, c, _, _, _ = synth_inputs(4, 48, 21)
    losses = []
    for seed in (5, 6):
        torch.manual_seed(seed)
        with torch.autocast("cpu", dtype=torch.bfloat16), torch.no_grad():
            l_ref = ora(x, a, c).item()
        after_ref = torch.rand(1).item()
        torch.manual_seed(seed)
        with torch.no_grad():
            l_new = new(x, a, c).item()
        after_new = torch.rand(1).item()
        assert abs(l_new - l_ref) <= 1e-2 * abs(l_ref), (seed, l_new, l_ref)
        assert after_new == after_ref          # the generator was advanced by exactly the same draws
        losses.append(l_ref)
    assert abs(losses[0] - losses[1]) > 1e-3   # the draws do matter: a different seed gives a different loss


@pytest.mark.parametrize("grad16", [True, False])
def test_grad_prescale_is_an_exact_power_of_two_shift(fake_abi, monkeypatch, grad16):
    """ddp.GradAllReducer takes the mean by pre-division: `unet.grad_prescale = 1 / world` scales the loss gradient at the start of
    backward and the buckets are all-reduced with SUM (the NVLS path has no AVG).  For a power-of-two world that is exact: every
    gradient is bit-for-bit the unscaled one times 1 / world.  Also covers the epilogue-emitted bf16 gradient copies (OF_GRAD16):
    both settings must give the same gradients as the cast pass they replace."""
    from oracle.synth import synth_inputs
    from osufusion_b200 import engine
    from osufusion_b200.modules import UNetFunction
    monkeypatch.setattr(engine, "GRAD16", grad16)
    _, new = _unet_pair("default")
    x, a, c, t, noise, keep = synth_inputs(2, 48, 11)

    def grads(scale):
        new.grad_prescale = scale
        new.zero_grad(set_to_none=True)
        y = UNetFunction.apply(new, x, a, t, c, keep, *list(new.parameters()))
        torch.nn.functional.mse_loss(y, noise).backward()
        return {k: p.grad.detach().clone() for k, p in new.named_parameters()}

    g1, g8 = grads(1.0), grads(0.125)
    new.grad_prescale = 1.0
    assert all(torch.equal(g8[k], g1[k] * 0.125) for k in g1), [k for k in g1 if not torch.equal(g8[k], g1[k] * 0.125)][:5]
    calls = fake_abi.CALLS
    assert ("of_cast_copy" in calls) or grad16


def test_epilogue_emitted_bf16_gradients_equal_the_cast_pass(fake_abi, monkeypatch):
    """OF_GRAD16: the dgrad GEMM that completes an activation's gradient also writes its bf16 copy; the values every consumer sees
    are those the separate fp32 -> bf16 cast pass produced (bf16 of the same fp32 sum), so all parameter gradients are identical."""
    from oracle.synth import synth_inputs
    from osufusion_b200 import engine
    from osufusion_b200.modules import UNetFunction
    _, new = _unet_pair("default")
    x, a, c, t, noise, keep = synth_inputs(2, 64, 12)
    out = {}
    for flag in (True, False):
        monkeypatch.setattr(engine, "GRAD16", flag)
        fake_abi.CALLS.clear()
        new.zero_grad(set_to_none=True)
        y = UNetFunction.apply(new, x, a, t, c, keep, *list(new.parameters()))
        torch.nn.functional.mse_loss(y, noise).backward()
        out[flag] = ({k: p.grad.detach().clone() for k, p in new.named_parameters()}, list(fake_abi.CALLS).count("of_cast_copy"))
    (g_on, n_on), (g_off, n_off) = out[True], out[False]
    assert n_on < n_off, (n_on, n_off)
    assert all(torch.equal(g_on[k], g_off[k]) for k in g_on)
