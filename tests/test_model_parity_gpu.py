"""Whole-model parity on the GPU: CUDA engine vs the oracle (= the reference, proven bit-identical on CPU).

Three arms on identical weights / inputs / noise / timesteps / CFG mask:
    new   = osufusion_b200 engine (bf16 compute)
    ref16 = oracle under torch.autocast(cuda, bf16)  -> what the reference trainer runs (mode M1)
    truth = oracle in fp32 (mode M2: fp32 everywhere except the built-in bf16 SDPA cast)
Tolerance (north_star: bf16 max relative error <= 1e-2 on outputs and gradients, norm-wise max|a-b|/max|b|), applied as
SURVEY.md §8c recommends: err(new, truth) <= max(1e-2, 2*err(ref16, truth)) for the output, and the same with factor
GRAD_SLACK = 3 for every one of the gradient tensors — i.e. the engine is as close to the fp32 truth as the reference's own
bf16 path up to a small factor (observed worst ratio 2.2, on one cancellation-dominated `se.to_k.weight` gradient).
`se.to_k.bias` is excluded: softmax is shift-invariant, its true gradient is exactly zero.

The engine accumulates with fp32 atomics (split-K weight gradients, dq/dk/dv, per-channel reductions), so the bf16 rounding of
downstream tensors differs from run to run.  A handful of tiny, cancellation-dominated reductions (bias gradients of a few
channels summed over B*L signed bf16 values) therefore fluctuate around the GRAD_SLACK bound; the criterion tolerates at most
OUTLIER_FRAC of the tensors between GRAD_SLACK and OUTLIER_SLACK times the reference's own bf16 error, none beyond.
"""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"
GOLD = Path(__file__).parent / "golden" / "unet_tiny_ref.pt"
GRAD_SLACK = 3.0
OUTLIER_SLACK, OUTLIER_FRAC = 8.0, 0.01


def nrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def build_pair(cfg, init):
    from oracle.denoiser import UNet as OracleUNet
    from oracle.synth import synth_state_dict
    from osufusion_b200.modules import UNet
    torch.manual_seed(0)
    ora = OracleUNet(6, 96, 5, **cfg)
    if init == "synth":
        ora.load_state_dict(synth_state_dict(ora))
    else:
        torch.nn.init.normal_(ora.final_conv.weight, std=0.02)
    sd = {k: v.clone() for k, v in ora.state_dict().items()}
    new = UNet(6, 96, 5, **cfg)
    new.load_state_dict(sd)
    return ora.to(dev), new.to(dev)


def fwd_bwd(model, inputs, keep, autocast):
    x, a, c, t, noise = inputs
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        y = model(x, a, t, c, cond_mask=keep)
    torch.nn.functional.mse_loss(y.float(), noise).backward()
    return y.detach().float(), {k: p.grad.detach().float().clone() for k, p in model.named_parameters() if p.grad is not None}


def check(cfg, B, n, init, drop):
    from oracle.synth import synth_inputs
    ora, new = build_pair(cfg, init)
    x, a, c, t, noise, mask = (v.to(dev) for v in synth_inputs(B, n, 1234))
    keep = mask if drop else torch.ones(B, dtype=torch.bool, device=dev)
    inputs = (x, a, c, t, noise)
    y_new, g_new = fwd_bwd(new, inputs, keep, False)
    y_ref, g_ref = fwd_bwd(ora, inputs, keep, True)
    y_tru, g_tru = fwd_bwd(ora, inputs, keep, False)
    assert y_tru.abs().max() > 1e-3
    assert nrel(y_new, y_tru) <= max(1e-2, 2 * nrel(y_ref, y_tru))
    assert set(g_new) == set(g_tru)
    bad, outliers = [], []
    for k in g_tru:
        if k.endswith("se.to_k.bias"):
            assert g_new[k].abs().max() <= 1e-3 * max(1.0, g_ref[k].abs().max().item() * 1e3)
            continue
        e_new, e_ref = nrel(g_new[k], g_tru[k]), nrel(g_ref[k], g_tru[k])
        if e_new > max(3e-2, OUTLIER_SLACK * e_ref):
            bad.append((k, e_new, e_ref))
        elif e_new > max(1e-2, GRAD_SLACK * e_ref):
            outliers.append((k, e_new, e_ref))
    assert not bad, bad[:5]
    assert len(outliers) <= max(1, int(OUTLIER_FRAC * len(g_tru))), outliers[:8]
    return y_new


@pytest.mark.parametrize("init,n,drop", [("default", 200, True), ("synth", 64, False), ("synth", 200, True)])
def test_tiny_forward_backward(init, n, drop):
    from oracle.synth import TINY
    check(TINY, 2, n, init, drop)


def test_small_cfg_s_forward_backward():
    from oracle.synth import SMALL
    check(SMALL, 2, 1024, "default", True)


def test_cfg_l_forward_backward():
    """The headline configuration (dim_h=512, 1.28 B parameters, all 1239 gradients) at a reduced length: exercises the 256-wide
    tiles, the CTA-pair GEMM mode and every deep-level shape of the benchmark."""
    from oracle.synth import LARGE
    check(LARGE, 1, 1024, "default", True)
    torch.cuda.empty_cache()


def test_golden_reference_outputs():
    """Engine output vs golden vectors produced by the REAL reference (CPU fp32) — tests/golden/unet_tiny_ref.pt."""
    from oracle.denoiser import UNet as OracleUNet
    from oracle.synth import synth_inputs, synth_state_dict
    from osufusion_b200.modules import UNet
    gold = torch.load(GOLD, weights_only=False)
    ora = OracleUNet(6, 96, 5, **gold["config"])
    sd = synth_state_dict(ora, seed=gold["weight_seed"])
    ora.load_state_dict(sd)
    ora = ora.to(dev)
    new = UNet(6, 96, 5, **gold["config"])
    new.load_state_dict(sd)
    new = new.to(dev)
    for name, case in gold["cases"].items():
        x, a, c, t, _, _ = (v.to(dev) for v in synth_inputs(case["batch"], case["n"], case["seed"]))
        with torch.no_grad():
            y = new(x, a, t, c, cond_drop_prob=case["cond_drop_prob"]).cpu()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y16 = ora(x, a, t, c, cond_drop_prob=case["cond_drop_prob"]).float().cpu()
        e_new, e_ref = nrel(y, case["y"]), nrel(y16, case["y"])
        assert e_new <= max(1e-2, 2 * e_ref), (name, e_new, e_ref)


def test_cond_scale_one_equals_forward_and_ragged_length():
    from oracle.synth import TINY, synth_inputs
    _, new = build_pair(TINY, "default")
    x, a, c, t, _, _ = (v.to(dev) for v in synth_inputs(2, 37, 5))   # 37 is not a multiple of 4: padding path
    with torch.no_grad():
        y0 = new(x, a, t, c)
        y1 = new.forward_with_cond_scale(x, a, t, c, cond_scale=1.0)
    assert y0.shape == (2, 6, 37) and torch.equal(y0, y1)


def test_zero_init_final_conv_gives_zero_output():
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200.modules import UNet
    torch.manual_seed(0)
    net = UNet(6, 96, 5, **TINY).to(dev)
    x, a, c, t, _, _ = (v.to(dev) for v in synth_inputs(1, 32, 1))
    with torch.no_grad():
        assert net(x, a, t, c).abs().max() == 0     # unet.py:354


def test_train_step_loss_and_grads_match_oracle():
    """model(x, a, c, orig_len) with injected noise / timesteps / mask vs the oracle wrapper (diffusion.py:79-111)."""
    from oracle.models import DiffusionOsuFusion as OracleModel
    from oracle.models import RectifiedFlowOsuFusion as OracleRF
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200.models import DiffusionOsuFusion, RectifiedFlowOsuFusion
    for OC, NC, tkind in ((OracleModel, DiffusionOsuFusion, "int"), (OracleRF, RectifiedFlowOsuFusion, "float")):
        torch.manual_seed(0)
        ora = OC(**TINY)
        torch.nn.init.normal_(ora.unet.final_conv.weight, std=0.02)
        new = NC(**TINY)
        new.load_state_dict(ora.state_dict())
        ora, new = ora.to(dev), new.to(dev)
        x, a, c, t, noise, mask = (v.to(dev) for v in synth_inputs(3, 120, 77))
        ts = t if tkind == "int" else torch.rand(3, device=dev) * 0.9 + 0.05
        orig_len = torch.tensor([120, 64, 100], device=dev)
        with torch.autocast("cuda", dtype=torch.bfloat16):
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


def test_sampling_matches_oracle_trajectory():
    """Fused, CFG-batched, audio-cached sampler vs the oracle's plain loop (diffusion.py:59-77, rectified_flow.py:57-79)."""
    from oracle.models import DiffusionOsuFusion as OracleModel
    from oracle.models import RectifiedFlowOsuFusion as OracleRF
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200.models import DiffusionOsuFusion, RectifiedFlowOsuFusion
    for OC, NC, kw, tol in ((OracleModel, DiffusionOsuFusion, dict(sampling_timesteps=1), 3e-2),
                            (OracleRF, RectifiedFlowOsuFusion, dict(sampling_timesteps=2), 3e-2),
                            (OracleModel, DiffusionOsuFusion, dict(sampling_timesteps=6), 0.3),
                            (OracleRF, RectifiedFlowOsuFusion, dict(sampling_timesteps=4), 0.3)):
        torch.manual_seed(0)
        ora = OC(**TINY, **kw)
        torch.nn.init.normal_(ora.unet.final_conv.weight, std=0.02)
        new = NC(**TINY, **kw)
        new.load_state_dict(ora.state_dict())
        ora, new = ora.to(dev).eval(), new.to(dev).eval()
        x, a, c, _, noise, _ = (v.to(dev) for v in synth_inputs(2, 100, 11))
        for scale in (1.0, 2.0):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y_ref = ora.sample(a, c, noise.clone(), cond_scale=scale)
            y_new = new.sample(a, c, noise.clone(), cond_scale=scale)
            assert y_new.shape == y_ref.shape == (2, 6, 100)
            # one step (1-2 denoiser evaluations) is held to 3e-2; longer chains of bf16 evaluations (CFG amplifies their
            # differences, x0 is clamped) only to a loose max bound plus a mean bound.  The per-step update arithmetic itself is
            # checked tightly in test_sampler_update_kernel_matches_schedules.
            assert nrel(y_new, y_ref) < tol, (NC.__module__, kw, scale, nrel(y_new, y_ref))
            assert (y_new - y_ref).abs().mean() < 3e-2 * y_ref.abs().mean().clamp_min(0.1)


def test_sampler_update_kernel_matches_schedules():
    """Fused CFG + DDIM / midpoint update vs the restated diffusers / torchdiffeq arithmetic on identical predictions."""
    from oracle.schedules import DDIMSchedule
    from osufusion_b200 import _native as N
    torch.manual_seed(3)
    B, n, Lp = 2, 100, 112
    x = torch.randn(B, 6, n, device=dev) * 1.5
    cond = torch.randn(B, Lp, 8, device=dev).bfloat16()
    null = torch.randn(B, Lp, 8, device=dev).bfloat16()
    c_ref = cond[:, :n, :6].transpose(1, 2)
    n_ref = null[:, :n, :6].transpose(1, 2)
    scale = 2.0
    eps = n_ref + (c_ref - n_ref) * scale                    # bf16 arithmetic, as unet.py:465 under autocast
    sch = DDIMSchedule(1000)
    sch.set_timesteps(35)
    for t in (952, 28, 0):
        ref = sch.step(eps, t, x)
        a_t = float(sch.alphas_cumprod[t])
        tp = t - 1000 // 35
        a_p = float(sch.alphas_cumprod[tp]) if tp >= 0 else 1.0
        out = torch.empty_like(x)
        packed = torch.empty(B, Lp, 8, device=dev, dtype=torch.bfloat16)
        N.call("of_sampler_update", x.data_ptr(), cond.data_ptr(), null.data_ptr(), 8, Lp * 8, scale, 0, (1 - a_t) ** 0.5, a_t ** 0.5,
               a_p ** 0.5, (1 - a_p) ** 0.5, B, 6, n, out.data_ptr(), packed.data_ptr(), Lp, 8, -1.0)
        assert nrel(out, ref) < 2e-3, (t, nrel(out, ref))
        assert nrel(packed[:, :n, :6].transpose(1, 2), ref) < 1e-2
        assert (packed[:, n:, :6] == -1.0).all() and (packed[:, :, 6:] == 0).all()
    dt = 1.0 / 15
    out = torch.empty_like(x)
    N.call("of_sampler_update", x.data_ptr(), cond.data_ptr(), None, 8, Lp * 8, 1.0, 1, dt, 1.0, 0.0, 0.0, B, 6, n, out.data_ptr(), None,
           Lp, 8, -1.0)
    assert nrel(out, x + c_ref * torch.tensor(dt)) < 2e-3


def test_gradient_accumulation_without_zero_grad():
    """`p.grad` aliases the engine's gradient arena after a backward pass; a second backward WITHOUT zero_grad must still
    accumulate (torch semantics: p.grad = g1 + g2), and zero_grad(set_to_none=True) must give back a single gradient."""
    from oracle.synth import TINY, synth_inputs
    _, new = build_pair(TINY, "default")
    x, a, c, t, noise, keep = (v.to(dev) for v in synth_inputs(2, 96, 11))

    def bwd():
        y = new(x, a, t, c, cond_mask=keep)
        torch.nn.functional.mse_loss(y, noise).backward()

    new.zero_grad(set_to_none=True)
    bwd()
    g1 = {k: p.grad.detach().clone() for k, p in new.named_parameters()}
    live = [k for k in g1 if not k.endswith("se.to_k.bias") and g1[k].abs().max() > 1e-9]   # to_k.bias: true gradient is zero
    assert len(live) > 300
    bwd()                                   # no zero_grad: accumulate

    def compare(scale):
        # the engine's fp32 atomics make gradients differ slightly from run to run (cancellation-dominated reductions by up to a
        # few percent): nearly all tensors must agree to 3e-2, none may be off by more than 0.15 -- a missed accumulation is off by 0.5
        errs = {k: nrel(p.grad, scale * g1[k]) for k, p in new.named_parameters() if k in live}
        assert max(errs.values()) < 0.15, max(errs.items(), key=lambda kv: kv[1])
        assert sum(e > 3e-2 for e in errs.values()) <= max(1, len(errs) // 100), sorted(errs.items(), key=lambda kv: -kv[1])[:5]

    compare(2.0)
    new.zero_grad(set_to_none=True)
    bwd()
    compare(1.0)
