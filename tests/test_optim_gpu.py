"""Fused clip + AdamW step against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW on the same
parameters and gradients (the reference's optimizer tail, trainer.py:302-309)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def _tiny():
    from oracle.synth import TINY
    from osufusion_b200.modules import UNet
    torch.manual_seed(0)
    net = UNet(6, 96, 5, **TINY).to(dev)
    torch.nn.init.normal_(net.final_conv.weight, std=0.02)
    return net


@pytest.mark.parametrize("max_norm", [1.0, None])
def test_fused_adamw_matches_torch(max_norm):
    from oracle.synth import synth_inputs
    from osufusion_b200.optim import FusedAdamW, cosine_schedule_with_warmup
    net = _tiny()
    ref = copy.deepcopy(net)
    opt = FusedAdamW(net, lr=1e-3, weight_decay=1e-2, max_grad_norm=max_norm)
    sched = cosine_schedule_with_warmup(opt, 2, 10)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=1e-2)
    rsched = cosine_schedule_with_warmup(ropt, 2, 10)
    x, a, c, t, noise, keep = (v.to(dev) for v in synth_inputs(2, 64, 7))
    for it in range(4):
        net.zero_grad(set_to_none=True)
        y = net(x, a, t, c, cond_mask=keep)
        torch.nn.functional.mse_loss(y, noise).backward()
        # feed the torch optimizer the SAME gradients (the engine's bf16 gradients are not bit-reproducible run to run)
        for (_, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            q.grad = p.grad.detach().clone()
        if max_norm is not None:
            tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        ropt.step()
        rsched.step()
        opt.step()
        sched.step()
        if max_norm is not None:
            assert abs(float(opt.grad_norm) - float(tn)) <= 1e-4 * float(tn)
        worst = max(((p - q).abs().max() / q.abs().max().clamp_min(1e-12)).item()
                    for p, q in zip(net.parameters(), ref.parameters()))
        assert worst < 1e-5, (it, worst)
    # the engine must see the updated weights: its output now differs from the initial one and equals a fresh model's
    with torch.no_grad():
        y1 = net(x, a, t, c, cond_mask=keep)
        fresh = _tiny()
        fresh.load_state_dict(net.state_dict())
        y2 = fresh(x, a, t, c, cond_mask=keep)
    assert torch.allclose(y1, y2, rtol=0, atol=2e-2 * y2.abs().max().item())


def test_train_loop_loss_decreases():
    """A few optimizer steps of the whole training wrapper on a fixed batch: the loss must go down."""
    from oracle.synth import TINY
    from osufusion_b200.models import DiffusionOsuFusion
    from osufusion_b200.optim import FusedAdamW
    torch.manual_seed(0)
    model = DiffusionOsuFusion(96, dim_h_mult=TINY["dim_h_mult"], num_layer_blocks=TINY["num_layer_blocks"],
                               num_middle_transformers=TINY["num_middle_transformers"], attn_dim_head=TINY["attn_dim_head"],
                               attn_heads=TINY["attn_heads"]).to(dev)
    torch.nn.init.normal_(model.unet.final_conv.weight, std=0.02)
    opt = FusedAdamW(model, lr=2e-3)
    g = torch.Generator(device="cpu").manual_seed(1)
    x, a, c = torch.randn(2, 6, 128, generator=g).to(dev), torch.randn(2, 96, 128, generator=g).to(dev), torch.randn(2, 5, generator=g).to(dev)
    noise = torch.randn(2, 6, 128, generator=g).to(dev)
    ts = torch.tensor([100, 700], device=dev)
    keep = torch.tensor([True, False], device=dev)
    losses = []
    for _ in range(12):
        model.zero_grad(set_to_none=True)
        loss = model(x, a, c, noise=noise, timesteps=ts, cond_mask=keep)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.7 * losses[0], losses
