"""Checkpoint layout parity: the files this repo writes have the reference's names / keys / shapes and
round-trip; a state_dict produced by the oracle (= the reference's module tree) loads into the engine's module tree unchanged."""
import copy

import torch
from safetensors.torch import load_file

from oracle.synth import TINY
from osufusion_b200 import checkpoint as ck
from osufusion_b200 import lora
from osufusion_b200.models import DiffusionOsuFusion


def _model():
    torch.manual_seed(0)
    return DiffusionOsuFusion(96, dim_h_mult=TINY["dim_h_mult"], num_layer_blocks=TINY["num_layer_blocks"],
                              num_middle_transformers=TINY["num_middle_transformers"], attn_dim_head=TINY["attn_dim_head"],
                              attn_heads=TINY["attn_heads"])


def test_model_safetensors_keys_match_reference_layout(tmp_path):
    from oracle.models import DiffusionOsuFusion as OracleModel
    m = _model()
    ck.save_model_sd(m, tmp_path)
    sd = load_file(str(tmp_path / "model.safetensors"))
    ora = OracleModel(96, dim_h_mult=TINY["dim_h_mult"], num_layer_blocks=TINY["num_layer_blocks"],
                      num_middle_transformers=TINY["num_middle_transformers"], attn_dim_head=TINY["attn_dim_head"],
                      attn_heads=TINY["attn_heads"])
    ref = ora.state_dict()
    assert set(sd) == set(ref) and all(k.startswith("unet.") for k in sd)
    assert all(sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype for k in ref)
    ora.load_state_dict(sd)                 # our file loads into the reference's module tree ...
    m2 = _model()
    ck.load_model(m2, tmp_path / "model.safetensors")
    m2.load_state_dict(ora.state_dict())    # ... and the reference's state_dict loads into ours
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_checkpoint_pt_roundtrip(tmp_path):
    m = _model()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-5)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    sched.step()
    d = ck.save_checkpoint(m, opt, sched, 41, tmp_path)
    assert d.name == "checkpoint-42" and (d / "checkpoint.pt").exists()
    raw = torch.load(d / "checkpoint.pt", weights_only=False)
    assert set(raw) == {"model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "rng_state"}
    m2 = _model()
    with torch.no_grad():
        for p in m2.parameters():
            p.add_(1.0)
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-5)
    sched2 = torch.optim.lr_scheduler.LambdaLR(opt2, lambda s: 1.0 / (1 + s))
    step = ck.load_checkpoint(m2, opt2, sched2, d)
    assert step == 42 and sched2.last_epoch == sched.last_epoch
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert ck.get_latest_checkpoint(tmp_path) == d
    assert ck.load_checkpoint(m2, opt2, sched2, d, reset_steps=True) == 0
    ck.load_model(m2, d / "checkpoint.pt")


def test_adapter_directory_roundtrip_and_merged_model(tmp_path):
    m = _model()
    base_keys = set(m.state_dict())
    names = lora.inject_adapters(m, r=8, lora_alpha=8, use_dora=True)
    with torch.no_grad():
        for n in names:
            ad = m.get_submodule(n)
            ad.lora_B["default"].weight.normal_(std=0.05)
            ad.magnitude().mul_(1.05)
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-5)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    d = ck.save_peft_checkpoint(m, opt, sched, 9, tmp_path, r=8, lora_alpha=8)
    assert d == tmp_path / "loras" / "checkpoint-10"
    ad_sd = load_file(str(d / "adapter_model.safetensors"))
    assert len(ad_sd) == 3 * len(names)
    assert all(k.startswith("base_model.model.unet.") and ".default" not in k for k in ad_sd)
    assert any(k.endswith("lora_magnitude_vector.weight") for k in ad_sd)
    assert set(torch.load(d / "checkpoint.pt", weights_only=False)) == {"optimizer_state_dict", "scheduler_state_dict", "rng_state"}
    m2 = _model()
    lora.inject_adapters(m2, r=8, lora_alpha=8, use_dora=True)
    cfg = ck.load_adapter(m2, d)
    assert cfg["r"] == 8 and cfg["use_dora"] is True
    for n in names:
        a, b = m.get_submodule(n), m2.get_submodule(n)
        assert torch.equal(a.lora_B["default"].weight, b.lora_B["default"].weight)
        assert torch.equal(a.magnitude(), b.magnitude())
    assert ck.load_peft_checkpoint(opt, sched, d, reset_steps=False) == 10
    assert ck.get_latest_checkpoint(tmp_path, "loras") == d
    # merged model: plain key set again, effective weights folded in
    ad0 = m.get_submodule(names[0])
    w_eff = ad0.effective_weight().detach().clone() if hasattr(ad0, "effective_weight") else None
    ck.save_merged_model_sd(copy.deepcopy(m), tmp_path)
    merged = load_file(str(tmp_path / "merged_model.safetensors"))
    assert set(merged) == base_keys
    if w_eff is not None:
        assert torch.allclose(merged[names[0] + ".weight"], w_eff, atol=1e-5)


def test_fused_adamw_state_dict_is_torch_format_both_directions(tmp_path):
    """optimizer checkpoints must move between `torch.optim.AdamW(model.parameters())` (what the reference saves,
    trainer.py:230,167) and FusedAdamW in both directions: per-parameter state in model.parameters() order, moments scattered to /
    gathered from the arenas.  Host logic only (no kernel launch)."""
    from oracle.models import DiffusionOsuFusion as OracleModel
    from osufusion_b200.optim import FusedAdamW
    m = _model()
    ref = OracleModel(96, dim_h_mult=TINY["dim_h_mult"], num_layer_blocks=TINY["num_layer_blocks"],
                      num_middle_transformers=TINY["num_middle_transformers"], attn_dim_head=TINY["attn_dim_head"],
                      attn_heads=TINY["attn_heads"])
    ref.load_state_dict(m.state_dict())
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-5)
    g = torch.Generator().manual_seed(3)
    for _ in range(2):
        for p in ref.parameters():
            p.grad = torch.randn(p.shape, generator=g)
        ropt.step()
    sd = ropt.state_dict()
    opt = FusedAdamW(m, lr=1e-5)
    assert opt.state_dict()["state"] == {}                      # like torch before the first step
    opt.load_state_dict(copy.deepcopy(sd))                       # reference -> ours
    assert opt._step == 2
    for p, q in zip(m.parameters(), ref.parameters()):
        mv, vv = opt._moment_views(p)
        assert torch.equal(mv, ropt.state[q]["exp_avg"]) and torch.equal(vv, ropt.state[q]["exp_avg_sq"])
    out = opt.state_dict()                                       # ours -> reference
    assert out["param_groups"][0]["params"] == sd["param_groups"][0]["params"]
    ropt2 = torch.optim.AdamW(ref.parameters(), lr=1e-5)
    ropt2.load_state_dict(out)
    for q in ref.parameters():
        assert torch.equal(ropt2.state[q]["exp_avg"], ropt.state[q]["exp_avg"])
        assert torch.equal(ropt2.state[q]["exp_avg_sq"], ropt.state[q]["exp_avg_sq"])
        assert float(ropt2.state[q]["step"]) == 2.0
    # through the reference's checkpoint.pt layout, with the ValueError / RuntimeError fallback exercised by a wrong-size optimizer dict
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    d = ck.save_checkpoint(m, opt, sched, 4, tmp_path)
    m2 = _model()
    opt2 = FusedAdamW(m2, lr=1e-5)
    assert ck.load_checkpoint(m2, opt2, torch.optim.lr_scheduler.LambdaLR(opt2, lambda s: 1.0), d) == 5
    assert opt2._step == 2 and torch.equal(opt2.exp_avg, opt.exp_avg)
    # PEFT: the reference's AdamW covers ALL parameters (frozen ones carry no state); ours covers the adapter tensors
    m3 = _model()
    lora.inject_adapters(m3, r=4, lora_alpha=4, use_dora=True)
    ropt3 = torch.optim.AdamW(m3.parameters(), lr=1e-5)
    for p in m3.parameters():
        if p.requires_grad:
            p.grad = torch.randn(p.shape, generator=g)
    ropt3.step()
    opt3 = FusedAdamW(m3, lr=1e-5)
    opt3.load_state_dict(copy.deepcopy(ropt3.state_dict()))
    assert opt3._step == 1
    for p in m3.parameters():
        if p.requires_grad:
            assert torch.equal(opt3._moment_views(p)[0], ropt3.state[p]["exp_avg"])
