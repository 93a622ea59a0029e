"""GPU parity on the configurations that bench.py actually measures (VERDICT r1 "What's weak" 1-3):

  * attention kernels at the level-0 shape of the headline step (B=4, H=16, KVH=1, D=64, L=4096, forward + backward) and at the
    long-song shape of config 4 (B=1, L=32 768, forward) against fp32 math on the same bf16 inputs (reference: attention.py:77-101);
  * the whole denoiser at the HEADLINE configuration CFG-L, B=4, N=4096 (all 1239 gradients) and CFG-S, B=1, N=32 768 (forward),
    three arms as in test_model_parity_gpu.py (engine / oracle under bf16 autocast / oracle fp32);
  * the 35-step DDIM loop at cond_scale 2.0 (and the 16-point midpoint loop) step by step: at every step BOTH implementations get
    the oracle's x_t, so the bound is a per-step bound, not a comparison of diverged chains;
  * the CUDA-graph sampler against the eager loop.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def nrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------ attention at benchmarked shapes
def _attn_ref_block(q, k, v, H, D, q0, q1):
    """fp32 attention of query rows q0:q1 of ONE batch item: q (L, H*D), k/v (L, D) bf16 -> out (q1-q0, H*D), lse2 (H, q1-q0)."""
    L = k.shape[0]
    qh = q[q0:q1].float().view(q1 - q0, H, D).transpose(0, 1)            # (H, m, D)
    s = (qh @ k.float().t()) / D ** 0.5                                  # (H, m, L)
    o = s.softmax(-1) @ v.float()                                        # (H, m, D)
    return o.transpose(0, 1).reshape(q1 - q0, H * D), torch.logsumexp(s, -1) * 1.4426950408889634


def test_attention_level0_shape_forward_backward():
    """B=4, H=16, KVH=1, D=64, L=4096: the 64 key-tile loop of the forward kernel and the 32 query-tile loop of the backward kernel."""
    from osufusion_b200 import ops_raw as R
    B, L, H, D = 4, 4096, 16, 64
    torch.manual_seed(11)
    qkv = torch.randn(B, L, (H + 2) * D, device=dev).bfloat16()
    q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:(H + 1) * D], qkv[:, :, (H + 1) * D:]
    out = torch.zeros(B, L, H * D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=dev)
    R.attn_fwd(q, k, v, out, lse, H=H, KVH=1, D=D)
    dout = torch.randn(B, L, H * D, device=dev).bfloat16()
    delta = torch.zeros(B, H, L, device=dev)
    dq = torch.zeros(B, L, H * D, device=dev)
    dkv = torch.zeros(B, L, 2 * D, device=dev)
    R.attn_bwd(q, k, v, out, lse, dout, delta, dq, dkv[:, :, :D], dkv[:, :, D:], H=H, KVH=1, D=D)
    for b in range(B):                                                   # fp32 reference one batch item at a time (1 GiB of scores)
        qf, kf, vf = (t[b].float().detach().clone().requires_grad_(True) for t in (q, k, v))
        qh = qf.view(L, H, D).transpose(0, 1)
        s = (qh @ kf.t()) / D ** 0.5
        o_ref = (s.softmax(-1) @ vf).transpose(0, 1).reshape(L, H * D)
        lse_ref = torch.logsumexp(s.detach(), -1) * 1.4426950408889634
        o_ref.backward(dout[b].float())
        assert nrel(out[b], o_ref) < 1e-2 and nrel(lse[b], lse_ref) < 1e-4, b
        assert nrel(dq[b], qf.grad) < 1e-2 and nrel(dkv[b, :, :D], kf.grad) < 1e-2 and nrel(dkv[b, :, D:], vf.grad) < 1e-2, b
        del s, o_ref, qh
    torch.cuda.empty_cache()


def test_attention_long_song_forward():
    """B=1, L=32 768 (config 4: 98 % of its FLOPs): 512 key tiles per query tile; fp32 reference chunked over query blocks."""
    from osufusion_b200 import ops_raw as R
    B, L, H, D = 1, 32768, 16, 64
    torch.manual_seed(12)
    qkv = torch.randn(B, L, (H + 2) * D, device=dev).bfloat16()
    q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:(H + 1) * D], qkv[:, :, (H + 1) * D:]
    out = torch.zeros(B, L, H * D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=dev)
    R.attn_fwd(q, k, v, out, lse, H=H, KVH=1, D=D)
    worst_o = worst_l = 0.0
    scale_o = 0.0
    for q0 in range(0, L, 2048):
        o_ref, lse_ref = _attn_ref_block(q[0], k[0], v[0], H, D, q0, q0 + 2048)
        worst_o = max(worst_o, (out[0, q0:q0 + 2048].float() - o_ref).abs().max().item())
        scale_o = max(scale_o, o_ref.abs().max().item())
        worst_l = max(worst_l, nrel(lse[0, :, q0:q0 + 2048], lse_ref))
    assert worst_o / scale_o < 1e-2 and worst_l < 1e-4, (worst_o / scale_o, worst_l)
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ whole model at benchmarked shapes
def test_cfg_l_headline_batch4_frames4096_forward_backward():
    """THE benchmarked configuration (bench.py default): CFG-L dim_h=512, 1.28 B parameters, batch 4 x 4096 frames, all gradients."""
    from oracle.synth import LARGE
    from test_model_parity_gpu import check
    check(LARGE, 4, 4096, "default", True)
    torch.cuda.empty_cache()


def test_cfg_s_long_song_forward():
    """Config 4's evaluation shape: CFG-S (dim_h=128), one song of 32 768 frames, forward only (inference)."""
    from oracle.synth import SMALL, synth_inputs
    from test_model_parity_gpu import build_pair
    ora, new = build_pair(SMALL, "default")
    x, a, c, t, _, _ = (v.to(dev) for v in synth_inputs(1, 32768, 77))
    with torch.no_grad():
        y_new = new(x, a, t, c)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y_ref = ora(x, a, t, c).float()
        y_tru = ora(x, a, t, c)
    assert y_tru.abs().max() > 1e-3
    assert nrel(y_new, y_tru) <= max(1e-2, 2 * nrel(y_ref, y_tru)), (nrel(y_new, y_tru), nrel(y_ref, y_tru))
    del ora, new
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ sampler, step by step
def _sampler_pair(OC, NC, cfg, **kw):
    torch.manual_seed(0)
    ora = OC(**cfg, **kw)
    torch.nn.init.normal_(ora.unet.final_conv.weight, std=0.02)
    new = NC(**cfg, **kw)
    new.load_state_dict(ora.state_dict())
    return ora.to(dev).eval(), new.to(dev).eval()


@pytest.mark.parametrize("cfg_name,n", [("TINY", 100), ("SMALL", 1024)])
def test_ddim_35_steps_per_step_bound(cfg_name, n):
    """diffusion.py:59-77 with the reference defaults (35 steps) at cond_scale 2.0: at EVERY step the engine's fused
    [cond; null] evaluation + CFG + DDIM update, fed with the oracle's x_t, must match the oracle's step on the same x_t:
    err(new, truth) <= max(3e-2, 2 * err(ref_bf16, truth)), truth = the fp32 oracle step."""
    import oracle.synth as S
    from oracle.models import DiffusionOsuFusion as OracleModel
    from osufusion_b200.models import DiffusionOsuFusion
    from osufusion_b200.modules import X_PAD_VALUE, _pack
    ora, new = _sampler_pair(OracleModel, DiffusionOsuFusion, getattr(S, cfg_name))
    _, a, c, _, noise, _ = (v.to(dev) for v in S.synth_inputs(2, n, 21))
    scale = 2.0
    with torch.inference_mode():
        s = new._sampler_setup(a, c, noise.clone(), scale)
        new.scheduler.set_timesteps(new.sampling_timesteps)
        ora.scheduler.set_timesteps(ora.sampling_timesteps)
        steps = ora.scheduler.timesteps.tolist()
        assert len(steps) == 35 and steps[0] == 952 and steps[-1] == 0
        x = noise.clone()
        worst = (0.0, 0.0, -1)
        for t in steps:
            tb = torch.full((2,), t, dtype=torch.int64, device=dev)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                x_ref = ora.scheduler.step(ora.unet.forward_with_cond_scale(x, a, tb, c, cond_scale=scale), t, x)
            x_tru = ora.scheduler.step(ora.unet.forward_with_cond_scale(x, a, tb, c, cond_scale=scale), t, x)
            cond16, null16 = new._eval_denoiser(s, _pack(x, 8, s.Lp, X_PAD_VALUE), tb)
            x_new, _ = new._update(s, x, cond16, null16, scale, 0, *new.scheduler.step_coeffs(t))
            e_new, e_ref = nrel(x_new, x_tru), nrel(x_ref.float(), x_tru)
            assert e_new <= max(3e-2, 2 * e_ref), (t, e_new, e_ref)
            if e_new > worst[0]:
                worst = (e_new, e_ref, t)
            x = x_ref.float()
    print(f"DDIM per-step worst err {worst[0]:.2e} (reference bf16 {worst[1]:.2e}) at t={worst[2]}")


def test_midpoint_16_points_per_step_bound():
    """rectified_flow.py:57-79 (torchdiffeq fixed-grid midpoint over linspace(0,1,16)): both half-steps of every interval."""
    import oracle.synth as S
    from oracle.models import RectifiedFlowOsuFusion as OracleRF
    from osufusion_b200.models import RectifiedFlowOsuFusion
    from osufusion_b200.modules import X_PAD_VALUE, _pack
    ora, new = _sampler_pair(OracleRF, RectifiedFlowOsuFusion, S.TINY)
    _, a, c, _, noise, _ = (v.to(dev) for v in S.synth_inputs(2, 100, 22))
    scale = 2.0
    times = torch.linspace(0.0, 1.0, 16)
    with torch.inference_mode():
        s = new._sampler_setup(a, c, noise.clone(), scale)
        y = noise.clone()
        for t0, t1 in zip(times[:-1].tolist(), times[1:].tolist()):
            dt = t1 - t0
            for tt, base, h in ((t0, None, 0.5 * dt), (t0 + 0.5 * dt, "mid", dt)):
                tb = torch.full((2,), tt, dtype=torch.float32, device=dev)
                xin = y if base is None else ymid
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    f_ref = ora.unet.forward_with_cond_scale(xin, a, tb, c, cond_scale=scale)
                f_tru = ora.unet.forward_with_cond_scale(xin, a, tb, c, cond_scale=scale)
                o_ref, o_tru = y + f_ref * h, y + f_tru * h
                cond16, null16 = new._eval_denoiser(s, _pack(xin, 8, s.Lp, X_PAD_VALUE), tb)
                o_new, _ = new._update(s, y, cond16, null16, scale, 1, h)
                e_new, e_ref = nrel(o_new, o_tru), nrel(o_ref.float(), o_tru)
                assert e_new <= max(3e-2, 2 * e_ref), (tt, e_new, e_ref)
                if base is None:
                    ymid = o_ref.float()
                else:
                    y = o_ref.float()


@pytest.mark.parametrize("kind", ["ddim", "midpoint"])
def test_graphed_sampler_step_matches_eager_step(kind):
    """One CUDA graph per sampler step (timestep / coefficients as device data) vs the eager launch sequence of the same kernels,
    step by step on the graph's own state (two runs of the engine differ by 1-2 bf16 ulps of the output, which a 35-step CFG chain
    amplifies, so whole chains are not comparable); the second `sample()` call (different inputs, cached graph) must refresh
    every static buffer; and the complete graphed loop stays within the loose trajectory bound of the eager loop."""
    import oracle.synth as S
    from osufusion_b200.models import DiffusionOsuFusion, RectifiedFlowOsuFusion
    NC = DiffusionOsuFusion if kind == "ddim" else RectifiedFlowOsuFusion
    torch.manual_seed(0)
    new = NC(**S.TINY).to(dev).eval()
    torch.nn.init.normal_(new.unet.final_conv.weight, std=0.02)
    scale = 2.0
    for seed in (31, 32):
        _, a, c, _, noise, _ = (v.to(dev) for v in S.synth_inputs(2, 100, seed))
        with torch.inference_mode():
            s = new._sampler_setup(a, c, noise.clone(), scale)
            if kind == "ddim":
                new.scheduler.set_timesteps(new.sampling_timesteps)
                steps = new.scheduler.timesteps.tolist()
                st, (graph,) = new._sampler_graphs(s, scale, "ddim", [(0, "x16", "x", "x16")])
                assert torch.equal(st.x, s.x) and torch.equal(st.x16, s.x16)          # static buffers refreshed from THIS call's inputs
                for t in steps:
                    coef = new.scheduler.step_coeffs(t)
                    tb = torch.full((s.b,), t, dtype=torch.int64, device=dev)
                    cond16, null16 = new._eval_denoiser(s, st.x16.clone(), tb)
                    x_e, p_e = new._update(s, st.x.clone(), cond16, null16, scale, 0, *coef)
                    st.t_buf.fill_(float(t))
                    st.coef.copy_(torch.tensor(coef, dtype=torch.float32))
                    graph.replay()
                    assert nrel(st.x, x_e) < 3e-2 and nrel(st.x16, p_e) < 3e-2, (seed, t, nrel(st.x, x_e))
            else:
                times = torch.linspace(0.0, 1.0, new.sample_timesteps)
                st, (g_half, g_full) = new._sampler_graphs(s, scale, "midpoint", [(1, "x16", "xtmp", "xmid16"), (1, "xmid16", "x", "x16")])
                assert torch.equal(st.x, s.x) and torch.equal(st.x16, s.x16)
                for t0, t1 in zip(times[:-1].tolist(), times[1:].tolist()):
                    dt = t1 - t0
                    for tt_, h, src, graph, dst in ((t0, 0.5 * dt, "x16", g_half, "xmid16"), (t0 + 0.5 * dt, dt, "xmid16", g_full, "x16")):
                        tb = torch.full((s.b,), tt_, dtype=torch.float32, device=dev)
                        cond16, null16 = new._eval_denoiser(s, getattr(st, src).clone(), tb)
                        x_e, p_e = new._update(s, st.x.clone(), cond16, null16, scale, 1, h)
                        st.t_buf.fill_(tt_)
                        st.coef.copy_(torch.tensor([h, 1.0, 0.0, 0.0]))
                        graph.replay()
                        got_x = st.xtmp if dst == "xmid16" else st.x
                        assert nrel(got_x, x_e) < 3e-2 and nrel(getattr(st, dst), p_e) < 3e-2, (seed, tt_)
        y_graph = new.sample(a, c, noise.clone(), cond_scale=scale)
        os.environ["OF_SAMPLER_GRAPH"] = "0"
        try:
            y_eager = new.sample(a, c, noise.clone(), cond_scale=scale)
        finally:
            os.environ.pop("OF_SAMPLER_GRAPH", None)
        # whole chains are not comparable element-wise (see above): shape, finiteness and the same overall statistics
        assert y_graph.shape == y_eager.shape and torch.isfinite(y_graph).all()
        assert abs(y_graph.abs().mean().item() - y_eager.abs().mean().item()) < 0.25 * y_eager.abs().mean().item()
    assert len(new._sgraphs) == 1
