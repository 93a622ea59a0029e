"""Debug probe: graphed sampler step vs eager step, step by step (TINY config)."""
import os
import sys

import torch

sys.path.insert(0, ".")
import oracle.synth as S  # noqa: E402
from osufusion_b200 import engine as E  # noqa: E402
from osufusion_b200.models import DiffusionOsuFusion  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
new = DiffusionOsuFusion(**S.TINY).to(dev).eval()
torch.nn.init.normal_(new.unet.final_conv.weight, std=0.02)
_, a, c, _, noise, _ = (v.to(dev) for v in S.synth_inputs(2, 100, 31))


def nrel(u, v):
    return ((u.float() - v.float()).abs().max() / v.float().abs().max().clamp_min(1e-30)).item()


for scale in (2.0, 1.0):
    with torch.inference_mode():
        s = new._sampler_setup(a, c, noise.clone(), scale)
        new.scheduler.set_timesteps(new.sampling_timesteps)
        steps = new.scheduler.timesteps.tolist()
        st, (graph,) = new._sampler_graphs(s, scale, "ddim", [(0, "x16", "x", "x16")])
        tt = torch.tensor(steps, dtype=torch.float32).to(dev)
        coefs = torch.tensor([new.scheduler.step_coeffs(t) for t in steps], dtype=torch.float32).to(dev)
        xcur, x16 = s.x.clone(), s.x16.clone()
        for i, t in enumerate(steps[:6]):
            # eager on the graph's current state
            tb = torch.full((s.b,), t, dtype=torch.int64, device=dev)
            cond16, null16 = new._eval_denoiser(s, st.x16.clone(), tb)
            xe, pe = new._update(s, st.x.clone(), cond16, null16, scale, 0, *new.scheduler.step_coeffs(t))
            st.t_buf.copy_(tt[i].expand(st.b))
            st.coef.copy_(coefs[i])
            graph.replay()
            torch.cuda.synchronize()
            print(f"scale {scale} step {i} t={t}: x graph-vs-eager {nrel(st.x, xe):.3e}  x16 {nrel(st.x16, pe):.3e}", flush=True)
