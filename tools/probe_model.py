"""GPU probe: whole-model forward/backward parity of the CUDA engine vs the oracle (bf16 autocast) and fp32 truth."""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle.denoiser import UNet as OracleUNet  # noqa: E402
from oracle.synth import LARGE, SMALL, TINY, synth_inputs, synth_state_dict  # noqa: E402
from osufusion_b200.modules import UNet  # noqa: E402

dev = "cuda"


def nrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def run(cfg_name, B, n, drop=0.0, check_grads=True):
    cfg = dict(TINY=TINY, SMALL=SMALL, LARGE=LARGE)[cfg_name]
    ora = OracleUNet(6, 96, 5, **cfg)
    sd = synth_state_dict(ora)
    ora.load_state_dict(sd)
    ora = ora.to(dev)
    new = UNet(6, 96, 5, **cfg)
    new.load_state_dict(sd)
    new = new.to(dev)
    x, a, c, t, noise, mask = (v.to(dev) for v in synth_inputs(B, n, 1234))
    keep = torch.ones(B, dtype=torch.bool, device=dev) if drop == 0.0 else mask

    def fwd_bwd(model, autocast):
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = model(x, a, t, c, cond_mask=keep)
        loss = torch.nn.functional.mse_loss(y.float(), noise)
        loss.backward()
        return y.detach().float(), {k: p.grad.detach().float().clone() for k, p in model.named_parameters() if p.grad is not None}

    t0 = time.time()
    y_new, g_new = fwd_bwd(new, False)
    torch.cuda.synchronize()
    print(f"[{cfg_name} B{B} n{n}] engine fwd+bwd ok in {time.time() - t0:.2f}s", flush=True)
    y_ref, g_ref = fwd_bwd(ora, True)
    y_tru, g_tru = fwd_bwd(ora, False)
    e_new, e_ref = nrel(y_new, y_tru), nrel(y_ref, y_tru)
    print(f"  output: err(new,truth)={e_new:.3e} err(ref_bf16,truth)={e_ref:.3e} err(new,ref)={nrel(y_new, y_ref):.3e}", flush=True)
    if not check_grads:
        return
    missing = set(g_tru) - set(g_new)
    print(f"  grads: {len(g_new)} produced, {len(missing)} missing {sorted(missing)[:5]}")
    rows = []
    for k in g_tru:
        if k in g_new:
            rows.append((nrel(g_new[k], g_tru[k]), nrel(g_ref[k], g_tru[k]), k))
    rows.sort(reverse=True)
    bad = [r for r in rows if r[0] > max(1e-2, 2 * r[1])]
    print(f"  worst new-vs-truth grads (new, ref_bf16, name); {len(bad)} exceed max(1e-2, 2*ref):")
    for r in rows[:12]:
        print(f"    {r[0]:.3e} {r[1]:.3e} {r[2]}")
    import statistics
    print(f"  median err new {statistics.median(r[0] for r in rows):.3e} ref {statistics.median(r[1] for r in rows):.3e}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    if which == "tiny":
        run("TINY", 2, 64)
        run("TINY", 2, 200, drop=0.5)
    elif which == "small":
        run("SMALL", 2, 1024)
    elif which == "large":
        run("LARGE", 1, 1024)
