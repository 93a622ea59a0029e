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


def run(cfg_name, B, n, drop=0.0, check_grads=True, init="synth"):
    cfg = dict(TINY=TINY, SMALL=SMALL, LARGE=LARGE)[cfg_name]
    torch.manual_seed(0)
    ora = OracleUNet(6, 96, 5, **cfg)
    if init == "synth":
        sd = synth_state_dict(ora)
        ora.load_state_dict(sd)
    else:  # torch default init (what the trainer starts from) with final_conv re-randomised
        torch.nn.init.normal_(ora.final_conv.weight, std=0.02)
        sd = {k: v.clone() for k, v in ora.state_dict().items()}
    print(f"--- init={init}")
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
        if k in g_new and not k.endswith("se.to_k.bias"):  # softmax shift invariance: the true gradient is exactly 0
            rows.append((nrel(g_new[k], g_tru[k]), nrel(g_ref[k], g_tru[k]), nrel(g_new[k], g_ref[k]), k))
    rows.sort(reverse=True)
    bad = [r for r in rows if r[0] > max(1e-2, 2 * r[1])]
    print(f"  worst new-vs-truth grads (new|truth, ref_bf16|truth, new|ref, name); {len(bad)} exceed max(1e-2, 2*ref):")
    for r in rows[:8]:
        print(f"    {r[0]:.3e} {r[1]:.3e} {r[2]:.3e} {r[3]}")
    import statistics
    print(f"  median: new|truth {statistics.median(r[0] for r in rows):.3e} ref|truth {statistics.median(r[1] for r in rows):.3e} "
          f"new|ref {statistics.median(r[2] for r in rows):.3e}; max new|ref {max(r[2] for r in rows):.3e}", flush=True)
    zk = [k for k in g_new if k.endswith("se.to_k.bias")]
    print(f"  to_k.bias grads (true value 0): max |new| {max(g_new[k].abs().max().item() for k in zk):.2e} max |ref| {max(g_ref[k].abs().max().item() for k in zk):.2e}")


def timeit(cfg_name, B, n):
    from osufusion_b200.models import DiffusionOsuFusion
    from oracle.models import DiffusionOsuFusion as OracleModel
    from osufusion_b200 import _native as NN
    cfg = dict(TINY=TINY, SMALL=SMALL, LARGE=LARGE)[cfg_name]
    x, a, c, t, noise, mask = (v.to(dev) for v in synth_inputs(B, n, 1234))
    for name, cls, ac in (("engine", DiffusionOsuFusion, False), ("oracle-eager-bf16", OracleModel, True)):
        torch.manual_seed(0)
        model = cls(**cfg).to(dev)
        torch.nn.init.normal_(model.unet.final_conv.weight, std=0.02)

        def step():
            model.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                loss = model(x, a, c)
            loss.backward()
            return loss
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        NN.lib().of_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 5
        for _ in range(K):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"[{cfg_name} B{B} n{n}] {name}: {ms:.1f} ms/step  {B / ms * 1e3:.2f} samples/s  loss {loss.item():.4f} "
              f"launches/step {NN.lib().of_launch_count() / K:.0f}  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        if name == "engine":
            from osufusion_b200.graphs import GraphedTrainStep
            del loss  # a live autograd graph pins AccumulateGrad nodes to the eager stream and breaks capture
            model.zero_grad(set_to_none=True)
            g = GraphedTrainStep(model, x, a, c)
            for _ in range(3):
                g()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(K):
                loss = g()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            gn = sum(float(p.grad.float().norm() ** 2) for p in model.parameters() if p.grad is not None) ** 0.5
            print(f"[{cfg_name} B{B} n{n}] engine+cudagraph: {ms:.1f} ms/step  {B / ms * 1e3:.2f} samples/s  loss {loss.item():.4f} gradnorm {gn:.4f}", flush=True)
            del g
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    if which == "tiny":
        run("TINY", 2, 64)
        run("TINY", 2, 200, drop=0.5)
        run("TINY", 2, 200, drop=0.5, init="default")
    elif which == "small":
        run("SMALL", 2, 1024)
        run("SMALL", 2, 4096, drop=0.5, init="default")
    elif which == "large":
        run("LARGE", 2, 4096)
    elif which == "time":
        timeit(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
