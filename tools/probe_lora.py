import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_lora_gpu import _build, nrel, dev
from oracle.synth import synth_inputs
for use_dora in (True, False):
    ora, new, names, leaves = _build(use_dora)
    x, a, c, t, noise, keep = (v.to(dev) for v in synth_inputs(2, 120, 5))
    def run_oracle(autocast):
        for A, Bm, mag in leaves.values():
            for v in (A, Bm, mag):
                if v is not None: v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = ora(x, a, t, c, cond_mask=keep)
        torch.nn.functional.mse_loss(y.float(), noise).backward()
        return y.detach().float(), {n: tuple(None if v is None else v.grad.detach().clone() for v in leaves[n]) for n in names}
    y_ref, g_ref = run_oracle(True); y_tru, g_tru = run_oracle(False)
    new.zero_grad(set_to_none=True)
    y_new = new(x, a, t, c, cond_mask=keep)
    torch.nn.functional.mse_loss(y_new, noise).backward()
    print("dora", use_dora, "out", nrel(y_new, y_tru), nrel(y_ref, y_tru))
    for n in names[:14]:
        ad = new.get_submodule(n)
        mine = (ad.lora_A["default"].weight.grad, ad.lora_B["default"].weight.grad, ad.magnitude().grad if use_dora else None)
        row = []
        for which, gm, gr, gt in zip("ABm", mine, g_ref[n], g_tru[n]):
            if gt is None: continue
            row.append(f"{which}: new {nrel(gm.view(gt.shape), gt):.3f} ref {nrel(gr, gt):.3f} |gt| {gt.abs().max().item():.2e}")
        print(f"  {n:55s} " + " | ".join(row))
