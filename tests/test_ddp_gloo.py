"""N>1 host logic on CPU: world_size-2 gloo runs of the gradient arena / bucket scheduler (osufusion_b200/ddp.py) — once with a
hand-driven stand-in for the backward tape, once with the REAL engine tape executed through the emulated C-ABI (tests/fake_native.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.synth import TINY
    from osufusion_b200.ddp import GradAllReducer, backward_param_order
    from osufusion_b200.modules import UNet

    torch.manual_seed(0)
    net = UNet(6, 96, 5, **TINY)
    order = backward_param_order(net)
    assert len(order) == len(list(net.parameters())) and len({id(p) for p in order}) == len(order)
    assert order[0] is net.final_conv.weight          # first gradient to complete in backward
    red = GradAllReducer(net, bucket_bytes=1 << 20)   # ~20 buckets on the tiny config
    assert len(red.buckets) > 4
    st = net._store
    n_ops = 40
    for step in range(2):                              # step 0 learns bucket readiness, step 1 overlaps
        st.begin_backward(net, False)
        net.grad_sync(n_ops + 1)
        per = max(1, len(order) // n_ops)
        k = 0
        for op in range(n_ops, -1, -1):                # emulate the backward tape: each op finalises a slice of params
            for p in order[k:k + per] if op > 0 else order[k:]:
                g = st.grad(p)
                g.add_(float(rank + 1) * (1 + step))
            k += per
            net.grad_sync(op)
        net.grad_finish()
        expect = (1 + step) * sum(r + 1 for r in range(world)) / world
        plist = list(net.parameters())
        for p_ in plist:
            p_.grad = None
        # packed conv gradients are assigned to `.grad` by take_grads itself (autograd would re-lay them out); the rest is returned
        grads = [g if g is not None else p_.grad for g, p_ in zip(st.take_grads(plist), plist)]
        assert all(torch.allclose(g, torch.full_like(g, expect)) for g in grads), (rank, step)
        if step == 1:
            assert red.ready_at is not None and len(red._launched) == len(red.buckets)
    out.put((rank, "ok", len(red.buckets)))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] == "ok" for r in res)


def _worker_real_tape(rank, world, port, out, kind="unet", prescale=None):
    """Every rank runs the real engine forward/backward (through the emulated C-ABI) on its own batch with the bucketed all-reduce
    hooked into the tape; the result must equal the mean of the per-rank gradients computed without any reducer — in the learning
    step (all buckets reduced at the end) and in the overlapped step (buckets launched from inside backward)."""
    import copy

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import fake_native
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200 import _native
    from osufusion_b200.ddp import GradAllReducer
    from osufusion_b200.modules import UNet, UNetFunction
    _native.call = fake_native.call

    torch.manual_seed(0)
    if kind == "unet":
        base = UNet(6, 96, 5, **TINY)
        torch.nn.init.normal_(base.final_conv.weight, std=0.02)
    else:                                   # the DiT / MMDiT backbones share the arena / hook protocol (osufusion_b200/backbones.py)
        from oracle.make_golden_backbones import MMDIT_TINY
        from oracle.synth import synth_state_dict
        from osufusion_b200.backbones import MMDiT
        base = MMDiT(6, 96, 5, **MMDIT_TINY)
        base.load_state_dict(synth_state_dict(base))

    def grads_of(net, seed):
        x, a, c, t, noise, keep = synth_inputs(2, 48, seed)
        net.zero_grad(set_to_none=True)
        y = UNetFunction.apply(net, x, a, t, c, keep, *list(net.parameters()))
        torch.nn.functional.mse_loss(y, noise).backward()
        return [p.grad.detach().clone() for p in net.parameters() if p.grad is not None]

    per_rank = [grads_of(copy.deepcopy(base), 100 + r) for r in range(world)]
    expect = [sum(gs) / world for gs in zip(*per_rank)]
    net = copy.deepcopy(base)
    red = GradAllReducer(net, bucket_bytes=1 << 18, prescale=prescale)
    assert len(red.buckets) > 4
    # mean by pre-division (loss gradient x 1/world, buckets reduced with SUM: the NVLS-capable form used on NCCL) or, by default on
    # gloo, SUM followed by a division — both must give the mean of the per-rank gradients
    assert red.prescale == bool(prescale) and net.grad_prescale == (1.0 / world if prescale else 1.0)
    for step in range(2):
        got = grads_of(net, 100 + rank)
        assert len(got) == len(expect)
        worst = max(((g - e).abs().max() / e.abs().max().clamp_min(1e-12)).item() for g, e in zip(got, expect))
        assert worst < 1e-5, (rank, step, worst)
        if step == 0:
            assert red.ready_at is not None
        else:
            assert len(red._launched) == len(red.buckets)
    out.put((rank, "ok", len(red.buckets)))
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("kind,prescale", [("unet", None), ("mmdit", None), ("unet", True)])
def test_real_backward_tape_allreduce_world2_gloo(kind, prescale):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_real_tape, args=(r, 2, port, out, kind, prescale)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] == "ok" for r in res)


def _worker_replica_sync(rank, world, port, out):
    """(a) replicas built from DIFFERENT seeds must be identical after GradAllReducer construction (DDP's constructor
    broadcast); (b) a change of the trainable set (adapter injection) rebuilds the arena — buckets and learned readiness must
    follow it instead of slicing the new arena with stale offsets."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import fake_native
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200 import _native, lora
    from osufusion_b200.ddp import GradAllReducer
    from osufusion_b200.modules import UNet, UNetFunction
    _native.call = fake_native.call

    torch.manual_seed(1000 + rank)                      # rank-dependent initialisation
    net = UNet(6, 96, 5, **TINY)
    torch.nn.init.normal_(net.final_conv.weight, std=0.02)
    red = GradAllReducer(net, bucket_bytes=1 << 18)
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas differ after construction"

    def grads_of(seed):
        x, a, c, t, noise, keep = synth_inputs(2, 48, seed)
        net.zero_grad(set_to_none=True)
        y = UNetFunction.apply(net, x, a, t, c, keep, *list(net.parameters()))
        torch.nn.functional.mse_loss(y, noise).backward()
        return {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}

    for _ in range(2):
        grads_of(5 + rank)
    n_buckets_full, key_full = len(red.buckets), red._plan_key
    assert red.ready_at is not None
    # change the trainable set: only adapter tensors keep requires_grad -> the engine rebuilds a (much smaller) arena
    lora.inject_adapters(net, r=4, lora_alpha=4, use_dora=True)
    with torch.no_grad():
        for m in net.modules():
            if hasattr(m, "lora_B"):
                m.lora_B["default"].weight.normal_(std=0.05)
    g1 = grads_of(9 + rank)
    assert red._plan_key != key_full and len(red.buckets) != n_buckets_full
    assert red.arena.numel() * 4 == sum((e - s) * 4 for s, e, _ in red.buckets)
    g2 = grads_of(9 + rank)                               # overlapped step with the re-learned readiness
    for k in g1:
        assert torch.allclose(g1[k], g2[k], rtol=1e-4, atol=1e-7), k
    flat = torch.cat([g2[k].flatten() for k in sorted(g2)])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "averaged adapter gradients differ across ranks"
    out.put((rank, "ok", len(red.buckets)))
    dist.destroy_process_group()


def test_replica_broadcast_and_arena_rebuild_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_replica_sync, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] == "ok" for r in res)
