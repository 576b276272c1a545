"""Host logic of the bucketed, overlapped gradient all-reduce (optim.FlatAdamW.reduce_params / all_reduce_grads) on two gloo
ranks: every parameter is reduced exactly once whatever the hand-over pattern, never-trained parameters sit outside the
updated range, and the optimizer speaks the torch.optim protocol (param_groups, state_dict round trip)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from torch import nn
    from mammoclip_b200.optim import FlatAdamW
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    tower_a = nn.ModuleList([nn.Linear(40, 30) for _ in range(5)])          # "image tower": handed over block by block, last first
    tower_b = nn.ModuleList([nn.Linear(17, 9) for _ in range(3)])           # "text tower": handed over at once
    head = nn.Linear(8, 4)                                                   # never handed over: reduced by all_reduce_grads
    model = nn.ModuleDict({"a": tower_a, "b": tower_b, "h": head})
    frozen = list(tower_b[2].parameters())
    opt = FlatAdamW(model.parameters(), lr=1e-3, frozen=frozen)
    opt.bucket_bytes = 6000                                                  # a few blocks per bucket
    assert opt.n_frozen == 2 and opt.n_active < opt.flat.numel()
    for step in range(2):
        opt.zero_grad()
        g = torch.Generator().manual_seed(10 * step + rank)
        local = {}
        for k, p in model.named_parameters():
            p.grad.copy_(torch.randn(p.shape, generator=g))
            local[k] = p.grad.clone()
        if step == 0:                                                        # pattern 1: towers hand ranges over early
            opt.reduce_params([p for m in list(tower_b)[:2] for p in m.parameters()], flush=True)
            for i in reversed(range(5)):
                opt.reduce_params(list(tower_a[i].parameters()), flush=(i == 0))
        scale = opt.all_reduce_grads(world)                                  # pattern 2 (step 1): nothing handed over
        assert scale == 1.0 / world
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: v for k, v in local.items()})
        for k, p in model.named_parameters():
            want = sum(gathered[r][k] for r in range(world))
            if any(p is q for q in frozen):
                continue                                                     # outside the reduced range: untouched
            assert torch.allclose(p.grad, want, atol=1e-6), (step, k)
        calls, nbytes = opt.comm_stats
        assert calls >= 1 and nbytes >= 4 * sum(p.numel() for p in model.parameters() if not any(p is q for q in frozen))
    # torch.optim protocol
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 0.5)
    assert abs(opt.lr - 5e-4) < 1e-12
    opt.exp_avg.normal_(); opt.exp_avg_sq.uniform_(); opt.steps = 7
    import copy
    sd = copy.deepcopy(opt.state_dict())                                     # what torch.save / torch.load hand back
    want = {k: (opt.state[p]["exp_avg"].clone(), opt.state[p]["exp_avg_sq"].clone()) for k, p in model.named_parameters()}
    opt.exp_avg.zero_(); opt.exp_avg_sq.zero_(); opt.steps = 0
    opt.load_state_dict(sd)
    assert opt.steps == 7
    for k, p in model.named_parameters():
        st = opt.state[p]
        assert torch.equal(st["exp_avg"], want[k][0]) and torch.equal(st["exp_avg_sq"], want[k][1]), k
        assert st["exp_avg"].data_ptr() >= opt.exp_avg.data_ptr() and st["exp_avg"].data_ptr() < opt.exp_avg.data_ptr() + 4 * opt.exp_avg.numel()
    assert all(p.grad.data_ptr() >= opt.grad.data_ptr() for p in model.parameters())
    dist.destroy_process_group()
    print("ok", rank)
''') % ROOT


def test_bucketed_allreduce_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0 and f"ok {r}" in out, out[-3000:]
