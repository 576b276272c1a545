"""AdamW on ONE flat fp32 buffer (one kernel launch per step) + flat gradient all-reduce for data parallelism.

Same update rule as the reference's optimizer (breastclip/optimizer/__init__.py:23-31 builds torch.optim.AdamW(lr,
weight_decay) over all parameters; its `no_decay` branch is dead code).  Parameters keep their identity
(`nn.Parameter` objects and state-dict names are untouched): only `.data` / `.grad` are re-pointed into the flat buffers."""
import torch
import torch.distributed as dist

from . import ops


class FlatAdamW:
    def __init__(self, params, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        align = 32                                   # floats: every parameter starts on a 128-byte boundary (vector loads, TMA)
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + align - 1) // align * align
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, offs):
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.steps = 0
        self.zero_count = 0

    def attach(self, *modules):
        """Modules whose backward may write gradients straight into the flat buffer (first backward after each zero_grad)."""
        for m in modules:
            for sub in m.modules():
                if (hasattr(sub, "_wcache") and hasattr(sub, "geom")) or getattr(sub, "_mclip_direct_grads", False):
                    object.__setattr__(sub, "_flat_optimizer", self)
        return self

    def zero_grad(self):
        """Gradients live in the flat buffer; autograd accumulates into the views, so clear instead of set_to_none."""
        self.grad.zero_()
        self.zero_count += 1

    def all_reduce_grads(self, world):
        """DDP's gradient averaging (trainer_ddp.py:134) as one NCCL all-reduce over the flat buffer."""
        if world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            return 1.0 / world
        return 1.0

    def step(self, grad_scale=1.0):
        self.steps += 1
        ops.adamw_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                       self.steps, grad_scale)
