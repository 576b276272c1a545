"""AdamW on ONE flat fp32 buffer (one kernel launch per step) + bucketed, overlapped gradient all-reduce.

Same update rule as the reference's optimizer (breastclip/optimizer/__init__.py:23-31 builds torch.optim.AdamW(lr,
weight_decay) over all parameters; its `no_decay` branch is dead code).  Parameters keep their identity
(`nn.Parameter` objects and state-dict names are untouched): only `.data` / `.grad` are re-pointed into the flat buffers.

* `FlatAdamW` IS a `torch.optim.Optimizer`: `param_groups` (LR schedulers, scheduler/__init__.py), `state_dict()` /
  `load_state_dict()` (the trainer checkpoints `optimizer.state_dict()`, trainer_ddp.py:237-248) work as usual.
* Parameters that can never receive a gradient on this path (the BERT pooler: computed by the reference, result discarded,
  SURVEY A9) are placed at the END of the flat buffers and left out of the update, like torch.optim.AdamW skips `grad is None`.
* Data parallel (DDP's gradient averaging, trainer_ddp.py:134): the towers hand finished, contiguous gradient ranges to
  `reduce_params()` while the rest of the backward still runs; each range is all-reduced on a side stream (NCCL), so only the
  last bucket (stem + first blocks, a few hundred KB) is exposed after the backward.
"""
import torch
import torch.distributed as dist

from . import ops


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, frozen=()):
        plist = [p for p in params if p.requires_grad]
        super().__init__(plist, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        frozen_ids = {id(p) for p in frozen}
        self._build(plist, frozen_ids)
        self.steps = 0
        self.zero_count = 0
        self._works, self._reduced, self._comm, self._bucket, self._bucket_bytes = [], set(), None, [], 0
        self.bucket_bytes = 32 << 20
        # MCLIP_AR_OVERLAP=0: reduce everything after the backward (one ncclAllReduce); 1: overlapped buckets (default)
        import os
        self.overlap = os.environ.get("MCLIP_AR_OVERLAP", "1") != "0"
        self._nbytes, self.comm_stats = 0, (0, 0)   # (all-reduce calls, bytes) of the last step, for bench.py

    # ------------------------------------------------------------------------------------------------ layout
    def _build(self, plist, frozen_ids):
        """Flat layout: updated parameters first (model order), never-updated ones last; 128-byte aligned slices."""
        self.params = [p for p in plist if id(p) not in frozen_ids] + [p for p in plist if id(p) in frozen_ids]
        self.n_frozen = sum(1 for p in plist if id(p) in frozen_ids)
        dev = self.params[0].device
        align = 32                                   # floats: every parameter starts on a 128-byte boundary (vector loads, TMA)
        self.offsets, n, self.n_active = {}, 0, 0
        for i, p in enumerate(self.params):
            self.offsets[id(p)] = (n, p.numel())
            n += (p.numel() + align - 1) // align * align
            if i + 1 == len(self.params) - self.n_frozen:
                self.n_active = n
        if self.n_frozen == 0:
            self.n_active = n
        old = getattr(self, "flat", None)
        flat = torch.zeros(n, dtype=torch.float32, device=dev)
        grad = torch.zeros(n, dtype=torch.float32, device=dev)
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        for p in self.params:
            off, k = self.offsets[id(p)]
            flat[off:off + k].copy_(p.data.reshape(-1))
            st = self.state.get(p)
            if old is not None and st:
                m[off:off + k].copy_(st["exp_avg"].reshape(-1)); v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            p.grad = grad[off:off + k].view(p.shape)
            # torch.optim state layout (so that state_dict() interoperates with torch.optim.AdamW checkpoints)
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": m[off:off + k].view(p.shape), "exp_avg_sq": v[off:off + k].view(p.shape)}
        self.flat, self.grad, self.exp_avg, self.exp_avg_sq = flat, grad, m, v

    def note_forward(self, module):
        """Towers call this in forward: gradients may be handed to reduce_params() early only if the tower ran ONCE since
        zero_grad() (the multi-view loss runs each tower twice; the second backward accumulates into the same views)."""
        if getattr(module, "_fwd_epoch", None) != self.zero_count:
            object.__setattr__(module, "_fwd_epoch", self.zero_count)
            object.__setattr__(module, "_fwd_calls", 0)
        object.__setattr__(module, "_fwd_calls", module._fwd_calls + 1)

    def single_use(self, module):
        return getattr(module, "_fwd_epoch", None) == self.zero_count and getattr(module, "_fwd_calls", 0) == 1

    def attach(self, *modules):
        """Modules whose backward may write gradients straight into the flat buffer (first backward after each zero_grad).
        Also finds the parameters this path never trains (BERT pooler) and moves them out of the updated range."""
        frozen = []
        for m in modules:
            for sub in m.modules():
                if (hasattr(sub, "_wcache") and hasattr(sub, "geom")) or getattr(sub, "_mclip_direct_grads", False):
                    object.__setattr__(sub, "_flat_optimizer", self)
                pool = getattr(sub, "pooler", None)
                if pool is not None and sub.__class__.__name__ == "BertModel":
                    frozen += [p for p in pool.parameters() if id(p) in self.offsets]
        if frozen and self.n_frozen == 0 and self.steps == 0:
            self._build(list(self.params), {id(p) for p in frozen})
        return self

    # ------------------------------------------------------------------------------------------------ torch.optim API
    def zero_grad(self, set_to_none=False):
        """Gradients live in the flat buffer; autograd accumulates into the views, so clear instead of set_to_none."""
        self.grad.zero_()
        self.zero_count += 1
        self._reduced.clear()

    def state_dict(self):
        for p in self.params:
            self.state[p]["step"] = torch.tensor(float(self.steps))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)          # replaces the state tensors by copies: move them back into the flat buffers
        steps = 0
        for p in self.params:
            st = self.state.get(p)
            if not st:
                continue
            off, k = self.offsets[id(p)]
            self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1)); self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            st["exp_avg"], st["exp_avg_sq"] = self.exp_avg[off:off + k].view(p.shape), self.exp_avg_sq[off:off + k].view(p.shape)
            steps = max(steps, int(float(st.get("step", 0))))
        self.steps = steps

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        for g in self.param_groups:
            g["lr"] = v

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        g = self.param_groups[0]
        self.steps += 1
        n = self.n_active
        ops.adamw_step(self.flat[:n], self.grad[:n], self.exp_avg[:n], self.exp_avg_sq[:n], g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                       g["weight_decay"], self.steps, grad_scale)

    # ------------------------------------------------------------------------------------------------ data parallel
    def _range(self, params):
        lo = min(self.offsets[id(p)][0] for p in params)
        hi = max(self.offsets[id(p)][0] + self.offsets[id(p)][1] for p in params)
        return lo, hi

    def _launch(self, lo, hi):
        if not self.flat.is_cuda:                    # gloo (CPU tests of the bucketing logic)
            self._works.append(dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
            self._nbytes += (hi - lo) * 4
            return
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.flat.device)
        ev = torch.cuda.current_stream().record_event()
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            self._works.append(dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        self._nbytes += (hi - lo) * 4

    def reduce_params(self, params, flush=False):
        """Called by a tower's backward (direct-gradient mode) with parameters whose gradients are FINAL for this step.  Ranges are
        coalesced into buckets of >= bucket_bytes and all-reduced asynchronously; `params` of successive calls must be
        adjacent in the flat buffer (a tower hands over its blocks in backward order)."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        if not self.overlap and not flush:
            return
        if not self.overlap:
            params = []
        if not self._works and not self._bucket:
            self._nbytes = 0
        params = [p for p in params if id(p) in self.offsets and id(p) not in self._reduced]
        self._bucket += params
        self._bucket_bytes += sum(p.numel() for p in params) * 4
        if self._bucket and (flush or self._bucket_bytes >= self.bucket_bytes):
            lo, hi = self._range(self._bucket)
            # only contiguous runs may be merged into one call (everything between lo and hi must belong to the bucket)
            covered = sum((self.offsets[id(p)][1] + 31) // 32 * 32 for p in self._bucket)
            if covered >= hi - lo:
                self._launch(lo, hi)
            else:
                for p in self._bucket:
                    o, k = self.offsets[id(p)]
                    self._launch(o, o + k)
            self._reduced.update(id(p) for p in self._bucket)
            self._bucket, self._bucket_bytes = [], 0

    def all_reduce_grads(self, world):
        """DDP's gradient averaging (trainer_ddp.py:134): reduces whatever the towers have not handed over yet, then makes the
        current stream wait for every outstanding bucket.  Returns the factor 1/world that `step(grad_scale=...)` applies."""
        if world <= 1:
            return 1.0
        if not self._works and not self._bucket:
            self._nbytes = 0
        self.reduce_params([], flush=True)
        run = []
        for p in self.params[:len(self.params) - self.n_frozen] + [None]:
            if p is not None and id(p) not in self._reduced:
                run.append(p)
                continue
            if run:
                self._launch(*self._range(run))
                run = []
        for w in self._works:
            w.wait()
        self.comm_stats = (len(self._works), self._nbytes)
        self._works = []
        return 1.0 / world
