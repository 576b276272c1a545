"""Projection heads with the reference's parameter names (projection.py:4-29).  `LinearProjectionHead` — the one the
shipped configs select (configs/model/clip_b5_det_clinical.yaml:20-23) — runs on the tcgen05 GEMM; `MLPProjectionHead` runs
Linear → GELU → Linear → dropout → residual → LayerNorm (forward and backward) on the same kernels as the text tower."""
import torch
from torch import nn

from ... import ops


class _LinearFn(torch.autograd.Function):
    """y = x @ W^T + b with bf16 operands / fp32 accumulation (what autocast does to nn.Linear, trainer_ddp.py:294)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        xb = ops.cast_bf16(x.detach().float().contiguous())
        wb = ops.cast_bf16(weight.detach().contiguous())
        y = ops.gemm_tn(xb, wb, bias=bias.detach().float() if bias is not None else None)
        ctx.save_for_backward(xb, wb)
        ctx.has_bias = bias is not None
        return y.float()

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        dyb = ops.cast_bf16(dy.contiguous().float())
        dx = ops.gemm_tn(dyb, wb.t().contiguous()).float()              # [M,K] = dy[M,N] @ W[N,K]
        dw = ops.gemm_wgrad(dyb, xb)                                     # [N,K]
        db = ops.colsum(dyb, torch.empty(dyb.shape[1], dtype=torch.float32, device=dy.device)) if ctx.has_bias else None
        return dx, dw, db


class _KernelLinear(nn.Linear):
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("mammoclip_b200 projection heads run on a B200 only (no CPU fallback)")
        shape = x.shape
        y = _LinearFn.apply(x.reshape(-1, shape[-1]), self.weight, self.bias)
        return y.reshape(*shape[:-1], self.out_features)


class LinearProjectionHead(nn.Module):
    def __init__(self, embedding_dim, projection_dim):
        super().__init__()
        self.projection = _KernelLinear(embedding_dim, projection_dim)

    def forward(self, x):
        return self.projection(x)


class _MLPHeadFn(torch.autograd.Function):
    """projection.py:13-20 in one node: p = Linear(x); y = LayerNorm(dropout(fc(gelu(p))) + p).  Forward: two tcgen05 GEMMs
    (erf-GELU epilogue that also keeps the pre-activation; dropout keep-mask + residual epilogue) and the LayerNorm kernel;
    backward: LayerNorm/GELU backward kernels, data gradients on `gemm_tn`, weight/bias gradients on `gemm_wgrad` / `colsum`."""

    @staticmethod
    def forward(ctx, x, wp, bp, wf, bf, gamma, beta, mask, drop_scale, eps):
        xb = ops.cast_bf16(x.detach().float().contiguous())
        wpb, wfb = ops.cast_bf16(wp.detach().contiguous()), ops.cast_bf16(wf.detach().contiguous())
        pre = torch.empty((xb.shape[0], wp.shape[0]), dtype=torch.bfloat16, device=x.device)
        h = ops.gemm_tn(xb, wpb, bias=bp.detach().float(), act=1, aux_pre=pre)                      # gelu(projected); pre = projected
        y = ops.gemm_tn(h, wfb, bias=bf.detach().float(), residual=pre, dropmask=mask, drop_scale=drop_scale)
        out = ops.layernorm(y, gamma.detach(), beta.detach(), eps)
        ctx.save_for_backward(xb, wpb, wfb, pre, h, y, gamma.detach(), mask)
        ctx.drop_scale, ctx.eps = drop_scale, eps
        return out.float()

    @staticmethod
    def backward(ctx, dout):
        xb, wpb, wfb, pre, h, y, gamma, mask = ctx.saved_tensors
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        # dy: gradient of the residual branch (d projected); dyd: through the dropout into fc's output
        dy, dyd = ops.layernorm_backward(y, ops.cast_bf16(dout.contiguous().float()), gamma, ctx.eps, dgamma, dbeta, dropmask=mask,
                                         drop_scale=ctx.drop_scale)
        dwf = ops.gemm_wgrad(dyd, h)
        dbf = ops.colsum(dyd, torch.empty(dyd.shape[1], dtype=torch.float32, device=dout.device))
        g = ops.gelu_backward(ops.gemm_tn(dyd, wfb.t().contiguous()), pre)                          # d projected through fc∘gelu
        # d projected = g + dy: both feed the first Linear, so its three gradients are accumulated over the two terms
        dx = ops.gemm_tn(g, wpb.t().contiguous(), residual=ops.gemm_tn(dy, wpb.t().contiguous()))
        dwp = ops.gemm_wgrad(dy, xb, out=ops.gemm_wgrad(g, xb), accumulate=True)
        dbp = ops.colsum(dy, ops.colsum(g, torch.empty(g.shape[1], dtype=torch.float32, device=dout.device)), accumulate=True)
        return dx.float(), dwp, dbp, dwf, dbf, dgamma, dbeta, None, None, None


class MLPProjectionHead(nn.Module):
    """Same parameters as the reference (`projection`, `fc`, `layer_norm`; projection.py:4-20); runs on the kernels above."""

    def __init__(self, embedding_dim, projection_dim, dropout):
        super().__init__()
        if projection_dim not in (512, 768):
            raise NotImplementedError("the B200 MLP head's LayerNorm kernels are built for projection_dim 512 / 768")
        self.projection = nn.Linear(embedding_dim, projection_dim)
        self.gelu = nn.GELU()
        self.fc = nn.Linear(projection_dim, projection_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(projection_dim)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("mammoclip_b200 projection heads run on a B200 only (no CPU fallback)")
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        p = self.dropout.p
        mask, scale = None, 1.0
        if self.training and p > 0:
            mask = (torch.rand((x2.shape[0], self.fc.out_features), device=x.device) >= p).to(torch.uint8)
            scale = 1.0 / (1.0 - p)
        y = _MLPHeadFn.apply(x2, self.projection.weight, self.projection.bias, self.fc.weight, self.fc.bias, self.layer_norm.weight,
                             self.layer_norm.bias, mask, scale, self.layer_norm.eps)
        return y.reshape(*shape[:-1], self.fc.out_features)
