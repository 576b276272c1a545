"""Projection heads with the reference's parameter names (projection.py:4-29).  `LinearProjectionHead` — the one the
shipped configs select (configs/model/clip_b5_det_clinical.yaml:20-23) — runs on the tcgen05 GEMM; `MLPProjectionHead`
keeps the reference structure on top of the same Linear kernel."""
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


class MLPProjectionHead(nn.Module):
    def __init__(self, embedding_dim, projection_dim, dropout):
        super().__init__()
        self.projection = _KernelLinear(embedding_dim, projection_dim)
        self.gelu = nn.GELU()
        self.fc = _KernelLinear(projection_dim, projection_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(projection_dim)

    def forward(self, x):
        projected = self.projection(x)
        x = self.fc(self.gelu(projected))
        x = self.dropout(x) + projected
        return self.layer_norm(x)
