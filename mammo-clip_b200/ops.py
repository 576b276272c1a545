"""Thin Python wrappers over the C ABI: argument marshalling only, all arithmetic happens in the CUDA library."""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, LossArgs, WgradArgs, check, lib, ptr, stream_ptr


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.MclipError("mammoclip_b200 kernels need CUDA tensors on a B200; there is no CPU fallback")


# ------------------------------------------------------------------------------------------------ GEMMs

def gemm_tn(a, b, out=None, bias=None, residual=None, act=0, want_stats=False):
    """out[b,m,n] = epi(sum_k a[b,m,k] * w[b|0,n,k]).  a: [M,K] or [Bt,M,K] bf16; b: [N,K] or [Bt,N,K] bf16.
    Returns out (bf16) or (out, stats[slots,2,N] fp32) when want_stats."""
    _require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    a3 = a if a.dim() == 3 else a.unsqueeze(0)
    bt, m, k = a3.shape
    b_batched = b.dim() == 3
    n = b.shape[-2]
    assert b.shape[-1] == k and a3.stride(2) == 1 and b.stride(-1) == 1
    if out is None:
        out = torch.empty((bt, m, n) if a.dim() == 3 else (m, n), dtype=torch.bfloat16, device=a.device)
    o3 = out if out.dim() == 3 else out.unsqueeze(0)
    g = GemmArgs()
    g.a, g.lda, g.a_batch_stride = a3.data_ptr(), a3.stride(1), a3.stride(0) if bt > 1 else 0
    g.b, g.ldb = b.data_ptr(), b.stride(-2)
    g.b_batch_stride = b.stride(0) if b_batched else 0
    g.d, g.ldd, g.d_batch_stride = o3.data_ptr(), o3.stride(1), o3.stride(0) if bt > 1 else 0
    g.m, g.n, g.k, g.batches = m, n, k, bt
    g.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        r3 = residual if residual.dim() == 3 else residual.unsqueeze(0)
        g.residual, g.ldr, g.r_batch_stride = r3.data_ptr(), r3.stride(1), r3.stride(0) if bt > 1 else 0
    g.act = act
    stats = None
    if want_stats:
        slots = lib().mclip_gemm_tn_stat_slots(m, n, bt)
        stats = torch.empty((slots, 2, n), dtype=torch.float32, device=a.device)
        g.stats, g.stat_slots = stats.data_ptr(), slots
    check(lib().mclip_gemm_tn(C.byref(g), stream_ptr()), "mclip_gemm_tn")
    return (out, stats) if want_stats else out


_wgrad_ws = {}


def gemm_wgrad(a, b, out=None, accumulate=False):
    """out[i,j] (+)= sum_r a[r,i] * b[r,j]; a: [R,I] bf16, b: [R,J] bf16, out fp32 [I,J]."""
    _require_cuda(a, b)
    r, i = a.shape
    j = b.shape[1]
    assert b.shape[0] == r and a.stride(1) == 1 and b.stride(1) == 1
    if out is None:
        out = torch.empty((i, j), dtype=torch.float32, device=a.device)
        accumulate = False
    need = lib().mclip_gemm_wgrad_workspace_bytes(r, i, j)
    key = a.device.index
    ws = _wgrad_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=a.device)
        _wgrad_ws[key] = ws
    g = WgradArgs()
    g.a, g.lda, g.b, g.ldb = a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0)
    g.out, g.ldo, g.accumulate = out.data_ptr(), out.stride(0), int(accumulate)
    g.r, g.i, g.j = r, i, j
    g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    check(lib().mclip_gemm_wgrad(C.byref(g), stream_ptr()), "mclip_gemm_wgrad")
    return out


# ------------------------------------------------------------------------------------------------ loss

_loss_ws = {}


def contrastive_loss_raw(local, pairs, scale, world=1, rank=0, symm=None):
    """local: list of [B,D] fp32 CUDA tensors; pairs: list of (a, b, w_row, w_col, eps).
    Returns (out[2+2P] fp32 device tensor, grads list).  `symm` carries the peer-mapped buffers when world > 1."""
    _require_cuda(*local)
    B, D = local[0].shape
    dev = local[0].device
    P, K = len(pairs), len(local)
    args = LossArgs()
    args.world, args.rank, args.batch, args.dim, args.n_tensors, args.n_pairs = world, rank, B, D, K, P
    grads = []
    for k, t in enumerate(local):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.shape == (B, D)
        g = torch.empty_like(t)
        grads.append(g)
        args.local[k], args.grad[k] = t.data_ptr(), g.data_ptr()
    for i, (a, b, wr, wc, eps) in enumerate(pairs):
        args.pair_a[i], args.pair_b[i], args.w_row[i], args.w_col[i], args.label_smoothing[i] = a, b, wr, wc, eps
    args.logit_scale = float(scale)
    need = lib().mclip_loss_workspace_bytes(world, B, D, P)
    key = (dev.index, need)
    ws = _loss_ws.get(key)
    if ws is None:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _loss_ws[key] = ws
    args.workspace, args.workspace_bytes = ws.data_ptr(), need
    out = torch.empty(2 + 2 * P, dtype=torch.float32, device=dev)
    args.out = out.data_ptr()
    if world > 1:
        symm.fill_args(args, K)
    check(lib().mclip_contrastive_loss(C.byref(args), stream_ptr()), "mclip_contrastive_loss")
    return out, grads
