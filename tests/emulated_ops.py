"""Test infrastructure: PyTorch-on-CPU stand-ins for the C-ABI wrappers that `model/modules/bert_kernels.py` calls, with bf16
rounding at the same storage points as the CUDA kernels.  They exist so that the HOST orchestration of the text tower's
hand-written forward/backward (which tensor feeds which GEMM, transposed weights, dropout masks, gradient routing) can be
checked against the oracle without a GPU.  Never imported by the product; the GPU tests run the real kernels."""
import math

import torch

BF = torch.bfloat16


def _r(x):
    return x.to(BF)


def gemm_tn(a, b, out=None, bias=None, residual=None, act=0, want_stats=None, dropmask=None, drop_scale=1.0, aux_pre=None):
    v = a.float() @ b.float().t()
    if bias is not None:
        v = v + bias.float()
    if aux_pre is not None:
        aux_pre.copy_(_r(v))
    if act == 1:
        v = torch.nn.functional.gelu(v)
    if dropmask is not None:
        v = v * (dropmask.view_as(v).float() * drop_scale)
    if residual is not None:
        v = v + residual.float()
    return _r(v)


def gemm_wgrad(a, b, out=None, accumulate=False):
    v = a.float().t() @ b.float()
    if out is None:
        return v
    out.copy_(out + v if accumulate else v)
    return out


def colsum(x, out, accumulate=False):
    v = x.float().sum(0)
    out.copy_(out + v if accumulate else v)
    return out


def cast_bf16(x):
    return _r(x)


def weight_prep(entries, device, dst_ld=0):
    return list(entries)


def weight_prep_run(table, n):
    for e in table:
        src, dst, dst_t = e[0], e[1], e[2]
        if dst is not None:
            dst.copy_(_r(src))
        if dst_t is not None:
            dst_t.copy_(_r(src).t())


def bert_embed_ln(ids, tts, word, pos, typ, gamma, beta, eps, dropmask=None, drop_scale=1.0):
    b, l = ids.shape
    v = word[ids] + pos[torch.arange(l)][None] + typ[tts if tts is not None else torch.zeros_like(ids)]
    y = torch.nn.functional.layer_norm(v, (v.shape[-1],), gamma, beta, eps).view(b * l, -1)
    if dropmask is not None:
        y = y * (dropmask.float() * drop_scale)
    return _r(y)


def layernorm(x, gamma, beta, eps):
    return _r(torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps))


def gelu_forward(x):
    return _r(torch.nn.functional.gelu(x.float()))


def gelu_backward(dy, x):
    f = x.float()
    cdf = 0.5 * (1 + torch.erf(f * 0.7071067811865476))
    pdf = 0.3989422804014327 * torch.exp(-0.5 * f * f)
    return _r(dy.float() * (cdf + f * pdf))


def _split(t, b, l, heads, d):
    return t.view(b, l, heads, d).transpose(1, 2)


def bert_attention(qkv, amask, b, l, heads, d, dropmask=None, drop_scale=1.0, want_lse=False):
    h = heads * d
    q, k, v = (_split(t, b, l, heads, d) for t in qkv.float().view(b, l, 3 * h).split(h, dim=-1))
    s = q @ k.transpose(-1, -2) / math.sqrt(d)
    s = s.masked_fill(amask[:, None, None, :] == 0, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - lse[..., None])
    if dropmask is not None:
        p = p * (dropmask.float() * drop_scale)
    out = _r((p @ v).transpose(1, 2).reshape(b * l, h))
    return (out, lse) if want_lse else out


def bert_attention_backward(qkv, d_out, lse, amask, b, l, heads, d, dropmask=None, drop_scale=1.0):
    h = heads * d
    q, k, v = (_split(t, b, l, heads, d) for t in qkv.float().view(b, l, 3 * h).split(h, dim=-1))
    do = _split(d_out.float().view(b, l, h), b, l, heads, d)
    s = (q @ k.transpose(-1, -2) / math.sqrt(d)).masked_fill(amask[:, None, None, :] == 0, float("-inf"))
    p = torch.exp(s - lse[..., None])
    m = dropmask.float() * drop_scale if dropmask is not None else torch.ones_like(p)
    dpd = do @ v.transpose(-1, -2)
    delta = (p * m * dpd).sum(-1, keepdim=True)
    ds = p * (dpd * m - delta)
    dq = ds @ k / math.sqrt(d)
    dk = ds.transpose(-1, -2) @ q / math.sqrt(d)
    dv = (p * m).transpose(-1, -2) @ do
    return _r(torch.cat([t.transpose(1, 2).reshape(b * l, h) for t in (dq, dk, dv)], dim=-1))


def _ln_bwd(x, dy, gamma, eps):
    xr = x.detach().float().clone().requires_grad_(True)
    g = gamma.detach().clone().requires_grad_(True)
    bta = torch.zeros_like(gamma).requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (x.shape[-1],), g, bta, eps).backward(dy.float())
    return xr.grad, g.grad, bta.grad


def layernorm_backward(x, dy, gamma, eps, dgamma, dbeta, dropmask=None, drop_scale=1.0, accumulate=False):
    with torch.enable_grad():
        dx, dg, db = _ln_bwd(x, dy, gamma, eps)
    dgamma.copy_(dgamma + dg if accumulate else dg)
    dbeta.copy_(dbeta + db if accumulate else db)
    dxb = _r(dx)
    if dropmask is None:
        return dxb, dxb
    return dxb, _r(dx * (dropmask.float() * drop_scale))


def bert_embed_backward(ids, tts, word, pos, typ, gamma, eps, dout, dword, dpos, dtype, dgamma, dbeta, dropmask=None, drop_scale=1.0,
                        accumulate=False):
    b, l = ids.shape
    d = dout.float()
    if dropmask is not None:
        d = d * (dropmask.float() * drop_scale)
    tt = tts if tts is not None else torch.zeros_like(ids)
    with torch.enable_grad():
        v = (word[ids] + pos[torch.arange(l)][None] + typ[tt]).view(b * l, -1)
        dv, dg, db = _ln_bwd(v, d, gamma, eps)
    gw, gp, gt = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros_like(typ)
    gw.index_add_(0, ids.reshape(-1), dv)
    gp[:l] = dv.view(b, l, -1).sum(0)
    gt.index_add_(0, tt.reshape(-1), dv)
    for dst, g in ((dword, gw), (dpos, gp), (dtype, gt), (dgamma, dg), (dbeta, db)):
        dst.copy_(dst + g if accumulate else g)


def install(monkeypatch, ops):
    for name in ("gemm_tn", "gemm_wgrad", "colsum", "cast_bf16", "weight_prep", "weight_prep_run", "bert_embed_ln", "layernorm", "gelu_forward",
                 "gelu_backward", "bert_attention", "bert_attention_backward", "layernorm_backward", "bert_embed_backward"):
        monkeypatch.setattr(ops, name, globals()[name])
