"""EfficientNet-B0..B7 image encoder on hand-written sm_100a kernels, drop-in for the reference class.

Mirrors `breastclip/model/modules/efficientnet_custom.py` (EfficientNet :143-411, MBConvBlock :36-140): same constructor
entry points (`from_name`, `from_pretrained`), same `forward` contract (tensor -> pooled [B,out_dim]; dict with "image"
-> (pooled, raw_features), :287-313), same state-dict keys and tensor shapes (checkpoints interchange).
All arithmetic runs in libmclip_b200.so through one torch.autograd.Function spanning the whole tower; activations are
NHWC bf16, only PRE-BatchNorm conv outputs are kept for backward (BN+swish are re-applied on load).
"""
import math
from dataclasses import dataclass
from typing import List, Tuple

import torch
from torch import nn

from ... import ops

# ----------------------------------------------------------------------------------------------- geometry
# (repeats, kernel, stride, expand, in, out): efficient_net_custom_utils.py:502-510
_STAGES = ((1, 3, 1, 1, 32, 16), (2, 3, 2, 6, 16, 24), (2, 5, 2, 6, 24, 40), (3, 3, 2, 6, 40, 80),
           (3, 5, 1, 6, 80, 112), (4, 5, 2, 6, 112, 192), (1, 3, 1, 6, 192, 320))
# name -> (width, depth, nominal res, dropout): efficient_net_custom_utils.py:466-478
_PARAMS = {"efficientnet-b0": (1.0, 1.0, 224, 0.2), "efficientnet-b1": (1.0, 1.1, 240, 0.2), "efficientnet-b2": (1.1, 1.2, 260, 0.3),
           "efficientnet-b3": (1.2, 1.4, 300, 0.3), "efficientnet-b4": (1.4, 1.8, 380, 0.4), "efficientnet-b5": (1.6, 2.2, 456, 0.4),
           "efficientnet-b6": (1.8, 2.6, 528, 0.5), "efficientnet-b7": (2.0, 3.1, 600, 0.5)}
VALID_MODELS = tuple(_PARAMS)
BN_MOMENTUM, BN_EPS, DROP_CONNECT_RATE, SE_RATIO = 0.01, 1e-3, 0.2, 0.25


def _round_filters(ch, width, divisor=8):
    ch = ch * width
    new = max(divisor, int(ch + divisor / 2) // divisor * divisor)
    if new < 0.9 * ch:
        new += divisor
    return int(new)


def _static_pad(size, k, s):
    """(before, after, out) of Conv2dStaticSamePadding along one axis (efficient_net_custom_utils.py:255-271)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2, out


@dataclass
class BlockGeom:
    cin: int
    cexp: int
    cout: int
    k: int
    s: int
    expand: bool
    cse: int
    pads: Tuple[int, int, int, int]     # (left, right, top, bottom), frozen from the NOMINAL resolution
    skip: bool


@dataclass
class NetGeom:
    stem_out: int
    stem_pads: Tuple[int, int, int, int]
    blocks: List[BlockGeom]
    head_out: int
    dropout: float


def net_geometry(name: str) -> NetGeom:
    width, depth, res, dropout = _PARAMS[name]
    h = w = res
    l, r, w2 = _static_pad(w, 3, 2)
    t, b, h2 = _static_pad(h, 3, 2)
    stem_pads = (l, r, t, b)
    h, w = h2, w2
    blocks = []
    for rep, k, s, e, cin, cout in _STAGES:
        cin, cout = _round_filters(cin, width), _round_filters(cout, width)
        for i in range(int(math.ceil(depth * rep))):
            bi, bs = (cin, s) if i == 0 else (cout, 1)
            l, r, wo = _static_pad(w, k, bs)
            t, b, ho = _static_pad(h, k, bs)
            blocks.append(BlockGeom(bi, bi * e, cout, k, bs, e != 1, max(1, int(bi * SE_RATIO)), (l, r, t, b), bs == 1 and bi == cout))
            h, w = ho, wo
    return NetGeom(_round_filters(32, width), stem_pads, blocks, _round_filters(1280, width), dropout)


# ----------------------------------------------------------------------------------------------- parameter holders
class _ConvParams(nn.Module):
    """Holds `weight` (OIHW, as nn.Conv2d) and optional `bias`; default init identical to nn.Conv2d."""

    def __init__(self, cout, cin_per_group, k, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin_per_group, k, k))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1 / math.sqrt(cin_per_group * k * k)
            nn.init.uniform_(self.bias, -bound, bound)


class _BNParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.momentum, self.eps = BN_MOMENTUM, BN_EPS


class MBConvBlock(nn.Module):
    """Parameter container with the reference's attribute names (efficientnet_custom.py:50-89)."""

    def __init__(self, g: BlockGeom):
        super().__init__()
        self.geom = g
        if g.expand:
            self._expand_conv = _ConvParams(g.cexp, g.cin, 1)
            self._bn0 = _BNParams(g.cexp)
        self._depthwise_conv = _ConvParams(g.cexp, 1, g.k)
        self._bn1 = _BNParams(g.cexp)
        self._se_reduce = _ConvParams(g.cse, g.cexp, 1, bias=True)
        self._se_expand = _ConvParams(g.cexp, g.cse, 1, bias=True)
        self._project_conv = _ConvParams(g.cout, g.cexp, 1)
        self._bn2 = _BNParams(g.cout)


# ----------------------------------------------------------------------------------------------- the engine
import os as _os
_FOLD_BN0 = [_os.environ.get("MCLIP_FOLD_BN0", "1") != "0"]      # fold the expand conv's BN backward into its dgrad / wgrad GEMM operands
_UPDATE_RUNNING = [True]      # False during the recompute pass of the multi-view memory plan (the statistics were already folded in)
_DROP_Y0 = [False]            # "lean" multi-view plan: the expand conv's output is not kept; the backward re-runs that one GEMM


def _bn_fin(part, count, bn: _BNParams, training):
    if training and not _UPDATE_RUNNING[0]:
        return ops.bn_finalize(part, count, bn.weight, bn.bias, bn.running_mean, bn.running_var, None, training, 0.0, bn.eps)
    return ops.bn_finalize(part, count, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, training, bn.momentum, bn.eps)


class _WeightCache:
    """bf16 copies (and transposes) of the 1x1-conv weights, refreshed by ONE table-driven kernel per step."""

    def __init__(self, net):
        dev = net._conv_stem.weight.device
        self.entries, self.bf16, self.bf16_t = [], {}, {}

        def add(key, w, straight, transposed):
            w2 = w.detach().view(w.shape[0], -1)
            d = torch.empty_like(w2, dtype=torch.bfloat16) if straight else None
            dt = torch.empty((w2.shape[1], w2.shape[0]), dtype=torch.bfloat16, device=dev) if transposed else None
            self.entries.append((w2, d, dt))
            self.bf16[key], self.bf16_t[key] = d, dt

        for i, blk in enumerate(net._blocks):
            if blk.geom.expand:
                add(("e", i), blk._expand_conv.weight, True, True)
            add(("p", i), blk._project_conv.weight, False, True)
        add(("h",), net._conv_head.weight, True, True)
        ws = net._conv_stem.weight
        self.bf16[("s",)] = torch.zeros((ws.shape[0], 32), dtype=torch.bfloat16, device=dev)      # K = 27 taps padded to 32
        self.entries.append((ws.detach().view(ws.shape[0], 27), self.bf16[("s",)], None, 32))
        self.table = ops.weight_prep(self.entries, dev)
        self.key = self._key(net)

    @staticmethod
    def _key(net):
        return tuple(p.data_ptr() for p in net.parameters())

    def refresh(self):
        ops.weight_prep_run(self.table, len(self.entries))


def _block_forward(net, i, x, pending, n, h, w, training, rowscale=None):
    """One MBConvBlock (efficientnet_custom.py:91-132).  `x`: materialised block input [N,H,W,Cin] bf16, or None when
    `pending` = (pre-BN tensor, BNState) of the stem whose BN+swish the depthwise loader applies.  Returns
    (block output [N,Ho,Wo,Cout] bf16, saved state for _block_backward)."""
    blk, wc = net._blocks[i], net._weights()
    g = blk.geom
    B = {"x_in": x, "h": h, "w": w}
    if g.expand:
        y0, st = ops.gemm_tn(x.view(n * h * w, g.cin), wc.bf16[("e", i)], want_stats=training)
        y0 = y0.view(n, h, w, g.cexp)
        bn0 = _bn_fin(st, n * h * w, blk._bn0, training)
        B["y0"], B["bn0"] = y0, bn0
        dw_in, dw_bn = y0, bn0
    elif pending is not None:
        dw_in, dw_bn = pending
        B["from_stem"] = True
    else:
        dw_in, dw_bn = x, None
    y1, st = ops.dwconv_forward(dw_in, blk._depthwise_conv.weight, g.k, g.s, g.pads, bn=dw_bn, want_stats=training)
    if g.expand and _DROP_Y0[0]:
        B["y0"] = y0 = dw_in = None       # 41 % of the saved state (37.9 of 92.9 GB per EN-B5 view at B = 64, 1520x912)
    ho, wo = y1.shape[1], y1.shape[2]
    bn1 = _bn_fin(st, n * ho * wo, blk._bn1, training)
    u, pool = ops.ew_forward(y1.view(n, ho * wo, g.cexp), bn=bn1, act=1, write=True, pool=True)
    w1, w2 = blk._se_reduce.weight.view(g.cse, g.cexp), blk._se_expand.weight.view(g.cexp, g.cse)
    pooled, z1, gate = ops.se_fc(pool, ho * wo, w1, blk._se_reduce.bias, w2, blk._se_expand.bias)
    wg = ops.se_scale_weights(blk._project_conv.weight.view(g.cout, g.cexp), gate)
    y2, st = ops.gemm_tn(u, wg, want_stats=training)                     # [N, ho*wo, cout], per-sample weights
    del u, wg
    bn2 = _bn_fin(st, n * ho * wo, blk._bn2, training)
    rs = rowscale if g.skip else None
    # per-channel sums of the block output ride along (pool partials): the next block's folded BN0 backward needs sum(X)
    x_out, xsum = ops.ew_forward(y2, bn=bn2, act=0, rowscale=rs, residual=x.view(n, h * w, g.cin) if g.skip else None, pool=training and _FOLD_BN0[0])
    B.update(y1=y1, bn1=bn1, pooled=pooled, z1=z1, gate=gate, y2=y2, bn2=bn2, rowscale=rs, ho=ho, wo=wo, out_sum_part=xsum)
    return x_out.view(n, ho, wo, g.cout), B


def _forward(net, images, training, drop_rowscales, dropout_mult, want_raw, save=True):
    """Returns (features [N,Chead] fp32, saved state for backward, raw head activation or None).  save=False keeps nothing
    for a backward pass (every intermediate is freed as soon as its consumer has run)."""
    geom = net.geom
    n, _, h_in, w_in = images.shape
    wc = net._weights()
    wc.refresh()
    S = {"images": images, "training": training, "blocks": []}
    lut = None
    if images.dtype == torch.uint8:           # input edge: raw grey levels, per-image min-max + mean/std applied on load
        _, lut = ops.image_norm_lut_u8(images, net.input_mean, net.input_std)
    y, st, patches = ops.stem_forward(images, net._conv_stem.weight, geom.stem_pads, want_stats=training, w_bf16=wc.bf16[("s",)], return_patches=True,
                                      norm_lut=lut)
    S["patches"] = patches
    h, w = y.shape[1], y.shape[2]
    bn = _bn_fin(st, n * h * w, net._bn0, training)
    S["stem"] = (y, bn)
    pending = (y, bn)      # a pre-BN tensor whose BN+swish is applied by the consumer
    x = None               # materialised block input [N,H,W,C] bf16
    prev_sum = None        # pool partials [N,chunks,C] of the pass that wrote x
    for i, blk in enumerate(net._blocks):
        if blk.geom.expand and pending is not None:
            # (never the case for B0-B7, whose first stage has expand_ratio 1) the expand GEMM needs a materialised input
            x, _ = ops.ew_forward(pending[0].view(n, h * w, -1), bn=pending[1], act=1)
            x, pending = x.view(n, h, w, -1), None
            S["stem_materialised"] = True
        x, B = _block_forward(net, i, x, pending, n, h, w, training, drop_rowscales.get(i) if drop_rowscales else None)
        B["x_sum_part"], prev_sum = prev_sum, B.pop("out_sum_part", None)
        h, w, pending = B["ho"], B["wo"], None
        if save:
            S["blocks"].append(B)
        else:
            if i == 0:
                S.pop("stem", None); S.pop("patches", None)
            del B
    yh, st = ops.gemm_tn(x.view(n * h * w, x.shape[-1]), wc.bf16[("h",)], want_stats=training)
    yh = yh.view(n, h * w, geom.head_out)
    bnh = _bn_fin(st, n * h * w, net._bn1, training)
    raw, pool = ops.ew_forward(yh, bn=bnh, act=1, write=want_raw, pool=True)
    feat = ops.pool_finalize(pool, h * w, dropout_mult)
    if save:
        S.update(x_last=x, yh=yh, bnh=bnh, h=h, w=w, dropout_mult=dropout_mult)
    else:
        S = None
    if want_raw:
        raw = raw.view(n, h, w, geom.head_out).permute(0, 3, 1, 2).float()
    return feat, S, raw


def _bn_backward(y3, bn, training, bnp: _BNParams, grads, prefix, act, du=None, dvec=None, gate=None, dpool=None, rowscale=None, G=torch.empty_like):
    """Two-pass BatchNorm(+swish) backward over y3 [N,HW,C]; returns dY (bf16) and fills grads of gamma/beta."""
    part = ops.ew_backward(0, y3, bn, act, du=du, dvec=dvec, gate=gate, dpool=dpool, rowscale=rowscale)
    dg, db = G(bnp.weight), G(bnp.bias)
    c1, c2 = ops.bn_bwd_finalize(part, bn.count, training, dg, db)
    grads[prefix + ".weight"], grads[prefix + ".bias"] = dg, db
    return ops.ew_backward(1, y3, bn, act, du=du, dvec=dvec, gate=gate, dpool=dpool, rowscale=rowscale, c1=c1, c2=c2)


def _block_backward(net, i, B, dx, n, training, grads, G, S=None):
    """Backward of one MBConvBlock.  dx: gradient w.r.t. the block output [N,Ho,Wo,Cout] bf16; returns the gradient w.r.t.
    the block input (None for the block fed by the stem, whose BN/conv gradients are written here instead; S carries the stem state)."""
    blk, wc, geom = net._blocks[i], net._weights(), net.geom
    g, pre = blk.geom, f"_blocks.{i}."
    h, w, ho, wo = B["h"], B["w"], B["ho"], B["wo"]
    dxo = dx.view(n, ho * wo, g.cout)
    dy2 = _bn_backward(B["y2"], B["bn2"], training, blk._bn2, grads, pre + "_bn2", 0, du=dxo, rowscale=B["rowscale"], G=G)
    da2 = ops.gemm_tn(dy2.view(n * ho * wo, g.cout), wc.bf16_t[("p", i)]).view(n, ho * wo, g.cexp)
    y1 = B["y1"].view(n, ho * wo, g.cexp)
    a2, dgp = ops.ew_backward(2, y1, B["bn1"], 1, du=da2, gate=B["gate"])
    gp = G(blk._project_conv.weight)
    ops.gemm_wgrad(dy2.view(n * ho * wo, g.cout), a2.view(n * ho * wo, g.cexp), out=gp.view(g.cout, g.cexp))
    grads[pre + "_project_conv.weight"] = gp
    del a2, dy2
    w1, w2 = blk._se_reduce.weight.view(g.cse, g.cexp), blk._se_expand.weight.view(g.cexp, g.cse)
    dw1, db1, dw2, db2 = (G(t) for t in (blk._se_reduce.weight, blk._se_reduce.bias, blk._se_expand.weight, blk._se_expand.bias))
    dpool = ops.se_fc_backward(dgp, ho * wo, w1, w2, B["pooled"], B["z1"], B["gate"], dw1, db1, dw2, db2)
    grads[pre + "_se_reduce.weight"], grads[pre + "_se_reduce.bias"] = dw1, db1
    grads[pre + "_se_expand.weight"], grads[pre + "_se_expand.bias"] = dw2, db2
    # BN1 backward sums come out of the SE pass-1 partials (no separate reduction pass over dA2 / Y1)
    bnp1 = ops.se_bn_combine(dgp, B["gate"], dpool)
    dg1, db1_ = G(blk._bn1.weight), G(blk._bn1.bias)
    c1, c2 = ops.bn_bwd_finalize(bnp1, B["bn1"].count, training, dg1, db1_)
    grads[pre + "_bn1.weight"], grads[pre + "_bn1.bias"] = dg1, db1_
    dy1 = ops.ew_backward(1, y1, B["bn1"], 1, du=da2, gate=B["gate"], dpool=dpool, c1=c1, c2=c2)
    del da2
    dy1 = dy1.view(n, ho, wo, g.cexp)
    ddw = G(blk._depthwise_conv.weight)
    grads[pre + "_depthwise_conv.weight"] = ddw
    if g.expand:
        y0, bn0 = B["y0"], B["bn0"]
        if y0 is None:                    # lean plan: the same GEMM on the same bf16 operands reproduces the forward's tensor bit for bit
            y0 = ops.gemm_tn(B["x_in"].view(n * h * w, g.cin), wc.bf16[("e", i)]).view(n, h, w, g.cexp)
        dv0, bnp = ops.dwconv_backward(y0, blk._depthwise_conv.weight, g.k, g.s, g.pads, dy1, ddw, bn=bn0)
        del y0
        dg0, db0 = G(blk._bn0.weight), G(blk._bn0.bias)
        c1, c2 = ops.bn_bwd_finalize(bnp, bn0.count, training, dg0, db0)
        grads[pre + "_bn0.weight"], grads[pre + "_bn0.bias"] = dg0, db0
        x_in = B["x_in"].view(n * h * w, g.cin)
        ge = G(blk._expand_conv.weight)
        grads[pre + "_expand_conv.weight"] = ge
        skip_grad = dx.view(n * h * w, g.cin) if g.skip else None
        if _FOLD_BN0[0] and g.cexp % 8 == 0 and g.cin % 8 == 0:
            # both consumers of dY0 are linear: the BN0-backward apply pass over the 6x-wide tensor folds into the GEMM operands
            part = B.get("x_sum_part")
            dx = ops.bn0_fold_backward(dv0.view(n * h * w, g.cexp), x_in, blk._expand_conv.weight.view(g.cexp, g.cin), wc.bf16[("e", i)], bn0, c1, c2,
                                       ge.view(g.cexp, g.cin), residual=skip_grad, sumx=None if part is None else part.sum((0, 1))).view(n, h, w, g.cin)
            del dv0
        else:
            dy0 = ops.ew_backward(1, y0.view(n, h * w, g.cexp), bn0, 0, du=dv0.view(n, h * w, g.cexp), dv_given=True, c1=c1, c2=c2)
            del dv0
            dy0 = dy0.view(n * h * w, g.cexp)
            ops.gemm_wgrad(dy0, x_in, out=ge.view(g.cexp, g.cin))
            dx = ops.gemm_tn(dy0, wc.bf16_t[("e", i)], residual=skip_grad).view(n, h, w, g.cin)
            del dy0
    elif B.get("from_stem"):
        ys, bns = S["stem"]
        dvs, bnp = ops.dwconv_backward(ys, blk._depthwise_conv.weight, g.k, g.s, g.pads, dy1, ddw, bn=bns)
        dgs, dbs = G(net._bn0.weight), G(net._bn0.bias)
        c1, c2 = ops.bn_bwd_finalize(bnp, bns.count, training, dgs, dbs)
        grads["_bn0.weight"], grads["_bn0.bias"] = dgs, dbs
        cs = geom.stem_out
        dys = ops.ew_backward(1, ys.view(n, h * w, cs), bns, 0, du=dvs.view(n, h * w, cs), dv_given=True, c1=c1, c2=c2)
        dws = G(net._conv_stem.weight)
        ops.stem_wgrad(S["images"], dys.view(n, h, w, cs), geom.stem_pads, dws, patches=S["patches"])
        grads["_conv_stem.weight"] = dws
        dx = None
    else:
        dxd, _ = ops.dwconv_backward(B["x_in"], blk._depthwise_conv.weight, g.k, g.s, g.pads, dy1, ddw, bn=None)
        if g.skip:
            dxd, _ = ops.ew_forward(dxd.view(n, h * w, g.cin), residual=dx.view(n, h * w, g.cin))
        dx = dxd.view(n, h, w, g.cin)
    del dy1
    return dx


def _backward(net, S, dfeat, direct=False):
    """dfeat: [N,Chead] fp32.  Returns {state-dict key: gradient tensor} for every trainable parameter.
    direct=True: kernels write straight into the (freshly zeroed) `.grad` views of a flat gradient buffer (FlatAdamW),
    so autograd has nothing to accumulate for this tower."""
    geom, wc, training = net.geom, net._weights(), S["training"]
    grads = {}

    def G(param):
        return param.grad if direct else torch.empty_like(param)

    n, h, w = dfeat.shape[0], S["h"], S["w"]
    dvec = dfeat * (1.0 / (h * w))
    if S["dropout_mult"] is not None:
        dvec = dvec * S["dropout_mult"]
    dvec = dvec.contiguous()
    x_last = S["x_last"]
    dyh = _bn_backward(S["yh"], S["bnh"], training, net._bn1, grads, "_bn1", 1, dvec=dvec, G=G)
    dyh2 = dyh.view(n * h * w, geom.head_out)
    gh = G(net._conv_head.weight)
    ops.gemm_wgrad(dyh2, x_last.view(n * h * w, -1), out=gh.view(gh.shape[0], -1))
    grads["_conv_head.weight"] = gh
    dx = ops.gemm_tn(dyh2, wc.bf16_t[("h",)]).view(x_last.shape)
    del dyh, dyh2
    # data parallel: finished gradient ranges are all-reduced on a side stream while the backward goes on (optim.FlatAdamW)
    opt = getattr(net, "_flat_optimizer", None) if direct else None
    early = opt is not None and opt.single_use(net)
    if early:
        opt.reduce_params([net._conv_head.weight, net._bn1.weight, net._bn1.bias])
    for i in reversed(range(len(net._blocks))):
        dx = _block_backward(net, i, S["blocks"][i], dx, n, training, grads, G, S)
        if early:
            opt.reduce_params(list(net._blocks[i].parameters()))
    if S.get("stem_materialised"):
        ys, bns = S["stem"]
        nn_, hs, ws, cs = ys.shape
        dys = _bn_backward(ys.view(nn_, hs * ws, cs), bns, training, net._bn0, grads, "_bn0", 1, du=dx.view(nn_, hs * ws, cs), G=G)
        dws = G(net._conv_stem.weight)
        ops.stem_wgrad(S["images"], dys.view(nn_, hs, ws, cs), geom.stem_pads, dws, patches=S["patches"])
        grads["_conv_stem.weight"] = dws
    if early:
        opt.reduce_params([net._conv_stem.weight, net._bn0.weight, net._bn0.bias], flush=True)
    return grads


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, images, drop_rowscales, dropout_mult, want_raw, *params):
        feat, S, raw = _forward(net, images, net.training, drop_rowscales, dropout_mult, want_raw)
        ctx.net, ctx.S = net, S
        if want_raw:
            ctx.mark_non_differentiable(raw)
            return feat, raw
        return feat

    @staticmethod
    def backward(ctx, dfeat, *unused):
        net = ctx.net
        # direct mode: a FlatAdamW owns zeroed flat `.grad` views and this is the tower's first backward since zero_grad()
        opt = getattr(net, "_flat_optimizer", None)
        direct = opt is not None and opt.zero_count != getattr(net, "_direct_written_at", -1) and all(p.grad is not None for p in net.parameters())
        grads = _backward(net, ctx.S, dfeat.contiguous().float(), direct=direct)
        ctx.S = None
        if direct:
            net._direct_written_at = opt.zero_count
            return (None,) * (5 + len(list(net.parameters())))
        out = [grads.get(name) for name, _ in net.named_parameters()]
        return (None, None, None, None, None, *out)


def choose_view_plan(keep_bytes, lean_bytes, free_bytes, reserve=40 << 30):
    """Memory plan of a multi-view step: the cheapest plan whose saved state leaves `reserve` bytes for the transients of the
    backward (U, dA2, A2, dY1, dV0 of the largest block: <= 35 GB at c3).  keep: every view's full state; lean: without the expand
    convs' outputs (re-made by one GEMM each in the backward); recompute: one view resident at a time (two extra forwards)."""
    if keep_bytes + reserve <= free_bytes:
        return "keep"
    if lean_bytes + reserve <= free_bytes:
        return "lean"
    return "recompute"


class _MultiViewFn(torch.autograd.Function):
    """Memory plan for the multi-view loss (loss/breast_clip.py: two image views per step, clip.py:103-112) at the metric
    scale: the saved pre-BN tensors of ONE EN-B5 view at B = 64, 1520x912 are 91.5 GB, two do not fit 180 GB.  The forward
    runs every view WITHOUT keeping anything (features only); the backward re-runs one view at a time with saving (same
    drop-connect / dropout draws, batch statistics bit-identical, running statistics not updated again) and backpropagates
    the feature gradient the loss delivered.  Peak = one view's state; cost = one extra forward per view."""

    @staticmethod
    def forward(ctx, net, views, draws, *params):
        feats = []
        for v, (scales, mult) in zip(views, draws):
            f, _, _ = _forward(net, v, net.training, scales, mult, False, save=False)
            feats.append(f)
        ctx.net, ctx.views, ctx.draws, ctx.training = net, views, draws, net.training
        return tuple(feats)

    @staticmethod
    def backward(ctx, *dfeats):
        net = ctx.net
        total = None
        _UPDATE_RUNNING[0] = False
        try:
            for v, (scales, mult), d in zip(ctx.views, ctx.draws, dfeats):
                if d is None:
                    continue
                _, S, _ = _forward(net, v, ctx.training, scales, mult, False, save=True)
                g = _backward(net, S, d.contiguous().float(), direct=False)
                del S
                if total is None:
                    total = g
                else:
                    for k, t in g.items():
                        total[k].add_(t)
        finally:
            _UPDATE_RUNNING[0] = True
        out = [None if total is None else total.get(name) for name, _ in net.named_parameters()]
        return (None, None, None, *out)


class EfficientNet(nn.Module):
    """Drop-in for the reference EfficientNet (efficientnet_custom.py:143).  `num_classes`, `include_top`, `advprop`,
    `weights_path` are accepted for signature compatibility; this copy has no `_fc` either (:211 is commented out)."""

    def __init__(self, model_name="efficientnet-b2", stochastic=True, geometry=None):
        super().__init__()
        if model_name not in VALID_MODELS:
            raise ValueError("model_name should be one of: " + ", ".join(VALID_MODELS))
        self.model_name = model_name
        self.geom = geometry if geometry is not None else net_geometry(model_name)   # `geometry`: truncated towers in tests
        g = self.geom
        self._conv_stem = _ConvParams(g.stem_out, 3, 3)
        self._bn0 = _BNParams(g.stem_out)
        self._blocks = nn.ModuleList([MBConvBlock(b) for b in g.blocks])
        self._conv_head = _ConvParams(g.head_out, g.blocks[-1].cout, 1)
        self._bn1 = _BNParams(g.head_out)
        self.out_dim = g.head_out
        self.stochastic = stochastic      # False: drop-connect / dropout off in train mode (parity runs)
        self._wcache = None
        # input edge: a [B,1,H,W] uint8 batch is normalised on load like datasets/imagetext.py:129-134 with the
        # mean / std of the shipped configs (configs/pre_train_b5_clip.yaml:23-24)
        self.input_mean, self.input_std = 0.3089279, 0.25053555408335154

    @classmethod
    def from_name(cls, model_name, in_channels=3, **override_params):
        if in_channels != 3:
            raise ValueError("the B200 stem kernel is specialised for 3 input channels")
        return cls(model_name)

    # file names of the ImageNet checkpoints the reference downloads (efficient_net_custom_utils.py:556-579), as torch.hub caches them
    _PRETRAINED = {"efficientnet-b0": "efficientnet-b0-355c32eb.pth", "efficientnet-b1": "efficientnet-b1-f1951068.pth",
                   "efficientnet-b2": "efficientnet-b2-8bb594d6.pth", "efficientnet-b3": "efficientnet-b3-5fb5a3c3.pth",
                   "efficientnet-b4": "efficientnet-b4-6ed6700e.pth", "efficientnet-b5": "efficientnet-b5-b6417697.pth",
                   "efficientnet-b6": "efficientnet-b6-c76e70fd.pth", "efficientnet-b7": "efficientnet-b7-dcc49843.pth"}

    @classmethod
    def _find_pretrained(cls, model_name, weights_path, advprop):
        """weights_path argument > $MCLIP_EFFICIENTNET_WEIGHTS (file, or directory holding the reference's file names) > the
        torch.hub checkpoint cache the reference itself fills (model_zoo.load_url, efficient_net_custom_utils.py:602)."""
        import os
        if isinstance(weights_path, str):
            return weights_path
        fname = ("adv-" if advprop else "") + cls._PRETRAINED[model_name]
        cands = []
        env = os.environ.get("MCLIP_EFFICIENTNET_WEIGHTS")
        if env:
            cands += [env] if os.path.isfile(env) else [os.path.join(env, fname)]
        try:
            cands.append(os.path.join(torch.hub.get_dir(), "checkpoints", fname))
        except Exception:
            pass
        if not advprop:
            return next((c for c in cands if os.path.isfile(c)), None)
        import glob        # advprop checkpoints carry other hashes: match by prefix
        for c in cands:
            hit = glob.glob(os.path.join(os.path.dirname(c), f"adv-{model_name}-*.pth"))
            if hit:
                return hit[0]
        return None

    @classmethod
    def from_pretrained(cls, model_name, weights_path=None, advprop=False, in_channels=3, num_classes=1000, **override_params):
        """Reference :340-373 loads ImageNet weights (downloaded on first use).  There is no network here: the checkpoint is
        taken from `weights_path`, $MCLIP_EFFICIENTNET_WEIGHTS or the torch.hub cache; if none exists the tower keeps its
        random initialisation and says so LOUDLY (training results differ from the reference's in that case)."""
        model = cls.from_name(model_name, in_channels=in_channels)
        path = cls._find_pretrained(model_name, weights_path, advprop)
        if path is not None:
            sd = torch.load(path, map_location="cpu")
            sd = {k: v for k, v in sd.items() if not k.startswith("_fc.")}
            ret = model.load_state_dict(sd, strict=False)
            assert not ret.unexpected_keys, f"unexpected keys in {path}: {ret.unexpected_keys[:5]}"
            assert not ret.missing_keys, f"missing keys in {path}: {ret.missing_keys[:5]}"
        else:
            import logging
            import warnings
            msg = (f"[mammoclip_b200] no ImageNet checkpoint found for {model_name}: the image tower starts from RANDOM weights, unlike "
                   f"the reference's EfficientNet.from_pretrained (efficientnet_custom.py:340-373).  Provide image_encoder.weights_path, "
                   f"set MCLIP_EFFICIENTNET_WEIGHTS, or place {cls._PRETRAINED[model_name]} in the torch.hub checkpoint cache.")
            warnings.warn(msg, stacklevel=2)
            logging.getLogger(__name__).warning(msg)
        return model

    def _weights(self):
        if self._wcache is None or self._wcache.key != _WeightCache._key(self):
            self._wcache = _WeightCache(self)
        return self._wcache

    def _stochastic_inputs(self, n, device):
        """Per-sample drop-connect scales mask/keep (efficient_net_custom_utils.py:145-154; rate = 0.2*idx/len, :277-279)
        and the dropout multiplier of the pooled features (:312)."""
        if not (self.training and self.stochastic):
            return None, None
        nb = len(self._blocks)
        idx = [i for i, blk in enumerate(self._blocks) if blk.geom.skip and DROP_CONNECT_RATE * float(i) / nb > 0]
        keep = getattr(self, "_dc_keep", None)
        if keep is None or keep.device != device:
            keep = torch.tensor([1.0 - DROP_CONNECT_RATE * float(i) / nb for i in idx], device=device).view(-1, 1)
            object.__setattr__(self, "_dc_keep", keep)
        allscales = torch.floor(keep + torch.rand(len(idx), n, device=device)) / keep      # one launch set for all blocks
        scales = {i: allscales[j] for j, i in enumerate(idx)}
        p = self.geom.dropout
        mult = (torch.rand(n, self.out_dim, device=device) >= p).float() / (1.0 - p)
        return scales, mult

    @staticmethod
    def _prep(images):
        if not images.is_cuda:
            raise RuntimeError("mammoclip_b200.EfficientNet runs on a B200 only (no CPU fallback)")
        if images.dim() != 4 or images.shape[1] not in (1, 3):
            raise ValueError(f"expected images [B,3,H,W] (or the single-channel input edge [B,1,H,W]), got {tuple(images.shape)}")
        if images.shape[1] == 3:
            images = images.float()           # the reference trainer's tensor (trainer_ddp.py:288-291)
        elif images.dtype not in (torch.float32, torch.float16, torch.bfloat16, torch.uint8):
            images = images.float()
        elif images.dtype == torch.uint8:
            images = images.contiguous()
        ops._require_cuda(images)
        return images

    def forward(self, inputs):
        as_dict = isinstance(inputs, dict) and "image" in inputs
        images = self._prep(inputs["image"] if as_dict else inputs)
        if getattr(self, "_flat_optimizer", None) is not None:
            self._flat_optimizer.note_forward(self)
        scales, mult = self._stochastic_inputs(images.shape[0], images.device)
        params = [p for _, p in self.named_parameters()]
        out = _EncoderFn.apply(self, images, scales, mult, as_dict, *params)
        return out if as_dict else out

    def saved_bytes(self, n, h, w, keep_y0=True):
        """bf16 bytes of the pre-BN tensors one training forward keeps for its backward (the memory plan's unit)."""
        g = self.geom
        pl, pr, pt, pb = g.stem_pads
        h, w = (h + pt + pb - 3) // 2 + 1, (w + pl + pr - 3) // 2 + 1
        elems = h * w * (g.stem_out + 32)                     # stem output + im2col patches
        for b in g.blocks:
            l, r, t, bb = b.pads
            ho, wo = (h + t + bb - b.k) // b.s + 1, (w + l + r - b.k) // b.s + 1
            elems += (h * w * b.cexp if b.expand and keep_y0 else 0) + ho * wo * (b.cexp + 2 * b.cout)
            h, w = ho, wo
        elems += h * w * g.head_out
        return 2 * n * elems

    def forward_views(self, views, plan="auto"):
        """Features of several image batches of ONE step (multi-view loss).  plan: "keep" = ordinary forwards (every view's
        state stays resident); "lean" = ordinary forwards that do not keep the expand convs' outputs (41 % of the state; each is
        re-made by one GEMM in the backward: 2 forwards + 2 backwards per step); "recompute" = _MultiViewFn (4 forwards + 2
        backwards, one view resident at a time); "auto" = the first of these whose state fits the free HBM with 40 GB to spare
        for the transients.  MCLIP_MVS_PLAN overrides "auto"."""
        views = [self._prep(v) for v in views]
        if plan == "auto":
            plan = _os.environ.get("MCLIP_MVS_PLAN", "auto")
        if plan == "auto":
            grad = self.training and torch.is_grad_enabled()
            need = sum(self.saved_bytes(v.shape[0], v.shape[2], v.shape[3]) for v in views) if grad else 0
            lean = sum(self.saved_bytes(v.shape[0], v.shape[2], v.shape[3], keep_y0=False) for v in views) if grad else 0
            free = torch.cuda.mem_get_info(views[0].device)[0] + torch.cuda.memory_reserved(views[0].device) - torch.cuda.memory_allocated(views[0].device)
            plan = choose_view_plan(need, lean, free)
        object.__setattr__(self, "last_plan", plan)
        if plan in ("keep", "lean") or not (self.training and torch.is_grad_enabled()):
            _DROP_Y0[0] = plan == "lean"
            try:
                return [self.forward(v) for v in views]
            finally:
                _DROP_Y0[0] = False
        opt = getattr(self, "_flat_optimizer", None)
        draws = []
        for v in views:
            if opt is not None:
                opt.note_forward(self)
            draws.append(self._stochastic_inputs(v.shape[0], v.device))
        params = [p for _, p in self.named_parameters()]
        return list(_MultiViewFn.apply(self, views, draws, *params))

    def extract_features(self, inputs):
        return self.forward({"image": inputs})[1]
