"""CPU/GPU-agnostic restatement (plain PyTorch, fp32 by default) of the reference hot path.

TEST INFRASTRUCTURE ONLY — the checker, never the product.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.  The product package
(`mammo-clip_b200/`) never imports it and has no CPU fallback.

Parity status: the reference ships no tests, golden vectors or checkpoints, so this
oracle is pinned against the reference ITSELF, imported in the build container by
oracle/ref_loader.py: oracle/make_goldens.py asserts port == reference on seeded inputs
(fwd + grads) and commits the resulting vectors under tests/golden/.  The text tower's arithmetic
lives in the third-party `transformers` BertModel (reference pins 4.41.1, R/environment.yml:208;
this image has 5.5.0) and is used through the same call as the reference (text_encoder.py:37,48).

Each function cites the reference lines it restates.  R = /root/reference/src/codebase/breastclip.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

# --------------------------------------------------------------------------------------------
# EfficientNet geometry  (R/model/modules/efficient_net_custom_utils.py:83-126, 466-478, 502-526;
#                         efficientnet_custom.py:50-89, 156-205)
# --------------------------------------------------------------------------------------------

# (repeats, kernel, stride, expand, in, out) of the B0 stage table, utils:502-510
_B0_STAGES = (
    (1, 3, 1, 1, 32, 16),
    (2, 3, 2, 6, 16, 24),
    (2, 5, 2, 6, 24, 40),
    (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112),
    (4, 5, 2, 6, 112, 192),
    (1, 3, 1, 6, 192, 320),
)
# name -> (width, depth, nominal resolution, dropout), utils:466-478
_COEFFS = {
    "efficientnet-b0": (1.0, 1.0, 224, 0.2),
    "efficientnet-b1": (1.0, 1.1, 240, 0.2),
    "efficientnet-b2": (1.1, 1.2, 260, 0.3),
    "efficientnet-b3": (1.2, 1.4, 300, 0.3),
    "efficientnet-b4": (1.4, 1.8, 380, 0.4),
    "efficientnet-b5": (1.6, 2.2, 456, 0.4),
    "efficientnet-b6": (1.8, 2.6, 528, 0.5),
    "efficientnet-b7": (2.0, 3.1, 600, 0.5),
}
BN_EPS = 1e-3        # utils:521
BN_MOMENTUM = 0.01   # 1 - 0.99, efficientnet_custom.py:53,166
DROP_CONNECT = 0.2   # utils:522
SE_RATIO = 0.25


def _scale_width(ch: int, width: float, divisor: int = 8) -> int:
    """utils:83-108 (round_filters)."""
    scaled = ch * width
    new = max(divisor, int(scaled + divisor / 2) // divisor * divisor)
    if new < 0.9 * scaled:
        new += divisor
    return int(new)


def _same_pad_1d(size: int, k: int, s: int) -> Tuple[int, int, int]:
    """utils:255-271: pad so that out = ceil(size/s); returns (before, after, out)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2, out


@dataclass
class BlockSpec:
    cin: int
    cexp: int
    cout: int
    k: int
    s: int
    expand: bool
    cse: int
    pad: Tuple[int, int, int, int]   # (left, right, top, bottom) as nn.ZeroPad2d stores it
    skip: bool


@dataclass
class NetSpec:
    name: str
    stem_out: int
    stem_pad: Tuple[int, int, int, int]
    blocks: List[BlockSpec]
    head_in: int
    head_out: int
    dropout: float


def effnet_spec(name: str) -> NetSpec:
    """Static geometry, including the pads that the reference freezes from the NOMINAL resolution
    (efficientnet_custom.py:170-171,192-193; utils:255-271), not from the real input size."""
    width, depth, res, dropout = _COEFFS[name]
    size = (res, res)
    pl, pr, _ = _same_pad_1d(size[1], 3, 2)
    pt, pb, _ = _same_pad_1d(size[0], 3, 2)
    stem_pad = (pl, pr, pt, pb)
    size = (-(-size[0] // 2), -(-size[1] // 2))
    blocks: List[BlockSpec] = []
    for (rep, k, s, e, cin, cout) in _B0_STAGES:
        cin, cout = _scale_width(cin, width), _scale_width(cout, width)
        rep = int(math.ceil(depth * rep))        # utils:111-126
        for r in range(rep):
            b_in = cin if r == 0 else cout
            b_s = s if r == 0 else 1
            pl, pr, ow = _same_pad_1d(size[1], k, b_s)
            pt, pb, oh = _same_pad_1d(size[0], k, b_s)
            blocks.append(BlockSpec(
                cin=b_in, cexp=b_in * e, cout=cout, k=k, s=b_s, expand=(e != 1),
                cse=max(1, int(b_in * SE_RATIO)),                 # efficientnet_custom.py:80
                pad=(pl, pr, pt, pb),
                skip=(b_s == 1 and b_in == cout),
            ))
            size = (oh, ow)
    # efficientnet_custom.py:127: the first block of a stage carries stride=[1] (a list) so `== 1` is
    # False there; every such block has cin != cout anyway, so the plain test above is equivalent.
    return NetSpec(name=name, stem_out=_scale_width(32, width), stem_pad=stem_pad, blocks=blocks,
                   head_in=blocks[-1].cout, head_out=_scale_width(1280, width), dropout=dropout)


# --------------------------------------------------------------------------------------------
# EfficientNet oracle module (state-dict compatible with the reference's EfficientNet)
# --------------------------------------------------------------------------------------------

def swish(x):
    """utils:64-80 (x * sigmoid(x); backward is the analytic derivative, which autograd reproduces)."""
    return x * torch.sigmoid(x)


class _Conv(nn.Module):
    """Weight holder + static-pad conv (utils:248-276)."""

    def __init__(self, cin, cout, k, s, groups, bias, pad):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        self.s, self.groups, self.pad = s, groups, pad
        # same default init as nn.Conv2d so seeded inits line up with the reference
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            fan_in = (cin // groups) * k * k
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        if any(self.pad):
            x = F.pad(x, self.pad)
        return F.conv2d(x, self.weight, self.bias, stride=self.s, groups=self.groups)


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM, eps=BN_EPS)


def _q(t, on):
    """bf16 storage emulation: round to bf16 and back (what autocast does after every op in the reference's AMP path,
    trainer_ddp.py:293-296, and what the CUDA path does at its storage points).  Off by default (pure fp32 oracle)."""
    return t.to(torch.bfloat16).to(t.dtype) if on else t


class OracleMBConv(nn.Module):
    """efficientnet_custom.py:36-132."""

    def __init__(self, b: BlockSpec):
        super().__init__()
        self.spec = b
        if b.expand:
            self._expand_conv = _Conv(b.cin, b.cexp, 1, 1, 1, False, (0, 0, 0, 0))
            self._bn0 = _bn(b.cexp)
        self._depthwise_conv = _Conv(b.cexp, b.cexp, b.k, b.s, b.cexp, False, b.pad)
        self._bn1 = _bn(b.cexp)
        self._se_reduce = _Conv(b.cexp, b.cse, 1, 1, 1, True, (0, 0, 0, 0))
        self._se_expand = _Conv(b.cse, b.cexp, 1, 1, 1, True, (0, 0, 0, 0))
        self._project_conv = _Conv(b.cexp, b.cout, 1, 1, 1, False, (0, 0, 0, 0))
        self._bn2 = _bn(b.cout)

    def forward(self, x, drop_rate: float, drop_mask: Optional[torch.Tensor] = None, emulate: bool = False):
        inp = x
        if self.spec.expand:
            y = F.conv2d(x, _q(self._expand_conv.weight, emulate)) if emulate else self._expand_conv(x)
            x = _q(swish(self._bn0(_q(y, emulate))), emulate)            # :104-107
        x = _q(self._depthwise_conv(x), emulate)                          # :109
        x = _q(swish(self._bn1(x)), emulate)                              # :110-111
        sq = x.mean(dim=(2, 3), keepdim=True)                             # :115
        sq = self._se_expand(swish(self._se_reduce(sq)))                  # :116-118
        if emulate:
            # same algebra as :119,122 with the gate folded into per-sample project weights (the CUDA path's form)
            wg = _q(self._project_conv.weight[None, :, :, 0, 0] * torch.sigmoid(sq)[:, None, :, 0, 0], True)
            x = _q(torch.einsum("nchw,noc->nohw", x, wg), True)
        else:
            x = torch.sigmoid(sq) * x                                     # :119
            x = self._project_conv(x)                                     # :122
        x = self._bn2(x)                                                  # :123
        if self.spec.skip:
            if drop_rate and self.training:                               # :129-130, utils:129-154
                keep = 1.0 - drop_rate
                if drop_mask is None:
                    drop_mask = torch.floor(keep + torch.rand(x.shape[0], 1, 1, 1, dtype=x.dtype, device=x.device))
                x = x / keep * drop_mask.view(-1, 1, 1, 1).to(x.dtype)
            x = x + inp                                                   # :131
        return _q(x, emulate)


class OracleEfficientNet(nn.Module):
    """efficientnet_custom.py:143-313.  `drop_masks` / `dropout_mask` let a test inject the Bernoulli draws so the
    stochastic train path can be compared; `emulate_bf16` rounds stored activations / GEMM weights to bf16."""

    def __init__(self, name: str, spec: Optional[NetSpec] = None):
        super().__init__()
        sp = spec if spec is not None else effnet_spec(name)     # `spec`: truncated towers for block-level tests
        self.spec = sp
        self._conv_stem = _Conv(3, sp.stem_out, 3, 2, 1, False, sp.stem_pad)
        self._bn0 = _bn(sp.stem_out)
        self._blocks = nn.ModuleList([OracleMBConv(b) for b in sp.blocks])
        self._conv_head = _Conv(sp.head_in, sp.head_out, 1, 1, 1, False, (0, 0, 0, 0))
        self._bn1 = _bn(sp.head_out)
        self.out_dim = sp.head_out
        self.stochastic = True   # set False to disable drop-connect/dropout in train mode (parity runs)
        self.emulate_bf16 = False

    def extract_features(self, x, drop_masks=None):
        e = self.emulate_bf16
        x = _q(swish(self._bn0(_q(self._conv_stem(x), e))), e)           # :273
        n = len(self._blocks)
        for i, blk in enumerate(self._blocks):
            rate = DROP_CONNECT * float(i) / n if self.stochastic else 0.0   # :277-279
            x = blk(x, rate, None if drop_masks is None else drop_masks.get(i), e)
        y = F.conv2d(x, _q(self._conv_head.weight, e)) if e else self._conv_head(x)
        return _q(swish(self._bn1(_q(y, e))), e)                         # :283

    def forward(self, inputs, drop_masks=None, dropout_mask=None):
        as_dict = isinstance(inputs, dict) and "image" in inputs       # :298-305
        raw = self.extract_features(inputs["image"] if as_dict else inputs, drop_masks)
        x = raw.mean(dim=(2, 3))                                       # :309-311
        if self.training and self.stochastic:
            p = self.spec.dropout
            if dropout_mask is None:
                x = F.dropout(x, p, True)                              # :312
            else:
                x = x * dropout_mask.to(x.dtype) / (1.0 - p)
        return (x, raw) if as_dict else x


# --------------------------------------------------------------------------------------------
# Text tower, projection heads, CLIP wrapper  (text_encoder.py:5-49, projection.py:4-29, clip.py:14-114)
# --------------------------------------------------------------------------------------------

BERT_BASE_CASED = dict(vocab_size=28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                       intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2,
                       hidden_act="gelu", layer_norm_eps=1e-12, hidden_dropout_prob=0.1,
                       attention_probs_dropout_prob=0.1)


def make_bert(**overrides):
    from transformers import BertConfig, BertModel
    cfg = dict(BERT_BASE_CASED)
    cfg.update(overrides)
    return BertModel(BertConfig(**cfg))


class OracleTextEncoder(nn.Module):
    """text_encoder.py:5-49 with `pretrained: False` (BertModel from a config; no hub access here)."""

    def __init__(self, **bert_overrides):
        super().__init__()
        self.text_encoder = make_bert(**bert_overrides)
        self.out_dim = self.text_encoder.config.hidden_size

    def forward(self, tokens):
        return self.text_encoder(**tokens)["last_hidden_state"]


class OracleLinearHead(nn.Module):
    """projection.py:23-29."""

    def __init__(self, dim, proj):
        super().__init__()
        self.projection = nn.Linear(dim, proj)

    def forward(self, x):
        return self.projection(x)


class OracleBreastClip(nn.Module):
    """clip.py:14-114 (cnn image encoder, eos pooling, linear heads)."""

    def __init__(self, enc_name: str, proj_dim: int = 512, temperature: float = 0.07, **bert_overrides):
        super().__init__()
        self.image_encoder = OracleEfficientNet(enc_name)
        self.text_encoder = OracleTextEncoder(**bert_overrides)
        self.image_projection = OracleLinearHead(self.image_encoder.out_dim, proj_dim)
        self.text_projection = OracleLinearHead(self.text_encoder.out_dim, proj_dim)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / temperature))   # :39-41

    def encode_image(self, image):
        return self.image_encoder(image)

    def encode_text(self, tokens):
        feats = self.text_encoder(tokens)
        eos = tokens["attention_mask"].sum(dim=-1) - 1                  # :65-68
        return feats[torch.arange(feats.shape[0], device=feats.device), eos]

    def _embed_img(self, img):
        e = self.image_projection(self.encode_image(img))
        return e / e.norm(dim=1, keepdim=True)                          # :90

    def _embed_txt(self, tok):
        e = self.text_projection(self.encode_text(tok))
        return e / e.norm(dim=1, keepdim=True)                          # :91

    def forward(self, batch, device=None):
        img = self._embed_img(batch["images"])
        txt = self._embed_txt(batch["text_tokens"])
        out = {"image_embeddings": img, "text_embeddings": txt,
               "labels": torch.arange(img.shape[0], device=img.device),
               "logit_scale": self.logit_scale.exp()}                   # :94-100
        if "text_tokens2" in batch and "image_views" in batch:         # :103-112
            out["text_embeddings2"] = self._embed_txt(batch["text_tokens2"])
            out["image_view_embeddings"] = self._embed_img(batch["image_views"])
        return out


# --------------------------------------------------------------------------------------------
# Loss  (loss/breast_clip_contrastive.py:9-59, loss/breast_clip.py:29-127, util/dist_autograd.py:4-26)
# --------------------------------------------------------------------------------------------

class _AllGatherWithGrad(torch.autograd.Function):
    """dist_autograd.py:4-26 with partial=False: fwd all_gather, bwd reduce_scatter(SUM)."""

    @staticmethod
    def forward(ctx, x):
        out = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
        dist.all_gather(out, x.contiguous())
        return tuple(out)

    @staticmethod
    def backward(ctx, *grads):
        g = torch.zeros_like(grads[0])
        dist.reduce_scatter(g, [t.contiguous() for t in grads], dist.ReduceOp.SUM)
        return g


def gather_all(x):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return torch.cat(_AllGatherWithGrad.apply(x), 0)
    return x


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _pair(a, all_b, b, all_a, scale, labels, eps):
    """One (rows, columns) pair: CE of (scale*a)@all_b.T and of (scale*b)@all_a.T."""
    return (F.cross_entropy(scale * a @ all_b.T, labels, label_smoothing=eps),
            F.cross_entropy(scale * b @ all_a.T, labels, label_smoothing=eps))


def contrastive_loss(image_embeddings, text_embeddings, labels, logit_scale, is_train,
                     label_smoothing=0.0, return_parts=False, **_):
    """breast_clip_contrastive.py:28-59."""
    all_i, all_t = gather_all(image_embeddings), gather_all(text_embeddings)
    labels = labels + _rank() * labels.size(0)
    eps = label_smoothing if is_train else 0.0
    i2t, t2i = _pair(image_embeddings, all_t, text_embeddings, all_i, logit_scale, labels, eps)
    total = 0.75 * i2t + 0.25 * t2i
    return (total, i2t, t2i) if return_parts else total


def mvs_loss(image_embeddings, text_embeddings, text_embeddings2, image_view_embeddings, labels, logit_scale,
             is_train, label_smoothing=0.0, i2i_weight=0.0, t2t_weight=0.0, return_parts=False, **_):
    """breast_clip.py:29-127."""
    i1, t1, t2, i2 = image_embeddings, text_embeddings, text_embeddings2, image_view_embeddings
    a_i1, a_t1, a_t2, a_i2 = gather_all(i1), gather_all(t1), gather_all(t2), gather_all(i2)
    labels = labels + _rank() * labels.size(0)
    eps = label_smoothing if is_train else 0.0
    i2t = t2i = 0.0
    for (im, a_im, tx, a_tx) in ((i1, a_i1, t1, a_t1), (i2, a_i2, t1, a_t1), (i1, a_i1, t2, a_t2), (i2, a_i2, t2, a_t2)):
        x, y = _pair(im, a_tx, tx, a_im, logit_scale, labels, eps)
        i2t, t2i = i2t + x, t2i + y
    i2t, t2i = i2t / 4.0, t2i / 4.0
    x, y = _pair(i1, a_i2, i2, a_i1, logit_scale, labels, 0.0)
    i2i = (x + y) / 2.0
    x, y = _pair(t2, a_t1, t1, a_t2, logit_scale, labels, 0.0)
    t2t = (x + y) / 2.0
    total = (i2t + t2i) / 2.0 + i2i * i2i_weight + t2t * t2t_weight
    return (total, i2t, t2i, i2i, t2t) if return_parts else total


# --------------------------------------------------------------------------------------------
# Synthetic inputs, shared by goldens, tests and bench
# --------------------------------------------------------------------------------------------

def synth_images(batch, h, w, seed=1234, identical_channels=True, device="cpu"):
    """[B,3,H,W] fp32 with NHWC strides, as trainer_ddp.py:288-291 delivers them."""
    g = torch.Generator().manual_seed(seed)
    if identical_channels:
        x = torch.randn(batch, h, w, 1, generator=g).expand(batch, h, w, 3).contiguous()
    else:
        x = torch.randn(batch, h, w, 3, generator=g)
    return x.to(device).permute(0, 3, 1, 2)


def synth_tokens(batch, length, seed=4321, vocab=28996, device="cpu", lo=1000):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min(8, length), length + 1, (batch,), generator=g)
    ids = torch.randint(min(lo, vocab // 2), vocab, (batch, length), generator=g)
    pos = torch.arange(length)[None, :]
    mask = (pos < lens[:, None]).long()
    ids[:, 0] = min(101, vocab - 2)
    ids[torch.arange(batch), lens - 1] = min(102, vocab - 1)
    ids = ids * mask
    return {"input_ids": ids.to(device), "token_type_ids": torch.zeros_like(ids).to(device),
            "attention_mask": mask.to(device)}


# --------------------------------------------------------------------------------------------
# Deterministic, init-order-independent weights (so goldens need not store 30 MB of parameters)
# --------------------------------------------------------------------------------------------

def _name_seed(name: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in (name + "#" + str(seed)).encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFFFFFFFFFF


@torch.no_grad()
def fill_deterministic(module: nn.Module, seed: int = 0) -> nn.Module:
    """Every tensor of the state dict becomes a function of (its name, seed) only: conv/linear weights
    ~ N(0, 1/fan_in) (x1.2 so activations neither die nor explode through 23-39 blocks), norm weights
    ~ U(0.6,1.4), biases ~ N(0,0.1), running_mean ~ N(0,0.1), running_var ~ U(0.6,1.4)."""
    for name, t in module.state_dict().items():
        g = torch.Generator().manual_seed(_name_seed(name, seed))
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t.zero_()
        elif leaf == "logit_scale":
            t.fill_(math.log(1 / 0.07))                # clip.py:39-41
        elif leaf == "running_var":
            t.copy_(torch.rand(t.shape, generator=g) * 0.8 + 0.6)
        elif leaf == "running_mean":
            t.copy_(torch.randn(t.shape, generator=g) * 0.1)
        elif leaf == "position_ids" or not t.is_floating_point():
            continue
        elif t.dim() <= 1:
            is_norm_w = leaf == "weight"
            if is_norm_w:
                t.copy_(torch.rand(t.shape, generator=g) * 0.8 + 0.6)
            else:
                t.copy_(torch.randn(t.shape, generator=g) * 0.1)
        elif "embeddings" in name and t.dim() == 2:
            t.copy_(torch.randn(t.shape, generator=g) * 0.05)
        else:
            fan_in = t[0].numel()
            t.copy_(torch.randn(t.shape, generator=g) * (1.2 / math.sqrt(fan_in)))
    return module
