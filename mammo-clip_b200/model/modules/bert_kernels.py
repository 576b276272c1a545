"""BERT encoder forward on the sm_100a kernels (embedding+LayerNorm, tcgen05 Linear GEMMs with fused
bias / GELU / dropout / residual epilogues, masked-softmax attention, LayerNorm), for a Hugging Face `BertModel`
parameter container (names and shapes untouched, so checkpoints and DDP see the same parameters).

Forward = what `BertModel(**tokens)["last_hidden_state"]` computes in the reference (text_encoder.py:47-49; transformers
modeling_bert.py BertEmbeddings / BertSelfAttention / BertSelfOutput / BertIntermediate / BertOutput), the unused pooler
excepted.  Dropout (p = 0.1 in train mode) uses explicit keep-masks so that the backward pass sees the same draws.

Backward (phase 1, SURVEY §0 "text tower" row): the reference trains every BERT parameter, the north star names only the
text *forward* for hand-written kernels; gradients are obtained by re-running the layer stack through PyTorch autograd with
the same dropout masks (`_bert_torch`).  BERT backward kernels are the first "next" row of SURVEY §8(f)."""
import math

import torch
import torch.nn.functional as F

from ... import ops


class _BertWeights:
    """bf16 copies of the Linear weights (QKV fused into one [3H,H] operand), refreshed once per forward."""

    def __init__(self, bert):
        self.key = None
        self.layers = []

    @staticmethod
    def _key(bert):
        return tuple(p.data_ptr() for p in bert.parameters())

    def build(self, bert):
        dev = bert.embeddings.word_embeddings.weight.device
        entries, self.layers = [], []
        for layer in bert.encoder.layer:
            a, so, it, ou = layer.attention.self, layer.attention.output, layer.intermediate, layer.output
            h = a.query.weight.shape[0]
            qkv = torch.empty((3 * h, a.query.weight.shape[1]), dtype=torch.bfloat16, device=dev)
            for i, lin in enumerate((a.query, a.key, a.value)):
                entries.append((lin.weight.detach(), qkv[i * h:(i + 1) * h], None))
            wo = torch.empty_like(so.dense.weight, dtype=torch.bfloat16)
            w1 = torch.empty_like(it.dense.weight, dtype=torch.bfloat16)
            w2 = torch.empty_like(ou.dense.weight, dtype=torch.bfloat16)
            entries += [(so.dense.weight.detach(), wo, None), (it.dense.weight.detach(), w1, None), (ou.dense.weight.detach(), w2, None)]
            self.layers.append(dict(qkv=qkv, wo=wo, w1=w1, w2=w2))
        self.entries = entries
        self.table = ops.weight_prep(entries, dev)
        self.key = self._key(bert)

    def refresh(self, bert):
        if self.key != self._key(bert):
            self.build(bert)
        ops.weight_prep_run(self.table, len(self.entries))


def _weights(bert):
    w = getattr(bert, "_mclip_weights", None)
    if w is None:
        w = _BertWeights(bert)
        object.__setattr__(bert, "_mclip_weights", w)
    return w


def _make_masks(bert, b, l, device):
    cfg = bert.config
    p_h, p_a = cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob
    hdim, heads = cfg.hidden_size, cfg.num_attention_heads

    def keep(shape, p):
        return (torch.rand(shape, device=device) >= p).to(torch.uint8) if p > 0 else None

    masks = {"emb": keep((b * l, hdim), p_h), "layers": []}
    for _ in bert.encoder.layer:
        masks["layers"].append({"probs": keep((b, heads, l, l), p_a), "attn_out": keep((b * l, hdim), p_h), "ffn_out": keep((b * l, hdim), p_h)})
    return masks


def _kernel_forward(bert, ids, tts, amask, masks):
    cfg = bert.config
    b, l = ids.shape
    hdim, heads = cfg.hidden_size, cfg.num_attention_heads
    eps = cfg.layer_norm_eps
    sh = 1.0 / (1.0 - cfg.hidden_dropout_prob) if masks else 1.0
    sa = 1.0 / (1.0 - cfg.attention_probs_dropout_prob) if masks else 1.0
    w = _weights(bert)
    w.refresh(bert)
    emb = bert.embeddings
    x = ops.bert_embed_ln(ids, tts, emb.word_embeddings.weight, emb.position_embeddings.weight, emb.token_type_embeddings.weight,
                          emb.LayerNorm.weight, emb.LayerNorm.bias, eps, masks["emb"] if masks else None, sh)
    for i, layer in enumerate(bert.encoder.layer):
        a, so, it, ou = layer.attention.self, layer.attention.output, layer.intermediate, layer.output
        lw = w.layers[i]
        mk = masks["layers"][i] if masks else {"probs": None, "attn_out": None, "ffn_out": None}
        bqkv = torch.cat([a.query.bias, a.key.bias, a.value.bias]).detach()
        qkv = ops.gemm_tn(x, lw["qkv"], bias=bqkv)
        ctx = ops.bert_attention(qkv, amask, b, l, heads, hdim // heads, mk["probs"], sa)
        h1 = ops.gemm_tn(ctx, lw["wo"], bias=so.dense.bias.detach(), residual=x, dropmask=mk["attn_out"], drop_scale=sh)
        x1 = ops.layernorm(h1, so.LayerNorm.weight, so.LayerNorm.bias, eps)
        inter = ops.gemm_tn(x1, lw["w1"], bias=it.dense.bias.detach(), act=1)
        h2 = ops.gemm_tn(inter, lw["w2"], bias=ou.dense.bias.detach(), residual=x1, dropmask=mk["ffn_out"], drop_scale=sh)
        x = ops.layernorm(h2, ou.LayerNorm.weight, ou.LayerNorm.bias, eps)
    return x.view(b, l, hdim)


def _bert_torch(bert, ids, tts, amask, masks):
    """The same function in PyTorch ops with explicit dropout masks (autograd recompute path for the backward)."""
    cfg = bert.config
    b, l = ids.shape
    hdim, heads = cfg.hidden_size, cfg.num_attention_heads
    d = hdim // heads
    sh = 1.0 / (1.0 - cfg.hidden_dropout_prob) if masks else 1.0
    sa = 1.0 / (1.0 - cfg.attention_probs_dropout_prob) if masks else 1.0
    emb = bert.embeddings
    pos = torch.arange(l, device=ids.device)
    x = emb.word_embeddings(ids) + emb.position_embeddings(pos)[None] + emb.token_type_embeddings(tts if tts is not None else torch.zeros_like(ids))
    x = emb.LayerNorm(x)
    if masks and masks["emb"] is not None:
        x = x * (masks["emb"].view(b, l, hdim) * sh)
    bias = (1.0 - amask[:, None, None, :].to(x.dtype)) * torch.finfo(torch.float32).min
    for i, layer in enumerate(bert.encoder.layer):
        a, so, it, ou = layer.attention.self, layer.attention.output, layer.intermediate, layer.output
        mk = masks["layers"][i] if masks else {"probs": None, "attn_out": None, "ffn_out": None}

        def split(t):
            return t.view(b, l, heads, d).transpose(1, 2)

        q, k, v = split(a.query(x)), split(a.key(x)), split(a.value(x))
        p = torch.softmax((q @ k.transpose(-1, -2)).float() / math.sqrt(d) + bias, dim=-1).to(v.dtype)
        if mk["probs"] is not None:
            p = p * (mk["probs"] * sa).to(p.dtype)
        ctx = (p @ v).transpose(1, 2).reshape(b, l, hdim)
        h = so.dense(ctx)
        if mk["attn_out"] is not None:
            h = h * (mk["attn_out"].view(b, l, hdim) * sh).to(h.dtype)
        x1 = so.LayerNorm(h + x)
        o = ou.dense(F.gelu(it.dense(x1)))
        if mk["ffn_out"] is not None:
            o = o * (mk["ffn_out"].view(b, l, hdim) * sh).to(o.dtype)
        x = ou.LayerNorm(o + x1)
    return x


class _BertFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bert, ids, tts, amask, masks, *params):
        out = _kernel_forward(bert, ids, tts, amask, masks)
        ctx.bert, ctx.args = bert, (ids, tts, amask, masks)
        return out.float()

    @staticmethod
    def backward(ctx, dout):
        bert = ctx.bert
        ids, tts, amask, masks = ctx.args
        params = [p for p in bert.parameters() if not _is_pooler(bert, p)]
        with torch.enable_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = _bert_torch(bert, ids, tts, amask, masks)
        grads = torch.autograd.grad(out, params, dout.to(out.dtype), allow_unused=True)
        gmap = {id(p): g for p, g in zip(params, grads)}
        return (None, None, None, None, None, *[gmap.get(id(p)) for p in bert.parameters()])


def _is_pooler(bert, p):
    pool = getattr(bert, "pooler", None)
    return pool is not None and any(p is q for q in pool.parameters())


def bert_forward(bert, input_ids, token_type_ids, attention_mask, training):
    """-> last_hidden_state [B, L, H] fp32 (differentiable w.r.t. the BertModel parameters)."""
    ids = input_ids.contiguous()
    tts = token_type_ids.contiguous() if token_type_ids is not None else None
    amask = attention_mask.contiguous().long()
    cfg = bert.config
    if cfg.hidden_size != 768 or cfg.hidden_size // cfg.num_attention_heads != 64 or cfg.hidden_act != "gelu":
        raise NotImplementedError("the B200 text tower is built for BERT-base geometry (hidden 768, head dim 64, erf-GELU)")
    use_drop = training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0)
    masks = _make_masks(bert, ids.shape[0], ids.shape[1], ids.device) if use_drop else None
    return _BertFn.apply(bert, ids, tts, amask, masks, *bert.parameters())
