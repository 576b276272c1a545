"""BERT encoder forward AND backward on the sm_100a kernels (embedding+LayerNorm, tcgen05 Linear GEMMs with fused
bias / GELU / dropout / residual epilogues, masked-softmax attention, LayerNorm and their gradients), for a Hugging Face
`BertModel` parameter container (names and shapes untouched, so checkpoints and DDP see the same parameters).

Forward = what `BertModel(**tokens)["last_hidden_state"]` computes in the reference (text_encoder.py:47-49; transformers
modeling_bert.py BertEmbeddings / BertSelfAttention / BertSelfOutput / BertIntermediate / BertOutput), the unused pooler
excepted.  Dropout (p = 0.1 in train mode) uses explicit keep-masks so that the backward pass sees the same draws.

Backward = the autograd of that graph (the reference trains every BERT parameter, optimizer/__init__.py:23-31), written
out by hand on the same library: LayerNorm / GELU / attention / embedding backward kernels (csrc/bert.cu), data gradients
on `mclip_gemm_tn` with transposed bf16 weights, weight gradients on `mclip_gemm_wgrad`, bias gradients on `mclip_colsum`.
With a `FlatAdamW` attached, parameter gradients are written straight into the flat gradient buffer."""
import torch

from ... import ops


class _BertWeights:
    """bf16 copies of the Linear weights (QKV fused into one [3H,H] operand) and their transposes for the data-gradient
    GEMMs, refreshed once per forward by one table-driven kernel."""

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
            h, hin = a.query.weight.shape
            qkv = torch.empty((3 * h, hin), dtype=torch.bfloat16, device=dev)
            qkv_t = torch.empty((hin, 3 * h), dtype=torch.bfloat16, device=dev)         # [in, 3*out]: B operand of dX = dQKV @ Wqkv
            for i, lin in enumerate((a.query, a.key, a.value)):
                entries.append((lin.weight.detach(), qkv[i * h:(i + 1) * h], qkv_t[:, i * h:(i + 1) * h], 0, 3 * h))
            wo = torch.empty_like(so.dense.weight, dtype=torch.bfloat16)
            w1 = torch.empty_like(it.dense.weight, dtype=torch.bfloat16)
            w2 = torch.empty_like(ou.dense.weight, dtype=torch.bfloat16)
            wo_t = torch.empty(so.dense.weight.shape[::-1], dtype=torch.bfloat16, device=dev)
            w1_t = torch.empty(it.dense.weight.shape[::-1], dtype=torch.bfloat16, device=dev)
            w2_t = torch.empty(ou.dense.weight.shape[::-1], dtype=torch.bfloat16, device=dev)
            entries += [(so.dense.weight.detach(), wo, wo_t), (it.dense.weight.detach(), w1, w1_t), (ou.dense.weight.detach(), w2, w2_t)]
            self.layers.append(dict(qkv=qkv, wo=wo, w1=w1, w2=w2, qkv_t=qkv_t, wo_t=wo_t, w1_t=w1_t, w2_t=w2_t))
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
    hdim, heads, nl = cfg.hidden_size, cfg.num_attention_heads, len(bert.encoder.layer)

    # one draw per mask family for the whole tower (3 launches instead of 3 per layer)
    def keep(shape, p):
        return (torch.rand(shape, device=device) >= p).to(torch.uint8) if p > 0 else None

    hid = keep((2 * nl + 1, b * l, hdim), p_h)
    att = keep((nl, b, heads, l, l), p_a)
    masks = {"emb": hid[2 * nl] if hid is not None else None, "layers": []}
    for i in range(nl):
        masks["layers"].append({"probs": att[i] if att is not None else None,
                                "attn_out": hid[2 * i] if hid is not None else None,
                                "ffn_out": hid[2 * i + 1] if hid is not None else None})
    return masks


_NO_MASKS = {"probs": None, "attn_out": None, "ffn_out": None}


def _kernel_forward(bert, ids, tts, amask, masks, save=None):
    """-> x [B*L, H] bf16.  `save`: list that receives, per layer, the activations the backward pass needs."""
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
        mk = masks["layers"][i] if masks else _NO_MASKS
        bqkv = torch.cat([a.query.bias, a.key.bias, a.value.bias]).detach()
        qkv = ops.gemm_tn(x, lw["qkv"], bias=bqkv)
        if save is not None:
            ctx, lse = ops.bert_attention(qkv, amask, b, l, heads, hdim // heads, mk["probs"], sa, want_lse=True)
        else:
            ctx = ops.bert_attention(qkv, amask, b, l, heads, hdim // heads, mk["probs"], sa)
        h1 = ops.gemm_tn(ctx, lw["wo"], bias=so.dense.bias.detach(), residual=x, dropmask=mk["attn_out"], drop_scale=sh)
        x1 = ops.layernorm(h1, so.LayerNorm.weight, so.LayerNorm.bias, eps)
        # the fused erf-GELU epilogue also stores the pre-activation (bf16) when the backward pass will need GELU'
        pre = torch.empty((x1.shape[0], lw["w1"].shape[0]), dtype=torch.bfloat16, device=x1.device) if save is not None else None
        inter = ops.gemm_tn(x1, lw["w1"], bias=it.dense.bias.detach(), act=1, aux_pre=pre)
        h2 = ops.gemm_tn(inter, lw["w2"], bias=ou.dense.bias.detach(), residual=x1, dropmask=mk["ffn_out"], drop_scale=sh)
        x_out = ops.layernorm(h2, ou.LayerNorm.weight, ou.LayerNorm.bias, eps)
        if save is not None:
            save.append(dict(x=x, qkv=qkv, lse=lse, ctx=ctx, h1=h1, x1=x1, pre=pre, inter=inter, h2=h2))
        x = x_out
    return x


def _kernel_backward(bert, ids, tts, amask, masks, saved, dout, G, accumulate):
    """dout: [B*L, H] bf16, gradient of the last hidden state.  G(param) -> the tensor that receives d/d param (written, or
    added to when `accumulate`).  Returns {id(param): gradient tensor}."""
    cfg = bert.config
    b, l = ids.shape
    hdim, heads = cfg.hidden_size, cfg.num_attention_heads
    eps = cfg.layer_norm_eps
    sh = 1.0 / (1.0 - cfg.hidden_dropout_prob) if masks else 1.0
    sa = 1.0 / (1.0 - cfg.attention_probs_dropout_prob) if masks else 1.0
    w = _weights(bert)
    grads = {}

    def out(param):
        g = G(param)
        grads[id(param)] = g
        return g

    def linear_grads(lin, dy, xin):
        """dW = dy^T xin, db = column sums of dy (dy may be a column slice of a wider tensor)."""
        ops.gemm_wgrad(dy, xin, out=out(lin.weight), accumulate=accumulate)
        ops.colsum(dy, out(lin.bias), accumulate=accumulate)

    dx = dout
    for i in reversed(range(len(bert.encoder.layer))):
        layer = bert.encoder.layer[i]
        a, so, it, ou = layer.attention.self, layer.attention.output, layer.intermediate, layer.output
        lw, S = w.layers[i], saved[i]
        mk = masks["layers"][i] if masks else _NO_MASKS
        # BertOutput: x_out = LN(dropout(dense(inter)) + x1)
        dh2, dh2d = ops.layernorm_backward(S["h2"], dx, ou.LayerNorm.weight, eps, out(ou.LayerNorm.weight), out(ou.LayerNorm.bias),
                                           dropmask=mk["ffn_out"], drop_scale=sh, accumulate=accumulate)
        d_inter = ops.gemm_tn(dh2d, lw["w2_t"])
        linear_grads(ou.dense, dh2d, S["inter"])
        # BertIntermediate: inter = gelu(dense(x1))
        d_pre = ops.gelu_backward(d_inter, S["pre"])
        dx1 = ops.gemm_tn(d_pre, lw["w1_t"], residual=dh2)
        linear_grads(it.dense, d_pre, S["x1"])
        # BertSelfOutput: x1 = LN(dropout(dense(ctx)) + x)
        dh1, dh1d = ops.layernorm_backward(S["h1"], dx1, so.LayerNorm.weight, eps, out(so.LayerNorm.weight), out(so.LayerNorm.bias),
                                           dropmask=mk["attn_out"], drop_scale=sh, accumulate=accumulate)
        d_ctx = ops.gemm_tn(dh1d, lw["wo_t"])
        linear_grads(so.dense, dh1d, S["ctx"])
        # BertSelfAttention
        dqkv = ops.bert_attention_backward(S["qkv"], d_ctx, S["lse"], amask, b, l, heads, hdim // heads, mk["probs"], sa)
        dx = ops.gemm_tn(dqkv, lw["qkv_t"], residual=dh1)
        for j, lin in enumerate((a.query, a.key, a.value)):
            linear_grads(lin, dqkv[:, j * hdim:(j + 1) * hdim], S["x"])
        saved[i] = None
    emb = bert.embeddings
    ops.bert_embed_backward(ids, tts, emb.word_embeddings.weight, emb.position_embeddings.weight, emb.token_type_embeddings.weight,
                            emb.LayerNorm.weight, eps, dx, out(emb.word_embeddings.weight), out(emb.position_embeddings.weight),
                            out(emb.token_type_embeddings.weight), out(emb.LayerNorm.weight), out(emb.LayerNorm.bias),
                            dropmask=masks["emb"] if masks else None, drop_scale=sh, accumulate=accumulate)
    return grads


class _BertFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, bert, ids, tts, amask, masks, need_grad, *params):
        saved = [] if need_grad else None
        out = _kernel_forward(bert, ids, tts, amask, masks, save=saved)
        ctx.owner, ctx.bert, ctx.args, ctx.saved = owner, bert, (ids, tts, amask, masks), saved
        return out.view(ids.shape[0], ids.shape[1], -1).float()

    @staticmethod
    def backward(ctx, dout):
        bert, owner = ctx.bert, ctx.owner
        ids, tts, amask, masks = ctx.args
        if ctx.saved is None:
            raise RuntimeError("mammoclip_b200 BERT: backward called on a forward that ran without gradient tracking")
        trainable = [p for p in bert.parameters() if not _is_pooler(bert, p)]
        # direct mode: a FlatAdamW owns zeroed flat `.grad` views and this is the tower's first backward since zero_grad():
        # the kernels write there and autograd has nothing left to accumulate (second backward of a step, e.g. the MVS loss'
        # second text: ordinary returned gradients)
        opt = getattr(owner, "_flat_optimizer", None) if owner is not None else None
        direct = opt is not None and opt.zero_count != getattr(owner, "_direct_written_at", -1) and all(p.grad is not None for p in trainable)

        def G(param):
            if direct:
                return param.grad
            # embedding tables are only written where touched: start from zeros
            return torch.zeros_like(param) if _is_table(bert, param) else torch.empty_like(param)

        d2 = ops.cast_bf16(dout.contiguous().float().view(-1, dout.shape[-1]))
        grads = _kernel_backward(bert, ids, tts, amask, masks, ctx.saved, d2, G, accumulate=False)
        ctx.saved = None
        if direct:
            owner._direct_written_at = opt.zero_count
            if hasattr(opt, "single_use") and opt.single_use(owner):            # data parallel: the text tower's gradients are final, reduce them while the image tower's backward runs
                opt.reduce_params(trainable, flush=True)
            return (None,) * (7 + len(list(bert.parameters())))
        return (None, None, None, None, None, None, None, *[grads.get(id(p)) for p in bert.parameters()])


def _is_table(bert, p):
    e = bert.embeddings
    return p is e.word_embeddings.weight or p is e.position_embeddings.weight or p is e.token_type_embeddings.weight


def _is_pooler(bert, p):
    pool = getattr(bert, "pooler", None)
    return pool is not None and any(p is q for q in pool.parameters())


def bert_forward(bert, input_ids, token_type_ids, attention_mask, training, owner=None):
    """-> last_hidden_state [B, L, H] fp32 (differentiable w.r.t. the BertModel parameters).  `owner`: the module a
    FlatAdamW may have been attached to (direct gradient writes)."""
    ids = input_ids.contiguous()
    tts = token_type_ids.contiguous() if token_type_ids is not None else None
    amask = attention_mask.contiguous().long()
    cfg = bert.config
    if cfg.hidden_size != 768 or cfg.hidden_size // cfg.num_attention_heads != 64 or cfg.hidden_act != "gelu":
        raise NotImplementedError("the B200 text tower is built for BERT-base geometry (hidden 768, head dim 64, erf-GELU)")
    use_drop = training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0)
    masks = _make_masks(bert, ids.shape[0], ids.shape[1], ids.device) if use_drop else None
    params = list(bert.parameters())
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    opt = getattr(owner, "_flat_optimizer", None) if owner is not None else None
    if opt is not None and need_grad and hasattr(opt, "note_forward"):
        opt.note_forward(owner)
    return _BertFn.apply(owner, bert, ids, tts, amask, masks, need_grad, *params)
