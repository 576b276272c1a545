"""Test-side checker shared by the CPU (emulated kernels) and GPU text-tower tests."""
import torch


def bert_torch_ref(bert, ids, tts, amask, masks):
    """fp32 PyTorch restatement of BertModel's last_hidden_state with EXPLICIT dropout keep-masks (test-side checker for
    the train-mode kernels: same function, same draws)."""
    import math
    import torch.nn.functional as F
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
        p = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d) + bias, dim=-1)
        if mk["probs"] is not None:
            p = p * (mk["probs"] * sa)
        ctx = (p @ v).transpose(1, 2).reshape(b, l, hdim)
        h = so.dense(ctx)
        if mk["attn_out"] is not None:
            h = h * (mk["attn_out"].view(b, l, hdim) * sh)
        x1 = so.LayerNorm(h + x)
        o = ou.dense(F.gelu(it.dense(x1)))
        if mk["ffn_out"] is not None:
            o = o * (mk["ffn_out"].view(b, l, hdim) * sh)
        x = ou.LayerNorm(o + x1)
    return x
