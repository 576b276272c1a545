"""BERT text tower on the CUDA kernels vs transformers' BertModel (the reference's text path, text_encoder.py:47-49)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _build(layers, dropout=0.0):
    from transformers import BertConfig
    from mammoclip_b200.model.modules.text_encoder import HuggingfaceTextEncoder
    from oracle import port
    cfg = BertConfig(**dict(port.BERT_BASE_CASED, num_hidden_layers=layers, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout))
    ours = HuggingfaceTextEncoder(config=cfg)
    ref = port.OracleTextEncoder(num_hidden_layers=layers, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout)
    port.fill_deterministic(ref, 0)
    ours.load_state_dict(ref.state_dict())
    return ours.cuda(), ref.cuda()


@pytest.mark.parametrize("layers,B,L", [(2, 4, 32), (12, 8, 64), (2, 3, 100), (12, 4, 256), (2, 5, 256)])   # 256 = the reference's text_max_length (pre_train_b5_clip.yaml:27)
def test_bert_forward_backward(layers, B, L):
    from oracle import port
    from transformers import BatchEncoding
    ours, ref = _build(layers)
    ours.eval(), ref.eval()
    tok = BatchEncoding(port.synth_tokens(B, L, seed=4321, device="cuda"))
    ho, hr = ours(tok), ref(tok)
    valid = tok["attention_mask"].bool()
    assert rel_err(ho[valid], hr[valid]) < 2e-2
    g = torch.Generator().manual_seed(5)
    probe = torch.randn(B, L, 768, generator=g).cuda() * valid[..., None]
    (ho * probe).sum().backward()
    (hr * probe).sum().backward()
    gr = dict(ref.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in gr.values() if p.grad is not None)
    for k, p in ours.named_parameters():
        if gr[k].grad is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, k
            continue
        d = (p.grad.double() - gr[k].grad.double()).abs().max().item()
        assert d < 5e-2 * max(gr[k].grad.abs().max().item(), 1e-2 * gmax), (k, d)


from bert_ref import bert_torch_ref as _bert_torch_ref  # noqa: E402


@pytest.mark.parametrize("B,L", [(4, 32), (3, 100), (3, 256)])
def test_bert_train_mode_forward_backward_with_shared_dropout_masks(monkeypatch, B, L):
    """Train mode (p=0.1): kernel forward + kernel backward vs fp32 autograd of the same function with the SAME keep-masks
    (embedding / attention-probability / sub-layer dropout), every trainable parameter's gradient."""
    import copy
    from oracle import port
    from transformers import BatchEncoding
    from mammoclip_b200.model.modules import bert_kernels as bk
    ours, _ = _build(2, dropout=0.1)
    ours.train()
    ref = copy.deepcopy(ours.text_encoder).float()
    tok = port.synth_tokens(B, L, seed=1, device="cuda")
    bert = ours.text_encoder
    torch.manual_seed(3)
    masks = bk._make_masks(bert, B, L, "cuda")
    monkeypatch.setattr(bk, "_make_masks", lambda *a, **k: masks)
    ho = ours(BatchEncoding(tok))
    hr = _bert_torch_ref(ref, tok["input_ids"], tok["token_type_ids"], tok["attention_mask"], masks)
    valid = tok["attention_mask"].bool()
    assert rel_err(ho[valid], hr[valid]) < 2e-2
    g = torch.Generator().manual_seed(5)
    probe = torch.randn(B, L, 768, generator=g).cuda() * valid[..., None]
    (ho * probe).sum().backward()
    (hr * probe).sum().backward()
    gr = dict(ref.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in gr.values() if p.grad is not None)
    for k, p in bert.named_parameters():
        if gr[k].grad is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, k
            continue
        assert p.grad is not None, k
        d = (p.grad.double() - gr[k].grad.double()).abs().max().item()
        assert d < 5e-2 * max(gr[k].grad.abs().max().item(), 1e-2 * gmax), (k, d)
    # a different draw is a different function
    monkeypatch.undo()
    with torch.no_grad():
        h2 = ours(BatchEncoding(tok))
    assert rel_err(h2[valid], hr[valid]) > 5e-2


def test_layernorm_backward_kernel():
    from mammoclip_b200 import ops
    torch.manual_seed(0)
    for rows, h in ((37, 768), (4096, 768), (64, 512)):
        x = (torch.randn(rows, h, device="cuda") * 2 + 0.5).bfloat16()
        dy = torch.randn(rows, h, device="cuda").bfloat16()
        gamma = torch.randn(h, device="cuda")
        beta = torch.randn(h, device="cuda")
        mask = (torch.rand(rows, h, device="cuda") >= 0.1).to(torch.uint8)
        xr = x.float().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(xr, (h,), gr, br, 1e-12).backward(dy.float())
        dg, db = torch.empty_like(gamma), torch.empty_like(beta)
        dx, dxd = ops.layernorm_backward(x, dy, gamma, 1e-12, dg, db, dropmask=mask, drop_scale=1.0 / 0.9)
        assert rel_err(dx, xr.grad) < 1e-2
        assert rel_err(dxd, xr.grad * mask / 0.9) < 1e-2
        assert rel_err(dg, gr.grad) < 2e-3 and rel_err(db, br.grad) < 2e-3
        dg2, db2 = dg.clone(), db.clone()
        dx2, dxd2 = ops.layernorm_backward(x, dy, gamma, 1e-12, dg2, db2, accumulate=True)
        assert dxd2 is dx2 and torch.equal(dx2, dx)
        assert rel_err(dg2, 2 * gr.grad) < 2e-3 and rel_err(db2, 2 * br.grad) < 2e-3


def test_gelu_kernels():
    from mammoclip_b200 import ops
    torch.manual_seed(0)
    x = (torch.randn(1000, 3072, device="cuda") * 2).bfloat16()
    dy = torch.randn(1000, 3072, device="cuda").bfloat16()
    xr = x.float().requires_grad_(True)
    y = torch.nn.functional.gelu(xr)
    y.backward(dy.float())
    assert rel_err(ops.gelu_forward(x), y) < 1e-2
    assert rel_err(ops.gelu_backward(dy, x), xr.grad) < 1e-2


@pytest.mark.parametrize("B,L,drop", [(2, 64, False), (3, 100, True), (2, 160, True), (1, 9, False), (2, 256, True), (1, 300, True)])   # 300 > 256: the SIMT kernels of bert.cu
def test_attention_backward_kernel(B, L, drop):
    """d(Q,K,V) of softmax(QK^T/8 + padding mask) (o keep-mask) V vs fp32 autograd; L > 64 exercises the multi-block path."""
    import math
    from mammoclip_b200 import ops
    torch.manual_seed(L)
    heads, d = 12, 64
    H = heads * d
    qkv = torch.randn(B * L, 3 * H, device="cuda").bfloat16()
    lens = torch.randint(max(1, L // 3), L + 1, (B,), device="cuda")
    lens[0] = L
    amask = (torch.arange(L, device="cuda")[None] < lens[:, None]).long()
    keep = (torch.rand(B, heads, L, L, device="cuda") >= 0.1).to(torch.uint8) if drop else None
    sa = 1.0 / 0.9 if drop else 1.0
    out, lse = ops.bert_attention(qkv, amask, B, L, heads, d, keep, sa, want_lse=True)
    qr = qkv.float().requires_grad_(True)
    q, k, v = (t.view(B, L, heads, d).transpose(1, 2) for t in qr.view(B, L, 3 * H).split(H, dim=-1))
    bias = (1.0 - amask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d) + bias, dim=-1)
    lse_ref = torch.logsumexp(q @ k.transpose(-1, -2) / math.sqrt(d) + bias, dim=-1)
    if drop:
        p = p * (keep * sa)
    o_ref = (p @ v).transpose(1, 2).reshape(B * L, H)
    valid = amask.bool().view(-1)
    assert rel_err(out[valid], o_ref[valid]) < 1e-2
    assert (lse - lse_ref).abs().max().item() < 1e-3
    do = (torch.randn(B * L, H, device="cuda") * valid[:, None]).bfloat16()
    o_ref.backward(do.float())
    dqkv = ops.bert_attention_backward(qkv, do, lse, amask, B, L, heads, d, keep, sa)
    ref = qr.grad
    for j, name in enumerate("qkv"):
        a, r = dqkv[:, j * H:(j + 1) * H].float(), ref[:, j * H:(j + 1) * H]
        assert rel_err(a, r) < 2e-2, name


def test_embedding_backward_kernel():
    """Duplicate ids (pad, CLS, SEP, repeated words) must sum in the word table; positions sum over the batch."""
    from mammoclip_b200 import ops
    torch.manual_seed(0)
    B, L, H, V = 5, 48, 768, 1000
    ids = torch.randint(0, 40, (B, L), device="cuda")          # many duplicates
    ids[:, 0] = 101
    ids[:, -5:] = 0
    tts = torch.randint(0, 2, (B, L), device="cuda")
    word, pos, typ = (torch.randn(n, H, device="cuda") for n in (V, 64, 2))
    gamma, beta = torch.randn(H, device="cuda"), torch.randn(H, device="cuda")
    mask = (torch.rand(B * L, H, device="cuda") >= 0.1).to(torch.uint8)
    dout = torch.randn(B * L, H, device="cuda").bfloat16()
    wr, pr, tr, gr, br = (t.clone().requires_grad_(True) for t in (word, pos, typ, gamma, beta))
    v = wr[ids] + pr[torch.arange(L, device="cuda")][None] + tr[tts]
    y = torch.nn.functional.layer_norm(v, (H,), gr, br, 1e-12).view(B * L, H) * (mask / 0.9)
    y.backward(dout.float())
    dw, dp, dt = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros_like(typ)
    dg, db = torch.empty_like(gamma), torch.empty_like(beta)
    ops.bert_embed_backward(ids, tts, word, pos, typ, gamma, 1e-12, dout, dw, dp, dt, dg, db, dropmask=mask, drop_scale=1.0 / 0.9)
    for a, r, name in ((dw, wr.grad, "word"), (dp, pr.grad, "pos"), (dt, tr.grad, "type"), (dg, gr.grad, "gamma"), (db, br.grad, "beta")):
        assert rel_err(a, r) < 2e-3, name
    ops.bert_embed_backward(ids, tts, word, pos, typ, gamma, 1e-12, dout, dw, dp, dt, dg, db, dropmask=mask, drop_scale=1.0 / 0.9, accumulate=True)
    assert rel_err(dw, 2 * wr.grad) < 2e-3 and rel_err(dp, 2 * pr.grad) < 2e-3 and rel_err(dg, 2 * gr.grad) < 2e-3
