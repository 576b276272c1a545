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


@pytest.mark.parametrize("layers,B,L", [(2, 4, 32), (12, 8, 64), (2, 3, 100)])
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


def test_bert_dropout_masks_are_shared_between_forward_and_backward():
    """Train mode (p=0.1): the kernel forward and the autograd recompute must use the same keep-masks."""
    from oracle import port
    from transformers import BatchEncoding
    from mammoclip_b200.model.modules import bert_kernels as bk
    ours, _ = _build(2, dropout=0.1)
    ours.train()
    tok = port.synth_tokens(4, 32, seed=1, device="cuda")
    bert = ours.text_encoder
    masks = bk._make_masks(bert, 4, 32, "cuda")
    a = bk._kernel_forward(bert, tok["input_ids"], tok["token_type_ids"], tok["attention_mask"], masks).float()
    with torch.no_grad():
        b = bk._bert_torch(bert, tok["input_ids"], tok["token_type_ids"], tok["attention_mask"], masks)
    valid = tok["attention_mask"].bool()
    assert rel_err(a[valid], b[valid]) < 2e-2
    masks2 = bk._make_masks(bert, 4, 32, "cuda")
    with torch.no_grad():
        c = bk._bert_torch(bert, tok["input_ids"], tok["token_type_ids"], tok["attention_mask"], masks2)
    assert rel_err(a[valid], c[valid]) > 5e-2     # different draws give a different function
