"""CPU suite: host orchestration of the text tower's hand-written forward/backward (bert_kernels.py) with the C-ABI wrappers
replaced by PyTorch stand-ins that round to bf16 where the kernels do (tests/emulated_ops.py).  Checks which tensor feeds
which GEMM, the transposed-weight operands, dropout-mask routing and gradient destinations against the oracle / an fp32
autograd restatement.  The kernels themselves are checked on the GPU (tests/test_gpu_text.py)."""
import copy

import pytest
import torch

import emulated_ops
from bert_ref import bert_torch_ref
from conftest import rel_err


def _build(layers, dropout):
    from transformers import BertConfig
    from mammoclip_b200.model.modules.text_encoder import HuggingfaceTextEncoder
    from oracle import port
    cfg = BertConfig(**dict(port.BERT_BASE_CASED, num_hidden_layers=layers, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout))
    ours = HuggingfaceTextEncoder(config=cfg)
    ref = port.OracleTextEncoder(num_hidden_layers=layers, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout)
    port.fill_deterministic(ref, 0)
    ours.load_state_dict(ref.state_dict())
    return ours, ref


def _compare_grads(ours_named, ref_named):
    gr = dict(ref_named)
    gmax = max(p.grad.abs().max().item() for p in gr.values() if p.grad is not None)
    for k, p in ours_named:
        if gr[k].grad is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, k
            continue
        assert p.grad is not None, k
        d = (p.grad.double() - gr[k].grad.double()).abs().max().item()
        assert d < 5e-2 * max(gr[k].grad.abs().max().item(), 1e-2 * gmax), (k, d)


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_bert_backward_orchestration_matches_autograd(monkeypatch, dropout):
    from mammoclip_b200 import ops
    from mammoclip_b200.model.modules import bert_kernels as bk
    from oracle import port
    emulated_ops.install(monkeypatch, ops)
    layers, B, L = 3, 4, 40
    ours, ref = _build(layers, dropout)
    ours.train(dropout > 0)
    ref.eval()
    bert = ours.text_encoder
    tok = port.synth_tokens(B, L, seed=4321)
    masks = None
    if dropout > 0:
        torch.manual_seed(3)
        masks = bk._make_masks(bert, B, L, "cpu")
        monkeypatch.setattr(bk, "_make_masks", lambda *a, **k: masks)
    ho = bk.bert_forward(bert, tok["input_ids"], tok["token_type_ids"], tok["attention_mask"], dropout > 0, owner=ours)
    if dropout > 0:
        rb = copy.deepcopy(ref.text_encoder)
        hr = bert_torch_ref(rb, tok["input_ids"], tok["token_type_ids"], tok["attention_mask"], masks)
        ref_named = [("text_encoder." + k, p) for k, p in rb.named_parameters()]
    else:
        hr = ref(tok)
        ref_named = list(ref.named_parameters())
    valid = tok["attention_mask"].bool()
    assert rel_err(ho[valid], hr[valid]) < 2e-2
    g = torch.Generator().manual_seed(5)
    probe = torch.randn(B, L, 768, generator=g) * valid[..., None]
    (ho * probe).sum().backward()
    (hr * probe).sum().backward()
    _compare_grads(list(ours.named_parameters()), ref_named)


def test_bert_second_backward_accumulates_and_flat_optimizer_gets_direct_writes(monkeypatch):
    """MVS runs the text tower twice per step: the first backward after zero_grad() writes straight into the flat gradient
    buffer, the second one is accumulated by autograd; both must add up to 2x the single gradient."""
    from mammoclip_b200 import ops
    from mammoclip_b200.model.modules import bert_kernels as bk
    from oracle import port
    emulated_ops.install(monkeypatch, ops)
    ours, _ = _build(1, 0.0)
    ours.eval()
    bert = ours.text_encoder
    tok = port.synth_tokens(2, 16, seed=1)

    class FakeOpt:
        zero_count = 1

    for p in ours.parameters():
        p.grad = torch.zeros_like(p)
    object.__setattr__(ours, "_flat_optimizer", FakeOpt())
    grads_before = {k: p.grad for k, p in ours.named_parameters()}

    def run():
        return bk.bert_forward(bert, tok["input_ids"], tok["token_type_ids"], tok["attention_mask"], False, owner=ours).square().sum()

    run().backward()
    assert ours._direct_written_at == 1
    single = {k: p.grad.clone() for k, p in ours.named_parameters()}
    for k, p in ours.named_parameters():
        assert p.grad is grads_before[k], k                # written in place, not replaced
    run().backward()                                       # same zero_count: ordinary accumulation
    for k, p in ours.named_parameters():
        if "pooler" in k:
            continue
        assert torch.allclose(p.grad, 2 * single[k], rtol=1e-5, atol=1e-6 * single[k].abs().max().item() + 1e-12), k


@pytest.mark.parametrize("drop", [False, True])
def test_mlp_projection_head_orchestration(monkeypatch, drop):
    """MLP head (projection.py:4-20) forward/backward routing on the emulated kernels vs fp32 autograd of the reference formula."""
    from mammoclip_b200 import ops
    from mammoclip_b200.model.modules.projection import MLPProjectionHead, _MLPHeadFn
    emulated_ops.install(monkeypatch, ops)
    torch.manual_seed(0)
    head = MLPProjectionHead(768, 512, 0.1)
    ref = copy.deepcopy(head)
    x = torch.randn(6, 768)
    mask = (torch.rand(6, 512) >= 0.1).to(torch.uint8) if drop else None
    scale = 1.0 / 0.9 if drop else 1.0
    xa = x.clone().requires_grad_(True)
    out = _MLPHeadFn.apply(xa, head.projection.weight, head.projection.bias, head.fc.weight, head.fc.bias, head.layer_norm.weight,
                           head.layer_norm.bias, mask, scale, head.layer_norm.eps)
    xb = x.clone().requires_grad_(True)
    p = ref.projection(xb)
    y = ref.fc(torch.nn.functional.gelu(p))
    if drop:
        y = y * (mask.float() * scale)
    r = ref.layer_norm(y + p)
    assert rel_err(out, r) < 2e-2
    g = torch.randn(6, 512)
    (out * g).sum().backward()
    (r * g).sum().backward()
    assert rel_err(xa.grad, xb.grad) < 3e-2
    for (k, a), (_, b) in zip(head.named_parameters(), ref.named_parameters()):
        assert rel_err(a.grad, b.grad) < 3e-2, k
