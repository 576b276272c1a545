"""Whole contrastive step through the reference-shaped public API (build_model / build_loss) vs the oracle and the
reference-generated c1 goldens (EN-B2 + 2-layer BERT, B=4, 224x224, L=32)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


class _Tok:
    vocab_size = 28996


def _cfgs(layers, mvs, dropout=0.0):
    from transformers import BertConfig
    from mammoclip_b200.model.modules.text_encoder import BERT_BASE_CASED
    bcfg = BertConfig(**dict(BERT_BASE_CASED, num_hidden_layers=layers, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout))
    cfg = {"name": "clip_custom",
           "image_encoder": {"source": "cnn", "name": "tf_efficientnetv2-detect", "pretrained": True, "model_type": "cnn"},
           "text_encoder": {"source": "huggingface", "name": "offline", "pretrained": False, "gradient_checkpointing": False, "pooling": "eos",
                            "cache_dir": "/tmp/none", "trust_remote_code": False, "config": bcfg},
           "projection_head": {"name": "linear", "proj_dim": 512, "dropout": 0.1}, "temperature": 0.07}
    key = "breast_clip" if mvs else "breast_clip_contrastive"
    return cfg, {key: {"label_smoothing": 0.1, "i2i_weight": 1.0, "t2t_weight": 0.5, "loss_ratio": 1.0}}


def _batch(B, h, w, L, mvs):
    from oracle import port
    from transformers import BatchEncoding
    b = {"images": port.synth_images(B, h, w, seed=1234, device="cuda"), "text_tokens": BatchEncoding(port.synth_tokens(B, L, seed=4321, device="cuda"))}
    if mvs:
        b["image_views"] = port.synth_images(B, h, w, seed=1235, device="cuda")
        b["text_tokens2"] = BatchEncoding(port.synth_tokens(B, L, seed=4322, device="cuda"))
    return b


@pytest.mark.parametrize("mvs", [False, True])
def test_clip_step_eval_vs_oracle(mvs):
    """eval-mode BatchNorm (running statistics): tight parity of embeddings, loss and every gradient."""
    from mammoclip_b200.loss import build_loss
    from mammoclip_b200.model import build_model
    from oracle import port
    cfg, lcfg = _cfgs(2, mvs)
    ours = build_model(cfg, lcfg, _Tok())
    ref = port.OracleBreastClip("efficientnet-b2", num_hidden_layers=2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    port.fill_deterministic(ref, 0)
    ours.load_state_dict(ref.state_dict())
    ours.cuda().eval(), ref.cuda().eval()
    batch = _batch(4, 96, 64, 32, mvs)
    out = ours(batch, "cuda")
    loss = build_loss(lcfg)(**out, is_train=True)["total"]
    loss.backward()
    o_ref = ref(batch)
    fn = port.mvs_loss if mvs else port.contrastive_loss
    l_ref = fn(**o_ref, is_train=True, label_smoothing=0.1, i2i_weight=1.0, t2t_weight=0.5)
    l_ref.backward()
    for k in o_ref:
        if k.endswith("embeddings") or k.endswith("embeddings2"):
            assert rel_err(out[k], o_ref[k]) < 2e-2, k
    assert abs(loss.item() - l_ref.item()) < 2e-2 * abs(l_ref.item())
    gr = dict(ref.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in gr.values() if p.grad is not None)
    for k, p in ours.named_parameters():
        if gr[k].grad is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, k      # BERT pooler: unused in both
            continue
        assert p.grad is not None, k
        d = (p.grad.double() - gr[k].grad.double()).abs().max().item()
        assert d < 6e-2 * max(gr[k].grad.abs().max().item(), 2e-2 * gmax), (k, d)


@pytest.mark.parametrize("tag,mvs", [("clip_c1_contrastive", False), ("clip_c1_mvs", True)])
def test_clip_c1_reference_golden(golden_dir, tag, mvs):
    """BASELINE config 1 fixture produced by the reference's own build_model/build_loss in TRAIN mode (fp32, CPU).
    Batch-statistics BN on 4 images in bf16: checked at the level PyTorch's own bf16 autocast reaches (see
    test_gpu_encoder._compare), plus structural facts (unused pooler, finite grads, loss within 10%)."""
    from mammoclip_b200.loss import build_loss
    from mammoclip_b200.model import build_model
    from oracle import port
    z = np.load(os.path.join(golden_dir, tag + ".npz"))
    meta = json.loads(str(z["meta"]))
    cfg, lcfg = _cfgs(meta["bert_layers"], mvs)
    ours = build_model(cfg, lcfg, _Tok())
    ref = port.OracleBreastClip("efficientnet-b2", num_hidden_layers=meta["bert_layers"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    port.fill_deterministic(ref, 0)
    ours.load_state_dict(ref.state_dict())
    ours.cuda().train()
    ours.image_encoder.stochastic = False
    batch = _batch(meta["batch"], meta["h"], meta["w"], meta["L"], mvs)
    out = ours(batch, "cuda")
    loss = build_loss(lcfg)(**out, is_train=True)["total"]
    loss.backward()
    # text tower has no BatchNorm: tight
    assert rel_err(out["text_embeddings"], torch.from_numpy(z["text_embeddings"])) < 2e-2
    # image tower: bf16 + batch statistics on 4 images
    ref.cuda().train()
    ref.image_encoder.stochastic = False
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        e_amp = rel_err(ref._embed_img(batch["images"]).float(), torch.from_numpy(z["image_embeddings"]))
    e_ours = rel_err(out["image_embeddings"], torch.from_numpy(z["image_embeddings"]))
    assert e_ours < max(2e-2, 1.5 * e_amp), (e_ours, e_amp)
    assert abs(loss.item() - float(z["loss"])) < 0.1 * abs(float(z["loss"]))
    unused = set(z["unused"].tolist())
    for k, p in ours.named_parameters():
        if k in unused:
            assert p.grad is None or p.grad.abs().max().item() == 0, k
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_stochastic_train_step_runs_and_is_finite():
    """drop-connect + dropout + BERT dropout on (the throughput configuration): finite loss/grads, stats updated."""
    from mammoclip_b200.loss import build_loss
    from mammoclip_b200.model import build_model
    cfg, lcfg = _cfgs(2, False, dropout=0.1)
    torch.manual_seed(0)
    ours = build_model(cfg, lcfg, _Tok()).cuda().train()
    batch = _batch(8, 64, 96, 32, False)
    out = ours(batch, "cuda")
    loss = build_loss(lcfg)(**out, is_train=True)["total"]
    loss.backward()
    assert torch.isfinite(loss)
    for k, p in ours.named_parameters():
        if "pooler" not in k:
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
    assert ours.image_encoder._bn0.num_batches_tracked.item() == 1


def test_flat_adamw_matches_torch_adamw_and_direct_grads():
    """FlatAdamW (one kernel over a flat buffer; the image tower writes its gradients straight into it) must give the
    same gradients as plain autograd accumulation and the same update as torch.optim.AdamW (optimizer/__init__.py:23-31)."""
    import copy
    from mammoclip_b200.loss import build_loss
    from mammoclip_b200.model import build_model
    from mammoclip_b200.optim import FlatAdamW
    cfg, lcfg = _cfgs(1, False)
    torch.manual_seed(0)
    a = build_model(cfg, lcfg, _Tok()).cuda().eval()          # eval-mode BN: deterministic comparison
    b = copy.deepcopy(a)
    loss_fn = build_loss(lcfg)
    batch = _batch(4, 64, 64, 16, False)
    opt_a = FlatAdamW(a.parameters(), lr=1e-3, weight_decay=1e-2).attach(a)
    opt_b = torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=1e-2)
    for step in range(2):
        opt_a.zero_grad()
        opt_b.zero_grad(set_to_none=True)
        loss_fn(**a(batch, "cuda"), is_train=True)["total"].backward()
        loss_fn(**b(batch, "cuda"), is_train=True)["total"].backward()
        gmax = max(p.grad.abs().max().item() for p in b.parameters() if p.grad is not None)
        for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            if pb.grad is None:
                assert pa.grad.abs().max().item() == 0, k
            else:
                # step 0: identical weights -> identical gradients; later steps: bf16 rounding flips after ~1e-6 weight drift.
                # structurally-zero gradients (e.g. attention key bias) are pure rounding noise: absolute floor.
                tol = 1e-5 if step == 0 else 2e-2
                d = (pa.grad - pb.grad).abs().max().item()
                assert d < tol * max(pb.grad.abs().max().item(), 1e-3 * gmax), (step, k, d)
        opt_a.step()
        opt_b.step()
        for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            if pb.grad is None:
                continue
            if step == 0:      # identical gradients -> the two AdamW implementations must agree
                assert (pa - pb).abs().max().item() < 2e-6 + 1e-5 * pb.abs().max().item(), (step, k)
            else:              # Adam normalises by sqrt(v): noise-level gradients may flip sign, |delta| <= 2*lr per step
                assert (pa - pb).abs().max().item() < 2.5e-3, (step, k)


@pytest.mark.parametrize("training", [False, True])
def test_mlp_projection_head_on_kernels(training, monkeypatch):
    """`load_projection_head({"name": "mlp"})` (projection.py:4-20): Linear -> GELU -> Linear -> dropout -> +residual -> LayerNorm
    forward and backward on the kernels vs fp32 PyTorch with the same keep-mask."""
    import copy
    from mammoclip_b200.model.modules import load_projection_head
    torch.manual_seed(0)
    head = load_projection_head(768, {"name": "mlp", "proj_dim": 512, "dropout": 0.1}).cuda()
    ref = copy.deepcopy(head)
    head.train(training)
    x = torch.randn(37, 768, device="cuda")
    keep = (torch.rand(37, 512, device="cuda") >= 0.1)
    if training:
        monkeypatch.setattr(torch, "rand", lambda *a, **k: keep.float())         # mask = (rand >= p) reproduces `keep`
    xa = x.clone().requires_grad_(True)
    out = head(xa)
    monkeypatch.undo()
    xb = x.clone().requires_grad_(True)
    p = ref.projection(xb)
    y = ref.fc(torch.nn.functional.gelu(p))
    if training:
        y = y * (keep.float() / 0.9)
    r = ref.layer_norm(y + p)
    assert rel_err(out, r) < 2e-2
    g = torch.randn(37, 512, device="cuda")
    (out * g).sum().backward()
    (r * g).sum().backward()
    assert rel_err(xa.grad, xb.grad) < 3e-2
    for (k, a), (_, b) in zip(head.named_parameters(), ref.named_parameters()):
        assert rel_err(a.grad, b.grad) < 3e-2, k


def test_mlp_projection_head_vs_reference_golden(golden_dir):
    """Kernel MLP head vs the fixture produced by the reference's own MLPProjectionHead (eval mode; oracle/make_goldens.py)."""
    from mammoclip_b200.model.modules import load_projection_head
    from oracle import port
    z = np.load(os.path.join(golden_dir, "mlp_head_768_512.npz"))
    head = load_projection_head(768, {"name": "mlp", "proj_dim": 512, "dropout": 0.1})
    port.fill_deterministic(head, 3)
    head.cuda().eval()
    x = torch.from_numpy(z["x"]).cuda().requires_grad_(True)
    out = head(x)
    (out * torch.from_numpy(z["probe"]).cuda()).sum().backward()
    assert rel_err(out, torch.from_numpy(z["out"])) < 2e-2 and rel_err(x.grad, torch.from_numpy(z["dx"])) < 3e-2
    for k, v in head.named_parameters():
        g = v.grad[::8, ::8] if v.dim() == 2 else v.grad
        assert rel_err(g, torch.from_numpy(z["grad." + k])) < 3e-2, k
