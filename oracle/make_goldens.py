"""Pin oracle/port.py against the UNMODIFIED reference and write tests/golden/*.npz.

Run in the build container only (needs /root/reference):  python -m oracle.make_goldens
Every fixture is produced BY THE REFERENCE CODE (breastclip.*), after asserting that the port
reproduces it on the same seeded inputs; fixtures are small (embeddings, losses, gradient
summaries), weights are regenerated from (name, seed) by port.fill_deterministic.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import port, ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
ENC_NAMES = {"efficientnet-b2": "tf_efficientnetv2-detect", "efficientnet-b5": "tf_efficientnet_b5_ns-detect"}


def _versions():
    import transformers
    return {"torch": torch.__version__, "transformers": transformers.__version__, "numpy": np.__version__}


def _close(a, b, tol, what, floor=1e-30):
    """max|a-b| <= tol * max(|b|max, floor).  `floor` guards gradients that are mathematically zero
    (e.g. a bias feeding a train-mode BN) and therefore pure rounding noise in both implementations."""
    a, b = a.detach().double(), b.detach().double()
    err = (a - b).abs().max().item()
    ref = max(b.abs().max().item(), floor)
    assert err <= tol * ref, f"{what}: port != reference, max err {err:.3e} vs scale {ref:.3e}"
    return err / ref


# ----------------------------------------------------------------------------- image encoder

def _ref_encoder(bc, name):
    from breastclip.model.modules import load_image_encoder
    enc = load_image_encoder({"source": "cnn", "name": ENC_NAMES[name], "pretrained": True, "model_type": "cnn"})
    # parity runs disable the stochastic ops (SURVEY finding 5)
    enc._global_params = enc._global_params._replace(drop_connect_rate=0.0)
    enc._dropout.p = 0.0
    return enc


PROBE_KEYS = ["_conv_stem.weight", "_bn0.weight", "_bn0.bias", "_blocks.0._depthwise_conv.weight",
              "_blocks.0._se_reduce.weight", "_blocks.0._se_reduce.bias", "_blocks.0._se_expand.weight",
              "_blocks.0._project_conv.weight", "_blocks.2._expand_conv.weight", "_blocks.2._bn0.weight",
              "_blocks.2._bn1.bias", "_blocks.2._bn2.weight", "_blocks.3._expand_conv.weight", "_blocks.3._bn0.bias",
              "_blocks.5._depthwise_conv.weight",
              "_bn1.weight", "_bn1.bias"]


def golden_encoder(bc, name, batch, h, w, tag):
    ref = _ref_encoder(bc, name)
    mine = port.OracleEfficientNet(name)
    mine.stochastic = False
    port.fill_deterministic(ref, 0)
    port.fill_deterministic(mine, 0)
    x = port.synth_images(batch, h, w, seed=1234, identical_channels=False)
    g = torch.Generator().manual_seed(99)
    probe = torch.randn(batch, ref.out_dim, generator=g)
    out = {}
    for mode in ("eval", "train"):
        ref.train(mode == "train"), mine.train(mode == "train")
        ref.zero_grad(), mine.zero_grad()
        fr, fm = ref(x), mine(x)
        _close(fm, fr, 2e-4, f"{tag}/{mode}/features")
        (fr * probe).sum().backward()
        (fm * probe).sum().backward()
        gr = dict(ref.named_parameters())
        gm = dict(mine.named_parameters())
        norms = []
        gmax = max(p.grad.abs().max().item() for p in gr.values())
        for k in gr:
            _close(gm[k].grad, gr[k].grad, 5e-3, f"{tag}/{mode}/grad {k}", floor=1e-2 * gmax)
            norms.append(gr[k].grad.double().norm().item())
        out[f"{mode}_features"] = fr.detach().numpy()
        out[f"{mode}_grad_norms"] = np.asarray(norms, dtype=np.float64)
        for k in PROBE_KEYS:
            if k in gr:
                out[f"{mode}_grad::{k}"] = gr[k].grad.detach().numpy()
    # running statistics after the single train-mode forward above
    sd_r, sd_m = ref.state_dict(), mine.state_dict()
    for k in ("_bn0.running_mean", "_bn0.running_var", "_blocks.2._bn1.running_mean", "_blocks.2._bn1.running_var",
              "_bn1.running_var"):
        _close(sd_m[k], sd_r[k], 1e-4, f"{tag}/{k}")
        out[f"after_train::{k}"] = sd_r[k].numpy()
    out["param_names"] = np.asarray(list(gr.keys()))
    # dict-input call form (efficientnet_custom.py:298-305)
    ref.eval()
    pooled, raw = ref({"image": x})
    out["dict_raw_shape"] = np.asarray(raw.shape)
    meta = dict(kind="encoder", encoder=name, batch=batch, h=h, w=w, image_seed=1234, weight_seed=0, probe_seed=99,
                identical_channels=False, **_versions())
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), meta=json.dumps(meta), **out)
    print("wrote", tag)


# ----------------------------------------------------------------------------- full CLIP step (c1)

def _bert_dir(layers, dropout):
    d = tempfile.mkdtemp(prefix="mclip_bert_")
    cfg = dict(port.BERT_BASE_CASED, num_hidden_layers=layers, hidden_dropout_prob=dropout,
               attention_probs_dropout_prob=dropout, model_type="bert", architectures=["BertModel"])
    json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
    return d


class _Tok:
    vocab_size = 28996


def golden_clip(bc, tag, enc="efficientnet-b2", layers=2, batch=4, h=224, w=224, L=32, mvs=False):
    from breastclip.model import build_model
    from breastclip.loss import build_loss
    cfg = {"name": "clip_custom",
           "image_encoder": {"source": "cnn", "name": ENC_NAMES[enc], "pretrained": True, "model_type": "cnn"},
           "text_encoder": {"source": "huggingface", "name": _bert_dir(layers, 0.0), "pretrained": False,
                            "gradient_checkpointing": False, "pooling": "eos", "cache_dir": "/tmp/none",
                            "trust_remote_code": False},
           "projection_head": {"name": "linear", "proj_dim": 512, "dropout": 0.1},
           "temperature": 0.07}
    loss_key = "breast_clip" if mvs else "breast_clip_contrastive"
    loss_cfg = {loss_key: {"label_smoothing": 0.1, "i2i_weight": 1.0, "t2t_weight": 0.5, "loss_ratio": 1.0}}
    ref = build_model(cfg, loss_cfg, _Tok())
    ref.image_encoder._global_params = ref.image_encoder._global_params._replace(drop_connect_rate=0.0)
    ref.image_encoder._dropout.p = 0.0
    ref_loss = build_loss(loss_cfg)
    mine = port.OracleBreastClip(enc, num_hidden_layers=layers, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    mine.image_encoder.stochastic = False
    port.fill_deterministic(ref, 0), port.fill_deterministic(mine, 0)
    ref.logit_scale.data.fill_(float(np.log(1 / 0.07))), mine.logit_scale.data.fill_(float(np.log(1 / 0.07)))
    from transformers import BatchEncoding
    batch_d = {"images": port.synth_images(batch, h, w, seed=1234, identical_channels=True),
               "text_tokens": BatchEncoding(port.synth_tokens(batch, L, seed=4321))}
    if mvs:
        batch_d["image_views"] = port.synth_images(batch, h, w, seed=1235, identical_channels=True)
        batch_d["text_tokens2"] = BatchEncoding(port.synth_tokens(batch, L, seed=4322))
    ref.train(), mine.train()
    o_r = ref(batch_d, "cpu")
    o_m = mine(batch_d)
    l_r = ref_loss(**o_r, is_train=True)["total"]
    fn = port.mvs_loss if mvs else port.contrastive_loss
    l_m = fn(**o_m, is_train=True, label_smoothing=0.1, i2i_weight=1.0, t2t_weight=0.5)
    out = {}
    for k in ("image_embeddings", "text_embeddings", "text_embeddings2", "image_view_embeddings"):
        if k in o_r:
            _close(o_m[k], o_r[k], 2e-4, f"{tag}/{k}")
            out[k] = o_r[k].detach().numpy()
    _close(l_m, l_r, 1e-4, f"{tag}/loss")
    l_r.backward(), l_m.backward()
    pr, pm = dict(ref.named_parameters()), dict(mine.named_parameters())
    names, norms, unused = [], [], []
    gmax = max(p.grad.abs().max().item() for p in pr.values() if p.grad is not None)
    for k, p in pr.items():
        if p.grad is None:
            unused.append(k)
            assert pm[k].grad is None, k
            continue
        _close(pm[k].grad, p.grad, 1e-2, f"{tag}/grad {k}", floor=1e-2 * gmax)
        names.append(k), norms.append(p.grad.double().norm().item())
    out["loss"] = np.asarray(l_r.item())
    out["logit_scale_grad"] = pr["logit_scale"].grad.numpy()
    out["grad_names"], out["grad_norms"], out["unused"] = np.asarray(names), np.asarray(norms), np.asarray(unused)
    for k in ("image_projection.projection.weight", "text_projection.projection.bias",
              "image_encoder._conv_stem.weight", "image_encoder._bn1.weight",
              "text_encoder.text_encoder.encoder.layer.0.attention.self.query.bias",
              "text_encoder.text_encoder.embeddings.LayerNorm.weight"):
        out[f"grad::{k}"] = pr[k].grad.numpy()[:8]      # leading rows only: fixtures stay small
    meta = dict(kind="clip", encoder=enc, bert_layers=layers, batch=batch, h=h, w=w, L=L, mvs=mvs, label_smoothing=0.1,
                i2i_weight=1.0, t2t_weight=0.5, image_seed=1234, token_seed=4321, weight_seed=0, **_versions())
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), meta=json.dumps(meta), **out)
    print("wrote", tag, "loss", l_r.item())


# ----------------------------------------------------------------------------- loss, W=1 and W=2 (gloo)

def _embeds(n, seed):
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(n, 512, generator=g)
    return e / e.norm(dim=1, keepdim=True)


def _loss_worker(rank, world, port_no, B, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    bc = ref_loader.load_reference()
    ref_loader.reset_global_env()
    from breastclip.loss import build_loss
    res = {}
    for mvs in (False, True):
        for eps in (0.0, 0.1):
            key = "breast_clip" if mvs else "breast_clip_contrastive"
            lf = build_loss({key: {"label_smoothing": eps, "i2i_weight": 1.0, "t2t_weight": 0.5, "loss_ratio": 1.0}})
            full = [_embeds(world * B, s) for s in (11, 12, 13, 14)]
            loc = [t[rank * B:(rank + 1) * B].clone().requires_grad_(True) for t in full]
            loc2 = [t.detach().clone().requires_grad_(True) for t in loc]
            scale = torch.tensor(14.2857, requires_grad=True)
            scale2 = torch.tensor(14.2857, requires_grad=True)
            kw = dict(image_embeddings=loc[0], text_embeddings=loc[1], labels=torch.arange(B), logit_scale=scale)
            kw2 = dict(image_embeddings=loc2[0], text_embeddings=loc2[1], labels=torch.arange(B), logit_scale=scale2)
            if mvs:
                kw.update(text_embeddings2=loc[2], image_view_embeddings=loc[3])
                kw2.update(text_embeddings2=loc2[2], image_view_embeddings=loc2[3])
            l_r = lf(**kw, is_train=True)["total"]
            fn = port.mvs_loss if mvs else port.contrastive_loss
            l_m = fn(**kw2, is_train=True, label_smoothing=eps, i2i_weight=1.0, t2t_weight=0.5)
            l_r.backward(), l_m.backward()
            _close(l_m, l_r, 1e-6, "loss")
            tag = f"{'mvs' if mvs else 'con'}_eps{eps}"
            res[f"{tag}::loss"] = l_r.detach().numpy()
            res[f"{tag}::dscale"] = scale.grad.numpy()
            _close(scale2.grad, scale.grad, 1e-5, "dscale")
            for i, nm in enumerate(("img", "txt", "txt2", "img2")[: 4 if mvs else 2]):
                _close(loc2[i].grad, loc[i].grad, 1e-5, f"d{nm}")
                res[f"{tag}::d{nm}"] = loc[i].grad.numpy()
    ret[rank] = res
    if world > 1:
        dist.destroy_process_group()


def golden_loss(world, B, tag, port_no):
    mgr = mp.Manager()
    ret = mgr.dict()
    if world == 1:
        _loss_worker(0, 1, port_no, B, ret)
    else:
        mp.spawn(_loss_worker, args=(world, port_no, B, ret), nprocs=world, join=True)
    out = {}
    for r in range(world):
        for k, v in ret[r].items():
            out[f"rank{r}::{k}"] = v
    meta = dict(kind="loss", world=world, B=B, D=512, seeds=[11, 12, 13, 14], logit_scale=14.2857,
                i2i_weight=1.0, t2t_weight=0.5, **_versions())
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), meta=json.dumps(meta), **out)
    print("wrote", tag)


# ----------------------------------------------------------------------------- MLP projection head

def golden_mlp_head(bc, tag="mlp_head_768_512"):
    """projection.py:4-20 of the reference, eval mode (dropout off), weights from (name, seed): output + every gradient."""
    from breastclip.model.modules import load_projection_head
    head = load_projection_head(768, {"name": "mlp", "proj_dim": 512, "dropout": 0.1})
    port.fill_deterministic(head, 3)
    head.eval()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(9, 768, generator=g).requires_grad_(True)
    probe = torch.randn(9, 512, generator=g)
    out = head(x)
    (out * probe).sum().backward()
    arrays = {"x": x.detach().numpy(), "probe": probe.numpy(), "out": out.detach().numpy(), "dx": x.grad.numpy()}
    for k, v in head.named_parameters():       # weight gradients: an 8x8-strided sample keeps the fixture small
        arrays["grad." + k] = (v.grad[::8, ::8] if v.dim() == 2 else v.grad).numpy().copy()
    meta = {"versions": _versions(), "seed": 3, "embedding_dim": 768, "proj_dim": 512}
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), meta=json.dumps(meta), **arrays)
    print("wrote", tag)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    bc = ref_loader.load_reference()
    golden_loss(1, 8, "loss_w1_b8", 29611)
    golden_loss(2, 6, "loss_w2_b6", 29612)
    golden_encoder(bc, "efficientnet-b2", 2, 96, 64, "enc_b2_96x64")
    golden_encoder(bc, "efficientnet-b5", 2, 80, 48, "enc_b5_80x48")
    golden_clip(bc, "clip_c1_contrastive", mvs=False)
    golden_clip(bc, "clip_c1_mvs", mvs=True)
    golden_mlp_head(bc)


if __name__ == "__main__":
    main()
