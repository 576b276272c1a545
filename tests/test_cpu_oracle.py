"""CPU suite: the oracle (oracle/port.py) against the committed goldens that the UNMODIFIED reference produced
(oracle/make_goldens.py), including the world-size-2 gloo path of the gather-with-gradient loss."""
import json
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err
from oracle import port


def _embeds(n, seed):
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(n, 512, generator=g)
    return e / e.norm(dim=1, keepdim=True)


def _loss_case(z, meta, rank, world, mvs, eps):
    B = meta["B"]
    full = [_embeds(world * B, s) for s in meta["seeds"]]
    loc = [t[rank * B:(rank + 1) * B].clone().requires_grad_(True) for t in full]
    scale = torch.tensor(meta["logit_scale"], requires_grad=True)
    kw = dict(image_embeddings=loc[0], text_embeddings=loc[1], labels=torch.arange(B), logit_scale=scale)
    if mvs:
        kw.update(text_embeddings2=loc[2], image_view_embeddings=loc[3])
    fn = port.mvs_loss if mvs else port.contrastive_loss
    loss = fn(**kw, is_train=True, label_smoothing=eps, i2i_weight=meta["i2i_weight"], t2t_weight=meta["t2t_weight"])
    loss.backward()
    tag = f"rank{rank}::{'mvs' if mvs else 'con'}_eps{eps}"
    assert abs(loss.item() - float(z[f"{tag}::loss"])) < 1e-5
    assert abs(scale.grad.item() - float(z[f"{tag}::dscale"])) < 1e-5
    for i, nm in enumerate(("img", "txt", "txt2", "img2")[: 4 if mvs else 2]):
        assert rel_err(loc[i].grad, torch.from_numpy(z[f"{tag}::d{nm}"])) < 1e-4, (tag, nm)


def test_loss_oracle_matches_reference_w1(golden_dir):
    z = np.load(os.path.join(golden_dir, "loss_w1_b8.npz"))
    meta = json.loads(str(z["meta"]))
    for mvs in (False, True):
        for eps in (0.0, 0.1):
            _loss_case(z, meta, 0, 1, mvs, eps)


def _w2_worker(rank, world, port_no, gdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.load(os.path.join(gdir, "loss_w2_b6.npz"))
        meta = json.loads(str(z["meta"]))
        for mvs in (False, True):
            for eps in (0.0, 0.1):
                _loss_case(z, meta, rank, world, mvs, eps)
    finally:
        dist.destroy_process_group()


def test_loss_oracle_matches_reference_w2_gloo(golden_dir):
    """all_gather forward / reduce_scatter backward (util/dist_autograd.py:4-26) on 2 gloo ranks."""
    mp.spawn(_w2_worker, args=(2, 29731, golden_dir), nprocs=2, join=True)


def test_comm_free_backward_identity(golden_dir):
    """The identity the CUDA kernel relies on: rank r's gradients under all_gather+reduce_scatter equal rows r*B..(r+1)*B
    of d(sum_r L_r)/dE computed on the concatenated batch, and the per-rank loss is the mean over its own rows."""
    z = np.load(os.path.join(golden_dir, "loss_w2_b6.npz"))
    meta = json.loads(str(z["meta"]))
    B, W, s = meta["B"], meta["world"], meta["logit_scale"]
    img, txt = (_embeds(W * B, sd).requires_grad_(True) for sd in meta["seeds"][:2])
    S = s * img @ txt.T
    lab = torch.arange(W * B)
    total = 0.0
    for r in range(W):
        rows = slice(r * B, (r + 1) * B)
        l_r = 0.75 * torch.nn.functional.cross_entropy(S[rows], lab[rows], label_smoothing=0.1) + \
            0.25 * torch.nn.functional.cross_entropy(S.T[rows], lab[rows], label_smoothing=0.1)
        assert abs(l_r.item() - float(z[f"rank{r}::con_eps0.1::loss"])) < 1e-5
        total = total + l_r
    total.backward()
    for r in range(W):
        assert rel_err(img.grad[r * B:(r + 1) * B], torch.from_numpy(z[f"rank{r}::con_eps0.1::dimg"])) < 1e-4
        assert rel_err(txt.grad[r * B:(r + 1) * B], torch.from_numpy(z[f"rank{r}::con_eps0.1::dtxt"])) < 1e-4


def _lse_exchange_worker(rank, world, port_no, golden_dir):
    """The data flow of csrc/loss.cu in exchange mode (ABI 5), restated with torch + gloo: a rank scores ONLY its rows x all columns
    and all rows x its columns, all-gathers the 2*B log-sum-exps it owns, and forms its gradients without any further exchange."""
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port_no}", rank=rank, world_size=world)
    z = np.load(os.path.join(golden_dir, "loss_w2_b6.npz"))
    meta = json.loads(str(z["meta"]))
    B, s, eps = meta["B"], meta["logit_scale"], 0.1
    img_all, txt_all = (_embeds(world * B, sd) for sd in meta["seeds"][:2])
    WB = world * B
    mine = slice(rank * B, (rank + 1) * B)
    # phase 0: the gather (here: everybody already holds both slabs); phase 1: the cross only
    S_rows = s * img_all[mine] @ txt_all.T                 # [B, WB]  local rows x all columns
    S_cols = s * img_all @ txt_all[mine].T                 # [WB, B]  all rows x local columns
    lse_row_mine, lse_col_mine = torch.logsumexp(S_rows, 1), torch.logsumexp(S_cols, 0)
    # phase 1b: exchange (2*B floats per rank)
    rows, cols = [torch.empty(B) for _ in range(world)], [torch.empty(B) for _ in range(world)]
    dist.all_gather(rows, lse_row_mine)
    dist.all_gather(cols, lse_col_mine)
    lse_row, lse_col = torch.cat(rows), torch.cat(cols)
    # phase 2: G for (local rows, all columns) and (all rows, local columns) from the exchanged LSEs
    w_row, w_col = 0.75, 0.25
    q = torch.full((WB, WB), eps / WB) + (1 - eps) * torch.eye(WB)

    def G(Sblk, rsel, csel):
        return (w_row * (torch.exp(Sblk - lse_row[rsel, None]) - q[rsel][:, csel]) + w_col * (torch.exp(Sblk - lse_col[None, csel]) - q[rsel][:, csel])) / B

    allr = slice(0, WB)
    dimg = s * G(S_rows, mine, allr) @ txt_all             # dE_a[i] = scale * sum_j G_ij E_b[j], local i
    dtxt = s * G(S_cols, allr, mine).T @ img_all           # dE_b[j] = scale * sum_i G_ij E_a[i], local j
    loss = (w_row * (lse_row_mine - (q[mine] * S_rows).sum(1)).mean() + w_col * (lse_col_mine - (q[:, mine] * S_cols).sum(0)).mean())
    assert abs(loss.item() - float(z[f"rank{rank}::con_eps0.1::loss"])) < 1e-5
    assert rel_err(dimg, torch.from_numpy(z[f"rank{rank}::con_eps0.1::dimg"])) < 1e-4
    assert rel_err(dtxt, torch.from_numpy(z[f"rank{rank}::con_eps0.1::dtxt"])) < 1e-4
    dist.destroy_process_group()


def test_lse_exchange_identity_world2(golden_dir):
    """World-size-2 gloo run of the loss kernel's LSE-exchange algorithm against the reference's own all_gather / reduce_scatter
    goldens (loss, dE_img, dE_txt per rank)."""
    import torch.multiprocessing as mp
    mp.spawn(_lse_exchange_worker, args=(2, 29741, golden_dir), nprocs=2, join=True)


@pytest.mark.parametrize("tag", ["enc_b2_96x64", "enc_b5_80x48"])
def test_encoder_oracle_matches_reference(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, tag + ".npz"))
    meta = json.loads(str(z["meta"]))
    m = port.OracleEfficientNet(meta["encoder"])
    m.stochastic = False
    port.fill_deterministic(m, meta["weight_seed"])
    x = port.synth_images(meta["batch"], meta["h"], meta["w"], seed=meta["image_seed"], identical_channels=False)
    g = torch.Generator().manual_seed(meta["probe_seed"])
    probe = torch.randn(meta["batch"], m.out_dim, generator=g)
    for mode in ("eval", "train"):
        m.train(mode == "train")
        m.zero_grad()
        f = m(x)
        assert rel_err(f, torch.from_numpy(z[f"{mode}_features"])) < 5e-4, mode
        (f * probe).sum().backward()
        grads = dict(m.named_parameters())
        gmax = float(z[f"{mode}_grad_norms"].max())
        for key in z.files:
            if key.startswith(f"{mode}_grad::"):
                k = key.split("::", 1)[1]
                ref = torch.from_numpy(z[key])
                d = (grads[k].grad.double() - ref.double()).abs().max().item()
                assert d < 5e-3 * max(ref.abs().max().item(), 1e-2 * gmax), (mode, k)
    sd = m.state_dict()
    for key in z.files:
        if key.startswith("after_train::"):
            assert rel_err(sd[key.split("::", 1)[1]], torch.from_numpy(z[key])) < 1e-4, key
    m.eval()
    pooled, raw = m({"image": x})
    assert list(raw.shape) == z["dict_raw_shape"].tolist()


def test_clip_c1_oracle_matches_reference(golden_dir):
    from transformers import BatchEncoding
    z = np.load(os.path.join(golden_dir, "clip_c1_contrastive.npz"))
    meta = json.loads(str(z["meta"]))
    torch.set_num_threads(8)
    m = port.OracleBreastClip(meta["encoder"], num_hidden_layers=meta["bert_layers"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m.image_encoder.stochastic = False
    port.fill_deterministic(m, meta["weight_seed"])
    m.train()
    batch = {"images": port.synth_images(meta["batch"], meta["h"], meta["w"], seed=meta["image_seed"]),
             "text_tokens": BatchEncoding(port.synth_tokens(meta["batch"], meta["L"], seed=meta["token_seed"]))}
    out = m(batch)
    loss = port.contrastive_loss(**out, is_train=True, label_smoothing=meta["label_smoothing"])
    assert rel_err(out["image_embeddings"], torch.from_numpy(z["image_embeddings"])) < 5e-4
    assert rel_err(out["text_embeddings"], torch.from_numpy(z["text_embeddings"])) < 5e-4
    assert abs(loss.item() - float(z["loss"])) < 1e-4
    loss.backward()
    named = dict(m.named_parameters())
    assert sorted(k for k, p in named.items() if p.grad is None) == sorted(z["unused"].tolist())
    assert abs(named["logit_scale"].grad.item() - float(z["logit_scale_grad"])) < 1e-3 * abs(float(z["logit_scale_grad"])) + 1e-6


def test_bf16_emulation_is_off_by_default_and_small_in_eval():
    m = port.OracleEfficientNet("efficientnet-b2")
    port.fill_deterministic(m, 0)
    m.eval()
    x = port.synth_images(1, 64, 64)
    with torch.no_grad():
        a = m(x)
        m.emulate_bf16 = True
        b = m(x)
    assert not torch.equal(a, b) and rel_err(b, a) < 2e-2


def _mlp_head_formula(head, x):
    """projection.py:13-20 written out (what the kernel path implements)."""
    p = head.projection(x)
    return head.layer_norm(head.dropout(head.fc(torch.nn.functional.gelu(p))) + p)


def test_mlp_head_formula_matches_the_reference_golden(golden_dir):
    """The fixture was produced by the reference's own MLPProjectionHead (oracle/make_goldens.py::golden_mlp_head)."""
    import numpy as np
    from oracle import port
    z = np.load(os.path.join(golden_dir, "mlp_head_768_512.npz"))
    head = torch.nn.Module()
    head.projection, head.gelu, head.fc = torch.nn.Linear(768, 512), torch.nn.GELU(), torch.nn.Linear(512, 512)
    head.dropout, head.layer_norm = torch.nn.Dropout(0.1), torch.nn.LayerNorm(512)
    port.fill_deterministic(head, 3)
    head.eval()
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    out = _mlp_head_formula(head, x)
    (out * torch.from_numpy(z["probe"])).sum().backward()
    assert torch.allclose(out, torch.from_numpy(z["out"]), atol=1e-5) and torch.allclose(x.grad, torch.from_numpy(z["dx"]), atol=1e-5)
    for k, v in head.named_parameters():
        g = v.grad[::8, ::8] if v.dim() == 2 else v.grad
        assert torch.allclose(g, torch.from_numpy(z["grad." + k]), atol=1e-4, rtol=1e-4), k
