"""CPU suite: C-ABI surface, struct layouts, host-side logic (geometry, factories, state-dict compatibility, no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from mammoclip_b200 import _lib
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "mclip.h")).read()
    names = sorted(set(re.findall(r"\b(mclip_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mclip.h but not exported by libmclip_b200.so"
    assert lib.mclip_version() >= 4
    assert lib.mclip_loss_workspace_bytes(8, 64, 512, 1) > 0


def test_ctypes_structs_match_the_header(tmp_path):
    from mammoclip_b200 import _lib
    pairs = [("mclip_loss_args", _lib.LossArgs), ("mclip_gemm_args", _lib.GemmArgs), ("mclip_wgrad_args", _lib.WgradArgs),
             ("mclip_dwconv_args", _lib.DwconvArgs), ("mclip_stem_args", _lib.StemArgs), ("mclip_bn_args", _lib.BnArgs),
             ("mclip_ew_args", _lib.EwArgs), ("mclip_se_args", _lib.SeArgs), ("mclip_ew_bwd_args", _lib.EwBwdArgs),
             ("mclip_prep_entry", _lib.PrepEntry), ("mclip_bert_embed_args", _lib.BertEmbedArgs),
             ("mclip_bert_embed_bwd_args", _lib.BertEmbedBwdArgs)]
    src = '#include "include/mclip.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){' + \
        "".join(f'printf("%zu\\n", sizeof({c}));' for c, _ in pairs) + \
        'printf("%zu\\n", offsetof(mclip_loss_args, out)); printf("%zu\\n", offsetof(mclip_gemm_args, stats)); return 0;}'
    cfile, exe = tmp_path / "sz.c", tmp_path / "sz"
    cfile.write_text(src)
    subprocess.check_call(["gcc", f"-I{ROOT}", "-o", str(exe), str(cfile)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    for (cname, st), v in zip(pairs, vals):
        assert C.sizeof(st) == v, cname
    assert _lib.LossArgs.out.offset == vals[-2] and _lib.GemmArgs.stats.offset == vals[-1]


def test_geometry_matches_oracle_and_survey_table():
    from mammoclip_b200.model.modules.efficientnet_custom import net_geometry
    from oracle import port
    for name in ("efficientnet-b0", "efficientnet-b2", "efficientnet-b5", "efficientnet-b7"):
        g, s = net_geometry(name), port.effnet_spec(name)
        assert (g.stem_out, g.stem_pads, g.head_out, g.dropout) == (s.stem_out, s.stem_pad, s.head_out, s.dropout)
        for a, b in zip(g.blocks, s.blocks):
            assert (a.cin, a.cexp, a.cout, a.k, a.s, a.expand, a.cse, a.pads, a.skip) == (b.cin, b.cexp, b.cout, b.k, b.s, b.expand, b.cse, b.pad, b.skip)
    b5, b2 = net_geometry("efficientnet-b5"), net_geometry("efficientnet-b2")
    assert len(b5.blocks) == 39 and len(b2.blocks) == 23 and b5.stem_out == 48 and b5.head_out == 2048 and b2.head_out == 1408
    # pads frozen from the nominal resolution (SURVEY finding 1): NOT TF-SAME at the real input size
    assert b5.blocks[13].pads == (1, 1, 1, 1) and b5.blocks[13].s == 2
    assert b5.blocks[8].pads == (1, 2, 1, 2) and b5.blocks[3].pads == (0, 1, 0, 1) and b5.blocks[27].pads == (2, 2, 2, 2)
    assert b2.blocks[5].pads == (2, 2, 2, 2) and b2.blocks[8].pads == (1, 1, 1, 1) and b2.blocks[2].pads == (0, 1, 0, 1)
    assert [b.cse for b in b5.blocks[:4]] == [12, 6, 6, 6] and not b5.blocks[0].expand and not b5.blocks[2].expand
    assert sum(b.skip for b in b5.blocks) == 32


def test_state_dict_is_interchangeable_with_the_reference_layout():
    from mammoclip_b200.model.modules.efficientnet_custom import EfficientNet
    from oracle import port
    for name, n_params, n_keys in (("efficientnet-b2", 7700994, 506), ("efficientnet-b5", 28340784, 852)):
        ours, ref = EfficientNet.from_name(name), port.OracleEfficientNet(name)
        assert sum(p.numel() for p in ours.parameters()) == n_params and len(ours.state_dict()) == n_keys
        ours.load_state_dict(ref.state_dict(), strict=True)
        ref.load_state_dict(ours.state_dict(), strict=True)
        assert "_blocks.0._expand_conv.weight" not in ours.state_dict()     # expand_ratio 1: no expand conv at all (A2)


def test_factories_follow_the_reference_contract():
    from transformers import BertConfig
    from mammoclip_b200.loss import build_loss
    from mammoclip_b200.model import build_model
    from mammoclip_b200.model.modules import load_image_encoder, load_projection_head, load_text_encoder
    from mammoclip_b200.model.modules.text_encoder import BERT_BASE_CASED
    enc = load_image_encoder({"source": "cnn", "name": "tf_efficientnet_b5_ns-detect", "pretrained": True, "model_type": "cnn"})
    assert enc.out_dim == 2048
    assert load_image_encoder({"source": "cnn", "name": "tf_efficientnetv2-detect", "pretrained": True, "model_type": "cnn"}).out_dim == 1408
    with pytest.raises(KeyError):
        load_image_encoder({"source": "cnn", "name": "resnet152", "pretrained": True})
    with pytest.raises(KeyError):
        load_text_encoder({"source": "nowhere"}, 10)
    with pytest.raises(KeyError):
        load_projection_head(8, {"name": "conv"})
    assert load_projection_head(768, {"name": "linear", "proj_dim": 512}).projection.weight.shape == (512, 768)
    with pytest.raises(KeyError):
        build_loss({"triplet": {"loss_ratio": 1.0}})
    cl = build_loss({"breast_clip": {"label_smoothing": 0.1, "i2i_weight": 1.0, "t2t_weight": 0.5, "loss_ratio": 1.0},
                     "breast_clip_contrastive": {"label_smoothing": 0.0, "i2i_weight": 0.0, "t2t_weight": 0.0, "loss_ratio": 0.0}})
    assert len(cl.loss_list) == 1 and cl.loss_list[0].name == "contrastive" and cl.loss_list[0].loss_ratio == 1.0
    with pytest.raises(KeyError):
        build_model({"name": "unknown"}, {}, None)
    bcfg = BertConfig(**dict(BERT_BASE_CASED, num_hidden_layers=1))
    cfg = {"name": "clip_custom", "image_encoder": {"source": "cnn", "name": "tf_efficientnetv2-detect", "pretrained": True, "model_type": "cnn"},
           "text_encoder": {"source": "huggingface", "name": "x", "pretrained": False, "gradient_checkpointing": False, "pooling": "eos",
                            "cache_dir": "/tmp/none", "trust_remote_code": False, "config": bcfg},
           "projection_head": {"name": "linear", "proj_dim": 512, "dropout": 0.1}, "temperature": 0.07}

    class Tok:
        vocab_size = 28996

    m = build_model(cfg, {}, Tok())
    keys = set(m.state_dict())
    assert {"logit_scale", "image_encoder._conv_stem.weight", "image_projection.projection.weight", "text_projection.projection.bias",
            "text_encoder.text_encoder.embeddings.word_embeddings.weight", "text_encoder.text_encoder.pooler.dense.weight"} <= keys
    assert abs(m.logit_scale.item() - 2.6593) < 1e-3 and m.projection is True
    for attr in ("encode_image", "encode_text", "encode_image_normalized", "image_projection", "text_projection", "tokenizer"):
        assert hasattr(m, attr)


def test_no_cpu_fallback():
    """The product path must fail loudly without a B200: CPU tensors are rejected, nothing routes through the oracle."""
    from mammoclip_b200 import ops
    from mammoclip_b200._lib import MclipError
    from mammoclip_b200.model.modules.efficientnet_custom import EfficientNet
    with pytest.raises(MclipError):
        ops.gemm_tn(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(MclipError):
        ops.contrastive_loss_raw([torch.zeros(4, 512), torch.zeros(4, 512)], [(0, 1, .75, .25, 0.)], 1.0)
    with pytest.raises(RuntimeError):
        EfficientNet.from_name("efficientnet-b2")(torch.zeros(1, 3, 32, 32))
    import mammoclip_b200
    pkg = os.path.join(ROOT, "mammo-clip_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} must not import the oracle"


def test_argument_validation_without_a_gpu():
    """Host-side argument checks of the C ABI return error codes (no kernel is launched)."""
    from mammoclip_b200 import _lib
    lib = _lib.lib()
    g = _lib.GemmArgs()
    assert lib.mclip_gemm_tn(C.byref(g), None) != 0 and b"null" in lib.mclip_last_error()
    a = _lib.LossArgs()
    a.world, a.rank, a.batch, a.dim, a.n_tensors, a.n_pairs = 1, 0, 4, 100, 2, 1
    assert lib.mclip_contrastive_loss(C.byref(a), None) != 0 and b"multiple of 16" in lib.mclip_last_error()
    d = _lib.DwconvArgs()
    d.k, d.stride, d.c = 7, 1, 8
    d.in_, d.weight = 1, 1
    assert lib.mclip_dwconv_forward(C.byref(d), None) != 0 and b"unsupported" in lib.mclip_last_error()


def test_global_env_contract():
    from mammoclip_b200.util import GlobalEnv
    env = GlobalEnv.reset()
    assert env.world_size == 1 and env.world_rank == 0 and env.master and env.summary_writer.global_step == 0
    assert env.summary_writer.train is None


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the agreed keys; c1 keeps it short."""
    import json
    import sys
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                                  text=True, timeout=600)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_multi_view_memory_plan_choice():
    """Host logic of the multi-view memory plan at the metric scale (EN-B5, B = 64, 1520x912, two views) on a 180 GB device."""
    from mammoclip_b200.model.modules.efficientnet_custom import EfficientNet, choose_view_plan
    enc = EfficientNet.from_name("efficientnet-b5", num_classes=1)
    keep, lean = enc.saved_bytes(64, 1520, 912), enc.saved_bytes(64, 1520, 912, keep_y0=False)
    assert 92e9 < keep < 94e9 and 54e9 < lean < 56e9            # 92.9 GB per view, 55.0 GB without the expand outputs
    free = 175 << 30
    assert choose_view_plan(keep, lean, free) == "keep"                      # one view
    assert choose_view_plan(2 * keep, 2 * lean, free) == "lean"              # the shipped YAML's two views
    assert choose_view_plan(4 * keep, 4 * lean, free) == "recompute"
    assert choose_view_plan(0, 0, free) == "keep"                            # eval / no_grad: nothing is saved
