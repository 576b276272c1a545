"""Parity at the scales the metric runs at (VERDICT r01 "weak" items 1-5).

* full-depth EN-B2 / EN-B5 towers, train AND eval mode, forward + every gradient + running statistics at B = 8, >= 256 px,
  against the fp32 oracle ON THE GPU, criterion SURVEY A13:
      err(ours_bf16, oracle_fp32) <= max(1e-2, err(oracle_bf16_autocast, oracle_fp32))
* the c3 GEOMETRY: EN-B5 at 1520x912, B = 4, eval and train mode, features + gradients vs the fp32 oracle
* EVERY MBConv block of EN-B5 at c3 geometry (and of EN-B2 at 456x456) in TRAIN mode, teacher-forced: fed the oracle's block
  input / output gradient and compared with fp32 autograd of that block (output 1e-2, input gradient 2e-2, parameter
  gradients 3e-2) -- the tight train-mode pin; through the full depth batch-statistics BN makes ANY two bf16-storage
  implementations (ours, PyTorch autocast, the fp32 oracle with bf16 rounding at the storage points) differ by 20-50 % in
  the gradients, independent of batch size (measured: scripts/diag_parity.py, profiles/r02_parity_diag.txt)
* layer-level kernels on tensors of MORE THAN 2^31 elements (the metric run has 3.2 G-element activations):
  depthwise k3 s2 at 760x456x144, the streaming passes and the 22 M-row small-K GEMM, checked chunk-wise vs fp32 torch
The oracle is true fp32 here: TF32 is disabled for cuDNN/cuBLAS by tests/conftest.py.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

REPORT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _report(name, obj):
    try:
        os.makedirs(REPORT_DIR, exist_ok=True)
        with open(os.path.join(REPORT_DIR, f"parity_{name}.json"), "w") as f:
            json.dump(obj, f, indent=1)
    except OSError:
        pass


def _tensor_errs(got, want, floor):
    """(max-norm, L2) error of one tensor relative to max(|want|_max, floor) / max(|want|_2, floor*sqrt(n))."""
    g, w = got.double(), want.double()
    d = g - w
    emax = d.abs().max().item() / max(w.abs().max().item(), floor)
    el2 = d.norm().item() / max(w.norm().item(), floor * w.numel() ** 0.5)
    return emax, el2


def _global_l2(g, ref, skip):
    num = sum(((g[k].double() - ref[k].double()) ** 2).sum().item() for k in ref if k not in skip)
    den = sum((ref[k].double() ** 2).sum().item() for k in ref if k not in skip)
    return (num / den) ** 0.5


def _run_tower(name, batch, h, w, mode, seed=1234):
    """ours (bf16 kernels) vs oracle fp32 vs oracle under bf16 autocast: features, gradients, running statistics."""
    from test_gpu_encoder import _build, _structurally_zero
    from oracle import port
    ours, ref = _build(name)
    train = mode == "train"
    ours.train(train), ref.train(train)
    x = port.synth_images(batch, h, w, seed=seed, identical_channels=False, device="cuda")
    g = torch.Generator().manual_seed(99)
    probe = torch.randn(batch, ref.out_dim, generator=g).cuda()
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}

    fo = ours(x)
    (fo * probe).sum().backward()
    g_ours = {k: p.grad.detach().clone() for k, p in ours.named_parameters()}
    stats_ours = {k: v.clone() for k, v in ours.state_dict().items() if "running_" in k}

    fr = ref(x)
    (fr * probe).sum().backward()
    g32 = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
    stats32 = {k: v.clone() for k, v in ref.state_dict().items() if "running_" in k}
    fr = fr.detach()

    ref.load_state_dict(sd0)
    ref.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fa = ref(x)
        (fa.float() * probe).sum().backward()
    g_amp = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
    fa = fa.detach().float()

    gmax = max(t.abs().max().item() for t in g32.values())
    floor = 1e-2 * gmax
    rows = []
    skip = {k for k in g32 if _structurally_zero(ours, k)}
    for k in g32:
        if k in skip:
            continue
        eo, eo2 = _tensor_errs(g_ours[k], g32[k], floor)
        ea, ea2 = _tensor_errs(g_amp[k], g32[k], floor)
        rows.append((k, eo, ea, eo2, ea2))
    res = {"name": name, "batch": batch, "h": h, "w": w, "mode": mode,
           "feat_ours": rel_err(fo, fr), "feat_amp": rel_err(fa, fr),
           "grad_worst_ours": max(rows, key=lambda r: r[1])[:3], "grad_worst_amp": max(rows, key=lambda r: r[2])[:3],
           "grad_l2_ours": max(r[3] for r in rows), "grad_l2_amp": max(r[4] for r in rows),
           "grad_global_ours": _global_l2(g_ours, g32, skip), "grad_global_amp": _global_l2(g_amp, g32, skip),
           "n_tensors": len(rows), "n_ours_gt_amp_and_1e-2": sum(1 for r in rows if r[1] > max(1e-2, r[2])),
           "stats_worst": max((rel_err(stats_ours[k], stats32[k]), k) for k in stats32) if train else None,
           "violations": [(r[0], round(r[1], 5), round(r[2], 5)) for r in rows if r[1] > max(1e-2, r[2])][:20]}
    return res, rows


def _assert_a13(res, rows, strict_single=True):
    """err(ours_bf16, oracle_fp32) <= max(1e-2, err(oracle_bf16_autocast, oracle_fp32)), asserted on the features, on
    the global gradient error (L2 over all parameters) and on the median per-tensor gradient error.  Per tensor, both bf16
    paths sit at the same noise level (the bf16-emulating fp32 oracle deviates from fp32 by the same
    amount), so the per-tensor statement is statistical: ours may exceed max(1e-2, autocast) on at most a quarter of the
    tensors and (eval mode) never by more than 3x."""
    import statistics
    print(json.dumps(res))
    _report(f"{res['name']}_{res['batch']}x{res['h']}x{res['w']}_{res['mode']}", res)
    slack = 1.0 if strict_single else 1.5     # single chaotic-regime inputs: the mean over inputs carries the A13 statement
    assert res["feat_ours"] <= max(1e-2, slack * res["feat_amp"]), f"features: ours {res['feat_ours']:.4g} vs autocast {res['feat_amp']:.4g}"
    assert res["grad_global_ours"] <= max(1e-2, slack * res["grad_global_amp"]), (res["grad_global_ours"], res["grad_global_amp"])
    med_o, med_a = statistics.median(r[1] for r in rows), statistics.median(r[2] for r in rows)
    assert med_o <= max(1e-2, med_a), (med_o, med_a)
    bad = [(k, eo, ea) for k, eo, ea, _, _ in rows if eo > max(1e-2, ea)]
    assert len(bad) <= len(rows) // 4, f"{len(bad)}/{len(rows)} gradient tensors worse than max(1e-2, autocast): {bad[:8]}"
    if res["mode"] == "eval":      # (train mode at full depth is in the chaotic regime: single tensors are noise on both sides)
        far = [(k, eo, ea) for k, eo, ea in bad if eo > 3 * max(1e-2, ea)]
        assert not far, far[:8]
    if res["stats_worst"] is not None:
        assert res["stats_worst"][0] < 1e-2, res["stats_worst"]


@pytest.mark.parametrize("name,batch,h,w", [("efficientnet-b2", 8, 320, 256), ("efficientnet-b5", 8, 456, 456)])
def test_full_depth_train_mode_a13(name, batch, h, w):
    """Train mode at full depth is chaotic (69/116 batch-statistics BN layers re-amplify every rounding): ours and PyTorch's own
    bf16 autocast both land 5-15 % from fp32 and which of the two is closer on ONE input is a coin flip (a change of the
    summation order of the BN partials flips it).  A13 is therefore asserted on the MEAN over three inputs (features and global
    gradient, 10 % statistical slack), and per input on the distributional statements of _assert_a13."""
    feats, gglob = [], []
    for seed in (1234, 1235, 1236):
        res, rows = _run_tower(name, batch, h, w, "train", seed=seed)
        feats.append((res["feat_ours"], res["feat_amp"]))
        gglob.append((res["grad_global_ours"], res["grad_global_amp"]))
        _assert_a13(res, rows, strict_single=False)
    mean = lambda xs, i: sum(x[i] for x in xs) / len(xs)
    print("train-mode A13 over 3 inputs: features ours/autocast", feats, "global gradient ours/autocast", gglob)
    assert mean(feats, 0) <= max(1e-2, 1.10 * mean(feats, 1)), feats
    assert mean(gglob, 0) <= max(1e-2, 1.10 * mean(gglob, 1)), gglob


@pytest.mark.parametrize("name,batch,h,w", [("efficientnet-b2", 8, 320, 256), ("efficientnet-b5", 8, 456, 456)])
def test_full_depth_eval_mode_tight(name, batch, h, w):
    """Running-statistics BN (no batch-statistics feedback): the north-star's 1e-2 on the features and on the global gradient,
    every gradient tensor within 5e-2 of its own scale (measured worst 2.7e-2 / 4.5e-2; autocast 4.3e-2 / 7.0e-2)."""
    res, rows = _run_tower(name, batch, h, w, "eval")
    _assert_a13(res, rows)
    assert res["feat_ours"] < 1e-2 and res["grad_global_ours"] < 1e-2
    assert max(r[1] for r in rows) < 5e-2, max(rows, key=lambda r: r[1])


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_c3_geometry_b5_1520x912(mode):
    """The metric config's geometry (every layer shape, tile counts, > 255-pixel rows, 48x29 last stage) at B = 4."""
    res, rows = _run_tower("efficientnet-b5", 4, 1520, 912, mode)
    _assert_a13(res, rows, strict_single=(mode == "eval"))     # train mode: one chaotic-regime input (see test_full_depth_train_mode_a13)
    if mode == "eval":
        assert res["feat_ours"] < 1e-2 and res["grad_global_ours"] < 1e-2
        assert max(r[1] for r in rows) < 5e-2, max(rows, key=lambda r: r[1])
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ teacher-forced blocks
def _to_nhwc_bf16(t):
    return t.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _blockwise(name, batch, h, w):
    """TRAIN-mode parity of every MBConv block at its true geometry without error compounding: the fp32 oracle runs the whole
    tower once; each block of ours is then fed the ORACLE's (bf16-rounded) block input and output gradient and compared with
    fp32 autograd of the oracle's block on the same tensors: output, input gradient, every parameter gradient."""
    from test_gpu_encoder import _build
    from mammoclip_b200 import ops
    from mammoclip_b200.model.modules import efficientnet_custom as E
    from oracle import port
    ours, ref = _build(name)
    ours.train(), ref.train()
    x = port.synth_images(batch, h, w, seed=1234, identical_channels=False, device="cuda")
    probe = torch.randn(batch, ref.out_dim, generator=torch.Generator().manual_seed(99)).cuda()
    xin, gout, hooks = {}, {}, []
    for i, blk in enumerate(ref._blocks):
        def fwd_hook(m, inp, out, i=i):
            xin[i] = inp[0].detach().to(torch.bfloat16)
            out.register_hook(lambda g, i=i: gout.__setitem__(i, g.detach().to(torch.bfloat16)))
        hooks.append(blk.register_forward_hook(fwd_hook))
    fr = ref(x)
    (fr * probe).sum().backward()
    for hk in hooks:
        hk.remove()
    ours._weights().refresh()
    n = batch
    report = []
    for i, blk in enumerate(ref._blocks):
        g_nchw = gout[i].float()
        grads = {}
        if i == 0:
            # block 0 is fed by the stem: BN0 + swish of the stem output are applied by the depthwise loader, so the stem is part
            # of the unit under test (reference: conv_stem -> bn0 -> swish -> block 0, fp32 autograd)
            wc = ours._weights()
            ys, st, patches = ops.stem_forward(x, ours._conv_stem.weight, ours.geom.stem_pads, want_stats=True, w_bf16=wc.bf16[("s",)], return_patches=True)
            hh, ww = ys.shape[1], ys.shape[2]
            bn = E._bn_fin(st, n * hh * ww, ours._bn0, True)
            S = {"images": x, "patches": patches, "stem": (ys, bn)}
            yo, B = E._block_forward(ours, 0, None, (ys, bn), n, hh, ww, True)
            dxo = E._block_backward(ours, 0, B, _to_nhwc_bf16(g_nchw), n, True, grads, torch.empty_like, S)
            assert dxo is None
            ref.zero_grad(set_to_none=True)
            a0 = port.swish(ref._bn0(ref._conv_stem(x)))
            yr = blk(a0, 0.0)
            yr.backward(g_nchw)
            gref = {f"_blocks.0.{k}": p.grad for k, p in blk.named_parameters()}
            gref.update({"_conv_stem.weight": ref._conv_stem.weight.grad, "_bn0.weight": ref._bn0.weight.grad, "_bn0.bias": ref._bn0.bias.grad})
            e_dx = 0.0
        else:
            xi = xin[i]
            hh, ww = xi.shape[2], xi.shape[3]
            yo, B = E._block_forward(ours, i, _to_nhwc_bf16(xi), None, n, hh, ww, True)
            dxo = E._block_backward(ours, i, B, _to_nhwc_bf16(g_nchw), n, True, grads, torch.empty_like)
            blk.zero_grad(set_to_none=True)
            xr = xi.float().requires_grad_(True)
            yr = blk(xr, 0.0)
            yr.backward(g_nchw)
            gref = {f"_blocks.{i}.{k}": p.grad.clone() for k, p in blk.named_parameters()}
            e_dx = rel_err(dxo.float().permute(0, 3, 1, 2), xr.grad)
            # the same block under PyTorch bf16 autocast (what the reference's AMP path computes), for the A13 bound
            blk.zero_grad(set_to_none=True)
            xa = xi.float().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ya = blk(xa, 0.0)
                ya.backward(g_nchw.to(ya.dtype))
            gamp = {f"_blocks.{i}.{k}": p.grad.clone() for k, p in blk.named_parameters()}
            amp = {"y": rel_err(ya.float(), yr), "dx": rel_err(xa.grad, xr.grad)}
        e_y = rel_err(yo.float().permute(0, 3, 1, 2), yr)
        gmax = max(t.abs().max().item() for t in gref.values())
        worst, worst_amp, per = ("", 0.0), ("", 0.0), {}
        for k, t in gref.items():
            scale = max(t.abs().max().item(), 1e-2 * gmax)
            d = (grads[k].double() - t.double()).abs().max().item() / scale
            per[k.split(".", 2)[-1]] = [d]
            if d > worst[1]:
                worst = (k, d)
            if i > 0:
                da = (gamp[k].double() - t.double()).abs().max().item() / scale
                per[k.split(".", 2)[-1]].append(da)
                if da > worst_amp[1]:
                    worst_amp = (k, da)
        report.append({"block": i, "geom": [blk.spec.cin, blk.spec.cexp, blk.spec.cout, blk.spec.k, blk.spec.s, hh, ww], "y": e_y, "dx": e_dx, "worst_grad": worst,
                       "amp": dict(amp, worst_grad=worst_amp) if i > 0 else None, "per_tensor": per})
        del B, yo, dxo, yr, grads, gref
    return report


@pytest.mark.parametrize("name,batch,h,w", [("efficientnet-b5", 4, 1520, 912), ("efficientnet-b2", 8, 456, 456)])
def test_every_block_train_mode_teacher_forced(name, batch, h, w):
    rep = _blockwise(name, batch, h, w)
    _report(f"blockwise_{name}_{batch}x{h}x{w}", rep)
    wy, wdx, wg = max(r["y"] for r in rep), max(r["dx"] for r in rep), max(rep, key=lambda r: r["worst_grad"][1])
    print(f"blockwise {name} {batch}x{h}x{w}: worst y {wy:.4f} dx {wdx:.4f} grad {wg['worst_grad']} (block {wg['block']})")
    import statistics
    for r in rep:
        if r["block"] == 0:                           # block 0 includes the stem conv + BN0 (two more bf16 storage points)
            assert r["y"] < 2e-2 and r["worst_grad"][1] < 5e-2, r
            continue
        assert r["y"] < 1e-2, {k: r[k] for k in ("block", "geom", "y")}         # north-star bf16 tolerance on every block output
        assert r["dx"] < 1e-2, {k: r[k] for k in ("block", "geom", "dx")}        # ... and on every block's input gradient
        # parameter gradients: weights are bf16 GEMM operands on both bf16 paths; the worst tensor (BN0 gamma: a cancelling
        # sum of dv*yhat) sits at 1-4e-2 for ours and 2-15e-2 for PyTorch autocast of the same block
        assert r["worst_grad"][1] < 5e-2, {k: r[k] for k in ("block", "geom", "worst_grad")}
    ours_w = [r["worst_grad"][1] for r in rep if r["block"] > 0]
    amp_w = [r["amp"]["worst_grad"][1] for r in rep if r["block"] > 0]
    # SURVEY A13 over the blocks: no worse than autocast in the median and in the worst block
    assert statistics.median(ours_w) <= max(1e-2, statistics.median(amp_w)), (statistics.median(ours_w), statistics.median(amp_w))
    assert max(ours_w) <= max(1e-2, max(amp_w)), (max(ours_w), max(amp_w))
    assert statistics.median(r["y"] for r in rep if r["block"] > 0) <= max(1e-2, statistics.median(r["amp"]["y"] for r in rep if r["block"] > 0))
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ > 2^31-element tensors
def _rand_bf16(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    out = torch.empty(shape, dtype=torch.bfloat16, device="cuda")
    n0 = shape[0]
    for i in range(0, n0, 4):                       # chunked: a 2.4 G-element fp32 temporary would be 10 GB
        out[i:i + 4] = (torch.randn((min(4, n0 - i),) + tuple(shape[1:]), generator=g, device="cuda") * scale).to(torch.bfloat16)
    return out


BIG_N, BIG_H, BIG_W, BIG_C = 44, 760, 456, 144       # 44*760*456*144 = 2.196e9 > 2^31 elements (block 3 of EN-B5 at c3)


def test_dwconv_k3s2_over_2g_elements():
    from mammoclip_b200 import ops
    from test_gpu_conv import _bn_state
    n, h, w, c, k, s, pads = BIG_N, BIG_H, BIG_W, BIG_C, 3, 2, (0, 1, 0, 1)
    assert n * h * w * c > 2 ** 31
    x = _rand_bf16((n, h, w, c), 1)
    wt = (torch.randn(c, 1, k, k, device="cuda") * 0.3).contiguous()
    bn = _bn_state(c, 3)
    y, stats = ops.dwconv_forward(x, wt, k, s, pads, bn=bn)
    dy = _rand_bf16(tuple(y.shape), 4)
    dwt = torch.empty_like(wt)
    dx, bnp = ops.dwconv_backward(x, wt, k, s, pads, dy, dwt, bn=bn)
    torch.cuda.synchronize()
    dw_ref = torch.zeros_like(wt, dtype=torch.float64)
    s0 = torch.zeros(c, dtype=torch.float64, device="cuda"); s1 = torch.zeros_like(s0); p0 = torch.zeros_like(s0); p1 = torch.zeros_like(s0)
    worst_y = worst_dx = 0.0
    for i in range(0, n, 4):                         # per-sample-chunk fp32 reference (the op is independent across samples)
        xs = x[i:i + 4].float()
        v = xs * bn.scale + bn.shift
        a = (v * torch.sigmoid(v)).to(torch.bfloat16).float().permute(0, 3, 1, 2).detach().requires_grad_(True)
        wr = wt.clone().requires_grad_(True)
        yr = F.conv2d(F.pad(a, pads), wr, stride=s, groups=c)
        yr.backward(dy[i:i + 4].float().permute(0, 3, 1, 2))
        worst_y = max(worst_y, rel_err(y[i:i + 4].float(), yr.permute(0, 2, 3, 1)))
        sg = torch.sigmoid(v)
        dv = a.grad.permute(0, 2, 3, 1) * (sg * (1 + v * (1 - sg)))
        worst_dx = max(worst_dx, rel_err(dx[i:i + 4].float(), dv))
        dw_ref += wr.grad.double()
        yd = y[i:i + 4].double().reshape(-1, c)
        s0 += yd.sum(0); s1 += (yd * yd).sum(0)
        dvq = dx[i:i + 4].double().reshape(-1, c)
        yh = ((xs - bn.mean) * bn.invstd).double().reshape(-1, c)
        p0 += dvq.sum(0); p1 += (dvq * yh).sum(0)
        del xs, v, a, yr, sg, dv, yd, dvq, yh
    print(f"dw >2^31: y {worst_y:.2e} dx {worst_dx:.2e}")
    assert worst_y < 8e-3 and worst_dx < 1e-2
    assert rel_err(dwt, dw_ref) < 5e-3
    st = stats.double().sum(0)
    assert rel_err(st[0], s0) < 1e-4 and rel_err(st[1], s1) < 1e-4
    p = bnp.double().sum(0)
    assert rel_err(p[0], p0) < 5e-3 and rel_err(p[1], p1) < 5e-3       # sums of the fp32 gradient vs sums of the stored bf16 values


def test_streaming_passes_over_2g_elements():
    from mammoclip_b200 import ops
    from test_gpu_conv import _bn_state
    n, hw, c = BIG_N, BIG_H * BIG_W, BIG_C
    y, du = _rand_bf16((n, hw, c), 5), _rand_bf16((n, hw, c), 6)
    st = _bn_state(c, 7)
    gate, dpool = torch.rand(n, c, device="cuda"), torch.randn(n, c, device="cuda") * 0.01
    out, pool = ops.ew_forward(y, bn=st, act=1, pool=True)
    part = ops.ew_backward(0, y, st, 1, du=du, gate=gate, dpool=dpool)
    c1, c2 = torch.randn(c, device="cuda") * 0.01, torch.randn(c, device="cuda") * 0.01
    dyo = ops.ew_backward(1, y, st, 1, du=du, gate=gate, dpool=dpool, c1=c1, c2=c2)
    a2, sep = ops.ew_backward(2, y, st, 1, du=du, gate=gate)
    torch.cuda.synchronize()
    acc = torch.zeros(2, c, dtype=torch.float64, device="cuda")
    for i in list(range(0, n, 11)) + [n - 1]:        # first, last and a few samples in between: offsets beyond 2^31 included
        v = y[i].float() * st.scale + st.shift
        sg = torch.sigmoid(v)
        u = v * sg
        assert rel_err(out[i].float(), u) < 8e-3, i
        assert rel_err(pool[i].sum(0), out[i].float().sum(0)) < 1e-4, i
        dv = (du[i].float() * gate[i] + dpool[i]) * (sg * (1 + v * (1 - sg)))
        yh = (y[i].float() - st.mean) * st.invstd
        ref_dy = st.scale * (dv - c1 - yh * c2)
        assert rel_err(dyo[i].float(), ref_dy) < 1e-2, i
        assert rel_err(a2[i].float(), u * gate[i]) < 8e-3, i
        assert rel_err(sep[i, :, 0].sum(0), (du[i].float() * u).sum(0)) < 5e-3, i
    for i in range(n):
        v = y[i].float() * st.scale + st.shift
        sg = torch.sigmoid(v)
        dv = (du[i].float() * gate[i] + dpool[i]) * (sg * (1 + v * (1 - sg)))
        yh = (y[i].float() - st.mean) * st.invstd
        acc[0] += dv.double().sum(0); acc[1] += (dv * yh).double().sum(0)
    p = part.double().sum(0)
    assert rel_err(p[0], acc[0]) < 5e-3 and rel_err(p[1], acc[1]) < 5e-3


def test_small_k_gemm_22m_rows_over_2g_outputs():
    """Expand conv of block 3 at c3: M = 64*760*456 = 22.2 M rows, K = 24 -> N = 144 (3.2 G outputs) with BN partials, and the
    matching weight-gradient GEMM over the same rows."""
    from mammoclip_b200 import ops
    m, k, n = 64 * 760 * 456, 24, 144
    assert m * n > 2 ** 31
    a = _rand_bf16((64, 760 * 456, k), 8).view(m, k)
    w = (torch.randn(n, k, device="cuda") / k ** 0.5).to(torch.bfloat16)
    out, stats = ops.gemm_tn(a, w, want_stats=True)
    torch.cuda.synchronize()
    s0 = torch.zeros(n, dtype=torch.float64, device="cuda"); s1 = torch.zeros_like(s0)
    step = 760 * 456 * 4
    worst = 0.0
    for r0 in range(0, m, step):
        ref = a[r0:r0 + step].float() @ w.float().T
        worst = max(worst, rel_err(out[r0:r0 + step].float(), ref))
        o = out[r0:r0 + step].double()
        s0 += o.sum(0); s1 += (o * o).sum(0)
        del ref, o
    assert worst < 6e-3, worst
    st = stats.double().sum(0)
    assert rel_err(st[0], s0) < 1e-4 and rel_err(st[1], s1) < 1e-4
    dwt = ops.gemm_wgrad(out, a)                     # [144, 24] = out^T a over 22 M rows
    ref = torch.zeros(n, k, dtype=torch.float64, device="cuda")
    for r0 in range(0, m, step):
        ref += (out[r0:r0 + step].float().T @ a[r0:r0 + step].float()).double()
    assert rel_err(dwt, ref) < 2e-3
