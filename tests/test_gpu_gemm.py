"""tcgen05 GEMMs vs a plain PyTorch fp32 reference on the same bf16 inputs (tolerance: bf16 output rounding)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


# (batches, M, N, K): MBConv expand/project shapes incl. hostile channel counts, BERT shapes, ragged M
TN_SHAPES = [
    (1, 128, 64, 64), (1, 256, 16, 24), (1, 1000, 144, 24), (1, 4096, 240, 40), (1, 777, 40, 240),
    (1, 4096, 768, 768), (1, 4096, 3072, 768), (1, 4096, 768, 3072), (1, 2784, 304, 1824), (1, 1392, 1056, 176),
    (1, 33, 512, 2048), (1, 20000, 24, 144), (1, 300, 2048, 512), (1, 128, 1408, 352),
    # many M tiles per CTA with a shared weight matrix whose [block_n, K] panel fits shared memory: the resident-panel mode
    (1, 89088, 1824, 304), (1, 60001, 1056, 176), (1, 40000, 768, 128), (3, 30000, 384, 304),
]


@pytest.mark.parametrize("bt,m,n,k", TN_SHAPES)
def test_gemm_tn(bt, m, n, k):
    from mammoclip_b200 import ops
    a, w = _mk((m, k), 1), _mk((n, k), 2, 1.0 / k ** 0.5)
    out = ops.gemm_tn(a, w)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T
    err = rel_err(out.float(), ref)
    assert err < 6e-3, f"gemm_tn {m}x{n}x{k}: rel err {err}"


def test_gemm_tn_batched_per_sample_weights():
    from mammoclip_b200 import ops
    bt, m, n, k = 5, 300, 40, 240          # project conv with SE-gated per-sample weights; M not a tile multiple
    a, w = _mk((bt, m, k), 3), _mk((bt, n, k), 4, 1.0 / k ** 0.5)
    out = ops.gemm_tn(a, w)
    ref = torch.einsum("bmk,bnk->bmn", a.float(), w.float())
    assert rel_err(out.float(), ref) < 6e-3
    out2 = ops.gemm_tn(a, w[0].contiguous())
    ref2 = torch.einsum("bmk,nk->bmn", a.float(), w[0].float())
    assert rel_err(out2.float(), ref2) < 6e-3


@pytest.mark.parametrize("m,n,k", [(1000, 144, 24), (5000, 384, 64), (4096, 1824, 304), (257, 96, 16), (89088, 1824, 304), (70000, 1056, 176)])
def test_gemm_tn_bn_stats(m, n, k):
    from mammoclip_b200 import ops
    a, w = _mk((m, k), 5), _mk((n, k), 6, 1.0 / k ** 0.5)
    out, stats = ops.gemm_tn(a, w, want_stats=True)
    s = stats.double().sum(0)
    o = out.double()
    assert rel_err(s[0], o.sum(0)) < 1e-4
    assert rel_err(s[1], (o * o).sum(0)) < 1e-4


def test_gemm_tn_epilogues():
    from mammoclip_b200 import ops
    m, n, k = 4096, 768, 768
    a, w = _mk((m, k), 7), _mk((n, k), 8, 1.0 / k ** 0.5)
    bias = torch.randn(n, device="cuda")
    res = _mk((m, n), 9)
    out = ops.gemm_tn(a, w, bias=bias, act=1)
    ref = torch.nn.functional.gelu(a.float() @ w.float().T + bias)
    assert rel_err(out.float(), ref) < 6e-3
    out = ops.gemm_tn(a, w, bias=bias, residual=res)
    ref = a.float() @ w.float().T + bias + res.float()
    assert rel_err(out.float(), ref) < 6e-3


@pytest.mark.parametrize("r,i,j", [(4096, 144, 24), (20000, 40, 240), (1000, 768, 768), (2784, 304, 1824), (100, 24, 24),
                                   (8192, 3072, 768), (50000, 48, 48)])
def test_gemm_wgrad(r, i, j):
    from mammoclip_b200 import ops
    a, b = _mk((r, i), 10), _mk((r, j), 11)
    out = ops.gemm_wgrad(a, b)
    ref = a.float().T @ b.float()
    err = rel_err(out, ref)
    assert err < 2e-3, f"wgrad {r}x{i}x{j}: rel err {err}"
    out2 = ops.gemm_wgrad(a, b, out=out.clone(), accumulate=True)
    assert rel_err(out2, 2 * ref) < 2e-3


@pytest.mark.parametrize("rows,cols,ld_extra", [(4096, 768, 0), (4096, 768, 1536), (513, 3072, 0), (7, 512, 0), (64, 6, 2)])
def test_colsum(rows, cols, ld_extra):
    """Bias gradients: column sums of a (possibly column-sliced) bf16 matrix, written or accumulated."""
    from mammoclip_b200 import ops
    torch.manual_seed(rows + cols)
    full = torch.randn(rows, cols + ld_extra, device="cuda").bfloat16()
    x = full[:, ld_extra // 2: ld_extra // 2 + cols]
    out = torch.empty(cols, device="cuda")
    ops.colsum(x, out)
    ref = x.double().sum(0)
    assert ((out.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
    ops.colsum(x, out, accumulate=True)
    assert ((out.double() - 2 * ref).abs().max() / ref.abs().max()).item() < 1e-5


def test_weight_prep_straight_and_transposed_copies():
    """fp32 master weights -> bf16 operands: row-strided destination (stem), transposed copy, fused-QKV layout with strides."""
    from mammoclip_b200 import ops
    torch.manual_seed(0)
    a, b, q = torch.randn(100, 27, device="cuda"), torch.randn(3072, 768, device="cuda"), [torch.randn(768, 768, device="cuda") for _ in range(3)]
    a_d = torch.zeros(100, 32, dtype=torch.bfloat16, device="cuda")
    b_d, b_t = torch.empty(3072, 768, dtype=torch.bfloat16, device="cuda"), torch.empty(768, 3072, dtype=torch.bfloat16, device="cuda")
    qkv, qkv_t = torch.empty(2304, 768, dtype=torch.bfloat16, device="cuda"), torch.empty(768, 2304, dtype=torch.bfloat16, device="cuda")
    entries = [(a, a_d, None, 32), (b, b_d, b_t), (b, None, b_t)]
    entries += [(q[i], qkv[i * 768:(i + 1) * 768], qkv_t[:, i * 768:(i + 1) * 768], 0, 2304) for i in range(3)]
    table = ops.weight_prep(entries, a.device)
    ops.weight_prep_run(table, len(entries))
    assert torch.equal(a_d[:, :27], a.bfloat16()) and a_d[:, 27:].abs().max().item() == 0
    assert torch.equal(b_d, b.bfloat16()) and torch.equal(b_t, b.bfloat16().t())
    assert torch.equal(qkv, torch.cat(q).bfloat16()) and torch.equal(qkv_t, torch.cat(q).bfloat16().t())


def test_gemm_tn_pre_activation_copy_and_dropout_epilogue():
    """aux_pre receives the bf16 value before GELU (BertIntermediate backward); dropout keep-mask + residual epilogue."""
    from mammoclip_b200 import ops
    m, n, k = 1000, 3072, 768
    a, w = _mk((m, k), 12), _mk((n, k), 13, 1.0 / k ** 0.5)
    bias = torch.randn(n, device="cuda")
    pre = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
    out = ops.gemm_tn(a, w, bias=bias, act=1, aux_pre=pre)
    ref_pre = a.float() @ w.float().T + bias
    assert rel_err(pre.float(), ref_pre) < 6e-3
    assert rel_err(out.float(), torch.nn.functional.gelu(ref_pre)) < 6e-3
    keep = (torch.rand(m, n, device="cuda") >= 0.1).to(torch.uint8)
    res = _mk((m, n), 14)
    out = ops.gemm_tn(a, w, bias=bias, residual=res, dropmask=keep, drop_scale=1.0 / 0.9)
    assert rel_err(out.float(), ref_pre * keep / 0.9 + res.float()) < 6e-3


@pytest.mark.parametrize("bt,m,n,k", [(1, 4096, 240, 40), (1, 20000, 24, 144), (1, 4096, 3072, 768), (1, 2784, 304, 1824), (3, 1392, 176, 1056), (1, 333, 40, 240)])
def test_gemm_tn_epilogue_instantiations_agree(bt, m, n, k, monkeypatch):
    """The specialised epilogues (plain / +residual, with / without BN partials) must reproduce the all-in-one epilogue bit
    for bit (same accumulators, same rounding points), and the residual path must match the fp32 reference."""
    from mammoclip_b200 import ops
    a = _mk((bt, m, k) if bt > 1 else (m, k), 15)
    w = _mk((bt, n, k) if bt > 1 else (n, k), 16, 1.0 / k ** 0.5)
    res = _mk((bt, m, n) if bt > 1 else (m, n), 17)
    o_plain, s_plain = ops.gemm_tn(a, w, want_stats=True)
    o_res = ops.gemm_tn(a, w, residual=res)
    monkeypatch.setenv("MCLIP_GEMM_GENERIC", "1")
    o_gen, s_gen = ops.gemm_tn(a, w, want_stats=True)
    o_gres = ops.gemm_tn(a, w, residual=res)
    assert torch.equal(o_plain, o_gen) and torch.equal(s_plain, s_gen) and torch.equal(o_res, o_gres)
    ref = torch.einsum("bmk,bnk->bmn", a.float(), w.float()) if bt > 1 else a.float() @ w.float().T
    assert rel_err(o_plain.float(), ref) < 6e-3 and rel_err(o_res.float(), ref + res.float()) < 6e-3


def test_gemm_tn_k_concatenated_second_operand():
    """A = [a | a2] along K with the weights laid out as [n, ceil64(k) + k2] (zeros in the padding columns)."""
    from mammoclip_b200 import ops
    for m, n, k, k2 in [(5000, 40, 240, 40), (777, 24, 144, 24), (4096, 304, 1824, 304), (300, 64, 384, 64)]:
        a, a2 = _mk((m, k), 21), _mk((m, k2), 22)
        w, w2 = _mk((n, k), 23, 1.0 / k ** 0.5), _mk((n, k2), 24, 1.0 / k2 ** 0.5)
        kp = (k + 63) // 64 * 64
        wcat = torch.zeros((n, kp + k2), dtype=torch.bfloat16, device="cuda")
        wcat[:, :k], wcat[:, kp:] = w, w2
        bias = torch.randn(n, device="cuda")
        res = _mk((m, n), 25)
        out = ops.gemm_tn(a, wcat, a2=a2, bias=bias, residual=res)
        ref = a.float() @ w.float().T + a2.float() @ w2.float().T + bias + res.float()
        assert rel_err(out.float(), ref) < 6e-3, (m, n, k, k2)


@pytest.mark.parametrize("m,cin,cexp,training", [(20000, 40, 240, True), (6000, 24, 144, True), (3000, 304, 1824, True), (5000, 64, 384, False)])
def test_bn0_fold_backward_matches_the_unfolded_path(m, cin, cexp, training):
    """Folded BatchNorm backward of the expand conv (dX = [dV0|X][aWe;G] + bias, dWe from dV0^T X and X^T X) vs fp32 autograd
    algebra of bn0(conv(x)) on the same bf16 tensors; X carries a non-zero mean so that the centring matters."""
    from mammoclip_b200 import ops
    torch.manual_seed(m + cin)
    x = (torch.randn(m, cin, device="cuda") * 0.7 + torch.randn(cin, device="cuda") * 0.8).to(torch.bfloat16)
    we = torch.randn(cexp, cin, device="cuda") / cin ** 0.5
    we_b = we.to(torch.bfloat16)
    y0 = (x.float() @ we_b.float().T).to(torch.bfloat16)                    # what the forward stored
    dv0 = (torch.randn(m, cexp, device="cuda") * 0.3 + torch.randn(cexp, device="cuda") * 0.05).to(torch.bfloat16)
    gamma = torch.rand(cexp, device="cuda") + 0.5
    yf = y0.float()
    mean, var = yf.mean(0), yf.var(0, unbiased=False)
    bn = ops.BNState(cexp, "cuda")
    bn.mean.copy_(mean); bn.invstd.copy_((var + 1e-3).rsqrt()); bn.scale.copy_(gamma * bn.invstd); bn.shift.zero_()
    yhat = (yf - bn.mean) * bn.invstd
    dvf = dv0.float()
    if training:
        c1, c2 = dvf.mean(0), (dvf * yhat).mean(0)
    else:
        c1, c2 = torch.zeros(cexp, device="cuda"), torch.zeros(cexp, device="cuda")
    dy0 = bn.scale * (dvf - c1 - yhat * c2)
    dx_ref = dy0 @ we                                                          # fp32 truth on the stored tensors
    dwe_ref = dy0.T @ x.float()
    skip = (torch.randn(m, cin, device="cuda") * 0.1).to(torch.bfloat16)
    dwe = torch.empty_like(we)
    dx = ops.bn0_fold_backward(dv0, x, we, we_b, bn, c1.contiguous(), c2.contiguous(), dwe, residual=skip)
    # today's path for comparison: dY0 rounded to bf16, bf16 weights
    dy0_b = dy0.to(torch.bfloat16)
    dx_old = ops.gemm_tn(dy0_b, we_b.t().contiguous(), residual=skip)
    e_new, e_old = rel_err(dx.float(), dx_ref + skip.float()), rel_err(dx_old.float(), dx_ref + skip.float())
    w_new, w_old = rel_err(dwe, dwe_ref), rel_err(ops.gemm_wgrad(dy0_b, x), dwe_ref)
    print(f"bn0 fold m={m} cin={cin} cexp={cexp}: dX err {e_new:.2e} (unfolded {e_old:.2e}), dWe err {w_new:.2e} (unfolded {w_old:.2e})")
    assert e_new < 1e-2 and w_new < 1e-2
