"""Depthwise / stem / BN / SE kernels vs plain PyTorch fp32 references on the same (bf16-rounded) inputs."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def _bn_state(c, seed):
    from mammoclip_b200 import ops
    st = ops.BNState(c, "cuda")
    g = torch.Generator(device="cuda").manual_seed(seed)
    st.scale.copy_(torch.rand(c, generator=g, device="cuda") + 0.5)
    st.shift.copy_(torch.randn(c, generator=g, device="cuda") * 0.3)
    st.mean.copy_(torch.randn(c, generator=g, device="cuda") * 0.2)
    st.invstd.copy_(torch.rand(c, generator=g, device="cuda") + 0.5)
    st.count, st.training = 1, True
    return st


DW_CASES = [  # (N,H,W,C,k,s,pads(l,r,t,b), with_bn)
    (2, 20, 24, 48, 3, 1, (1, 1, 1, 1), True), (2, 33, 17, 144, 3, 2, (0, 1, 0, 1), True), (1, 19, 23, 240, 5, 2, (1, 2, 1, 2), True),
    (2, 24, 16, 384, 5, 1, (2, 2, 2, 2), True), (2, 31, 29, 64, 3, 2, (1, 1, 1, 1), True), (1, 29, 29, 88, 5, 2, (2, 2, 2, 2), True),
    (3, 12, 40, 24, 3, 1, (1, 1, 1, 1), False), (1, 7, 7, 3072, 3, 1, (1, 1, 1, 1), True), (2, 57, 57, 16, 3, 1, (1, 1, 1, 1), False),
    # narrow k3 s1 layers on 16-channel lane groups (5-column sub-strips, 80-column CTA strips): several strips, ragged right edge
    (2, 30, 170, 24, 3, 1, (1, 1, 1, 1), False), (1, 40, 95, 48, 3, 1, (1, 1, 1, 1), True), (2, 26, 81, 32, 3, 1, (1, 1, 1, 1), True),
    (1, 50, 7, 16, 3, 1, (1, 1, 1, 1), True), (2, 64, 240, 24, 3, 1, (1, 1, 1, 1), True),
]


def _dw_ref(x, w, k, s, pads, bn):
    xf = x.float().permute(0, 3, 1, 2)
    if bn is not None:
        v = xf * bn.scale.view(1, -1, 1, 1) + bn.shift.view(1, -1, 1, 1)
        xf = v * torch.sigmoid(v)                                   # BN + swish in fp32 registers (stride 2 still stages a bf16 tile)
    xf = xf.detach().requires_grad_(True)
    y = F.conv2d(F.pad(xf, pads), w, stride=s, groups=w.shape[0])
    return xf, y


@pytest.mark.parametrize("n,h,w,c,k,s,pads,with_bn", DW_CASES)
def test_dwconv_forward_backward(n, h, w, c, k, s, pads, with_bn):
    from mammoclip_b200 import ops
    x = _rand((n, h, w, c), 1)
    wt = (_rand((c, 1, k, k), 2, 0.3, torch.float32)).contiguous()
    bn = _bn_state(c, 3) if with_bn else None
    y, stats = ops.dwconv_forward(x, wt, k, s, pads, bn=bn)
    wref = wt.clone().requires_grad_(True)
    xf, yref = _dw_ref(x, wref, k, s, pads, bn)
    yr = yref.permute(0, 2, 3, 1)
    assert y.shape == yr.shape
    assert rel_err(y.float(), yr) < 8e-3
    # BatchNorm partials: sums of the fp32 accumulators (the bf16 rounding of the stored tensor is unbiased), i.e. the fp32
    # reference's own statistics; against the stored bf16 values they agree to the rounding noise of the sum
    st = stats.double().sum(0)
    yd, yrd = y.double().reshape(-1, c), yr.detach().double().reshape(-1, c)
    assert rel_err(st[0], yrd.sum(0)) < 3e-3 and rel_err(st[1], (yrd * yrd).sum(0)) < 3e-3
    assert rel_err(st[0], yd.sum(0)) < 5e-3 and rel_err(st[1], (yd * yd).sum(0)) < 5e-3
    # backward
    dy = _rand(y.shape, 4)
    dwt = torch.empty_like(wt)
    dx, bnp = ops.dwconv_backward(x, wt, k, s, pads, dy, dwt, bn=bn)
    yref.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_err(dwt, wref.grad) < 5e-3, "dweight"
    dA = xf.grad.permute(0, 2, 3, 1)
    if bn is None:
        assert rel_err(dx.float(), dA) < 8e-3, "dx"
    else:
        v = x.float() * bn.scale + bn.shift
        sg = torch.sigmoid(v)
        dv = dA * (sg * (1 + v * (1 - sg)))
        assert rel_err(dx.float(), dv) < 1e-2, "dv"
        # BN-backward partials: sums of the fp32 gradient (stride 1) / of the stored bf16 values (stride 2); both within the
        # rounding noise of the sum of each other and of the fp32 reference
        yh = ((x.float() - bn.mean) * bn.invstd).double().reshape(-1, c)
        p = bnp.double().sum(0)
        dvq, dvr = dx.double().reshape(-1, c), dv.double().reshape(-1, c)
        assert rel_err(p[0], dvr.sum(0)) < 3e-3 and rel_err(p[1], (dvr * yh).sum(0)) < 3e-3
        assert rel_err(p[0], dvq.sum(0)) < 5e-3 and rel_err(p[1], (dvq * yh).sum(0)) < 5e-3


@pytest.mark.parametrize("n,h,w,c,pads,nhwc", [(2, 64, 48, 32, (0, 1, 0, 1), True), (3, 31, 45, 48, (0, 1, 0, 1), False), (1, 96, 64, 40, (1, 1, 1, 1), True)])
def test_stem(n, h, w, c, pads, nhwc):
    from mammoclip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.randn(n, h, w, 3, generator=g, device="cuda").permute(0, 3, 1, 2) if nhwc else torch.randn(n, 3, h, w, generator=g, device="cuda")
    wt = _rand((c, 3, 3, 3), 6, 0.3, torch.float32)
    y, stats = ops.stem_forward(img, wt, pads)
    # the stem runs as im2col (bf16 patches, like autocast) + tcgen05 GEMM with bf16 weights
    wref = wt.to(torch.bfloat16).float().requires_grad_(True)
    imgq = img.to(torch.bfloat16).float()
    yref = F.conv2d(F.pad(imgq, pads), wref, stride=2)
    assert rel_err(y.float(), yref.permute(0, 2, 3, 1)) < 6e-3
    assert rel_err(y.float(), F.conv2d(F.pad(img, pads), wt, stride=2).permute(0, 2, 3, 1)) < 2e-2      # vs the fp32 conv
    st = stats.double().sum(0)
    yd = y.double().reshape(-1, c)
    assert rel_err(st[0], yd.sum(0)) < 1e-4 and rel_err(st[1], (yd * yd).sum(0)) < 1e-4
    dy = _rand(y.shape, 7)
    dwt = torch.empty_like(wt)
    ops.stem_wgrad(img, dy, pads, dwt)
    yref.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_err(dwt, wref.grad) < 2e-3


def test_stem_input_edge_single_channel_and_uint8_are_bit_identical():
    """the data pipeline's image is ONE grey channel replicated three times (datasets/imagetext.py:121) and
    normalised per image on the CPU (:129-134).  The single-channel fp32 input and the raw uint8 input (normalised on load in
    the reference's fp32 operation order) must reproduce the 3-identical-channel fp32 path bit for bit: patches, stem output,
    BN partials and the stem weight gradient."""
    from mammoclip_b200 import ops
    n, h, w, c, pads = 3, 70, 52, 48, (0, 1, 0, 1)
    g = torch.Generator(device="cuda").manual_seed(11)
    u8 = torch.randint(3, 250, (n, 1, h, w), generator=g, device="cuda", dtype=torch.uint8)
    mean, std = 0.3089279, 0.25053555408335154
    t = u8.float()
    t = t - t.amin(dim=(1, 2, 3), keepdim=True)                 # image -= image.min()
    t = t / t.amax(dim=(1, 2, 3), keepdim=True)                 # image /= image.max()
    x1 = (t - mean) / std                                       # fp32, like numpy float32 with python-float scalars
    x3 = x1.expand(n, 3, h, w)                                  # what convert('RGB') + the trainer's permute deliver
    wt = _rand((c, 3, 3, 3), 6, 0.3, torch.float32)
    dy = None
    outs = []
    mm, lut = ops.image_norm_lut_u8(u8, mean, std)
    for img, kw in ((x3, {}), (x1, {}), (u8, dict(norm_lut=lut))):
        y, stats, patches = ops.stem_forward(img, wt, pads, return_patches=True, **kw)
        if dy is None:
            dy = _rand(y.shape, 7)
        dwt = torch.empty_like(wt)
        ops.stem_wgrad(img, dy, pads, dwt, patches=patches)
        outs.append((patches, y, stats, dwt))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert torch.equal(a, b)
    assert torch.equal(mm[:, 0], u8.float().amin(dim=(1, 2, 3))) and torch.equal(mm[:, 1], (u8.float().amax(dim=(1, 2, 3)) - u8.float().amin(dim=(1, 2, 3))))
    # fp16 single channel: same kernel path, values rounded to fp16 by the producer
    y16, _ = ops.stem_forward(x1.half(), wt, pads)
    y32, _ = ops.stem_forward(x1.half().float(), wt, pads)
    assert torch.equal(y16, y32)


@pytest.fixture(params=[0, 15], ids=["regs", "cp_async_ring"])
def ew_async(request):
    """Both staging variants of the streaming passes (registers / per-thread cp.async ring)."""
    from mammoclip_b200 import _lib
    old = _lib.lib().mclip_set_ew_async(request.param)
    yield request.param
    _lib.lib().mclip_set_ew_async(old)


def test_ew_staging_variants_are_bit_identical():
    from mammoclip_b200 import _lib, ops
    lib = _lib.lib()
    n, hw, c = 3, 2011, 240
    y, du, res = _rand((n, hw, c), 1), _rand((n, hw, c), 2), _rand((n, hw, c), 3)
    st = ops.BNState(c, "cuda")
    st.scale.uniform_(0.5, 1.5); st.shift.normal_(0, 0.3); st.mean.normal_(0, 0.2); st.invstd.uniform_(0.5, 1.5)
    gate, dpool, rs = torch.rand(n, c, device="cuda"), torch.randn(n, c, device="cuda") * 0.01, torch.rand(n, device="cuda") + 0.5
    c1 = torch.randn(c, device="cuda") * 0.01
    outs = []
    old = lib.mclip_set_ew_async(0)
    try:
        for mask in (0, 15):
            lib.mclip_set_ew_async(mask)
            o, pool = ops.ew_forward(y, bn=st, act=1, rowscale=rs, residual=res, pool=True)
            part = ops.ew_backward(0, y, st, 1, du=du, gate=gate, dpool=dpool)
            dy = ops.ew_backward(1, y, st, 1, du=du, gate=gate, dpool=dpool, c1=c1, c2=c1)
            a2, sep = ops.ew_backward(2, y, st, 1, du=du, gate=gate)
            outs.append((o, pool, part, dy, a2, sep))
    finally:
        lib.mclip_set_ew_async(old)
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,hw,c", [(2, 500, 144), (3, 1392, 1824), (1, 77, 3072), (4, 3000, 24), (2, 999, 1056)])
def test_bn_and_elementwise(n, hw, c, ew_async):
    from mammoclip_b200 import ops
    y = _rand((n, hw, c), 8)
    # statistics -> finalize vs torch batch_norm
    yf = y.float().reshape(-1, c)
    part = torch.stack([yf.sum(0), (yf * yf).sum(0)])[None].contiguous()
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    rm, rv, nb = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda"), torch.zeros((), dtype=torch.long, device="cuda")
    st = ops.bn_finalize(part, n * hw, gamma, beta, rm, rv, nb, True)
    rm2, rv2 = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    ref = F.batch_norm(yf, rm2, rv2, gamma, beta, True, 0.01, 1e-3)
    assert rel_err(yf * st.scale + st.shift, ref) < 1e-4
    assert rel_err(rm, rm2) < 1e-4 and rel_err(rv, rv2) < 1e-4 and nb.item() == 1
    # forward pass: swish(bn(y)) * rowscale + residual, pooling partials
    res, rs = _rand((n, hw, c), 9), torch.rand(n, device="cuda") + 0.5
    out, pool = ops.ew_forward(y, bn=st, act=1, rowscale=rs, residual=res, pool=True)
    v = y.float() * st.scale + st.shift
    refo = v * torch.sigmoid(v) * rs.view(n, 1, 1) + res.float()
    assert rel_err(out.float(), refo) < 8e-3
    assert rel_err(pool.sum(1), out.float().sum(1)) < 1e-4
    # backward passes vs autograd of the same expression
    du = _rand((n, hw, c), 10)
    gate, dpool = torch.rand(n, c, device="cuda"), torch.randn(n, c, device="cuda") * 0.01
    yv = y.float().clone().requires_grad_(True)
    g2, b2 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.batch_norm(yv.reshape(-1, c), None, None, g2, b2, True, 0.0, 1e-3).reshape(n, hw, c)
    u = z * torch.sigmoid(z)
    up = (du.float() * gate.view(n, 1, c) + dpool.view(n, 1, c))
    (u * up).sum().backward()
    part = ops.ew_backward(0, y, st, 1, du=du, gate=gate, dpool=dpool)
    dg, db = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    c1, c2 = ops.bn_bwd_finalize(part, n * hw, True, dg, db)
    assert rel_err(dg, g2.grad) < 5e-3 and rel_err(db, b2.grad) < 5e-3
    dy = ops.ew_backward(1, y, st, 1, du=du, gate=gate, dpool=dpool, c1=c1, c2=c2)
    assert rel_err(dy.float(), yv.grad) < 1e-2
    a2, sep = ops.ew_backward(2, y, st, 1, du=du, gate=gate)
    assert rel_err(a2.float(), (u * gate.view(n, 1, c)).detach()) < 8e-3
    assert sep.shape[2] == 5
    assert rel_err(sep[:, :, 0].sum(1), (du.float() * u.detach()).sum(1)) < 5e-3
    # the four extra sums reproduce the BN-backward reduction without a pass of their own
    bnp = ops.se_bn_combine(sep, gate, dpool)
    assert rel_err(bnp.sum(0), part.sum(0)) < 5e-3


@pytest.mark.parametrize("n,c,cse,hw", [(4, 144, 6, 100), (2, 3072, 128, 50), (3, 48, 12, 64)])
def test_se_fc(n, c, cse, hw):
    from mammoclip_b200 import ops
    part = torch.randn(n, 3, c, device="cuda")
    w1, b1 = torch.randn(cse, c, device="cuda") * 0.1, torch.randn(cse, device="cuda") * 0.1
    w2, b2 = torch.randn(c, cse, device="cuda") * 0.1, torch.randn(c, device="cuda") * 0.1
    pooled, z1, gate = ops.se_fc(part, hw, w1, b1, w2, b2)
    s = (part.sum(1) / hw).requires_grad_(True)
    W1, B1, W2, B2 = (t.clone().requires_grad_(True) for t in (w1, b1, w2, b2))
    z = s @ W1.T + B1
    h = z * torch.sigmoid(z)
    gr = torch.sigmoid(h @ W2.T + B2)
    assert rel_err(gate, gr) < 1e-4 and rel_err(pooled, s) < 1e-5
    dgp = torch.randn(n, 2, c, device="cuda")
    gr.backward(dgp.sum(1))
    dw1, db1, dw2, db2 = (torch.empty_like(t) for t in (w1, b1, w2, b2))
    dpool = ops.se_fc_backward(dgp, hw, w1, w2, pooled, z1, gate, dw1, db1, dw2, db2)
    assert rel_err(dpool, s.grad / hw) < 1e-3
    for a, b in ((dw1, W1.grad), (db1, B1.grad), (dw2, W2.grad), (db2, B2.grad)):
        assert rel_err(a, b) < 1e-3
    wp = torch.randn(40, c, device="cuda")
    wg = ops.se_scale_weights(wp, gate)
    assert rel_err(wg.float(), wp[None] * gate[:, None, :]) < 5e-3
