"""Thin Python wrappers over the C ABI: argument marshalling only, all arithmetic happens in the CUDA library."""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, LossArgs, WgradArgs, call, check, lib, ptr, stream_ptr


def _require_cuda(*tensors):
    """Tensors must live on the CURRENT CUDA device: kernels launch on that device's current stream (a model moved
    to cuda:1 without torch.cuda.set_device(1) would otherwise launch on device 0 with device-1 pointers)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.MclipError("mammoclip_b200 kernels need CUDA tensors on a B200; there is no CPU fallback")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.MclipError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: call torch.cuda.set_device({t.device.index}) "
                                  "(kernels launch on the current device's stream)")


# ------------------------------------------------------------------------------------------------ GEMMs

def gemm_tn(a, b, out=None, bias=None, residual=None, act=0, want_stats=None, dropmask=None, drop_scale=1.0, aux_pre=None, a2=None):
    """out[b,m,n] = epi(sum_k a[b,m,k] * w[b|0,n,k]).  a: [M,K] or [Bt,M,K] bf16; b: [N,K] or [Bt,N,K] bf16.
    want_stats=None: returns out (bf16).  want_stats=True/False: returns (out, stats[slots,2,N] fp32 or None)."""
    _require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    a3 = a if a.dim() == 3 else a.unsqueeze(0)
    bt, m, k = a3.shape
    b_batched = b.dim() == 3
    n = b.shape[-2]
    if a2 is not None:       # A = [a | a2] along K, b = [n, ceil64(k) + k2]
        a23 = a2 if a2.dim() == 3 else a2.unsqueeze(0)
        assert a23.shape[:2] == a3.shape[:2] and a23.dtype == torch.bfloat16 and a23.stride(2) == 1
        assert b.shape[-1] == (k + 63) // 64 * 64 + a23.shape[2] and b.stride(-1) == 1
    else:
        assert b.shape[-1] == k and b.stride(-1) == 1
    assert a3.stride(2) == 1
    if out is None:
        out = torch.empty((bt, m, n) if a.dim() == 3 else (m, n), dtype=torch.bfloat16, device=a.device)
    o3 = out if out.dim() == 3 else out.unsqueeze(0)
    g = GemmArgs()
    g.a, g.lda, g.a_batch_stride = a3.data_ptr(), a3.stride(1), a3.stride(0) if bt > 1 else 0
    g.b, g.ldb = b.data_ptr(), b.stride(-2)
    g.b_batch_stride = b.stride(0) if b_batched else 0
    g.d, g.ldd, g.d_batch_stride = o3.data_ptr(), o3.stride(1), o3.stride(0) if bt > 1 else 0
    g.m, g.n, g.k, g.batches = m, n, k, bt
    if a2 is not None:
        g.a2, g.lda2, g.a2_batch_stride, g.k2 = a23.data_ptr(), a23.stride(1), a23.stride(0) if bt > 1 else 0, a23.shape[2]
    g.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        r3 = residual if residual.dim() == 3 else residual.unsqueeze(0)
        g.residual, g.ldr, g.r_batch_stride = r3.data_ptr(), r3.stride(1), r3.stride(0) if bt > 1 else 0
    g.act = act
    if dropmask is not None:
        assert dropmask.dtype == torch.uint8 and dropmask.is_contiguous() and dropmask.numel() == bt * m * n
        g.dropmask, g.drop_scale = dropmask.data_ptr(), drop_scale
    if aux_pre is not None:       # bf16 [bt*m, n]: receives the pre-activation value
        assert aux_pre.dtype == torch.bfloat16 and aux_pre.is_contiguous() and aux_pre.numel() == bt * m * n
        g.aux_pre, g.ld_aux = aux_pre.data_ptr(), n
    stats = None
    if want_stats:
        slots = lib().mclip_gemm_tn_stat_slots(m, n, bt)
        stats = torch.empty((slots, 2, n), dtype=torch.float32, device=a.device)
        g.stats, g.stat_slots = stats.data_ptr(), slots
    call("mclip_gemm_tn", C.byref(g), nbytes=2 * (bt * m * k + n * k * (bt if b_batched else 1) + bt * m * n * (2 if residual is not None else 1)))
    return out if want_stats is None else (out, stats)


_wgrad_ws = {}


def gemm_wgrad(a, b, out=None, accumulate=False):
    """out[i,j] (+)= sum_r a[r,i] * b[r,j]; a: [R,I] bf16, b: [R,J] bf16, out fp32 [I,J]."""
    _require_cuda(a, b)
    r, i = a.shape
    j = b.shape[1]
    assert b.shape[0] == r and a.stride(1) == 1 and b.stride(1) == 1
    if out is None:
        out = torch.empty((i, j), dtype=torch.float32, device=a.device)
        accumulate = False
    need = lib().mclip_gemm_wgrad_workspace_bytes(r, i, j)
    key = a.device.index
    ws = _wgrad_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=a.device)
        _wgrad_ws[key] = ws
    g = WgradArgs()
    g.a, g.lda, g.b, g.ldb = a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0)
    g.out, g.ldo, g.accumulate = out.data_ptr(), out.stride(0), int(accumulate)
    g.r, g.i, g.j = r, i, j
    g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    call("mclip_gemm_wgrad", C.byref(g), nbytes=2 * r * (i + j) + 4 * i * j)
    return out


def bn0_fold_backward(dv0, x, we, we_bf16, bn, c1, c2, dwe_out, residual=None, sumx=None):
    """Folded BatchNorm backward of the expand convolution (include/mclip.h, mclip_bn0_fold): from dV0 [M,Cexp] (gradient
    w.r.t. the BN output, swish' applied), X [M,Cin] (bf16), the conv weight We (fp32 [Cexp,Cin] + its bf16 copy) and the BN
    backward means (c1, c2) -> dX [M,Cin] bf16 (+ residual) and dWe (written to dwe_out) WITHOUT materialising dY0."""
    from ._lib import Bn0FoldArgs
    m, cexp = dv0.shape
    cin = x.shape[1]
    dev = dv0.device
    k1pad = (cexp + 63) // 64 * 64
    ldw = k1pad + cin
    wcat = torch.empty((cin, ldw), dtype=torch.bfloat16, device=dev)
    twe = torch.empty((cexp, cin), dtype=torch.bfloat16, device=dev)
    bias = torch.empty(cin, dtype=torch.float32, device=dev)
    if sumx is None:                    # normally handed over by the forward (pool partials of the pass that wrote X)
        sumx = torch.empty(cin, dtype=torch.float32, device=dev)
        colsum(x, sumx)
    f = Bn0FoldArgs()
    f.cexp, f.cin, f.k1pad, f.ldw, f.count = cexp, cin, k1pad, ldw, float(m)
    f.we, f.scale, f.invstd, f.c1, f.c2 = we.data_ptr(), bn.scale.data_ptr(), bn.invstd.data_ptr(), c1.data_ptr(), c2.data_ptr()
    f.wcat, f.twe, f.bias, f.sumx = wcat.data_ptr(), twe.data_ptr(), bias.data_ptr(), sumx.data_ptr()
    call("mclip_bn0_fold", C.byref(f), 0)
    g = gemm_wgrad(twe, we_bf16)                               # G = (T We)^T We  [cin, cin]
    f.g = g.data_ptr()
    call("mclip_bn0_fold", C.byref(f), 1)
    dx = gemm_tn(dv0, wcat, a2=x, bias=bias, residual=residual)
    gemm_wgrad(dv0, x, out=dwe_out)                            # dV0^T X
    xtx = gemm_wgrad(x, x)                                     # X^T X
    gc = torch.empty((cin, cin), dtype=torch.bfloat16, device=dev)
    f.g, f.gc = xtx.data_ptr(), gc.data_ptr()
    call("mclip_bn0_fold", C.byref(f), 2)
    q = gemm_tn(we_bf16, gc)                                   # We (X^T X - sum sum^T / M)  [cexp, cin]
    f.dwe, f.q = dwe_out.data_ptr(), q.data_ptr()
    call("mclip_bn0_fold", C.byref(f), 3)
    return dx


# ------------------------------------------------------------------------------------------------ loss

_loss_ws = {}
_loss_status = {}     # device index -> (device int32 status word, pinned host mirror)


def _loss_status_words(dev):
    """Status word of the fused loss kernel (peer time-out) and its pinned host mirror.  The mirror is refreshed by an async
    copy after every call and inspected before the next one: a rank whose peer never arrived raises here instead of the
    kernel trapping the CUDA context."""
    st = _loss_status.get(dev.index)
    if st is None:
        st = (torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32).pin_memory())
        _loss_status[dev.index] = st
    code = int(st[1][0])
    if code != 0:
        st[0].zero_(); st[1].zero_()
        raise _lib.MclipError(f"fused contrastive loss: the embeddings of rank {code - 1} never arrived (MCLIP_PEER_TIMEOUT_S); "
                              "the loss of that step is NaN")
    return st


def contrastive_loss_raw(local, pairs, scale, world=1, rank=0, symm=None):
    """local: list of [B,D] fp32 CUDA tensors; pairs: list of (a, b, w_row, w_col, eps).
    Returns (out[2+2P] fp32 device tensor, grads list).  `symm` carries the peer-mapped buffers when world > 1."""
    _require_cuda(*local)
    B, D = local[0].shape
    dev = local[0].device
    P, K = len(pairs), len(local)
    args = LossArgs()
    args.world, args.rank, args.batch, args.dim, args.n_tensors, args.n_pairs = world, rank, B, D, K, P
    grads = []
    for k, t in enumerate(local):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.shape == (B, D)
        g = torch.empty_like(t)
        grads.append(g)
        args.local[k], args.grad[k] = t.data_ptr(), g.data_ptr()
    for i, (a, b, wr, wc, eps) in enumerate(pairs):
        args.pair_a[i], args.pair_b[i], args.w_row[i], args.w_col[i], args.label_smoothing[i] = a, b, wr, wc, eps
    if torch.is_tensor(scale):
        assert scale.is_cuda and scale.dtype == torch.float32 and scale.numel() == 1
        args.logit_scale_dev = scale.data_ptr()
    else:
        args.logit_scale = float(scale)
    need = lib().mclip_loss_workspace_bytes(world, B, D, P)
    key = (dev.index, need)
    ws = _loss_ws.get(key)
    if ws is None:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _loss_ws[key] = ws
    args.workspace, args.workspace_bytes = ws.data_ptr(), need
    out = torch.empty(2 + 2 * P, dtype=torch.float32, device=dev)
    args.out = out.data_ptr()
    if world > 1:
        symm.fill_args(args, K)
        st_dev, st_host = _loss_status_words(dev)
        args.status = st_dev.data_ptr()
    call("mclip_contrastive_loss", C.byref(args))
    if world > 1:
        st_host.copy_(st_dev, non_blocking=True)
    return out, grads


# ------------------------------------------------------------------------------------------------ conv / BN / SE passes
from ._lib import BnArgs, DwconvArgs, EwArgs, EwBwdArgs, PrepEntry, SeArgs, StemArgs  # noqa: E402


def _p(t):
    return None if t is None else t.data_ptr()


class BNState:
    """Affine (scale, shift) that consumers apply on load + batch statistics kept for the backward pass."""
    __slots__ = ("scale", "shift", "mean", "invstd", "count", "training")

    def __init__(self, c, device):
        buf = torch.empty(4, c, dtype=torch.float32, device=device)
        self.scale, self.shift, self.mean, self.invstd = buf[0], buf[1], buf[2], buf[3]


def bn_finalize(partials, count, gamma, beta, running_mean, running_var, num_batches, training, momentum=0.01, eps=1e-3):
    c = gamma.numel()
    st = BNState(c, gamma.device)
    st.count, st.training = count, training
    a = BnArgs()
    a.c, a.training, a.count = c, int(training), count
    if training:
        a.slots, a.partials = partials.shape[0], partials.data_ptr()
    a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
    a.running_mean, a.running_var = _p(running_mean), _p(running_var)
    a.num_batches_tracked = _p(num_batches)
    a.momentum, a.eps = momentum, eps
    a.scale, a.shift, a.mean, a.invstd = st.scale.data_ptr(), st.shift.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr()
    call("mclip_bn_finalize", C.byref(a))
    return st


_STEM_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.uint8: 3}


def image_norm_lut_u8(images, mean, std):
    """uint8 [N,1,H,W] (contiguous) -> (fp32 [N,2] per-image (min, max - min), bf16 [N,256] normalisation table):
    lut[n,u] = bf16(((u - min) / (max - min) - mean) / std), the fp32 arithmetic of datasets/imagetext.py:129-134 per grey level."""
    _require_cuda(images)
    assert images.dtype == torch.uint8 and images.is_contiguous()
    n = images.shape[0]
    mm = torch.empty((n, 2), dtype=torch.float32, device=images.device)
    lut = torch.empty((n, 256), dtype=torch.bfloat16, device=images.device)
    call("mclip_image_norm_lut_u8", ptr(images), n, C.c_longlong(images[0].numel()), C.c_float(mean), C.c_float(std), ptr(mm), ptr(lut))
    return mm, lut


def stem_im2col(images, pads, norm_lut=None):
    """images: [N,3,H,W] fp32 (any strides), or the input edge [N,1,H,W] fp32/fp16/bf16/uint8 (one grey channel, written to
    the three tap groups: bit-identical to three identical channels) -> (patches [N*Ho*Wo, 32] bf16, Ho, Wo);
    taps t = ci*9 + ky*3 + kx, zero padded.  norm_lut (bf16 [N,256] from image_norm_lut_u8) applies imagetext.py:129-134 on load."""
    _require_cuda(images)
    n, ci, h, w = images.shape
    assert (ci == 3 and images.dtype == torch.float32) or (ci == 1 and images.dtype in _STEM_DTYPES), (ci, images.dtype)
    pl, pr, pt, pb = pads
    ho, wo = (h + pt + pb - 3) // 2 + 1, (w + pl + pr - 3) // 2 + 1
    out = torch.empty((n * ho * wo, 32), dtype=torch.bfloat16, device=images.device)
    a = StemArgs()
    a.n, a.h, a.w, a.ho, a.wo = n, h, w, ho, wo
    a.pad_left, a.pad_right, a.pad_top, a.pad_bottom = pl, pr, pt, pb
    a.in_ = images.data_ptr()
    a.stride_n, a.stride_c, a.stride_h, a.stride_w = images.stride()
    a.out = out.data_ptr()
    a.in_channels, a.in_dtype = ci, _STEM_DTYPES[images.dtype]
    if norm_lut is not None:
        assert images.dtype == torch.uint8 and norm_lut.shape == (n, 256) and norm_lut.dtype == torch.bfloat16 and norm_lut.is_contiguous()
        a.norm_lut = norm_lut.data_ptr()
    call("mclip_stem_im2col", C.byref(a), nbytes=images.element_size() * images.numel() + 2 * out.numel())
    return out, ho, wo


def stem_weight_bf16(weight):
    """[C,3,3,3] fp32 -> [C,32] bf16 (K padded with zeros)."""
    c = weight.shape[0]
    w = torch.zeros((c, 32), dtype=torch.bfloat16, device=weight.device)
    table = weight_prep([(weight.detach().view(c, 27), w, None)], weight.device, dst_ld=32)
    weight_prep_run(table, 1)
    return w


def stem_forward(images, weight, pads, want_stats=True, w_bf16=None, return_patches=False, norm_lut=None):
    """Stem conv = im2col + tcgen05 GEMM.  -> (Y [N,Ho,Wo,C] bf16, stats) (+ patches for the weight gradient)."""
    patches, ho, wo = stem_im2col(images, pads, norm_lut=norm_lut)
    wb = w_bf16 if w_bf16 is not None else stem_weight_bf16(weight)
    y, stats = gemm_tn(patches, wb, want_stats=bool(want_stats))
    y = y.view(images.shape[0], ho, wo, weight.shape[0])
    return (y, stats, patches) if return_patches else (y, stats)


def stem_wgrad(images, dy, pads, dweight, patches=None):
    """dW[c, ci,ky,kx] = sum_pixels dY[pix,c] * patch[pix,t] on the tcgen05 wgrad GEMM."""
    if patches is None:
        patches, _, _ = stem_im2col(images, pads)
    c = dy.shape[-1]
    tmp = gemm_wgrad(dy.reshape(-1, c), patches)            # [C, 32]
    dweight.view(c, 27).copy_(tmp[:, :27])
    return dweight


def _dw_args(x, weight, k, stride, pads, bn):
    n, h, w, c = x.shape
    pl, pr, pt, pb = pads
    a = DwconvArgs()
    a.n, a.h, a.w, a.c, a.k, a.stride = n, h, w, c, k, stride
    a.ho, a.wo = (h + pt + pb - k) // stride + 1, (w + pl + pr - k) // stride + 1
    a.pad_left, a.pad_right, a.pad_top, a.pad_bottom = pl, pr, pt, pb
    a.in_ = x.data_ptr()
    if bn is not None:
        a.in_scale, a.in_shift, a.in_act = bn.scale.data_ptr(), bn.shift.data_ptr(), 1
    a.weight = weight.data_ptr()
    return a


def dwconv_forward(x, weight, k, stride, pads, bn=None, want_stats=True):
    """x: [N,H,W,C] bf16 (pre-BN output of the producer when `bn` is given: swish(bn(x)) is applied on load)."""
    _require_cuda(x, weight)
    a = _dw_args(x, weight, k, stride, pads, bn)
    out = torch.empty((a.n, a.ho, a.wo, a.c), dtype=torch.bfloat16, device=x.device)
    a.out = out.data_ptr()
    stats = None
    if want_stats:
        slots = lib().mclip_dwconv_slots(C.byref(a), 0)
        stats = torch.empty((slots, 2, a.c), dtype=torch.float32, device=x.device)
        a.stats, a.stat_slots = stats.data_ptr(), slots
    call("mclip_dwconv_forward", C.byref(a), nbytes=2 * (x.numel() + out.numel()))
    return out, stats


def dwconv_backward(x, weight, k, stride, pads, dy, dweight, bn=None):
    """Returns (dx, bn_partials).  dx = d/d(pre-activation x) when `bn` is given (swish' applied), else d/dx."""
    a = _dw_args(x, weight, k, stride, pads, bn)
    dx = torch.empty_like(x)
    slots = lib().mclip_dwconv_slots(C.byref(a), 1)
    dwp = torch.empty((slots, k * k, a.c), dtype=torch.float32, device=x.device)
    a.stat_slots, a.dy, a.dx, a.dweight, a.accumulate, a.dw_partials = slots, dy.data_ptr(), dx.data_ptr(), dweight.data_ptr(), 0, dwp.data_ptr()
    bnp = None
    if bn is not None:
        bnp = torch.empty((slots, 2, a.c), dtype=torch.float32, device=x.device)
        a.bn_partials, a.in_mean, a.in_invstd = bnp.data_ptr(), bn.mean.data_ptr(), bn.invstd.data_ptr()
    call("mclip_dwconv_backward", C.byref(a), nbytes=2 * (2 * x.numel() + dy.numel()))
    return dx, bnp


def ew_chunks(n, hw, c):
    r = lib().mclip_ew_chunks(n, hw, c)
    if r < 1:
        raise _lib.MclipError(f"mclip_ew_chunks({n},{hw},{c}) failed")
    return r


def ew_forward(y, bn=None, act=0, rowscale=None, residual=None, write=True, pool=False):
    """y: [N,HW,C] bf16.  Returns (out or None, pool_partials or None)."""
    n, hw, c = y.shape
    a = EwArgs()
    a.n, a.hw, a.c, a.act = n, hw, c, act
    a.y = y.data_ptr()
    if bn is not None:
        a.scale, a.shift = bn.scale.data_ptr(), bn.shift.data_ptr()
    a.rowscale, a.residual = _p(rowscale), _p(residual)
    out = torch.empty_like(y) if write else None
    a.out = _p(out)
    part = None
    if pool:
        a.chunks = ew_chunks(n, hw, c)
        part = torch.empty((n, a.chunks, c), dtype=torch.float32, device=y.device)
        a.pool_partials = part.data_ptr()
    call("mclip_ew_forward", C.byref(a), nbytes=2 * y.numel() * (1 + int(write) + int(residual is not None)))
    return out, part


def pool_finalize(part, hw, mult=None):
    n, chunks, c = part.shape
    out = torch.empty((n, c), dtype=torch.float32, device=part.device)
    call("mclip_pool_finalize", ptr(part), n, chunks, c, hw, ptr(mult), ptr(out))
    return out


def se_fc(part, hw, w1, b1, w2, b2):
    n, chunks, c = part.shape
    cse = w1.shape[0]
    dev = part.device
    pooled = torch.empty((n, c), dtype=torch.float32, device=dev)
    z1 = torch.empty((n, cse), dtype=torch.float32, device=dev)
    gate = torch.empty((n, c), dtype=torch.float32, device=dev)
    a = SeArgs()
    a.n, a.hw, a.c, a.cse, a.chunks = n, hw, c, cse, chunks
    a.pool_partials, a.w1, a.b1, a.w2, a.b2 = part.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
    a.pooled, a.z1, a.gate = pooled.data_ptr(), z1.data_ptr(), gate.data_ptr()
    call("mclip_se_fc", C.byref(a))
    return pooled, z1, gate


def se_fc_backward(dg_part, hw, w1, w2, pooled, z1, gate, dw1, db1, dw2, db2):
    """dg_part: [N,chunks,C] or the SE pass-1 partials [N,chunks,5,C] (component 0 = d gate sums)."""
    n, chunks, c = dg_part.shape[0], dg_part.shape[1], dg_part.shape[-1]
    cse = w1.shape[0]
    dev = dg_part.device
    dz2 = torch.empty((n, c), dtype=torch.float32, device=dev)
    dz1 = torch.empty((n, cse), dtype=torch.float32, device=dev)
    dpool = torch.empty((n, c), dtype=torch.float32, device=dev)
    a = SeArgs()
    a.n, a.hw, a.c, a.cse, a.chunks, a.accumulate = n, hw, c, cse, chunks, 0
    a.w1, a.w2, a.pooled, a.z1, a.gate = w1.data_ptr(), w2.data_ptr(), pooled.data_ptr(), z1.data_ptr(), gate.data_ptr()
    a.dgate_partials, a.dz2, a.dz1, a.dpool = dg_part.data_ptr(), dz2.data_ptr(), dz1.data_ptr(), dpool.data_ptr()
    a.dgate_chunk_stride = dg_part.stride(1)
    a.dw1, a.db1, a.dw2, a.db2 = dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr()
    call("mclip_se_fc_backward", C.byref(a))
    return dpool


def se_scale_weights(w, gate):
    """w: [Cout,Cexp] fp32, gate [N,Cexp] fp32 -> [N,Cout,Cexp] bf16."""
    cout, cexp = w.shape[0], w.shape[1]
    n = gate.shape[0]
    out = torch.empty((n, cout, cexp), dtype=torch.bfloat16, device=w.device)
    call("mclip_se_scale_weights", ptr(w), ptr(gate), ptr(out), n, cout, cexp)
    return out


def ew_backward(mode, y, bn, act, du=None, dvec=None, gate=None, dpool=None, rowscale=None, dv_given=False, c1=None, c2=None):
    """mode 0 -> partials [N*chunks,2,C]; mode 1 -> dY bf16; mode 2 -> (A2 bf16, dgate partials [N,chunks,C])."""
    n, hw, c = y.shape
    a = EwBwdArgs()
    a.n, a.hw, a.c, a.act, a.mode, a.dv_given = n, hw, c, act, mode, int(dv_given)
    a.y = y.data_ptr()
    if bn is not None:
        a.scale, a.shift, a.mean, a.invstd = bn.scale.data_ptr(), bn.shift.data_ptr(), bn.mean.data_ptr(), bn.invstd.data_ptr()
    a.du, a.dvec, a.gate, a.dpool, a.rowscale, a.c1, a.c2 = _p(du), _p(dvec), _p(gate), _p(dpool), _p(rowscale), _p(c1), _p(c2)
    part = out = None
    if mode != 1:
        a.chunks = ew_chunks(n, hw, c)
        shape = (n * a.chunks, 2, c) if mode == 0 else (n, a.chunks, 5, c)
        part = torch.empty(shape, dtype=torch.float32, device=y.device)
        a.partials = part.data_ptr()
    if mode != 0:
        out = torch.empty_like(y)
        a.out = out.data_ptr()
    call("mclip_ew_backward", C.byref(a), nbytes=2 * y.numel() * (1 + int(du is not None) + int(mode != 0)))
    return part if mode == 0 else out if mode == 1 else (out, part)


def se_bn_combine(se_part, gate, dpool):
    """SE pass-1 partials [N,chunks,5,C] + gate/dpool [N,C] -> BN-backward partials [N,2,C]."""
    n, chunks, _, c = se_part.shape
    out = torch.empty((n, 2, c), dtype=torch.float32, device=se_part.device)
    call("mclip_se_bn_combine", ptr(se_part), n, chunks, c, ptr(gate), ptr(dpool), ptr(out))
    return out


def bn_bwd_finalize(partials, count, training, dgamma, dbeta):
    slots, _, c = partials.shape
    cc = torch.empty((2, c), dtype=torch.float32, device=partials.device)
    call("mclip_bn_bwd_finalize", ptr(partials), slots, c, C.c_longlong(count), int(training), ptr(dgamma), ptr(dbeta), 0,
                                      ptr(cc[0]), ptr(cc[1]))
    return cc[0], cc[1]


def weight_prep(entries, device, dst_ld=0):
    """entries: list of (src fp32 2-D, dst bf16 or None, dst_t bf16 or None[, dst_ld[, dst_t_ld]]). Returns the device table (keep it alive)."""
    import numpy as np
    arr = (PrepEntry * len(entries))()
    for i, e in enumerate(entries):
        src, dst, dst_t = e[0], e[1], e[2]
        arr[i].src, arr[i].dst, arr[i].dst_t = src.data_ptr(), _p(dst), _p(dst_t)
        arr[i].rows, arr[i].cols = src.shape[0], src[0].numel()
        arr[i].dst_ld = e[3] if len(e) > 3 else dst_ld
        arr[i].dst_t_ld = e[4] if len(e) > 4 else 0
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)


def weight_prep_run(table, n_entries):
    call("mclip_weight_prep", ptr(table), n_entries)


def cast_bf16(x):
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    call("mclip_cast_bf16", ptr(x), ptr(out), C.c_longlong(x.numel()))
    return out


def l2norm_forward(x):
    rows, d = x.shape
    e = torch.empty((rows, d), dtype=torch.float32, device=x.device)
    nrm = torch.empty((rows,), dtype=torch.float32, device=x.device)
    call("mclip_l2norm_forward", ptr(x), ptr(e), ptr(nrm), rows, d)
    return e, nrm


def l2norm_backward(e, de, nrm):
    rows, d = e.shape
    dx = torch.empty((rows, d), dtype=torch.bfloat16, device=e.device)
    call("mclip_l2norm_backward", ptr(e), ptr(de), ptr(nrm), ptr(dx), rows, d)
    return dx


_colsum_ws = {}


def colsum(x, out, accumulate=False):
    rows, cols = x.shape
    assert x.stride(1) == 1
    need = lib().mclip_colsum_workspace_bytes(rows, cols)
    ws = _colsum_ws.get(x.device.index)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=x.device)
        _colsum_ws[x.device.index] = ws
    call("mclip_colsum", ptr(x), ptr(out), rows, cols, C.c_longlong(x.stride(0)), int(accumulate), ptr(ws), C.c_longlong(ws.numel()))
    return out


def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    call("mclip_adamw_step", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), C.c_longlong(param.numel()), C.c_float(lr), C.c_float(beta1),
                                 C.c_float(beta2), C.c_float(eps), C.c_float(weight_decay), C.c_longlong(step), C.c_float(grad_scale))


# ------------------------------------------------------------------------------------------------ BERT pieces
from ._lib import BertEmbedArgs  # noqa: E402


def bert_embed_ln(ids, tts, word, pos, typ, gamma, beta, eps, dropmask=None, drop_scale=1.0):
    b, l = ids.shape
    h = word.shape[1]
    out = torch.empty((b * l, h), dtype=torch.bfloat16, device=ids.device)
    a = BertEmbedArgs()
    a.batch, a.seq_len, a.hidden, a.vocab, a.max_positions = b, l, h, word.shape[0], pos.shape[0]
    a.input_ids, a.token_type_ids = ids.data_ptr(), _p(tts)
    a.word, a.pos, a.type, a.gamma, a.beta, a.eps = word.data_ptr(), pos.data_ptr(), typ.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps
    a.dropmask, a.drop_scale, a.out = _p(dropmask), drop_scale, out.data_ptr()
    call("mclip_bert_embed_ln", C.byref(a))
    return out


def layernorm(x, gamma, beta, eps):
    rows, h = x.shape
    out = torch.empty_like(x)
    call("mclip_layernorm", ptr(x), ptr(gamma), ptr(beta), C.c_float(eps), ptr(out), rows, h)
    return out


def bert_attention(qkv, attention_mask, batch, seq_len, heads, head_dim, dropmask=None, drop_scale=1.0, want_lse=False):
    out = torch.empty((batch * seq_len, heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((batch, heads, seq_len), dtype=torch.float32, device=qkv.device) if want_lse else None
    call("mclip_bert_attention", ptr(qkv), ptr(attention_mask), ptr(dropmask), C.c_float(drop_scale), ptr(out), ptr(lse), batch, seq_len, heads, head_dim)
    return (out, lse) if want_lse else out


# ------------------------------------------------------------------------------------------------ BERT backward pieces
from ._lib import BertEmbedBwdArgs  # noqa: E402


def layernorm_backward(x, dy, gamma, eps, dgamma, dbeta, dropmask=None, drop_scale=1.0, accumulate=False):
    """x (pre-LN input), dy: [rows,H] bf16.  Returns (dx, dx_drop); dx_drop is dx when no mask is given.  dgamma/dbeta are written (or +=)."""
    rows, h = x.shape
    assert x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape and dy.dtype == torch.bfloat16
    dx = torch.empty_like(x)
    dxd = torch.empty_like(x) if dropmask is not None else None
    slots = lib().mclip_layernorm_backward_slots(rows)
    part = torch.empty((slots, 2, h), dtype=torch.float32, device=x.device)
    call("mclip_layernorm_backward", ptr(x), ptr(dy), ptr(gamma), C.c_float(eps), ptr(dropmask), C.c_float(drop_scale), ptr(dx), ptr(dxd), ptr(part), slots,
         ptr(dgamma), ptr(dbeta), int(accumulate), rows, h)
    return dx, (dxd if dxd is not None else dx)


def gelu_forward(x):
    out = torch.empty_like(x)
    call("mclip_gelu_forward", ptr(x), ptr(out), C.c_longlong(x.numel()))
    return out


def gelu_backward(dy, x):
    out = torch.empty_like(x)
    call("mclip_gelu_backward", ptr(dy), ptr(x), ptr(out), C.c_longlong(x.numel()))
    return out


def bert_attention_backward(qkv, d_out, lse, attention_mask, batch, seq_len, heads, head_dim, dropmask=None, drop_scale=1.0):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    call("mclip_bert_attention_backward", ptr(qkv), ptr(d_out), ptr(lse), ptr(attention_mask), ptr(dropmask), C.c_float(drop_scale), ptr(delta), ptr(dqkv),
         batch, seq_len, heads, head_dim)
    return dqkv


def bert_embed_backward(ids, tts, word, pos, typ, gamma, eps, dout, dword, dpos, dtype, dgamma, dbeta, dropmask=None, drop_scale=1.0, accumulate=False):
    """dout: [B*L,H] bf16.  Writes (accumulate=False: touched rows only, the tables must be zeroed) or adds the embedding-table gradients."""
    b, l = ids.shape
    h = word.shape[1]
    a = BertEmbedBwdArgs()
    a.batch, a.seq_len, a.hidden, a.vocab, a.max_positions, a.n_types = b, l, h, word.shape[0], pos.shape[0], typ.shape[0]
    a.slots, a.accumulate = lib().mclip_layernorm_backward_slots(b * l), int(accumulate)
    dv = torch.empty((b * l, h), dtype=torch.float32, device=ids.device)
    part = torch.empty((a.slots, 2, h), dtype=torch.float32, device=ids.device)
    a.input_ids, a.token_type_ids = ids.data_ptr(), _p(tts)
    a.word, a.pos, a.type, a.gamma, a.eps = word.data_ptr(), pos.data_ptr(), typ.data_ptr(), gamma.data_ptr(), eps
    a.dropmask, a.drop_scale, a.dout = _p(dropmask), drop_scale, dout.data_ptr()
    a.dv, a.partials = dv.data_ptr(), part.data_ptr()
    a.dword, a.dpos, a.dtype, a.dgamma, a.dbeta = dword.data_ptr(), dpos.data_ptr(), dtype.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr()
    call("mclip_bert_embed_backward", C.byref(a))
