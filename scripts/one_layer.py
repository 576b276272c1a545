"""Run one kernel class on one EN-B5 layer shape a few times (target for ncu)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammoclip_b200 import ops
from mammoclip_b200.model.modules.efficientnet_custom import net_geometry
from layer_bench import bn_state

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="dw_bwd")
ap.add_argument("--block", type=int, default=4)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
g = net_geometry("efficientnet-b5")
pl, pr, pt, pb = g.stem_pads
h, w = (1520 + pt + pb - 3) // 2 + 1, (912 + pl + pr - 3) // 2 + 1
for i, b in enumerate(g.blocks):
    l, r, t, bb = b.pads
    ho, wo = (h + t + bb - b.k) // b.s + 1, (w + l + r - b.k) // b.s + 1
    if i == a.block:
        break
    h, w = ho, wo
n = a.batch
y0 = torch.randn(n, h, w, b.cexp, device="cuda").to(torch.bfloat16)
bn0 = bn_state(b.cexp)
wdw = torch.randn(b.cexp, 1, b.k, b.k, device="cuda") * 0.2
y1, _ = ops.dwconv_forward(y0, wdw, b.k, b.s, b.pads, bn=bn0)
y1v = y1.view(n, ho * wo, b.cexp)
dy1 = torch.randn_like(y1)
dwg = torch.empty_like(wdw)
gate = torch.rand(n, b.cexp, device="cuda")
c1 = torch.zeros(b.cexp, device="cuda")
xe = torch.randn(n * h * w, b.cin, device="cuda").to(torch.bfloat16)
we = (torch.randn(b.cexp, b.cin, device="cuda") * 0.1).to(torch.bfloat16)
wg = (torch.randn(n, b.cout, b.cexp, device="cuda") * 0.1).to(torch.bfloat16)
dy2 = torch.randn(n * ho * wo, b.cout, device="cuda").to(torch.bfloat16)
wpt = (torch.randn(b.cexp, b.cout, device="cuda") * 0.1).to(torch.bfloat16)
fns = {
    "dw_fwd": lambda: ops.dwconv_forward(y0, wdw, b.k, b.s, b.pads, bn=bn0),
    "dw_bwd": lambda: ops.dwconv_backward(y0, wdw, b.k, b.s, b.pads, dy1, dwg, bn=bn0),
    "ew_fwd": lambda: ops.ew_forward(y1v, bn=bn0, act=1, pool=True),
    "ew_red": lambda: ops.ew_backward(0, y1v, bn0, 1, du=dy1.view_as(y1v), gate=gate, dpool=gate),
    "ew_apply": lambda: ops.ew_backward(1, y1v, bn0, 1, du=dy1.view_as(y1v), gate=gate, dpool=gate, c1=c1, c2=c1),
    "expand": lambda: ops.gemm_tn(xe, we, want_stats=True),
    "project": lambda: ops.gemm_tn(y1v, wg, want_stats=True),
    "proj_dgrad": lambda: ops.gemm_tn(dy2, wpt),
    "proj_wgrad": lambda: ops.gemm_wgrad(dy2, y1v.view(n * ho * wo, b.cexp)),
}
for _ in range(a.reps):
    fns[a.op]()
torch.cuda.synchronize()
print("done", a.op, a.block, (n, h, w, b.cexp), b.k, b.s)
