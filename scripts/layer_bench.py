"""Per-layer kernel timing on the real geometry (EN-B5, B=64, 1520x912 by default): CUDA-event time and algorithmic
GB/s of every kernel class for every distinct layer shape.  Output: a table for profiles/."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mammoclip_b200 import ops
from mammoclip_b200.model.modules.efficientnet_custom import net_geometry


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bn_state(c):
    st = ops.BNState(c, "cuda")
    st.scale.uniform_(0.5, 1.5); st.shift.normal_(0, 0.3); st.mean.normal_(0, 0.2); st.invstd.uniform_(0.5, 1.5)
    st.count, st.training = 1, True
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="efficientnet-b5")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--h", type=int, default=1520)
    ap.add_argument("--w", type=int, default=912)
    ap.add_argument("--only", default="")
    ap.add_argument("--ew-async", type=int, default=-1, help="mclip_set_ew_async mask (-1: library default)")
    args = ap.parse_args()
    if args.ew_async >= 0:
        from mammoclip_b200 import _lib
        _lib.lib().mclip_set_ew_async(args.ew_async)
    g = net_geometry(args.name)
    n = args.batch
    pl, pr, pt, pb = g.stem_pads
    h, w = (args.h + pt + pb - 3) // 2 + 1, (args.w + pl + pr - 3) // 2 + 1
    seen = {}
    rows = []
    for i, b in enumerate(g.blocks):
        l, r, t, bb = b.pads
        ho, wo = (h + t + bb - b.k) // b.s + 1, (w + l + r - b.k) // b.s + 1
        key = (h, w, b.cin, b.cexp, b.cout, b.k, b.s, b.pads)
        if key in seen:
            seen[key].append(i)
            h, w = ho, wo
            continue
        seen[key] = [i]
        rows.append((i, b, h, w, ho, wo, key))
        h, w = ho, wo
    tot = {}
    print(f"{'blk':>4} {'xN':>3} {'shape':>28} {'op':>14} {'ms':>8} {'GB/s':>8} {'ms*N':>8}")
    for i, b, h, w, ho, wo, key in rows:
        mult = len(seen[key])
        res = []
        x_in = torch.randn(n, h, w, b.cexp if not b.expand else b.cin, device="cuda").to(torch.bfloat16)
        y0 = torch.randn(n, h, w, b.cexp, device="cuda").to(torch.bfloat16)
        bn0, bn1 = bn_state(b.cexp), bn_state(b.cexp)
        wdw = torch.randn(b.cexp, 1, b.k, b.k, device="cuda") * 0.2
        gb = lambda *ts: sum(t.numel() * t.element_size() for t in ts) / 1e9
        if b.expand and (not args.only or "gemm" in args.only):
            we = (torch.randn(b.cexp, b.cin, device="cuda") * 0.1).to(torch.bfloat16)
            xe = x_in.view(n * h * w, b.cin)
            ms = timeit(lambda: ops.gemm_tn(xe, we, want_stats=True))
            res.append(("expand_gemm", ms, gb(xe, y0)))
            wet = we.t().contiguous()
            ms = timeit(lambda: ops.gemm_tn(y0.view(n * h * w, b.cexp), wet))
            res.append(("expand_dgrad", ms, gb(xe, y0)))
            ms = timeit(lambda: ops.gemm_wgrad(y0.view(n * h * w, b.cexp), xe))
            res.append(("expand_wgrad", ms, gb(xe, y0)))
        if not args.only or "dw" in args.only:
            y1, _ = ops.dwconv_forward(y0, wdw, b.k, b.s, b.pads, bn=bn0)
            ms = timeit(lambda: ops.dwconv_forward(y0, wdw, b.k, b.s, b.pads, bn=bn0))
            res.append(("dw_fwd", ms, gb(y0, y1)))
            dy1 = torch.randn_like(y1)
            dwg = torch.empty_like(wdw)
            ms = timeit(lambda: ops.dwconv_backward(y0, wdw, b.k, b.s, b.pads, dy1, dwg, bn=bn0))
            res.append(("dw_bwd", ms, gb(y0, y0, dy1)))
        else:
            y1 = torch.randn(n, ho, wo, b.cexp, device="cuda").to(torch.bfloat16)
        y1v = y1.view(n, ho * wo, b.cexp)
        if not args.only or "ew" in args.only:
            ms = timeit(lambda: ops.ew_forward(y1v, bn=bn1, act=1, pool=True))
            res.append(("ew_fwd_pool", ms, 2 * gb(y1v)))
            du = torch.randn_like(y1v)
            gate = torch.rand(n, b.cexp, device="cuda")
            ms = timeit(lambda: ops.ew_backward(0, y1v, bn1, 1, du=du, gate=gate, dpool=gate))
            res.append(("ew_bwd_red", ms, 2 * gb(y1v)))
            c1 = torch.zeros(b.cexp, device="cuda")
            ms = timeit(lambda: ops.ew_backward(1, y1v, bn1, 1, du=du, gate=gate, dpool=gate, c1=c1, c2=c1))
            res.append(("ew_bwd_apply", ms, 3 * gb(y1v)))
            ms = timeit(lambda: ops.ew_backward(2, y1v, bn1, 1, du=du, gate=gate))
            res.append(("ew_bwd_se1", ms, 3 * gb(y1v)))
        if not args.only or "gemm" in args.only:
            wg = (torch.randn(n, b.cout, b.cexp, device="cuda") * 0.1).to(torch.bfloat16)
            ms = timeit(lambda: ops.gemm_tn(y1v, wg, want_stats=True))
            res.append(("project_gemm", ms, gb(y1v) + n * ho * wo * b.cout * 2 / 1e9))
            dy2 = torch.randn(n * ho * wo, b.cout, device="cuda").to(torch.bfloat16)
            wpt = (torch.randn(b.cexp, b.cout, device="cuda") * 0.1).to(torch.bfloat16)
            ms = timeit(lambda: ops.gemm_tn(dy2, wpt))
            res.append(("project_dgrad", ms, gb(y1v, dy2)))
            ms = timeit(lambda: ops.gemm_wgrad(dy2, y1v.view(n * ho * wo, b.cexp)))
            res.append(("project_wgrad", ms, gb(y1v, dy2)))
        shape = f"{h}x{w}x{b.cin}>{b.cexp}>{b.cout} k{b.k}s{b.s}"
        for op, ms, gbytes in res:
            print(f"{i:>4} {mult:>3} {shape:>28} {op:>14} {ms:8.3f} {gbytes / (ms * 1e-3):8.0f} {ms * mult:8.2f}")
            tot[op] = tot.get(op, 0.0) + ms * mult
        del x_in, y0, y1
        torch.cuda.empty_cache()
    print("totals (ms per step over all blocks):")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"  {k:>14} {v:8.2f}")
    print(f"  {'sum':>14} {sum(tot.values()):8.2f}")


if __name__ == "__main__":
    main()
