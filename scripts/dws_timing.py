"""debug: where do the streaming depthwise warps spend their cycles (needs a -DDWS_TIMING build via MCLIP_LIB)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammoclip_b200 import ops, _lib
from mammoclip_b200.model.modules.efficientnet_custom import net_geometry
from layer_bench import bn_state
lib = _lib.lib()
g = net_geometry("efficientnet-b5")
pl, pr, pt, pb = g.stem_pads
h, w = (1520 + pt + pb - 3) // 2 + 1, (912 + pl + pr - 3) // 2 + 1
for i, b in enumerate(g.blocks):
    l, r, t, bb = b.pads
    ho, wo = (h + t + bb - b.k) // b.s + 1, (w + l + r - b.k) // b.s + 1
    if i in (4, 9, 14, 28) :
        y0 = torch.randn(64, h, w, b.cexp, device="cuda").to(torch.bfloat16)
        bn0 = bn_state(b.cexp)
        wdw = torch.randn(b.cexp, 1, b.k, b.k, device="cuda") * 0.2
        ops.dwconv_forward(y0, wdw, b.k, b.s, b.pads, bn=bn0)
        torch.cuda.synchronize()
        out = (ctypes.c_ulonglong * 8)()
        lib.mclip_dws_timing(out, 1)
        ops.dwconv_forward(y0, wdw, b.k, b.s, b.pads, bn=bn0)
        lib.mclip_dws_timing(out, 0)
        wait, comp, prod, nblk, tot, nw = out[0], out[1], out[2], out[3], out[4], out[5]
        print(f"block {i} k{b.k}: warps {nw}, kernel cycles per warp {tot / nw:.0f}, warp-blocks per warp {nblk / nw:.0f}; per warp-block cycles: wait_full {wait / nblk:.0f}, "
              f"compute {comp / nblk:.0f}, produce+syncwarp {prod / nblk:.0f}; share of warp time: wait {wait / tot:.2f} compute {comp / tot:.2f} produce {prod / tot:.2f}")
        del y0
    h, w = ho, wo
