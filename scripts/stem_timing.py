import sys, os
sys.path.insert(0, "/root/repo")
import torch
from mammoclip_b200 import ops
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
B,H,W=64,1520,912
x3=torch.randn(B,H,W,1,device="cuda").expand(B,H,W,3).contiguous().permute(0,3,1,2)
u8=torch.randint(0,256,(B,1,H,W),device="cuda",dtype=torch.uint8)
x1=torch.randn(B,1,H,W,device="cuda")
pads=(0,1,0,1)
print("im2col fp32x3", t(lambda: ops.stem_im2col(x3,pads)))
print("im2col fp32x1", t(lambda: ops.stem_im2col(x1,pads)))
mm,lut=ops.image_norm_lut_u8(u8,0.3,0.25)
print("lut kernel", t(lambda: ops.image_norm_lut_u8(u8,0.3,0.25)))
print("im2col u8+lut", t(lambda: ops.stem_im2col(u8,pads,norm_lut=lut)))
print("im2col u8 raw", t(lambda: ops.stem_im2col(u8,pads)))
