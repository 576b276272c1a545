"""A few self-attention forward+backward calls at BERT-base geometry (target for ncu): python scripts/attention_one.py [B] [L]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammoclip_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 64
heads, d = 12, 64
H = heads * d
qkv = torch.randn(B * L, 3 * H, device="cuda").bfloat16()
amask = torch.ones(B, L, device="cuda", dtype=torch.long)
keep = (torch.rand(B, heads, L, L, device="cuda") >= 0.1).to(torch.uint8)
do = torch.randn(B * L, H, device="cuda").bfloat16()
for _ in range(3):
    out, lse = ops.bert_attention(qkv, amask, B, L, heads, d, keep, 1 / 0.9, want_lse=True)
    ops.bert_attention_backward(qkv, do, lse, amask, B, L, heads, d, keep, 1 / 0.9)
torch.cuda.synchronize()
print("done")
