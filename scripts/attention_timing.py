"""CUDA-event time of the self-attention kernels (forward, backward) at BERT-base geometry: python scripts/attention_timing.py [B] [L]
(MCLIP_ATT_TC=0 selects the SIMT kernels of bert.cu for comparison)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammoclip_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for L in ([int(sys.argv[2])] if len(sys.argv) > 2 else [64, 128, 256]):
    heads, d = 12, 64
    H = heads * d
    qkv = torch.randn(B * L, 3 * H, device="cuda").bfloat16()
    amask = torch.ones(B, L, device="cuda", dtype=torch.long)
    keep = (torch.rand(B, heads, L, L, device="cuda") >= 0.1).to(torch.uint8)
    do = torch.randn(B * L, H, device="cuda").bfloat16()
    out, lse = ops.bert_attention(qkv, amask, B, L, heads, d, keep, 1 / 0.9, want_lse=True)

    def t(fn, reps=20):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    f = t(lambda: ops.bert_attention(qkv, amask, B, L, heads, d, keep, 1 / 0.9, want_lse=True))
    bw = t(lambda: ops.bert_attention_backward(qkv, do, lse, amask, B, L, heads, d, keep, 1 / 0.9))
    flops = 4 * B * heads * L * L * d
    print(f"B={B} L={L} tc={os.environ.get('MCLIP_ATT_TC', '1')}: forward {f:8.1f} us ({flops / f / 1e6:7.1f} TFLOP/s)   backward {bw:8.1f} us ({2.5 * flops / bw / 1e6:7.1f} TFLOP/s)")
