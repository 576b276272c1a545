"""CPU check for DESIGN.md §9 item 3: BatchNorm statistics of the expand convolution's output from the INPUT's Gram matrix.

y = x We^T (1x1 conv, K = Cin narrow, N = 6 Cin wide).  Instead of accumulating sum(y), sum(y^2) over the wide output in the
GEMM epilogue (≈ 1/3 of its instructions), use  sum_p y[p,n] = We[n,:] . colsum(x)  and  sum_p y[p,n]^2 = We[n,:] G We[n,:]^T
with G = x^T x (a [Cin,Cin] wgrad-style GEMM over the narrow input, fp32 accumulation).  The reference normalises the
bf16-ROUNDED y; this script measures what ignoring that rounding (and fp32 accumulation of G) does to mean / invstd."""
import torch


def run(P, cin, cexp, dc, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(P, cin, generator=g) + 0.0)
    x[:, 0] = x[:, 0] * 0.05 + dc
    x = x.to(torch.bfloat16)
    We = (torch.randn(cexp, cin, generator=g) / cin ** 0.5).to(torch.bfloat16)
    y = (x.float() @ We.float().T).to(torch.bfloat16).double()          # what the GEMM stores
    mean_t, var_t = y.mean(0), y.var(0, unbiased=False)
    G = (x.float().T @ x.float()).double()                               # fp32-accumulated Gram (tensor-core like), then fp64
    cs = x.float().sum(0).double()
    W = We.double()
    mean_g = (W @ cs) / P
    var_g = ((W @ G) * W).sum(1) / P - mean_g ** 2
    inv_t, inv_g = 1 / torch.sqrt(var_t + 1e-3), 1 / torch.sqrt(var_g.clamp_min(0) + 1e-3)
    return ((mean_g - mean_t).abs() * inv_t).max().item(), ((inv_g - inv_t).abs() / inv_t).max().item(), (mean_t.abs() * inv_t).max().item()


if __name__ == "__main__":
    print(f"{'P':>8} {'cin':>4} {'cexp':>5} {'|mean|/std':>10} | {'d mean / std':>12} {'rel d invstd':>12}")
    for P, cin, cexp, dc in ((200000, 40, 240, 0.0), (200000, 40, 240, 3.0), (500000, 24, 144, 10.0), (50000, 176, 1056, 3.0)):
        dm, di, m = run(P, cin, cexp, dc, 0)
        print(f"{P:8d} {cin:4d} {cexp:5d} {m:10.2f} | {dm:12.2e} {di:12.2e}")
