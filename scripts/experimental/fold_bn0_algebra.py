"""CPU check of the algebra behind DESIGN.md §9 item 2 (no kernels involved; nothing here is imported by the product).

Today the backward of an MBConv block materialises dY0 = a0 * (dv0 - c1 - yhat0 * c2) with a pass over the 6x-wide tensor
(ew_backward mode 1, dv_given) and feeds it to two GEMMs (dX = dY0 We, dWe = dY0^T x).  Both consumers are linear, so the
BatchNorm-backward correction can be folded into the GEMM operands:

    dX  = dv0 (diag(a0) We) - Y0 (diag(k2) We) + 1 (t0^T We)
    dWe = diag(a0) (dv0^T x) - diag(k2) (We (x^T x)) + t0 (colsum x)^T          with Y0^T x ~= We (x^T x)
    k2 = a0 c2 invstd,  t0 = -a0 c1 + k2 mean

This script measures, against an fp64 reference, the error of (a) the current path with bf16 dY0 and (b) the folded path
with bf16 GEMM operands, for MBConv-like shapes and statistics.  Run: python scripts/experimental/fold_bn0_algebra.py"""
import torch


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


def run(P, cin, cexp, mean_over_std, seed):
    g = torch.Generator().manual_seed(seed)
    x = bf(torch.randn(P, cin, generator=g, dtype=torch.float64))
    We = torch.randn(cexp, cin, generator=g, dtype=torch.float64) / cin ** 0.5
    We[:, 0] += We.std()                                      # a DC input channel gives every expanded channel a mean of
    x[:, 0] = bf(x[:, 0] * 0.05 + mean_over_std)               # about `mean_over_std` standard deviations
    Y0 = bf(x @ bf(We).T)                                     # stored pre-BN conv output
    mean, var = Y0.mean(0), Y0.var(0, unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-3)
    gamma = torch.rand(cexp, generator=g, dtype=torch.float64) * 0.8 + 0.6
    a0 = gamma * invstd
    dv0 = bf(torch.randn(P, cexp, generator=g, dtype=torch.float64) * 1e-3)
    yhat = (Y0 - mean) * invstd
    c1, c2 = dv0.mean(0), (dv0 * yhat).mean(0)
    dY0 = a0 * (dv0 - c1 - yhat * c2)                         # exact BN backward
    dX_true, dWe_true = dY0 @ bf(We), dY0.T @ x
    # (a) current path: dY0 rounded to bf16, bf16 GEMM operands
    dX_a, dWe_a = bf(dY0) @ bf(We), bf(dY0).T @ x
    # (b) folded path
    k2 = a0 * c2 * invstd
    t0 = -a0 * c1 + k2 * mean
    B1, B2 = bf(a0[:, None] * We), bf(k2[:, None] * We)       # scaled weights as bf16 GEMM operands
    dX_b = dv0 @ B1 - Y0 @ B2 + (t0 @ bf(We))[None, :]
    gram, cs = x.T @ x, x.sum(0)
    dWe_b = a0[:, None] * (dv0.T @ x) - k2[:, None] * (bf(We) @ gram) + t0[:, None] * cs[None, :]
    rel = lambda u, v: ((u - v).abs().max() / v.abs().max()).item()
    return rel(dX_a, dX_true), rel(dX_b, dX_true), rel(dWe_a, dWe_true), rel(dWe_b, dWe_true), (mean.abs() * invstd).max().item()


if __name__ == "__main__":
    print(f"{'P':>8} {'cin':>4} {'cexp':>5} {'|mean|/std':>10} | {'dX cur':>9} {'dX fold':>9} | {'dWe cur':>9} {'dWe fold':>9}")
    for P, cin, cexp, mos in ((20000, 40, 240, 0.0), (20000, 40, 240, 3.0), (50000, 24, 144, 10.0), (8000, 176, 1056, 3.0), (4000, 304, 1824, 1.0)):
        ea, eb, wa, wb, m = run(P, cin, cexp, mos, 0)
        print(f"{P:8d} {cin:4d} {cexp:5d} {m:10.2f} | {ea:9.2e} {eb:9.2e} | {wa:9.2e} {wb:9.2e}")
