import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from mammoclip_b200 import ops
    from layer_bench import timeit
    for (name, m, k, n) in [("expand b3", 64*760*456, 24, 144), ("expand b4", 64*380*228, 40, 240), ("proj_dgrad b0", 64*760*456, 24, 48), ("proj_dgrad b4", 64*380*228, 40, 240), ("expand b9", 64*190*114, 64, 384)]:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16); w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
        ms = timeit(lambda: ops.gemm_tn(a, w, want_stats=True))
        gb = 2*(m*k+m*n)/1e9
        print(f"  dbg={os.environ.get('MCLIP_GEMM_DEBUG','0')} {name:14s} M={m} K={k} N={n}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s")
        del a, w
else:
    for dbg in ("0", "1", "3", "7", "5"):
        env = dict(os.environ, MCLIP_GEMM_DEBUG=dbg)
        subprocess.run([sys.executable, __file__, "run"], env=env)
