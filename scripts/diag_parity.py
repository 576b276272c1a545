"""Diagnostic (GPU box): full-depth tower errors of ours / autocast / bf16-emulating oracle vs the fp32 oracle."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from conftest import rel_err
from test_gpu_encoder import _build, _structurally_zero
from oracle import port


def grads(m):
    return {k: p.grad.detach().clone() for k, p in m.named_parameters()}


def errs(g, g32, floor, skip):
    out = {}
    for k in g32:
        if k in skip: continue
        d = (g[k].double() - g32[k].double())
        out[k] = (d.abs().max().item() / max(g32[k].abs().max().item(), floor), d.norm().item() / max(g32[k].double().norm().item(), floor * d.numel() ** 0.5))
    return out


def run(name, B, h, w, mode):
    ours, ref = _build(name)
    tr = mode == "train"
    ours.train(tr), ref.train(tr)
    x = port.synth_images(B, h, w, seed=1234, identical_channels=False, device="cuda")
    probe = torch.randn(B, ref.out_dim, generator=torch.Generator().manual_seed(99)).cuda()
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    fo = ours(x); (fo * probe).sum().backward(); go = grads(ours)
    fr = ref(x); (fr * probe).sum().backward(); g32 = grads(ref); fr = fr.detach()
    ref.load_state_dict(sd0); ref.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fa = ref(x); (fa.float() * probe).sum().backward()
    ga = grads(ref); fa = fa.detach().float()
    ref.load_state_dict(sd0); ref.zero_grad(set_to_none=True); ref.emulate_bf16 = True
    fe = ref(x); (fe * probe).sum().backward(); ge = grads(ref); fe = fe.detach()
    skip = {k for k in g32 if _structurally_zero(ours, k)}
    gmax = max(t.abs().max().item() for t in g32.values()); floor = 1e-2 * gmax
    eo, ea, ee = errs(go, g32, floor, skip), errs(ga, g32, floor, skip), errs(ge, g32, floor, skip)
    eoe = errs(go, ge, floor, skip)
    tot = lambda g, r: (sum(((g[k].double() - r[k].double()) ** 2).sum().item() for k in r if k not in skip) / sum((r[k].double() ** 2).sum().item() for k in r if k not in skip)) ** 0.5
    import statistics as st
    print(json.dumps({"cfg": [name, B, h, w, mode], "feat": {"ours": rel_err(fo, fr), "amp": rel_err(fa, fr), "emu": rel_err(fe, fr), "ours_vs_emu": rel_err(fo, fe)},
                      "grad_max": {"ours": max(v[0] for v in eo.values()), "amp": max(v[0] for v in ea.values()), "emu": max(v[0] for v in ee.values()), "ours_vs_emu": max(v[0] for v in eoe.values())},
                      "grad_median": {"ours": st.median(v[0] for v in eo.values()), "amp": st.median(v[0] for v in ea.values()), "emu": st.median(v[0] for v in ee.values()), "ours_vs_emu": st.median(v[0] for v in eoe.values())},
                      "grad_global_l2": {"ours": tot(go, g32), "amp": tot(ga, g32), "emu": tot(ge, g32), "ours_vs_emu": tot(go, ge)},
                      "ours_worse_than_amp": sum(1 for k in eo if eo[k][0] > max(1e-2, ea[k][0])), "n": len(eo)}))
    if os.environ.get("DIAG_DUMP"):
        rows = sorted(((k, eo[k][0], ea[k][0], ee[k][0], eoe[k][0], g32[k].abs().max().item() / gmax) for k in eo), key=lambda r: -r[1] / max(1e-2, r[2]))
        with open(os.path.join(ROOT, "gpurun_out", "diag_rows_%s_%dx%dx%d_%s.txt" % (name, B, h, w, mode)), "w") as f:
            f.write("tensor  ours  amp  emu  ours_vs_emu  |g|max/gmax\n")
            for r in rows:
                f.write("%-48s %.4f %.4f %.4f %.4f %.3g\n" % r)
    del ours, ref
    torch.cuda.empty_cache()


if __name__ == "__main__":
    cfgs = [("efficientnet-b5", 4, 1520, 912, "eval"), ("efficientnet-b5", 8, 456, 456, "eval")] if os.environ.get("DIAG_DUMP") else None
    for cfg in cfgs or [("efficientnet-b2", 8, 320, 256, "eval"), ("efficientnet-b2", 8, 320, 256, "train"), ("efficientnet-b2", 16, 448, 448, "train"),
                ("efficientnet-b5", 8, 456, 456, "eval"), ("efficientnet-b5", 8, 456, 456, "train"), ("efficientnet-b5", 4, 1520, 912, "eval"),
                ("efficientnet-b5", 4, 1520, 912, "train")]:
        run(*cfg)
