"""GPU diagnostic: per-block relative error of the CUDA tower vs the oracle (train and eval)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import port
from mammoclip_b200.model.modules import efficientnet_custom as E


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item(), ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run(name, batch, h, w, mode, emulate=False):
    ours = E.EfficientNet.from_name(name); ours.stochastic = False
    ref = port.OracleEfficientNet(name); ref.stochastic = False
    port.fill_deterministic(ref, 0)
    ours.load_state_dict(ref.state_dict()); ours.cuda(); ref.cuda()
    ours.train(mode == "train"); ref.train(mode == "train"); ref.emulate_bf16 = emulate
    x = port.synth_images(batch, h, w, seed=1234, identical_channels=False, device="cuda")
    acts = {}
    for i, blk in enumerate(ref._blocks):
        blk.register_forward_hook(lambda m, inp, out, i=i: acts.__setitem__(i, out.detach()))
    with torch.no_grad():
        fr = ref(x)
        feat, S, _ = E._forward(ours, x.float(), ours.training, None, None, False)
    print(f"== {name} {mode} emulate={emulate} B={batch} {h}x{w}: features max-rel {rel(feat, fr)[0]:.4f} l2-rel {rel(feat, fr)[1]:.4f}")
    nb = len(ours._blocks)
    for i in range(nb):
        xo = S["blocks"][i + 1]["x_in"] if i + 1 < nb else S["x_last"]
        r = acts[i].permute(0, 2, 3, 1)
        m, l2 = rel(xo.float(), r)
        print(f"  block {i:2d} out {tuple(xo.shape)} max-rel {m:.4f} l2-rel {l2:.4f}")


if __name__ == "__main__":
    run("efficientnet-b2", 2, 96, 64, "train", True)
    run("efficientnet-b2", 4, 224, 224, "train", True)
    run("efficientnet-b5", 2, 160, 96, "train", True)
