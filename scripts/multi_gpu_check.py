"""Run under torchrun on N GPUs of one box:  python -m torch.distributed.run --nproc-per-node N scripts/multi_gpu_check.py
Checks the fused NVLink all-gather + InfoNCE kernel against the oracle's NCCL all_gather / reduce_scatter path
(util/dist_autograd.py semantics) for the contrastive and the MVS loss, over several steps (double-buffer / epoch logic),
then times both for the 64..512 per-GPU batch sweep."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mammoclip_b200.loss import build_loss
    from mammoclip_b200.util import GlobalEnv
    from oracle import port
    GlobalEnv.reset()
    dev = torch.device("cuda", local)
    worst = 0.0
    for mvs in (False, True):
        key = "breast_clip" if mvs else "breast_clip_contrastive"
        lf = build_loss({key: {"label_smoothing": 0.1, "i2i_weight": 1.0, "t2t_weight": 0.5, "loss_ratio": 1.0}})
        for step in range(5):
            B = (64, 100, 64, 72, 128)[step]          # not all multiples of the 32 x 64 score tiles: rank boundaries cut through tiles
            g = torch.Generator(device=dev).manual_seed(100 * step + rank)
            embs = [torch.nn.functional.normalize(torch.randn(B, 512, generator=g, device=dev), dim=1) for _ in range(4)]
            ours = [e.clone().requires_grad_(True) for e in embs]
            ref = [e.clone().requires_grad_(True) for e in embs]
            s1 = torch.tensor(14.2857, device=dev, requires_grad=True)
            s2 = torch.tensor(14.2857, device=dev, requires_grad=True)
            kw = lambda t, s: dict(image_embeddings=t[0], text_embeddings=t[1], labels=torch.arange(B, device=dev), logit_scale=s,
                                   **(dict(text_embeddings2=t[2], image_view_embeddings=t[3]) if mvs else {}))
            lo = lf(**kw(ours, s1), is_train=True)["total"]
            lo.backward()
            fn = port.mvs_loss if mvs else port.contrastive_loss
            lr = fn(**kw(ref, s2), is_train=True, label_smoothing=0.1, i2i_weight=1.0, t2t_weight=0.5)
            lr.backward()
            errs = [abs(lo.item() - lr.item()) / abs(lr.item()), abs(s1.grad.item() - s2.grad.item()) / (abs(s2.grad.item()) + 1e-6)]
            for a, b in zip(ours[: 4 if mvs else 2], ref):
                errs.append(((a.grad - b.grad).abs().max() / b.grad.abs().max()).item())
            worst = max(worst, max(errs))
            assert max(errs) < 1e-3, (rank, mvs, step, errs)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[multi_gpu_check] W={world}: fused P2P loss == all_gather/reduce_scatter oracle, worst rel err {t.item():.2e}")
    if os.environ.get("MCLIP_CHECK_NO_SWEEP") == "1":          # the parity part only (tests/test_gpu_multi.py)
        dist.destroy_process_group()
        return
    # ---- timing sweep: fused kernel vs NCCL all_gather + reduce_scatter (+ torch loss) ----
    lf = build_loss({"breast_clip_contrastive": {"label_smoothing": 0.0, "i2i_weight": 0.0, "t2t_weight": 0.0, "loss_ratio": 1.0}})
    for B in (64, 128, 256, 512):
        embs = [torch.nn.functional.normalize(torch.randn(B, 512, device=dev), dim=1).requires_grad_(True) for _ in range(2)]
        s = torch.tensor(14.2857, device=dev, requires_grad=True)
        lab = torch.arange(B, device=dev)

        def ours_step():
            lf(image_embeddings=embs[0], text_embeddings=embs[1], labels=lab, logit_scale=s, is_train=True)["total"].backward()

        def ref_step():
            port.contrastive_loss(embs[0], embs[1], lab, s, True).backward()

        res = {}
        for name, fn in (("fused", ours_step), ("nccl+torch", ref_step)):
            for _ in range(5):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                fn()
            b.record()
            dist.barrier(); torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b) / 20], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[name] = ms.item()
        if rank == 0:
            recv = (world - 1) * 2 * B * 512 * 4
            print(f"[multi_gpu_check] W={world} B={B}/GPU: fused fwd+bwd {res['fused'] * 1e3:.1f} us, NCCL gather/reduce-scatter + torch loss "
                  f"{res['nccl+torch'] * 1e3:.1f} us, gather payload {recv / 1e6:.2f} MB/rank ({recv / 770e9 * 1e6:.2f} us at 770 GB/s)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
