"""Drop-in proof: the reference trainer's own step (trainer_ddp.py:134,140-156,270-308) restated around OUR factories.

    python -m torch.distributed.run --nproc-per-node W scripts/dropin_train_check.py        (W = 1 works too)

build_model / build_loss come from mammoclip_b200 (the three-import patch of INTEGRATION.md); everything else is what the
reference does: DDP(find_unused_parameters=True), torch.optim.AdamW over model.parameters() (optimizer/__init__.py:23-31),
a LambdaLR scheduler, torch.cuda.amp.autocast + GradScaler, a DataLoader whose images arrive as [B,1,H,W,3] float32 and are
`squeeze(1).permute(0,3,1,2)`-ed by the loop, loss_dict["total"].backward() through scaler.scale().
Asserts: finite decreasing-or-stable loss over 3 steps, the BERT pooler never gets a gradient (why DDP needs
find_unused_parameters), every other parameter does and moves, BN buffers follow DDP's buffer broadcast, gradients identical
on all ranks (DDP averaging)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
from torch.utils.data import DataLoader, Dataset
from transformers import BatchEncoding, BertConfig


class _Tok:
    vocab_size = 28996


class _Synthetic(Dataset):
    """What datasets/imagetext.py + the collate deliver: image [1,H,W,3] float32 (3 identical channels), token tensors."""

    def __init__(self, n, h, w, L, seed):
        g = torch.Generator().manual_seed(seed)
        self.img = torch.randn(n, 1, h, w, 1, generator=g).expand(n, 1, h, w, 3).contiguous()
        lens = torch.randint(8, L + 1, (n,), generator=g)
        ids = torch.randint(1000, 28996, (n, L), generator=g)
        self.mask = (torch.arange(L)[None, :] < lens[:, None]).long()
        ids[:, 0] = 101
        ids[torch.arange(n), lens - 1] = 102
        self.ids = ids * self.mask

    def __len__(self):
        return self.img.shape[0]

    def __getitem__(self, i):
        return {"images": self.img[i], "input_ids": self.ids[i], "attention_mask": self.mask[i]}


def _collate(items):
    tok = BatchEncoding({"input_ids": torch.stack([i["input_ids"] for i in items]), "attention_mask": torch.stack([i["attention_mask"] for i in items]),
                         "token_type_ids": torch.zeros(len(items), items[0]["input_ids"].shape[0], dtype=torch.long)})
    return {"images": torch.stack([i["images"] for i in items]), "text_tokens": tok}


def main():
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29877")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    from mammoclip_b200.loss import build_loss            # <- INTEGRATION.md: the reference imports these from breastclip.*
    from mammoclip_b200.model import build_model
    from mammoclip_b200.model.modules.text_encoder import BERT_BASE_CASED
    from mammoclip_b200.util import GlobalEnv
    GlobalEnv.reset()
    bcfg = BertConfig(**dict(BERT_BASE_CASED, num_hidden_layers=2))
    cfg = {"name": "clip_custom",
           "image_encoder": {"source": "cnn", "name": "tf_efficientnetv2-detect", "pretrained": True, "model_type": "cnn"},
           "text_encoder": {"source": "huggingface", "name": "offline-bert", "pretrained": False, "gradient_checkpointing": False, "pooling": "eos",
                            "cache_dir": "/tmp/none", "trust_remote_code": False, "config": bcfg},
           "projection_head": {"name": "linear", "proj_dim": 512, "dropout": 0.1}, "temperature": 0.07}
    loss_cfg = {"breast_clip_contrastive": {"label_smoothing": 0.0, "i2i_weight": 1.0, "t2t_weight": 0.5, "loss_ratio": 1.0}}
    torch.manual_seed(0)
    model = build_model(cfg, loss_cfg, _Tok()).to(device)
    model = DDP(model, device_ids=[local], find_unused_parameters=True)                      # trainer_ddp.py:134
    loss_func = build_loss(loss_cfg)                                                         # :140
    optimizer = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=1e-4)            # optimizer/__init__.py:23-31
    scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda s: min(1.0, (s + 1) / 10))
    scaler = torch.cuda.amp.GradScaler()                                                     # :156
    loader = DataLoader(_Synthetic(24, 128, 96, 32, seed=1234 + rank), batch_size=8, shuffle=False, drop_last=True, num_workers=0, collate_fn=_collate)
    before = {k: v.detach().clone() for k, v in model.module.state_dict().items()}
    model.train()
    losses = []
    for idx, batch in enumerate(loader):                                                     # train(), :270-308
        optimizer.zero_grad(set_to_none=True)
        batch["images"] = batch["images"].squeeze(1).permute(0, 3, 1, 2)
        with torch.cuda.amp.autocast():
            outputs = model(batch, device)
            loss_dict = loss_func(**outputs, is_train=True)
        total_loss = loss_dict["total"]
        scaler.scale(total_loss).backward()
        if idx == 0:
            named = dict(model.module.named_parameters())
            unused = [k for k, p in named.items() if p.grad is None]
            assert unused and all("pooler" in k for k in unused), unused
            for k, p in named.items():
                if p.grad is not None:
                    assert torch.isfinite(p.grad).all(), k
                    ref = p.grad.detach().clone()
                    dist.broadcast(ref, 0)
                    assert torch.equal(ref, p.grad), f"gradient of {k} differs across ranks (DDP averaging)"
        scaler.step(optimizer)
        scaler.update()
        scheduler.step()
        loss_dict = {key: value.detach().cpu() for key, value in loss_dict.items()}
        losses.append(float(loss_dict["total"]))
    assert len(losses) == 3 and all(l == l and abs(l) < 1e4 for l in losses), losses
    after = model.module.state_dict()
    moved = [k for k in before if before[k].dtype.is_floating_point and "running_" not in k and not torch.equal(before[k], after[k])]
    frozen = [k for k in before if before[k].dtype.is_floating_point and "running_" not in k and "pooler" not in k and torch.equal(before[k], after[k])]
    assert not frozen, frozen[:5]
    assert after["image_encoder._bn0.num_batches_tracked"].item() == 3
    # DDP(broadcast_buffers=True) hands rank 0's running statistics to every rank at the START of each forward; the last step's
    # local update then differs per rank (per-rank batch statistics, no SyncBN: SURVEY A1).  One more forward re-synchronises:
    model.eval()
    with torch.no_grad():
        model(batch, device)
    for k in ("image_encoder._bn0.running_mean", "image_encoder._blocks.5._bn1.running_var"):
        mine = model.module.state_dict()[k]
        ref = mine.detach().clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, mine), k
    if rank == 0:
        print(f"[dropin_train_check] W={world}: reference train() step on mammoclip_b200 factories ok; losses {['%.4f' % l for l in losses]}; "
              f"{len(moved)} tensors updated, pooler untouched")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
