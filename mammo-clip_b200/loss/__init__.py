"""`build_loss` with the reference's config contract (loss/__init__.py:9-28)."""
from typing import Dict

from .combined_loss import CombinedLoss
from .contrastive import BreastClip, BreastClip_contrastive


def build_loss(all_loss_config: Dict) -> CombinedLoss:
    loss_list = []
    for loss_config in all_loss_config:
        cfg = all_loss_config[loss_config]
        if cfg["loss_ratio"] == 0.0:
            continue
        if loss_config == "breast_clip":
            loss = BreastClip(**cfg)
        elif loss_config == "breast_clip_contrastive":
            loss = BreastClip_contrastive(**cfg)
        elif loss_config == "classification":
            raise KeyError("classification loss belongs to the downstream fine-tuning path, which is out of scope here")
        else:
            raise KeyError(f"Unknown loss: {loss_config}")
        loss_list.append(loss)
    return CombinedLoss(loss_list)
