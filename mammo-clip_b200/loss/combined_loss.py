"""Same contract as the reference's loss/combined_loss.py:6-29: sums `loss * loss_ratio`, returns a dict keyed by each
loss's `.name` plus "total"; `.loss_list` is iterated by the trainer (trainer_ddp.py:276-277)."""
from typing import List

import torch.nn as nn


class CombinedLoss(nn.Module):
    def __init__(self, loss_list: List[nn.Module]):
        super().__init__()
        self.loss_list = loss_list

    def forward(self, **kwargs):
        loss_dict = dict()
        total_loss = 0.0
        for loss in self.loss_list:
            cur_loss = loss(**kwargs)
            loss_dict[loss.name] = cur_loss
            total_loss += cur_loss * loss.loss_ratio
        loss_dict["total"] = total_loss
        return loss_dict
