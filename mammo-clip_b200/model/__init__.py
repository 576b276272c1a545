"""`build_model` with the reference's contract (model/__init__.py:10-21)."""
from typing import Dict

from torch import nn

from .clip import BreastClip


def build_model(model_config: Dict, loss_config: Dict, tokenizer=None) -> nn.Module:
    name = model_config["name"].lower()
    if name == "clip_custom":
        return BreastClip(model_config, loss_config, tokenizer)
    if name in ("finetune_classification", "pretrained_classifier"):
        raise KeyError(f"{model_config['name']} is a downstream fine-tuning model, outside the pre-training hot path")
    raise KeyError(f"Not supported model: {model_config['name']}")
