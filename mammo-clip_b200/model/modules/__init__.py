"""Factories with the reference's string-keyed contract (model/modules/__init__.py:11-89): the drop-in boundary."""
import os
from typing import Dict

from .efficientnet_custom import EfficientNet
from .projection import LinearProjectionHead, MLPProjectionHead
from .text_encoder import HuggingfaceTextEncoder


def load_image_encoder(config_image_encoder: Dict):
    source, name = config_image_encoder["source"].lower(), config_image_encoder["name"].lower()
    if source == "cnn" and name == "tf_efficientnetv2-detect":          # -> EfficientNet-B2, out_dim 1408 (:35-40)
        enc = EfficientNet.from_pretrained("efficientnet-b2", num_classes=1, weights_path=config_image_encoder.get("weights_path"))
        enc.out_dim = 1408
    elif source == "cnn" and name == "tf_efficientnet_b5_ns-detect":    # -> EfficientNet-B5, out_dim 2048 (:41-46)
        enc = EfficientNet.from_pretrained("efficientnet-b5", num_classes=1, weights_path=config_image_encoder.get("weights_path"))
        enc.out_dim = 2048
    else:
        # HuggingfaceImageEncoder / EfficientNet_Mammo (timm) / ResNet are not selected by any shipped pre-training
        # config (configs/pre_train_b{2,5}_clip.yaml) and are out of the hot path's scope.
        raise KeyError(f"Not supported image encoder: {config_image_encoder}")
    return enc


def load_text_encoder(config_text_encoder: Dict, vocab_size: int):
    if config_text_encoder["source"].lower() == "huggingface":
        cache_dir = config_text_encoder["cache_dir"]
        return HuggingfaceTextEncoder(
            name=config_text_encoder["name"], vocab_size=vocab_size, pretrained=config_text_encoder["pretrained"],
            gradient_checkpointing=config_text_encoder["gradient_checkpointing"], cache_dir=cache_dir,
            local_files_only=os.path.exists(os.path.join(cache_dir, f'models--{config_text_encoder["name"].replace("/", "--")}')),
            trust_remote_code=config_text_encoder["trust_remote_code"], config=config_text_encoder.get("config"))
    raise KeyError(f"Not supported text encoder: {config_text_encoder}")


def load_projection_head(embedding_dim: int, config_projection_head: Dict):
    name = config_projection_head["name"].lower()
    if name == "mlp":
        return MLPProjectionHead(embedding_dim=embedding_dim, projection_dim=config_projection_head["proj_dim"], dropout=config_projection_head["dropout"])
    if name == "linear":
        return LinearProjectionHead(embedding_dim=embedding_dim, projection_dim=config_projection_head["proj_dim"])
    raise KeyError(f"Not supported text encoder: {config_projection_head}")
