"""CLIP wrapper with the reference's interface (model/clip.py:14-114): attributes `image_encoder`, `text_encoder`,
`image_projection`, `text_projection`, `projection`, `logit_scale`, `tokenizer`; methods `encode_image`, `encode_text`,
`encode_image_normalized`; `forward(batch, device) -> dict` consumed by `CombinedLoss(**outputs, is_train=...)`."""
import logging
from typing import Dict

import numpy as np
import torch
from torch import nn

from .. import ops
from .modules import load_image_encoder, load_projection_head, load_text_encoder

log = logging.getLogger(__name__)


class _L2NormFn(torch.autograd.Function):
    """x / ||x||_2 per row, no epsilon (clip.py:90-91)."""

    @staticmethod
    def forward(ctx, x):
        xb = ops.cast_bf16(x.detach().float().contiguous())     # the projection output is bf16 under autocast
        e, nrm = ops.l2norm_forward(xb)
        ctx.save_for_backward(e, nrm)
        return e

    @staticmethod
    def backward(ctx, de):
        e, nrm = ctx.saved_tensors
        return ops.l2norm_backward(e, de.contiguous().float(), nrm).float()


class BreastClip(nn.Module):
    def __init__(self, model_config: Dict, all_loss_config: Dict, tokenizer=None):
        super().__init__()
        self.tokenizer = tokenizer
        self.image_encoder = load_image_encoder(model_config["image_encoder"])
        self.text_encoder = load_text_encoder(model_config["text_encoder"], vocab_size=getattr(tokenizer, "vocab_size", None))
        self.text_pooling = model_config["text_encoder"]["pooling"]
        self.model_config = model_config
        self.loss_config = {k: v for k, v in all_loss_config.items()}
        self.projection = "projection_head" in model_config
        if self.projection:
            self.image_projection = load_projection_head(embedding_dim=self.image_encoder.out_dim, config_projection_head=model_config["projection_head"])
            self.text_projection = load_projection_head(embedding_dim=self.text_encoder.out_dim, config_projection_head=model_config["projection_head"])
        else:
            assert self.image_encoder.out_dim == self.text_encoder.out_dim, \
                "Without 'projection_head', embedding_dim of the image and text encoder must be the same."
        self.temperature = model_config["temperature"] if "temperature" in model_config else None
        if self.temperature:
            self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / self.temperature))
        else:
            self.logit_scale = torch.tensor(1, dtype=torch.float32)
            log.warning("[Mammo-CLIP] missing temperature scaling factor")

    def encode_image(self, image):
        image_features = self.image_encoder(image)
        if self.model_config["image_encoder"]["model_type"].lower() == "cnn":
            return image_features
        return image_features[:, 0]

    def encode_image_normalized(self, image):
        img_emb = self.encode_image(image)
        img_emb = self.image_projection(img_emb) if self.projection else img_emb
        return _L2NormFn.apply(img_emb)

    def encode_text(self, text_tokens):
        text_features = self.text_encoder(text_tokens)
        if self.text_pooling == "eos":
            eos_token_indices = text_tokens["attention_mask"].sum(dim=-1) - 1
            text_features = text_features[torch.arange(text_features.shape[0], device=text_features.device), eos_token_indices]
        elif self.text_pooling == "bos":
            text_features = text_features[:, 0]
        elif self.text_pooling == "mean":
            m = text_tokens["attention_mask"].unsqueeze(-1).expand(text_features.size()).float()
            text_features = torch.sum(text_features * m, axis=1) / torch.clamp(m.sum(axis=1), min=1e-9)
        else:
            raise NotImplementedError("Not supported pooling method : %s", self.text_pooling)
        return text_features

    def _embed(self, feats, head):
        emb = head(feats) if self.projection else feats
        return _L2NormFn.apply(emb)

    def forward(self, batch, device=None):
        device = batch["images"].device if device is None else device
        mvs = "text_tokens2" in batch and "image_views" in batch
        image_view_encode = None
        if mvs and hasattr(self.image_encoder, "forward_views") and self.model_config["image_encoder"]["model_type"].lower() == "cnn":
            # two image batches in one step: let the tower choose its memory plan (both views' state resident, or recompute)
            image_features_g, image_view_encode = self.image_encoder.forward_views([batch["images"].to(device), batch["image_views"].to(device)],
                                                                                   plan=getattr(self, "mvs_memory_plan", "auto"))
        else:
            image_features_g = self.encode_image(batch["images"].to(device))
        text_features_g = self.encode_text(batch["text_tokens"].to(device))
        image_embeddings = self._embed(image_features_g, self.image_projection if self.projection else None)
        text_embeddings = self._embed(text_features_g, self.text_projection if self.projection else None)
        labels = torch.arange(image_embeddings.shape[0], device=device)
        out = {"image_embeddings": image_embeddings, "text_embeddings": text_embeddings, "labels": labels,
               "logit_scale": self.logit_scale.exp()}
        if "text_tokens2" in batch and "image_views" in batch:
            text_features_g2 = self.encode_text(batch["text_tokens2"].to(device))
            # clip.py:105 falls back to text_features_g (the FIRST text) without a projection head; kept as is
            out["text_embeddings2"] = self._embed(text_features_g2 if self.projection else text_features_g,
                                                  self.text_projection if self.projection else None)
            if image_view_encode is None:
                image_view_encode = self.encode_image(batch["image_views"].to(device))
            out["image_view_embeddings"] = self._embed(image_view_encode, self.image_projection if self.projection else None)
        return out
