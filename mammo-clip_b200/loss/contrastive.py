"""In-batch InfoNCE losses on the fused CUDA kernel, drop-in for the reference loss modules.

`BreastClip_contrastive` mirrors loss/breast_clip_contrastive.py:19-59 and `BreastClip` mirrors loss/breast_clip.py:20-127:
same constructor arguments, `.name`, `.loss_ratio`, and `forward(image_embeddings, text_embeddings, ..., labels,
logit_scale, is_train, **kwargs) -> scalar` with autograd to the embeddings and logit_scale.  The all-gather with gradient
(util/dist_autograd.py) is part of the kernel: no NCCL call, no reduce-scatter (csrc/loss.cu)."""
import torch
import torch.nn as nn

from .. import ops, util
from ..util.symm import SymmetricGather


class _FusedInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, pairs, logit_scale, *embeds):
        env = util.GlobalEnv.get()
        local = [e.detach().contiguous().float() for e in embeds]
        scale = logit_scale.detach().float().reshape(1).contiguous()
        symm = owner._symm(len(local), local[0].shape[0], local[0].shape[1]) if env.world_size > 1 else None
        out, grads = ops.contrastive_loss_raw(local, pairs, scale, world=env.world_size, rank=env.world_rank, symm=symm)
        ctx.grads, ctx.dscale, ctx.dtypes = grads, out[1], [e.dtype for e in embeds]
        ctx.scale_shape = logit_scale.shape
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, dloss, _dout):
        gs = [(g * dloss).to(dt) for g, dt in zip(ctx.grads, ctx.dtypes)]
        dscale = (ctx.dscale * dloss).reshape(ctx.scale_shape) if ctx.needs_input_grad[2] else None
        return (None, None, dscale, *gs)


class _FusedLossBase(nn.Module):
    def __init__(self, label_smoothing=0.0, i2i_weight=0.0, t2t_weight=0.0, loss_ratio=1.0):
        super().__init__()
        self.name = "contrastive"
        self.label_smoothing = label_smoothing
        self.loss_ratio = loss_ratio
        self.i2i_weight = i2i_weight
        self.t2t_weight = t2t_weight
        self._symm_cache = {}
        self.last_components = None     # device tensor [2+2P]: loss, dscale, per-pair (row CE, col CE); no host sync

    def _symm(self, k, b, d):
        key = (k, b, d)
        if key not in self._symm_cache:
            self._symm_cache[key] = SymmetricGather(k, b, d)
        return self._symm_cache[key]

    def _run(self, pairs, logit_scale, embeds):
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=embeds[0].device)
        elif logit_scale.device != embeds[0].device:
            # model config without `temperature`: clip.py keeps logit_scale as a CPU 0-dim constant, like the reference (:39-41),
            # whose `logit_scale * emb` accepts it; it carries no gradient
            logit_scale = logit_scale.detach().to(embeds[0].device)
        loss, comps = _FusedInfoNCE.apply(self, pairs, logit_scale, *embeds)
        self.last_components = comps
        return loss

    @staticmethod
    def _log(scalars):
        """The reference writes TensorBoard scalars from inside forward (breast_clip_contrastive.py:49-55); kept, but
        only when a writer is installed (each add_scalar is a device->host sync)."""
        sw = util.GlobalEnv.get().summary_writer
        if sw.train is not None:
            for tag, val in scalars:
                sw.train.add_scalar(tag, val, sw.global_step)


class BreastClip_contrastive(_FusedLossBase):
    def forward(self, image_embeddings, text_embeddings, labels, logit_scale, is_train, **kwargs):
        eps = float(self.label_smoothing) if is_train else 0.0
        # logits_per_image = rows of S(img,txt) (weight 0.75), logits_per_text = its columns (0.25): :42-43,58
        loss = self._run([(0, 1, 0.75, 0.25, eps)], logit_scale, [image_embeddings, text_embeddings])
        if is_train:
            c = self.last_components
            self._log([("loss/contrastive/steps_i2t", c[2]), ("loss/contrastive/steps_t2i", c[3])])
        return loss


class BreastClip(_FusedLossBase):
    """Multi-view / multi-text loss: 4 image-text pairs (label smoothing), image-image and text-text pairs (none)."""

    def forward(self, image_embeddings, text_embeddings, text_embeddings2, image_view_embeddings, labels, logit_scale, is_train, **kwargs):
        eps = float(self.label_smoothing) if is_train else 0.0
        iw, tw = float(self.i2i_weight), float(self.t2t_weight)
        # tensors: 0 = I1, 1 = T1, 2 = T2, 3 = I2.  total = (sum4 i2t/4 + sum4 t2i/4)/2 + iw*(i2i)/2 + tw*(t2t)/2   (:42-125)
        pairs = [(0, 1, .125, .125, eps), (3, 1, .125, .125, eps), (0, 2, .125, .125, eps), (3, 2, .125, .125, eps),
                 (0, 3, iw / 2, iw / 2, 0.0), (2, 1, tw / 2, tw / 2, 0.0)]
        loss = self._run(pairs, logit_scale, [image_embeddings, text_embeddings, text_embeddings2, image_view_embeddings])
        if is_train:
            c = self.last_components
            self._log([("loss/contrastive/steps_i2t", (c[2] + c[4] + c[6] + c[8]) / 4), ("loss/contrastive/steps_t2i", (c[3] + c[5] + c[7] + c[9]) / 4),
                       ("loss/contrastive/steps_i2i", (c[10] + c[11]) / 2), ("loss/contrastive/steps_t2t", (c[12] + c[13]) / 2),
                       ("params/logit_scale", logit_scale), ("params/temperature", 1.0 / logit_scale)])
        return loss
