"""Text tower with the reference's contract (text_encoder.py:5-49): parameters live under `.text_encoder.*` with the
Hugging Face BERT names, `out_dim = hidden_size`, `forward(BatchEncoding) -> last_hidden_state [B,L,H]`.

Forward and backward on CUDA run the hand-written kernels of `bert_kernels.py` (embedding+LayerNorm, tcgen05 QKV/out/FFN
GEMMs with fused bias/GELU/residual epilogues, masked softmax attention, and their gradients)."""
import torch
from torch import nn
from transformers import AutoConfig, BertModel


# Bio_ClinicalBERT == BERT-base-cased geometry (no hub access here: weights are random-init unless a local path is given)
BERT_BASE_CASED = dict(vocab_size=28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                       max_position_embeddings=512, type_vocab_size=2, hidden_act="gelu", layer_norm_eps=1e-12,
                       hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)


class HuggingfaceTextEncoder(nn.Module):
    def __init__(self, name="bert-base-uncased", vocab_size=None, pretrained=True, gradient_checkpointing=False,
                 cache_dir="~/.cache/huggingface/hub", local_files_only=False, trust_remote_code=False, config=None):
        super().__init__()
        if config is not None:                       # offline construction from a BertConfig (tests / bench)
            self.text_encoder = BertModel(config)
        elif pretrained:
            from transformers import AutoModel
            self.text_encoder = AutoModel.from_pretrained(name, ignore_mismatched_sizes=True, cache_dir=cache_dir,
                                                          local_files_only=local_files_only, trust_remote_code=trust_remote_code)
        else:
            model_config = AutoConfig.from_pretrained(name, ignore_mismatched_sizes=True, cache_dir=cache_dir,
                                                      local_files_only=local_files_only, trust_remote_code=trust_remote_code)
            if type(model_config).__name__ != "BertConfig":
                raise NotImplementedError(f"Not support training from scratch : {type(model_config).__name__}")
            self.text_encoder = BertModel(model_config)
        if not isinstance(self.text_encoder, BertModel):
            raise NotImplementedError("the B200 text tower implements the BERT architecture only")
        self.out_dim = self.text_encoder.config.hidden_size
        self.use_kernels = True
        self._mclip_direct_grads = True      # FlatAdamW.attach(): the backward may write gradients straight into the flat buffer

    def forward(self, x):
        ids = x["input_ids"]
        if not ids.is_cuda:
            raise RuntimeError("mammoclip_b200 text encoder runs on a B200 only (no CPU fallback)")
        from . import bert_kernels
        return bert_kernels.bert_forward(self.text_encoder, x["input_ids"], x.get("token_type_ids"), x["attention_mask"], self.training, owner=self)
