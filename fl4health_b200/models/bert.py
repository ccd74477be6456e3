"""Compact BERT-style encoder for sequence classification (the reference's ``bert_finetuning_example`` fine-tunes
``bert-base-cased`` from the HuggingFace hub; there is no hub access here, so the architecture ships in-tree and is
randomly initialised or loaded from a local state dict with HF-compatible shapes).

Inputs follow the HF convention and arrive as a dict (``input_ids``, ``attention_mask``[, ``token_type_ids``]), which
exercises the clients' dict-input path.  On a B200 in bf16 every block is in-house: attention (sequence <= 128, head
dimension 64) is one tcgen05 kernel per direction reading the packed QKV projection in place (``ops/attention.py``),
the sub-layer epilogues ``LayerNorm(x + dropout(.))`` are one kernel each (``ops/layer_norm.py``), and the projections
are ``LinearAct`` modules (tcgen05 GEMM with fused bias(+GELU / ReLU) epilogue).  Elsewhere the same modules run the
stock PyTorch composition.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from fl4health_b200.models.fused_layers import LinearAct, ResidualLayerNorm
from fl4health_b200.ops.attention import packed_self_attention


@dataclass
class BertConfig:
    vocab_size: int = 28996  # bert-base-cased
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    hidden_dropout_prob: float = 0.1
    activation: str = "gelu"  # "gelu" (erf, as HF BERT) or "relu": fused into the first feed-forward GEMM's epilogue

    @classmethod
    def tiny(cls, vocab_size: int = 1000) -> BertConfig:
        return cls(vocab_size=vocab_size, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                   max_position_embeddings=64)


class BertEmbeddings(nn.Module):
    def __init__(self, cfg: BertConfig) -> None:
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)

    def forward(self, input_ids: torch.Tensor, token_type_ids: torch.Tensor | None) -> torch.Tensor:
        positions = torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
        types = token_type_ids if token_type_ids is not None else torch.zeros_like(input_ids)
        x = self.word_embeddings(input_ids) + self.position_embeddings(positions) + self.token_type_embeddings(types)
        x = self.dropout(self.LayerNorm(x))
        if torch.is_autocast_enabled() and x.is_cuda:  # the encoder's activations travel in the compute dtype
            x = x.to(torch.get_autocast_dtype("cuda"))
        return x


class BertLayer(nn.Module):
    def __init__(self, cfg: BertConfig) -> None:
        super().__init__()
        self.num_heads = cfg.num_attention_heads
        self.qkv = LinearAct(cfg.hidden_size, 3 * cfg.hidden_size)
        self.attn_out = LinearAct(cfg.hidden_size, cfg.hidden_size)
        self.attn_norm = ResidualLayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps, dropout=cfg.hidden_dropout_prob)
        assert cfg.activation in ("gelu", "relu")
        self.ffn_in = LinearAct(cfg.hidden_size, cfg.intermediate_size, activation=cfg.activation)  # activation in the epilogue
        self.act = nn.Identity()
        self.ffn_out = LinearAct(cfg.intermediate_size, cfg.hidden_size)
        self.ffn_norm = ResidualLayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps, dropout=cfg.hidden_dropout_prob)

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None) -> torch.Tensor:
        # the packed projection goes to the attention kernel as is ([B, T, 3H]: Q, K, V are addressed inside it by TMA)
        context = packed_self_attention(self.qkv(x), mask, self.num_heads)
        x = self.attn_norm(self.attn_out(context), residual=x)  # LN(x + dropout(.)): one kernel
        return self.ffn_norm(self.ffn_out(self.act(self.ffn_in(x))), residual=x)


class BertEncoder(nn.Module):
    def __init__(self, cfg: BertConfig) -> None:
        super().__init__()
        self.config = cfg
        self.embeddings = BertEmbeddings(cfg)
        self.layers = nn.ModuleList(BertLayer(cfg) for _ in range(cfg.num_hidden_layers))

    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor | None = None,
                token_type_ids: torch.Tensor | None = None) -> torch.Tensor:
        mask = None
        if attention_mask is not None:  # [B, T] of {0,1}: key padding mask, one byte per token
            mask = attention_mask.to(torch.uint8).contiguous()
        x = self.embeddings(input_ids, token_type_ids)
        for layer in self.layers:
            x = layer(x, mask)
        return x


class BertForSequenceClassification(nn.Module):
    """[CLS] pooling -> tanh pooler -> dropout -> classifier; returns logits ``[B, num_labels]``."""

    def __init__(self, cfg: BertConfig, num_labels: int) -> None:
        super().__init__()
        self.bert = BertEncoder(cfg)
        self.pooler = LinearAct(cfg.hidden_size, cfg.hidden_size)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)
        self.classifier = LinearAct(cfg.hidden_size, num_labels)

    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor | None = None,
                token_type_ids: torch.Tensor | None = None) -> torch.Tensor:
        hidden = self.bert(input_ids, attention_mask, token_type_ids)
        pooled = torch.tanh(self.pooler(hidden[:, 0]))
        return self.classifier(self.dropout(pooled))
