"""LlamaGen GPT as a PARAMETER TREE for the sm_100a engine.

Same constructor arguments, attribute names and state-dict keys as the reference's llamagen/llamagen.py
(`ModelArgs` :56-83, `Transformer` :297-335, model zoo :474-504), so `GPT_models[name](...)` followed by
`load_state_dict(checkpoint)` works unchanged (test_llamagen.py:74-100).  There is deliberately no PyTorch forward:
the decoder stack runs in libsjd_b200.so (accelerating-t2i-ar-with-sjd_b200/hf_api.pack_llamagen packs these
parameters); calling the module raises instead of silently computing on the wrong path.
"""
from __future__ import annotations

import os as _os
import sys as _sys
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))


@dataclass
class ModelArgs:
    dim: int = 4096
    n_layer: int = 32
    n_head: int = 32
    n_kv_head: Optional[int] = None
    multiple_of: int = 256
    ffn_dim_multiplier: Optional[float] = None
    rope_base: float = 10000
    norm_eps: float = 1e-5
    initializer_range: float = 0.02
    token_dropout_p: float = 0.1
    attn_dropout_p: float = 0.0
    resid_dropout_p: float = 0.1
    ffn_dropout_p: float = 0.1
    drop_path_rate: float = 0.0
    num_classes: int = 1000
    caption_dim: int = 2048
    class_dropout_prob: float = 0.1
    model_type: str = "c2i"
    vocab_size: int = 16384
    cls_token_num: int = 1
    block_size: int = 256
    max_batch_size: int = 32
    max_seq_len: int = 2048


def ffn_hidden_dim(cfg: ModelArgs) -> int:
    """SwiGLU width rule of the reference FeedForward (llamagen.py:187-192)."""
    hidden = int(2 * (4 * cfg.dim) / 3)
    if cfg.ffn_dim_multiplier is not None:
        hidden = int(cfg.ffn_dim_multiplier * hidden)
    return cfg.multiple_of * ((hidden + cfg.multiple_of - 1) // cfg.multiple_of)


class _Weight(nn.Module):
    """RMSNorm parameter holder (`.weight`)."""

    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _Attention(nn.Module):
    def __init__(self, cfg: ModelArgs):
        super().__init__()
        n_kv = cfg.n_kv_head if cfg.n_kv_head is not None else cfg.n_head
        head_dim = cfg.dim // cfg.n_head
        self.wqkv = nn.Linear(cfg.dim, (cfg.n_head + 2 * n_kv) * head_dim, bias=False)
        self.wo = nn.Linear(cfg.dim, cfg.dim, bias=False)


class _FeedForward(nn.Module):
    def __init__(self, cfg: ModelArgs):
        super().__init__()
        hidden = ffn_hidden_dim(cfg)
        self.w1 = nn.Linear(cfg.dim, hidden, bias=False)
        self.w3 = nn.Linear(cfg.dim, hidden, bias=False)
        self.w2 = nn.Linear(hidden, cfg.dim, bias=False)


class _Block(nn.Module):
    def __init__(self, cfg: ModelArgs):
        super().__init__()
        self.attention = _Attention(cfg)
        self.feed_forward = _FeedForward(cfg)
        self.attention_norm = _Weight(cfg.dim)
        self.ffn_norm = _Weight(cfg.dim)


class LabelEmbedder(nn.Module):
    """Class-label embedding with one extra row for the CFG null class (llamagen.py:89-116)."""

    def __init__(self, num_classes, hidden_size, dropout_prob):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), hidden_size)
        self.num_classes = num_classes
        self.dropout_prob = dropout_prob

    def forward(self, labels, train=False, force_drop_ids=None):
        return self.embedding_table(labels).unsqueeze(1)


class _CapProj(nn.Module):
    def __init__(self, in_features, hidden_features, out_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features, bias=False)
        self.act = nn.GELU(approximate="tanh")
        self.fc2 = nn.Linear(hidden_features, out_features, bias=False)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class CaptionEmbedder(nn.Module):
    """T5-feature projection for text-conditional GPTs (llamagen.py:121-147).  This small MLP runs once per image,
    before the hot path, and stays in PyTorch."""

    def __init__(self, in_channels, hidden_size, uncond_prob, token_num=120):
        super().__init__()
        self.cap_proj = _CapProj(in_channels, hidden_size, hidden_size)
        self.register_buffer("uncond_embedding", torch.randn(token_num, in_channels) / in_channels ** 0.5)
        self.uncond_prob = uncond_prob

    def forward(self, caption, train=False, force_drop_ids=None):
        return self.cap_proj(caption)


class Transformer(nn.Module):
    def __init__(self, config: ModelArgs):
        super().__init__()
        self.config = config
        self.vocab_size = config.vocab_size
        self.n_layer = config.n_layer
        self.block_size = config.block_size
        self.num_classes = config.num_classes
        self.model_type = config.model_type
        self.cls_token_num = config.cls_token_num
        if self.model_type == "c2i":
            self.cls_embedding = LabelEmbedder(config.num_classes, config.dim, config.class_dropout_prob)
        elif self.model_type == "t2i":
            self.cls_embedding = CaptionEmbedder(config.caption_dim, config.dim, config.class_dropout_prob)
        else:
            raise Exception("please check model type")
        self.tok_embeddings = nn.Embedding(config.vocab_size, config.dim)
        self.layers = nn.ModuleList(_Block(config) for _ in range(config.n_layer))
        self.norm = _Weight(config.dim)
        self.output = nn.Linear(config.dim, config.vocab_size, bias=False)
        grid = int(self.block_size ** 0.5)
        assert grid * grid == self.block_size
        self.max_batch_size = -1
        self.max_seq_length = -1
        self.initialize_weights()

    def initialize_weights(self):
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, mean=0.0, std=std)
        nn.init.constant_(self.output.weight, 0)   # like the reference (llamagen.py:342)

    def setup_caches(self, max_batch_size, max_seq_length, dtype=None):
        """The engine owns the (static) KV cache; only the sizes are recorded (reference: llamagen.py:353-367)."""
        self.max_batch_size, self.max_seq_length = max_batch_size, max_seq_length

    def forward(self, *args, **kwargs):
        raise RuntimeError("the LlamaGen decoder stack runs in the sm_100a engine (libsjd_b200.so); use "
                           "llamagen.llamagen_solver.LlamaGenSolver.generate — there is no PyTorch forward")


def GPT_7B(**kw): return Transformer(ModelArgs(n_layer=32, n_head=32, dim=4096, **kw))
def GPT_3B(**kw): return Transformer(ModelArgs(n_layer=24, n_head=32, dim=3200, **kw))
def GPT_1B(**kw): return Transformer(ModelArgs(n_layer=22, n_head=32, dim=2048, **kw))
def GPT_XXXL(**kw): return Transformer(ModelArgs(n_layer=48, n_head=40, dim=2560, **kw))
def GPT_XXL(**kw): return Transformer(ModelArgs(n_layer=48, n_head=24, dim=1536, **kw))
def GPT_XL(**kw): return Transformer(ModelArgs(n_layer=36, n_head=20, dim=1280, **kw))
def GPT_L(**kw): return Transformer(ModelArgs(n_layer=24, n_head=16, dim=1024, **kw))
def GPT_B(**kw): return Transformer(ModelArgs(n_layer=12, n_head=12, dim=768, **kw))


GPT_models = {"GPT-B": GPT_B, "GPT-L": GPT_L, "GPT-XL": GPT_XL, "GPT-XXL": GPT_XXL, "GPT-XXXL": GPT_XXXL,
              "GPT-1B": GPT_1B, "GPT-3B": GPT_3B, "GPT-7B": GPT_7B}
