"""Decoder-stack shapes of the model families the reference drives, RoPE tables, and random-init weights at
those shapes (no checkpoints are reachable offline).

Shapes: Lumina-mGPT-7B / Chameleon-7B  lumina_mgpt/model/chameleon/configuration_chameleon.py:191-219
        LlamaGen GPT-B .. GPT-7B      llamagen/llamagen.py:474-504 (ffn dim: :187-192)
        Emu3-Gen                      emu3/mllm/configuration_emu3.py:128-155
"""
from __future__ import annotations

import torch

from .model import StackShape


def lumina_7b() -> StackShape:
    return StackShape(n_layers=32, d_model=4096, n_heads=32, n_kv_heads=32, head_dim=128, d_ff=11008, vocab=65536,
                      rms_eps=1e-5, qk_norm=True, rope_interleaved=False)


def emu3_gen() -> StackShape:
    return StackShape(n_layers=32, d_model=4096, n_heads=32, n_kv_heads=8, head_dim=128, d_ff=14336, vocab=184622,
                      rms_eps=1e-5, qk_norm=False, rope_interleaved=False)


def _llamagen_ffn(dim: int, multiple_of: int = 256) -> int:
    hidden = int(2 * (4 * dim) / 3)
    return hidden if hidden % multiple_of == 0 else hidden + multiple_of - hidden % multiple_of


_LLAMAGEN = {"GPT-B": (12, 12, 768), "GPT-L": (24, 16, 1024), "GPT-XL": (36, 20, 1280), "GPT-XXL": (48, 24, 1536),
             "GPT-XXXL": (48, 40, 2560), "GPT-1B": (22, 32, 2048), "GPT-3B": (24, 32, 3200), "GPT-7B": (32, 32, 4096)}


def llamagen(name: str = "GPT-B", vocab: int = 16384) -> StackShape:
    n_layer, n_head, dim = _LLAMAGEN[name]
    return StackShape(n_layers=n_layer, d_model=dim, n_heads=n_head, n_kv_heads=n_head, head_dim=dim // n_head,
                      d_ff=_llamagen_ffn(dim), vocab=vocab, rms_eps=1e-5, qk_norm=False, rope_interleaved=True)


def rope_rotate_half(head_dim: int, n_pos: int, theta: float = 10000.0, round_bf16: bool = True):
    """cos/sin [n_pos, head_dim/2] (modeling_chameleon.py:97-110); the reference casts them to bf16."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = torch.arange(n_pos, dtype=torch.float32)[:, None] * inv_freq[None, :]
    cos, sin = freqs.cos(), freqs.sin()
    if round_bf16:
        cos, sin = cos.bfloat16().float(), sin.bfloat16().float()
    return cos.contiguous(), sin.contiguous()


def rope_llamagen_2d(grid_size: int, head_dim: int, base: float = 10000.0, cls_token_num: int = 1):
    """cos/sin [cls_token_num + grid^2, head_dim/2] (llamagen/llamagen.py:441-454); condition rows are zero."""
    half = head_dim // 2
    freqs = 1.0 / (base ** (torch.arange(0, half, 2)[: half // 2].float() / half))
    f = torch.outer(torch.arange(grid_size), freqs)
    grid = torch.concat([f[:, None, :].expand(-1, grid_size, -1), f[None, :, :].expand(grid_size, -1, -1)], dim=-1)
    z = torch.zeros(cls_token_num, head_dim // 2)
    return (torch.cat([z, torch.cos(grid).flatten(0, 1)]).contiguous(),
            torch.cat([z, torch.sin(grid).flatten(0, 1)]).contiguous())


def random_weights(shape: StackShape, seed: int = 0, std: float = 0.02, device="cuda:0", embed: bool = True) -> dict:
    """N(0, std) bf16 weights generated directly on the device (13.5 GB for the 7B shapes)."""
    g = torch.Generator(device=device).manual_seed(seed)
    bf = torch.bfloat16

    def mat(*s):
        return (torch.randn(*s, generator=g, device=device, dtype=torch.float32) * std).to(bf)

    def ones(*s):
        return (1.0 + 0.05 * torch.randn(*s, generator=g, device=device)).to(bf)

    H, Hkv, Dh, d = shape.n_heads, shape.n_kv_heads, shape.head_dim, shape.d_model
    w = {"embed": mat(shape.vocab, d) if embed else None, "final_norm": ones(d), "lm_head": mat(shape.vocab, d),
         "layers": []}
    for _ in range(shape.n_layers):
        L = {"attn_norm": ones(d), "wqkv": mat((H + 2 * Hkv) * Dh, d), "wo": mat(d, H * Dh), "ffn_norm": ones(d),
             "w_gate_up": mat(2 * shape.d_ff, d), "w_down": mat(d, shape.d_ff)}
        if shape.qk_norm:
            L.update(q_norm_w=ones(H, Dh), q_norm_b=(0.05 * torch.randn(H, Dh, generator=g, device=device)).to(bf),
                     k_norm_w=ones(Hkv, Dh), k_norm_b=(0.05 * torch.randn(Hkv, Dh, generator=g, device=device)).to(bf))
        w["layers"].append(L)
    return w
