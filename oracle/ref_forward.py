"""ORACLE (test infrastructure, not product code): plain-PyTorch restatement of the reference's
decoder-stack forward for one Jacobi draft window over a KV cache.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
It restates, in fp32 with optional emulation of the bf16 rounding points of a bf16 checkpoint:

  * Chameleon / Lumina-mGPT decoder layer   lumina_mgpt/model/chameleon/modeling_chameleon.py:59-76 (RMSNorm),
    :198-219 (per-head QK LayerNorm), :84-110,:153-177 (rotate-half RoPE), :499-581 (attention over the cache),
    :181-195 (SwiGLU), :606-666 (residual wiring), :1358,:1560-1561 (final norm, fp32 logits)
  * LlamaGen GPT block                      llamagen/llamagen.py:170-181, :222-278, :184-200, :281-294, :441-467
    (2-D RoPE on adjacent pairs, zero rotation for condition tokens)
  * Emu3 (Llama-style GQA)                  emu3/mllm/modeling_emu3.py:106-120, :214-239, :660-744, :754-824
  * the Jacobi window mask                  scheduler/jacobi_iteration_lumina_mgpt.py:1256-1336 — key j is visible
    to the window query at cache slot t iff kv_lo[row] <= j <= t (3-D 0/1 mask AND `j <= cache_position`).

Pinned against the reference's own model code run on CPU in the build container by oracle/mint_golden.py
(tests/golden/forward_*.npz); see DESIGN.md §5.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch


@dataclass
class StackConfig:
    n_layers: int
    d_model: int
    n_heads: int
    n_kv_heads: int
    head_dim: int
    d_ff: int
    vocab: int
    rms_eps: float = 1e-5
    qk_norm: bool = False          # Chameleon per-head LayerNorm on q and k
    rope_interleaved: bool = False  # LlamaGen: rotate adjacent pairs; else rotate-half
    rope_theta: float = 10000.0
    family: str = "chameleon"      # 'chameleon' | 'llamagen' | 'emu3'
    extra: dict = field(default_factory=dict)


def bf16r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def rope_tables_rotate_half(head_dim: int, n_pos: int, theta: float, round_bf16: bool):
    """cos/sin [n_pos, head_dim/2] as ChameleonRotaryEmbedding builds them (modeling_chameleon.py:97-110);
    the reference casts them to the activation dtype (bf16) before use."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(n_pos, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    cos, sin = freqs.cos(), freqs.sin()
    if round_bf16:
        cos, sin = bf16r(cos), bf16r(sin)
    return cos.contiguous(), sin.contiguous()


def rope_tables_llamagen_2d(grid_size: int, head_dim: int, base: float, cls_token_num: int):
    """cos/sin [cls_token_num + grid^2, head_dim/2] following llamagen/llamagen.py:441-454; rows of condition
    tokens are all-zero (cos = sin = 0), which zeroes their q and k."""
    half_dim = head_dim // 2
    freqs = 1.0 / (base ** (torch.arange(0, half_dim, 2)[: (half_dim // 2)].float() / half_dim))
    t = torch.arange(grid_size)
    freqs = torch.outer(t, freqs)
    grid = torch.concat([
        freqs[:, None, :].expand(-1, grid_size, -1),
        freqs[None, :, :].expand(grid_size, -1, -1),
    ], dim=-1)
    cos = torch.cos(grid).flatten(0, 1)
    sin = torch.sin(grid).flatten(0, 1)
    z = torch.zeros(cls_token_num, head_dim // 2)
    return torch.cat([z, cos]).contiguous(), torch.cat([z, sin]).contiguous()


def random_weights(cfg: StackConfig, seed: int = 0, std: float = 0.02, device="cpu") -> dict:
    """Random-init stack at the given shapes (N(0, std), values exactly representable in bf16)."""
    g = torch.Generator().manual_seed(seed)

    def mat(*shape, s=std):
        return bf16r(torch.randn(*shape, generator=g) * s).to(device)

    H, Hkv, Dh, d = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.d_model
    w = {"embed": mat(cfg.vocab, d), "final_norm": bf16r(1.0 + 0.1 * torch.randn(d, generator=g)).to(device),
         "lm_head": mat(cfg.vocab, d), "layers": []}
    for _ in range(cfg.n_layers):
        L = {
            "attn_norm": bf16r(1.0 + 0.1 * torch.randn(d, generator=g)).to(device),
            "wqkv": mat((H + 2 * Hkv) * Dh, d),
            "wo": mat(d, H * Dh),
            "ffn_norm": bf16r(1.0 + 0.1 * torch.randn(d, generator=g)).to(device),
            "w_gate_up": mat(2 * cfg.d_ff, d),
            "w_down": mat(d, cfg.d_ff),
        }
        if cfg.qk_norm:
            L["q_norm_w"] = bf16r(1.0 + 0.1 * torch.randn(H, Dh, generator=g)).to(device)
            L["q_norm_b"] = bf16r(0.1 * torch.randn(H, Dh, generator=g)).to(device)
            L["k_norm_w"] = bf16r(1.0 + 0.1 * torch.randn(Hkv, Dh, generator=g)).to(device)
            L["k_norm_b"] = bf16r(0.1 * torch.randn(Hkv, Dh, generator=g)).to(device)
        w["layers"].append(L)
    return w


class RefStack:
    """fp32 reference of the decoder stack with a static KV cache and the Jacobi window mask."""

    def __init__(self, cfg: StackConfig, weights: dict, rope_cos: torch.Tensor, rope_sin: torch.Tensor,
                 rows: int, max_len: int, emulate_bf16: bool = True, logits_round_bf16: bool = True):
        self.cfg, self.w = cfg, weights
        self.cos, self.sin = rope_cos.float(), rope_sin.float()
        self.rows, self.max_len = rows, max_len
        self.emulate = emulate_bf16
        self.logits_round = logits_round_bf16
        dev = weights["lm_head"].device
        self.k = [torch.zeros(rows, cfg.n_kv_heads, max_len, cfg.head_dim, device=dev) for _ in range(cfg.n_layers)]
        self.v = [torch.zeros(rows, cfg.n_kv_heads, max_len, cfg.head_dim, device=dev) for _ in range(cfg.n_layers)]

    def r(self, x):
        return bf16r(x) if self.emulate else x

    def rmsnorm(self, x, w):
        # modeling_chameleon.py:68-73 / llamagen.py:176-181: normalise in fp32, cast down, then scale
        xf = x.float()
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.cfg.rms_eps)
        return self.r(w * self.r(y))

    def rope(self, x, pos):
        # x [rows, W, heads, Dh]; pos [rows, W]
        cos = self.cos.to(x.device)[pos][:, :, None, :]
        sin = self.sin.to(x.device)[pos][:, :, None, :]
        if self.cfg.rope_interleaved:  # llamagen.py:457-467
            xs = x.reshape(*x.shape[:-1], -1, 2)
            o = torch.stack([xs[..., 0] * cos - xs[..., 1] * sin, xs[..., 1] * cos + xs[..., 0] * sin], dim=-1)
            return o.flatten(3)
        h = x.shape[-1] // 2  # modeling_chameleon.py:146-177
        x1, x2 = x[..., :h], x[..., h:]
        return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)

    @torch.no_grad()
    def forward(self, ids=None, embeds=None, rope_pos=None, kv_len: int = 0, kv_lo=None, cache_pos=None,
                n_logit_tokens=None):
        cfg, w = self.cfg, self.w
        H, Hkv, Dh = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        x = w["embed"][ids] if embeds is None else embeds.float()
        rows, W, _ = x.shape
        dev = x.device
        if kv_lo is None:
            kv_lo = [0] * rows
        if cache_pos is None:
            cache_pos = torch.arange(kv_len, kv_len + W, device=dev)[None].expand(rows, W)
        T = kv_len + W
        j = torch.arange(T, device=dev)
        t_q = torch.arange(kv_len, kv_len + W, device=dev)
        lo = torch.tensor(kv_lo, device=dev)
        vis = (j[None, None, :] <= t_q[None, :, None]) & (j[None, None, :] >= lo[:, None, None])  # [rows, W, T]
        x = self.r(x)
        for l, L in enumerate(w["layers"]):
            xn = self.rmsnorm(x, L["attn_norm"])
            qkv = self.r(xn @ L["wqkv"].T)
            q = qkv[..., : H * Dh].reshape(rows, W, H, Dh)
            k = qkv[..., H * Dh: (H + Hkv) * Dh].reshape(rows, W, Hkv, Dh)
            v = qkv[..., (H + Hkv) * Dh:].reshape(rows, W, Hkv, Dh)
            if cfg.qk_norm:  # modeling_chameleon.py:216-219 (autocast keeps layer_norm in fp32)
                q = torch.nn.functional.layer_norm(q, (Dh,), None, None, 1e-5) * L["q_norm_w"] + L["q_norm_b"]
                k = torch.nn.functional.layer_norm(k, (Dh,), None, None, 1e-5) * L["k_norm_w"] + L["k_norm_b"]
            q = self.r(self.rope(q, rope_pos))
            k = self.r(self.rope(k, rope_pos))
            for b in range(rows):
                self.k[l][b, :, cache_pos[b]] = k[b].transpose(0, 1)
                self.v[l][b, :, cache_pos[b]] = v[b].transpose(0, 1)
            kk = self.k[l][:, :, :T].repeat_interleave(H // Hkv, dim=1)  # [rows, H, T, Dh]
            vv = self.v[l][:, :, :T].repeat_interleave(H // Hkv, dim=1)
            s = torch.einsum("bwhd,bhtd->bhwt", q, kk) / math.sqrt(Dh)
            s = s.masked_fill(~vis[:, None], float("-inf"))
            p = torch.softmax(s, dim=-1)
            p = torch.nan_to_num(p, nan=0.0)  # fully hidden query rows (CFG prefix) produce no output
            o = torch.einsum("bhwt,bhtd->bwhd", self.r(p), vv).reshape(rows, W, H * Dh)
            x = self.r(x + self.r(self.r(o) @ L["wo"].T))
            xn = self.rmsnorm(x, L["ffn_norm"])
            gu = self.r(xn @ L["w_gate_up"].T)
            g, u = gu[..., : cfg.d_ff], gu[..., cfg.d_ff:]
            act = self.r(self.r(torch.nn.functional.silu(g)) * u)
            x = self.r(x + self.r(act @ L["w_down"].T))
        xn = self.rmsnorm(x, w["final_norm"])
        if n_logit_tokens is not None:
            xn = xn[:, -n_logit_tokens:]
        logits = xn @ w["lm_head"].T
        return bf16r(logits) if (self.emulate and self.logits_round) else logits
