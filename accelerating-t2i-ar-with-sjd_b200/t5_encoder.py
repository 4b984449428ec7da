"""Input side of LlamaGen's text-to-image path: the T5 text encoder behind
`T5Embedder.get_text_embeddings` (llamagen/language/t5.py:62-83 — `self.model(input_ids, attention_mask)
['last_hidden_state']` of HF's `T5EncoderModel`, google/flan-t5-xl) with every linear layer on this repository's
weight-streaming tcgen05 GEMM (`sjd_gemm_bf16`, csrc/gemm_fused.cu) instead of nn.Linear -> cuBLAS.

The encoder is an HBM-bound weight stream at these row counts (one or two captions of <= 120 tokens: 120-240 token rows
against 1.2 B parameters for flan-t5-xl), exactly the regime the GEMM kernel was built for: q/k/v fused into one [3 inner, d]
projection, wi_0 / wi_1 into one [2 d_ff, d] projection, fp32 accumulation, bf16 operands.  The small per-head attention
([T, T] scores with T5's bucketed relative position bias and the padding mask), the RMS layer norms, gelu_new and the
residual stream (fp32) are library / element-wise torch ops.

What is mirrored is HF's T5Stack encoder (transformers `modeling_t5.py`: T5LayerNorm, T5Attention without 1/sqrt(d) scaling,
`_relative_position_bucket(bidirectional=True)` shared from block 0, T5DenseGatedActDense / T5DenseActDense, final_layer_norm),
driven from the module's own state dict.  Tokenisation and caption cleaning (t5.py:85-200: AutoTokenizer, ftfy / bs4)
stay the reference's code.  No CPU path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import _lib

_MAX_ROWS = _lib.SJD_MAX_TOKENS   # token rows per GEMM launch


def _round16(n: int) -> int:
    return (n + 15) // 16 * 16


class T5EncoderB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], *, num_heads: int, d_kv: int,
                 relative_attention_num_buckets: int = 32, relative_attention_max_distance: int = 128,
                 layer_norm_epsilon: float = 1e-6, gated: Optional[bool] = None, act: str = "gelu_new", device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("T5EncoderB200 runs on the GPU only (sjd_gemm_bf16 has no CPU fallback)")
        self.device, self.H, self.dkv = dev, int(num_heads), int(d_kv)
        self.n_buckets, self.max_dist, self.eps, self.act = relative_attention_num_buckets, relative_attention_max_distance, layer_norm_epsilon, act
        sd = {k[len("encoder."):] if k.startswith("encoder.") else k: v for k, v in state_dict.items()}
        g = lambda k: sd[k].detach().to(dev)
        self.embed = g("embed_tokens.weight" if "embed_tokens.weight" in sd else "shared.weight").float()
        self.d = self.embed.shape[1]
        n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("block."))
        if gated is None:
            gated = "block.0.layer.1.DenseReluDense.wi_0.weight" in sd
        self.gated = gated
        self.rel_bias = g("block.0.layer.0.SelfAttention.relative_attention_bias.weight").float()   # [buckets, heads]
        self.layers = []
        bf = lambda t: t.to(torch.bfloat16).contiguous()
        for i in range(n_layers):
            a, f = f"block.{i}.layer.0.", f"block.{i}.layer.1."
            qkv = torch.cat([g(a + "SelfAttention.q.weight"), g(a + "SelfAttention.k.weight"), g(a + "SelfAttention.v.weight")], 0)
            if gated:
                wi = torch.cat([g(f + "DenseReluDense.wi_0.weight"), g(f + "DenseReluDense.wi_1.weight")], 0)
            else:
                wi = g(f + "DenseReluDense.wi.weight")
            self.layers.append(dict(ln1=g(a + "layer_norm.weight").float(), qkv=bf(qkv), o=bf(g(a + "SelfAttention.o.weight")),
                                    ln2=g(f + "layer_norm.weight").float(), wi=bf(wi), wo=bf(g(f + "DenseReluDense.wo.weight"))))
        self.final_ln = g("final_layer_norm.weight").float()
        self.inner = self.H * self.dkv
        self.d_ff = self.layers[0]["wo"].shape[1]
        for w in (self.d, self.inner, self.d_ff):
            if w % 64:
                raise ValueError(f"sjd_gemm_bf16 needs reduction dims that are multiples of 64, got {w}")
        L = _lib.lib()
        shapes = [(3 * self.inner, self.d), (self.d, self.inner), ((2 if gated else 1) * self.d_ff, self.d), (self.d, self.d_ff)]
        ws_bytes = max(L.sjd_gemm_workspace_bytes(N, K, _MAX_ROWS, 0) for N, K in shapes)
        self.ws = torch.zeros(ws_bytes, device=dev, dtype=torch.uint8)     # zero before first use; the kernel re-arms its counters
        self.xbuf = torch.zeros(_MAX_ROWS, max(self.d, self.inner, self.d_ff), device=dev, dtype=torch.bfloat16)

    @classmethod
    def from_module(cls, t5: torch.nn.Module, device=None) -> "T5EncoderB200":
        """From HF's T5EncoderModel (what T5Embedder holds as `.model`)."""
        c = t5.config
        if device is None:
            device = next(t5.parameters()).device
        return cls(t5.state_dict(), num_heads=c.num_heads, d_kv=c.d_kv,
                   relative_attention_num_buckets=c.relative_attention_num_buckets,
                   relative_attention_max_distance=getattr(c, "relative_attention_max_distance", 128),
                   layer_norm_epsilon=c.layer_norm_epsilon, gated=bool(getattr(c, "is_gated_act", False)),
                   act=getattr(c, "dense_act_fn", "relu"), device=device)

    # ---------------------------------------------------------------------------------------------- pieces
    def _linear(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """y[M, N] (fp32) = x[M, K] (rounded to bf16) @ w[N, K]^T on the tcgen05 weight streamer, <= 256 rows per launch."""
        M, K = x.shape
        N = w.shape[0]
        out = torch.empty(M, N, device=self.device, dtype=torch.float32)
        L, st = _lib.lib(), torch.cuda.current_stream(self.device).cuda_stream
        for r0 in range(0, M, _MAX_ROWS):
            m = min(_MAX_ROWS, M - r0)
            m_tile = _round16(m)
            xb = self.xbuf.view(-1)[: m_tile * K].view(m_tile, K)
            xb[:m].copy_(x[r0:r0 + m])
            if m_tile > m:
                xb[m:].zero_()
            _lib.check(L.sjd_gemm_bf16(w.data_ptr(), N, K, xb.data_ptr(), m_tile, m, out[r0:r0 + m].data_ptr(), 1, 0,
                                       self.ws.data_ptr(), 0, st), "sjd_gemm_bf16")
        return out

    def _ln(self, h: torch.Tensor, w: torch.Tensor) -> torch.Tensor:          # T5LayerNorm: no mean subtraction, no bias
        return w * (h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + self.eps))

    def _position_bias(self, T: int) -> torch.Tensor:                          # [H, T, T], bidirectional buckets
        ctx = torch.arange(T, device=self.device)[:, None]
        mem = torch.arange(T, device=self.device)[None, :]
        rel = mem - ctx
        nb = self.n_buckets // 2
        bucket = (rel > 0).long() * nb
        rel = rel.abs()
        max_exact = nb // 2
        large = max_exact + (torch.log(rel.float().clamp(min=1) / max_exact) / math.log(self.max_dist / max_exact)
                             * (nb - max_exact)).long()
        large = torch.minimum(large, torch.full_like(large, nb - 1))
        bucket = bucket + torch.where(rel < max_exact, rel, large)
        return self.rel_bias[bucket].permute(2, 0, 1)

    def _act(self, x: torch.Tensor) -> torch.Tensor:
        if self.act in ("gelu_new", "gelu_pytorch_tanh"):
            return F.gelu(x, approximate="tanh")
        if self.act == "gelu":
            return F.gelu(x)
        return F.relu(x)

    # ---------------------------------------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, graph: bool = True) -> torch.Tensor:
        """last_hidden_state [B, T, d] (fp32).  graph=True: the ~700 launches of a 24-layer forward are host-bound when
        enqueued one by one (7.7 ms for flan-t5-xl's shape against 0.36 ms of weight streaming), so the forward of a given
        [B, T] is captured into a CUDA graph on its second call and replayed afterwards (T5Embedder pads every caption to
        model_max_length, so the shape repeats)."""
        ids = input_ids.to(self.device)
        B, T = ids.shape
        mask = torch.ones(B, T, device=self.device) if attention_mask is None else attention_mask.to(self.device).float()
        if not graph:
            return self._forward_eager(ids, mask)
        key = (B, T)
        ent = self._graphs.get(key) if hasattr(self, "_graphs") else None
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        if ent is None:                              # first call: eager (also runs every lazy initialisation outside a capture)
            self._graphs[key] = "warm"
            return self._forward_eager(ids, mask)
        if ent == "warm":
            s_ids, s_mask = ids.clone(), mask.clone()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._forward_eager(s_ids, s_mask)   # allocator warm-up on the capture stream
            torch.cuda.current_stream(self.device).wait_stream(side)
            with torch.cuda.graph(g):
                s_out = self._forward_eager(s_ids, s_mask)
            ent = self._graphs[key] = (g, s_ids, s_mask, s_out)
        g, s_ids, s_mask, s_out = ent
        s_ids.copy_(ids)
        s_mask.copy_(mask)
        g.replay()
        return s_out.clone()

    def _forward_eager(self, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        B, T = ids.shape
        bias = self._position_bias(T)[None] + (1.0 - mask)[:, None, None, :] * torch.finfo(torch.float32).min   # [B, H, T, T]
        h = self.embed[ids].reshape(B * T, self.d)
        H, dk = self.H, self.dkv
        for ly in self.layers:
            qkv = self._linear(self._ln(h, ly["ln1"]), ly["qkv"]).view(B, T, 3, H, dk)
            q, k, v = (qkv[:, :, j].transpose(1, 2) for j in range(3))                       # [B, H, T, dk]
            p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) + bias, dim=-1)           # T5: no 1/sqrt(d_kv)
            ctx = torch.matmul(p, v).transpose(1, 2).reshape(B * T, H * dk)
            h = h + self._linear(ctx, ly["o"])
            u = self._linear(self._ln(h, ly["ln2"]), ly["wi"])
            u = self._act(u[:, : self.d_ff]) * u[:, self.d_ff:] if self.gated else self._act(u)
            h = h + self._linear(u, ly["wo"])
        return self._ln(h, self.final_ln).view(B, T, self.d)

    __call__ = forward

    def get_text_embeddings(self, input_ids: torch.Tensor, attention_mask: torch.Tensor):
        """The model half of T5Embedder.get_text_embeddings (t5.py:78-83): (embeddings, mask)."""
        return self.forward(input_ids, attention_mask), attention_mask.to(self.device)

