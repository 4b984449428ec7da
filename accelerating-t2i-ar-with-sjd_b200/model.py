"""Device-side model context: packs a decoder stack's weights into the layout the sm_100a kernels
stream (bf16, nn.Linear [out, in] rows, fused qkv / gate-up) and drives ``sjd_ctx_forward``.

Host-side mirror of what the reference does inside ``outputs = self(**model_inputs)``
(scheduler/jacobi_iteration_lumina_mgpt.py:1107): one call = one draft-window forward of every CFG row
over the static KV cache.  Roll-back of rejected drafts (reference: delete_false_key_value, :47-54) needs no
device work here — the next call simply passes a smaller ``kv_len`` and overwrites the stale slots.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class StackShape:
    n_layers: int
    d_model: int
    n_heads: int
    n_kv_heads: int
    head_dim: int
    d_ff: int
    vocab: int
    rms_eps: float = 1e-5
    qk_norm: bool = False
    rope_interleaved: bool = False


class DeviceStack:
    """Owns the C context + the torch tensors whose storage the context points into."""

    def __init__(self, shape: StackShape, weights: dict, rope_cos: torch.Tensor, rope_sin: torch.Tensor,
                 rows: int, max_len: int, device="cuda:0", logits_round_bf16: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("sjd_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.lib()
        self.shape, self.rows, self.max_len = shape, rows, max_len
        self.device = torch.device(device)
        dev = self.device
        bf = torch.bfloat16

        def put(t, dtype=bf):   # staging only: the context copies / re-lays out everything it is given
            return t.detach().to(device=dev, dtype=dtype).contiguous()

        self.rope_cos = put(rope_cos, torch.float32)
        self.rope_sin = put(rope_sin, torch.float32)
        assert self.rope_cos.shape[1] == shape.head_dim // 2
        self.n_rope_pos = int(self.rope_cos.shape[0])
        cfg = _lib.ModelCfg(shape.n_layers, shape.d_model, shape.n_heads, shape.n_kv_heads, shape.head_dim,
                            shape.d_ff, shape.vocab, shape.rms_eps, int(shape.qk_norm),
                            int(shape.rope_interleaved), rows, max_len, self.rope_cos.shape[0],
                            int(logits_round_bf16))
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self.lib.sjd_ctx_create(C.byref(cfg), C.byref(h)), "sjd_ctx_create")
        self.ctx = h
        with torch.cuda.device(dev):
            for l, L in enumerate(weights["layers"]):
                lw = _lib.LayerWeights()
                names = ["attn_norm", "wqkv", "wo", "ffn_norm", "w_gate_up", "w_down"]
                if shape.qk_norm:
                    names += ["q_norm_w", "q_norm_b", "k_norm_w", "k_norm_b"]
                staged = {name: put(L[name]) for name in names}
                for name, t in staged.items():
                    setattr(lw, name, t.data_ptr())
                _lib.check(self.lib.sjd_ctx_set_layer(self.ctx, l, C.byref(lw)), "sjd_ctx_set_layer")  # syncs
                del staged
            embed = put(weights["embed"]) if weights.get("embed") is not None else None
            final_norm, lm_head = put(weights["final_norm"]), put(weights["lm_head"])
            _lib.check(self.lib.sjd_ctx_set_globals(
                self.ctx, embed.data_ptr() if embed is not None else None, final_norm.data_ptr(),
                lm_head.data_ptr(), self.rope_cos.data_ptr(), self.rope_sin.data_ptr()), "sjd_ctx_set_globals")
            torch.cuda.synchronize(dev)
        self.logits_buf = torch.empty(_lib.SJD_MAX_TOKENS, shape.vocab, dtype=torch.float32, device=dev)

    def device_bytes(self) -> int:
        return int(self.lib.sjd_ctx_device_bytes(self.ctx))

    def forward(self, W: int, rope_pos: torch.Tensor, cache_pos: torch.Tensor, kv_len: int, kv_lo,
                ids: torch.Tensor | None = None, embeds: torch.Tensor | None = None,
                n_logit_tokens: int | None = None, out: torch.Tensor | None = None,
                stream_handle: int | None = None) -> torch.Tensor:
        """ids / rope_pos / cache_pos: int32 device tensors [rows*W] (row-major).  Returns fp32 logits
        [rows, n_logit_tokens, vocab] (a view of an internal buffer unless ``out`` is given).  stream_handle: the raw
        cudaStream_t to launch on (default: torch's current stream; looking it up costs ~8 us of host time per call, which
        the decode loop — where it sits between one iteration's result and the next iteration's first kernel — avoids)."""
        n = W if n_logit_tokens is None else n_logit_tokens
        a = _lib.ForwardArgs()
        a.W = W
        a.ids = ids.data_ptr() if ids is not None else None
        a.embeds = embeds.data_ptr() if embeds is not None else None
        a.rope_pos = rope_pos.data_ptr()
        a.cache_pos = cache_pos.data_ptr()
        a.kv_len = kv_len
        for b in range(self.rows):
            a.kv_lo[b] = int(kv_lo[b])
        a.n_logit_tokens = n
        buf = self.logits_buf if out is None else out
        a.logits = buf.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream if stream_handle is None else stream_handle
        _lib.check(self.lib.sjd_ctx_forward(self.ctx, C.byref(a), C.c_void_p(stream)), "sjd_ctx_forward")
        return buf[: self.rows * n].view(self.rows, n, self.shape.vocab)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.sjd_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
