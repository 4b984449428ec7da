"""Output side of the decode loop: image-token ids -> pixels with the VQGAN decoders the reference
ships, driven from the weights of the reference's own modules / checkpoints.

    LlamaGen   `vq_model.decode_code(index_sample, qzshape)`  (test_llamagen.py:182;
               llamagen/tokenizer/tokenizer_image/vq_model.py:52-55 decode_code, :126-190 Decoder, :261-275 codebook)
    Chameleon / Lumina-mGPT   `get_codebook_entry(tokens, (1, h, w, emb)) -> decode`
               (lumina_mgpt/model/chameleon_vae_ori/image_tokenizer.py:116-121; vqgan.py:131-146, :410-530 Decoder, :589-592)

Both decoders are the taming-transformers design (conv_in, residual / attention middle, up-sampling levels of residual
blocks, GroupNorm + swish + conv_out) under two different module naming schemes.  `VQDecoder` reads the STRUCTURE from the
state dict's keys (levels, blocks per level, where attention sits, shortcuts), so it needs no config object, and runs

  * the token side in ONE hand-written kernel through the C ABI (`sjd_vq_lookup`, csrc/vq_lookup.cu: codebook gather, the
    LlamaGen codebook's L2 normalisation, the [B, C, h, w] layout and post_quant_conv's 1x1 convolution), and
  * the convolution stack as plain library calls (cuDNN convolutions / GroupNorm / batched GEMMs through
    torch.nn.functional), fp32 like the reference.

    Emu3       `Emu3VisionVQModel.decode(codes)` for images  (emu3/tokenizer/modeling_emu3visionvq.py:790-815 decode,
               :596-722 Emu3VisionVQDecoder: a temporal stack of causal 3-D convolutions in front of a taming-style 2-D
               decoder whose normalisations are conditioned on the quantised latents, MoVQ style) -> `Emu3VQDecoder`

No CPU path: the kernel needs the CUDA library; tensors must live on the GPU.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib


def _swish(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(x)


class VQDecoder:
    """Token ids -> images in [-1, 1].  Build with `from_module(vq_model)` or from a state dict."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, l2_norm: Optional[bool] = None):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("VQDecoder runs on the GPU only (sjd_vq_lookup has no CPU fallback)")
        self.device = dev
        sd = {k: v.detach().to(dev, torch.float32).contiguous() for k, v in state_dict.items()
              if k.startswith(("decoder.", "post_quant_conv.", "quantize.embedding."))}
        if "quantize.embedding.weight" not in sd or "post_quant_conv.weight" not in sd or "decoder.conv_in.weight" not in sd:
            raise ValueError("state dict holds no VQGAN decoder (quantize.embedding / post_quant_conv / decoder.conv_in)")
        self.sd = sd
        self.layout = "llamagen" if any(k.startswith("decoder.conv_blocks.") for k in sd) else "chameleon"
        # LlamaGen's codebook is L2-normalised at look-up time (vq_model.py:261-265), taming's / Chameleon's is not
        self.l2_norm = (self.layout == "llamagen") if l2_norm is None else bool(l2_norm)
        self.codebook = sd["quantize.embedding.weight"]
        self.pq_w = sd["post_quant_conv.weight"].reshape(sd["post_quant_conv.weight"].shape[0], -1).contiguous()
        self.pq_b = sd["post_quant_conv.bias"]
        self.program = self._plan()

    # ------------------------------------------------------------------------------------------ construction
    @classmethod
    def from_module(cls, vq_model: torch.nn.Module, device=None, l2_norm: Optional[bool] = None) -> "VQDecoder":
        """From the reference's VQModel (either family).  `l2_norm` defaults to the module's own config when it has one."""
        if l2_norm is None:
            cfg = getattr(vq_model, "config", None)
            if cfg is not None and hasattr(cfg, "codebook_l2_norm"):
                l2_norm = bool(cfg.codebook_l2_norm)
        if device is None:
            device = next(vq_model.parameters()).device
        return cls(vq_model.state_dict(), device, l2_norm)

    def _plan(self) -> List[Tuple[str, str]]:
        """The decoder as a list of (op, key prefix): op in res | attn | up."""
        keys = list(self.sd)

        def idx(pattern: str) -> List[int]:
            return sorted({int(m.group(1)) for k in keys for m in [re.match(pattern, k)] if m})

        prog: List[Tuple[str, str]] = []
        if self.layout == "llamagen":   # decoder.mid.{0,1,2}; decoder.conv_blocks.{i} stored in execution order
            prog += [("res", "decoder.mid.0"), ("attn", "decoder.mid.1"), ("res", "decoder.mid.2")]
            for i in idx(r"decoder\.conv_blocks\.(\d+)\."):
                base = f"decoder.conv_blocks.{i}"
                attn = set(idx(rf"decoder\.conv_blocks\.{i}\.attn\.(\d+)\."))
                for j in idx(rf"decoder\.conv_blocks\.{i}\.res\.(\d+)\."):
                    prog.append(("res", f"{base}.res.{j}"))
                    if j in attn:
                        prog.append(("attn", f"{base}.attn.{j}"))
                if f"{base}.upsample.conv.weight" in self.sd:
                    prog.append(("up", f"{base}.upsample"))
        else:                           # decoder.mid.block_1 / attn_1 / block_2; decoder.up.{level}, executed from the last level down
            prog += [("res", "decoder.mid.block_1"), ("attn", "decoder.mid.attn_1"), ("res", "decoder.mid.block_2")]
            for i in reversed(idx(r"decoder\.up\.(\d+)\.")):
                base = f"decoder.up.{i}"
                attn = set(idx(rf"decoder\.up\.{i}\.attn\.(\d+)\."))
                for j in idx(rf"decoder\.up\.{i}\.block\.(\d+)\."):
                    prog.append(("res", f"{base}.block.{j}"))
                    if j in attn:
                        prog.append(("attn", f"{base}.attn.{j}"))
                if f"{base}.upsample.conv.weight" in self.sd:
                    prog.append(("up", f"{base}.upsample"))
        return prog

    # ------------------------------------------------------------------------------------------ building blocks
    def _conv(self, x, prefix, padding):
        return F.conv2d(x, self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], padding=padding)

    def _gn(self, x, prefix):
        return F.group_norm(x, 32, self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], eps=1e-6)

    def _res(self, x, p):
        h = self._conv(_swish(self._gn(x, p + ".norm1")), p + ".conv1", 1)
        h = self._conv(_swish(self._gn(h, p + ".norm2")), p + ".conv2", 1)
        if p + ".nin_shortcut.weight" in self.sd:
            x = self._conv(x, p + ".nin_shortcut", 0)
        elif p + ".conv_shortcut.weight" in self.sd:
            x = self._conv(x, p + ".conv_shortcut", 1)
        return x + h

    def _attn(self, x, p):
        b, c, hh, ww = x.shape
        n = self._gn(x, p + ".norm")
        q = self._conv(n, p + ".q", 0).reshape(b, c, hh * ww)
        k = self._conv(n, p + ".k", 0).reshape(b, c, hh * ww)
        v = self._conv(n, p + ".v", 0).reshape(b, c, hh * ww)
        w_ = torch.softmax(torch.bmm(q.transpose(1, 2), k) * (int(c) ** -0.5), dim=2)   # [b, query, key]
        o = torch.bmm(v, w_.transpose(1, 2)).reshape(b, c, hh, ww)
        return x + self._conv(o, p + ".proj_out", 0)

    def _up(self, x, p):
        return self._conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), p + ".conv", 1)

    # ------------------------------------------------------------------------------------------ the two entry points
    @torch.no_grad()
    def latents(self, codes: torch.Tensor, batch: int, h: int, w: int) -> torch.Tensor:
        """[batch, z, h, w] = post_quant_conv(codebook[codes]) in one kernel."""
        codes = torch.as_tensor(codes)
        if codes.numel() != batch * h * w:
            raise ValueError(f"{codes.numel()} codes for a {batch} x {h} x {w} latent grid")
        n_e, e_dim = self.codebook.shape
        c64 = codes.reshape(-1).to(self.device, torch.int64)
        if int(c64.min()) < 0 or int(c64.max()) >= n_e:
            raise ValueError(f"image-token id outside the codebook [0, {n_e})")
        c32 = c64.to(torch.int32).contiguous()
        z = self.pq_w.shape[0]
        out = torch.empty(batch, z, h, w, device=self.device, dtype=torch.float32)
        L = _lib.lib()
        _lib.check(L.sjd_vq_lookup(c32.data_ptr(), batch * h * w, h * w, self.codebook.data_ptr(), n_e, e_dim,
                                   int(self.l2_norm), self.pq_w.data_ptr(), self.pq_b.data_ptr(), z, out.data_ptr(),
                                   torch.cuda.current_stream(self.device).cuda_stream), "sjd_vq_lookup")
        return out

    @torch.no_grad()
    def decode_latents(self, z: torch.Tensor) -> torch.Tensor:
        x = self._conv(z, "decoder.conv_in", 1)
        for op, p in self.program:
            x = self._res(x, p) if op == "res" else (self._attn(x, p) if op == "attn" else self._up(x, p))
        return self._conv(_swish(self._gn(x, "decoder.norm_out")), "decoder.conv_out", 1)

    def decode_code(self, code_b, shape: Optional[Sequence[int]] = None, channel_first: bool = True) -> torch.Tensor:
        """LlamaGen's signature (vq_model.py:52): `shape` = (B, C, h, w) if channel_first else (B, h, w, C)."""
        if shape is None:
            raise ValueError("decode_code needs the latent shape")
        b, h, w = (shape[0], shape[2], shape[3]) if channel_first else (shape[0], shape[1], shape[2])
        return self.decode_latents(self.latents(code_b, int(b), int(h), int(w)))

    def decode_tokens(self, tokens, h_latent_dim: int, w_latent_dim: int) -> torch.Tensor:
        """Chameleon's path (image_tokenizer.py:116-121): one image's tokens on an h x w latent grid -> [1, 3, H, W]."""
        return self.decode_latents(self.latents(tokens, 1, int(h_latent_dim), int(w_latent_dim)))


class Emu3VQDecoder:
    """Emu3's vision tokenizer, decode side, for IMAGES (codes [B, h, w] -> [B, 3, 8h, 8w]): what
    `Emu3VisionVQModel.decode` returns for a 3-D code tensor, i.e. the first of the frames its temporal stack produces.

    Token side in the hand-written kernel (twice: the raw codebook rows `zq` that condition every normalisation, and
    `post_quant_conv` of them — for a single frame the causal (3, 1, 1) convolution sees two frames of zero padding in
    front of the only real one, so only its LAST temporal tap acts and the convolution is the 1x1 `sjd_vq_lookup`
    computes); the temporal stack (BatchNorm3d in inference mode, causal 3x3x3 convolutions, nearest x2 up-sampling in
    time) and the 2-D decoder as library calls, structure read from the state dict's keys."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("Emu3VQDecoder runs on the GPU only (sjd_vq_lookup has no CPU fallback)")
        self.device = dev
        sd = {k: v.detach().to(dev, torch.float32).contiguous() for k, v in state_dict.items()
              if k.startswith(("decoder.", "post_quant_conv.", "quantize.embedding.")) and v.is_floating_point()}
        if "decoder.time_res_stack.0.conv1.conv.weight" not in sd or "post_quant_conv.conv.weight" not in sd:
            raise ValueError("state dict holds no Emu3VisionVQ decoder")
        self.sd = sd
        self.codebook = sd["quantize.embedding.weight"]
        e_dim = self.codebook.shape[1]
        pq = sd["post_quant_conv.conv.weight"]                       # [z, e, 3, 1, 1]
        self.pq_w = pq[:, :, -1, 0, 0].contiguous()                  # the tap that sees the single real frame
        self.pq_b = sd["post_quant_conv.conv.bias"]
        self.eye_w = torch.eye(e_dim, device=dev, dtype=torch.float32).contiguous()
        self.zero_b = torch.zeros(e_dim, device=dev, dtype=torch.float32)
        self.program = self._plan()

    @classmethod
    def from_module(cls, vq_model: torch.nn.Module, device=None) -> "Emu3VQDecoder":
        if device is None:
            device = next(vq_model.parameters()).device
        return cls(vq_model.state_dict(), device)

    def _plan(self):
        keys = list(self.sd)

        def idx(pattern):
            return sorted({int(m.group(1)) for k in keys for m in [re.match(pattern, k)] if m})

        prog = [("res", "decoder.mid.block_1"), ("attn", "decoder.mid.attn_1"), ("res", "decoder.mid.block_2")]
        for i in reversed(idx(r"decoder\.up\.(\d+)\.")):
            base = f"decoder.up.{i}"
            attn = set(idx(rf"decoder\.up\.{i}\.attn\.(\d+)\."))
            for j in idx(rf"decoder\.up\.{i}\.block\.(\d+)\."):
                prog.append(("res", f"{base}.block.{j}"))
                if j in attn:
                    prog.append(("attn", f"{base}.attn.{j}"))
            if f"{base}.upsample.conv.weight" in self.sd:
                prog.append(("up", f"{base}.upsample"))
        self.n_time_res = len(idx(r"decoder\.time_res_stack\.(\d+)\."))
        self.n_time_up = len(idx(r"decoder\.time_conv\.(\d+)\."))
        return prog

    # ---- temporal stack ([N, C, T, H, W]) ----
    def _cconv3(self, x, p):                                          # causal: two frames of padding in front, none behind
        w = self.sd[p + ".conv.weight"]
        kh, kw = w.shape[3], w.shape[4]
        x = F.pad(x, ((kw - 1) // 2 + (kw - 1) % 2, (kw - 1) // 2, (kh - 1) // 2 + (kh - 1) % 2, (kh - 1) // 2, 2, 0))
        return F.conv3d(x, w, self.sd[p + ".conv.bias"])

    def _bn3(self, x, p):
        return F.batch_norm(x, self.sd[p + ".running_mean"], self.sd[p + ".running_var"], self.sd[p + ".weight"],
                            self.sd[p + ".bias"], training=False, eps=1e-5)

    def _tres(self, x, p):
        h = self._cconv3(_swish(self._bn3(x, p + ".norm1")), p + ".conv1")
        h = self._cconv3(_swish(self._bn3(h, p + ".norm2")), p + ".conv2")
        if p + ".nin_shortcut.weight" in self.sd:
            x = F.conv3d(x, self.sd[p + ".nin_shortcut.weight"], self.sd[p + ".nin_shortcut.bias"])
        elif p + ".conv_shortcut.conv.weight" in self.sd:
            x = self._cconv3(x, p + ".conv_shortcut")
        return x + h

    def _tup(self, x, p):
        return self._cconv3(torch.repeat_interleave(x, 2, dim=2), p + ".conv")   # nearest x2 along time

    # ---- 2-D decoder with latent-conditioned normalisation ----
    def _conv(self, x, prefix, padding):
        return F.conv2d(x, self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], padding=padding)

    def _snorm(self, x, zq, p):
        zq = F.interpolate(zq, size=x.shape[-2:], mode="nearest")
        if p + ".conv.weight" in self.sd:
            zq = self._conv(zq, p + ".conv", 1)
        n = F.group_norm(x, 32, self.sd[p + ".norm_layer.weight"], self.sd[p + ".norm_layer.bias"], eps=1e-6)
        return n * self._conv(zq, p + ".conv_y", 0) + self._conv(zq, p + ".conv_b", 0)

    def _res(self, x, zq, p):
        h = self._conv(_swish(self._snorm(x, zq, p + ".norm1")), p + ".conv1", 1)
        h = self._conv(_swish(self._snorm(h, zq, p + ".norm2")), p + ".conv2", 1)
        if p + ".nin_shortcut.weight" in self.sd:
            x = self._conv(x, p + ".nin_shortcut", 0)
        elif p + ".conv_shortcut.weight" in self.sd:
            x = self._conv(x, p + ".conv_shortcut", 1)
        return x + h

    def _attn(self, x, zq, p):
        b, c, hh, ww = x.shape
        n = self._snorm(x, zq, p + ".norm")
        q = self._conv(n, p + ".q", 0).reshape(b, c, hh * ww)
        k = self._conv(n, p + ".k", 0).reshape(b, c, hh * ww)
        v = self._conv(n, p + ".v", 0).reshape(b, c, hh * ww)
        w_ = torch.softmax(torch.bmm(q.transpose(1, 2), k) / (c ** 0.5), dim=2)
        o = torch.bmm(v, w_.transpose(1, 2)).reshape(b, c, hh, ww)
        return x + self._conv(o, p + ".proj_out", 0)

    def _lookup(self, c32, n_pix, hw, w, bias, out):
        n_e, e_dim = self.codebook.shape
        L = _lib.lib()
        _lib.check(L.sjd_vq_lookup(c32.data_ptr(), n_pix, hw, self.codebook.data_ptr(), n_e, e_dim, 0, w.data_ptr(),
                                   bias.data_ptr(), w.shape[0], out.data_ptr(),
                                   torch.cuda.current_stream(self.device).cuda_stream), "sjd_vq_lookup")

    @torch.no_grad()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        codes = torch.as_tensor(codes)
        if codes.ndim != 3:
            raise NotImplementedError("Emu3VQDecoder decodes images (codes [B, h, w]); video codes [B, t, h, w] stay with "
                                      "the reference's Emu3VisionVQModel.decode")
        B, h, w = codes.shape
        n_e, e_dim = self.codebook.shape
        c64 = codes.reshape(-1).to(self.device, torch.int64)
        if int(c64.min()) < 0 or int(c64.max()) >= n_e:
            raise ValueError(f"image-token id outside the codebook [0, {n_e})")
        c32 = c64.to(torch.int32).contiguous()
        z = torch.empty(B, self.pq_w.shape[0], h, w, device=self.device, dtype=torch.float32)
        zq = torch.empty(B, e_dim, h, w, device=self.device, dtype=torch.float32)
        self._lookup(c32, B * h * w, h * w, self.pq_w, self.pq_b, z)        # post_quant_conv(embedding(codes)), single frame
        self._lookup(c32, B * h * w, h * w, self.eye_w, self.zero_b, zq)    # embedding(codes) as [B, C, h, w]
        return self.decode_latents(z, zq)

    @torch.no_grad()
    def decode_latents(self, z: torch.Tensor, zq: torch.Tensor) -> torch.Tensor:
        """z = post_quant_conv(embedding(codes)), zq = embedding(codes), both [B, C, h, w] (one frame) -> [B, 3, H, W]."""
        B = z.shape[0]
        x = torch.cat([z, zq], dim=0).unsqueeze(2)                          # [2B, C, T=1, h, w]
        for i in range(self.n_time_res):
            x = self._tres(x, f"decoder.time_res_stack.{i}")
        for i in range(self.n_time_up):
            x = _swish(self._tup(x, f"decoder.time_conv.{i}"))
        x = x.permute(0, 2, 1, 3, 4)                                        # [2B, T, C, h, w]
        hh, zz = torch.chunk(x, 2, dim=0)
        T = hh.shape[1]
        hh = hh.reshape(B * T, *hh.shape[2:])
        zz = zz.reshape(B * T, *zz.shape[2:])
        y = self._conv(hh, "decoder.conv_in", 1)
        for op, p in self.program:
            y = self._res(y, zz, p) if op == "res" else (self._attn(y, zz, p) if op == "attn" else
                                                         self._conv(F.interpolate(y, scale_factor=2.0, mode="nearest"), p + ".conv", 1))
        y = self._conv(_swish(self._snorm(y, zz, "decoder.norm_out")), "decoder.conv_out", 1)
        return y.reshape(B, T, *y.shape[1:])[:, 0]
