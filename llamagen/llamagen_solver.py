"""Drop-in for the reference's llamagen/llamagen_solver.py: `renew_llamagen`, `LlamaGenSolver` — the LlamaGen entry to
the SJD hot path (test_llamagen.py:85-88, :151-169), served by the sm_100a engine.

reference                                                   here
------------------------------------------------------------------------------------------------------------
renew_llamagen (llamagen_solver.py:196-339): HF-style forward,     renew_llamagen: marker subclass; the static KV cache
  static<->DynamicCache mirroring (two self-copies / layer / trip)   lives in the engine, roll-back is an integer
LlamaGenSolver.generate (:370-456): cond prefill + first token,    same call signature and RNG consumption: the first
  then model._sample                                                  image token comes from the GLOBAL torch generator
create_logits_processor (:458-470)                                 same (TopKLogitsWarper + TopPLogitsWarper3d holders)
"""
from __future__ import annotations

import os as _os
import sys as _sys

import torch

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sjd_b200  # noqa: E402,F401
from sjd_b200.hf_api import TopPLogitsWarper3d  # noqa: E402

try:
    from transformers import GenerationConfig
    from transformers.generation.logits_process import LogitsProcessorList, TopKLogitsWarper
    from transformers.generation.stopping_criteria import StoppingCriteria, StoppingCriteriaList
except Exception as _e:  # pragma: no cover
    raise ImportError("llamagen_solver needs transformers for the GenerationConfig / criteria plumbing") from _e


def renew_llamagen(model_class):
    class WrappedLLamaGen(model_class):
        def _init_new_params(self, *args, **kwargs):
            pass

        def clear_kvcache(self):
            """Reference zeroes its static cache (:206-209); here stale slots are never read (kv_len bounds the mask)."""
            self._sjd_kv_len0 = 0

    return WrappedLLamaGen


class MaxlenCriteria(StoppingCriteria):
    """llamagen_solver.py:341-347: stop once the id row holds `max_seq_length` tokens."""

    def __init__(self, max_seq_length):
        super().__init__()
        self.max_length = max_seq_length   # read by hf_api._criteria_limits
        self.max_seq_length = max_seq_length

    def __call__(self, input_ids, scores, **kwargs):
        return input_ids.shape[-1] >= self.max_seq_length


def _first_token(logits_rows: torch.Tensor, cfg_scale: float, temperature=1.0, top_k=0, top_p=1.0, sample_logits=True):
    """prefill() + sample() of the reference (llamagen_solver.py:95-104, :75-84) on device: one [1, V] row."""
    if cfg_scale > 1.0:
        c, u = logits_rows[0:1], logits_rows[1:2]
        lg = u + (c - u) * cfg_scale
    else:
        lg = logits_rows[0:1]
    lg = lg / max(temperature, 1e-5)
    if top_k > 0:
        k = min(max(top_k, 1), lg.shape[-1])
        lg = lg.masked_fill(lg < torch.topk(lg, k)[0][..., -1, None], float("-inf"))
    if top_p < 1.0:
        # the first token uses the gpt-fast nucleus of top_k_top_p_filtering (:56-71), not TopPLogitsWarper3d: descending
        # order, drop what lies beyond the first entry whose running probability exceeds top_p
        srt, order = torch.sort(lg, descending=True)
        beyond = torch.cumsum(torch.softmax(srt, dim=-1), dim=-1) > top_p
        beyond = torch.cat([torch.zeros_like(beyond[..., :1]), beyond[..., :-1]], dim=-1)
        lg = lg.masked_fill(beyond.scatter(1, order, beyond), float("-inf"))
    probs = torch.softmax(lg, dim=-1)
    return torch.multinomial(probs, num_samples=1) if sample_logits else torch.topk(probs, k=1, dim=-1)[1]


class LlamaGenSolver:
    def __init__(self, model, image_top_k, image_top_p):
        self.model = model
        self.image_top_k = image_top_k
        self.image_top_p = image_top_p

    def create_logits_processor(self):
        return LogitsProcessorList([TopKLogitsWarper(top_k=self.image_top_k), TopPLogitsWarper3d(top_p=self.image_top_p)])

    @torch.no_grad()
    def generate(self, cond, max_new_tokens, emb_masks=None, cfg_scale=1.0, cfg_interval=-1, **sampling_kwargs):
        model = self.model
        device = cond.device
        if device.type != "cuda":
            raise RuntimeError("the SJD engine needs the model and its inputs on a CUDA device (no CPU fallback)")
        if cond.shape[0] != 1:
            raise ValueError("the SJD sampler decodes one prompt per call (the reference's B>1 path is broken too)")
        do_cfg = cfg_scale > 1.0
        if model.model_type == "c2i":
            cond_combined = torch.cat([cond, torch.ones_like(cond) * model.num_classes]) if do_cfg else cond
            T = 1
            cond_embeds = model.cls_embedding(cond_combined)[:, : model.cls_token_num]
        elif model.model_type == "t2i":
            if do_cfg:
                cond_combined = torch.cat([cond, torch.zeros_like(cond) + model.cls_embedding.uncond_embedding])
            else:
                cond_combined = cond
            T = cond.shape[1]
            cond_embeds = model.cls_embedding(cond_combined)[:, : model.cls_token_num]
            if emb_masks is not None and not bool(emb_masks.bool().all()):
                raise NotImplementedError("caption padding masks (emb_masks with zeros) are not implemented on the SJD path")
        else:
            raise Exception("please check model type")
        if do_cfg != (bool(getattr(model, "do_cfg", True)) and getattr(model, "guidance_scale", cfg_scale) != 1):
            raise ValueError("cfg_scale and the sampler's do_cfg / guidance_scale disagree")
        rows = 2 if do_cfg else 1
        model.setup_caches(max_batch_size=rows, max_seq_length=T + max_new_tokens, dtype=model.tok_embeddings.weight.dtype)
        # ---- condition prefill on the engine: logits of the last condition position, then the first image token ----
        cap = T + max_new_tokens + int(model.max_num_new_tokens) + 8
        stack = model._sjd_stack(rows, int(-(-cap // 64) * 64), device)
        pos = torch.arange(T, dtype=torch.int32, device=device).repeat(rows)
        logits = stack.forward(T, pos, pos, 0, [0] * rows, embeds=cond_embeds.to(torch.bfloat16).contiguous(),
                               n_logit_tokens=1)
        next_token = _first_token(logits[:, 0].float(), cfg_scale, **sampling_kwargs)
        model._sjd_kv_len0 = T
        stopping_criteria = StoppingCriteriaList([MaxlenCriteria(max_new_tokens)])
        generation_config = GenerationConfig(max_new_tokens=T + max_new_tokens, max_length=T + max_new_tokens,
                                             temperature=1.0, top_k=None, do_sample=True, return_dict_in_generate=False)
        outputs = model._sample(input_ids=next_token.to(torch.long), logits_processor=self.create_logits_processor(),
                                stopping_criteria=stopping_criteria, generation_config=generation_config,
                                synced_gpus=False, streamer=None, logits_warper=None, use_cache=True,
                                attention_mask=torch.ones((1, T + 1), device=device))
        generated = outputs[:, -max_new_tokens:]
        model.clear_kvcache()
        return generated


@torch.no_grad()
def generate(model, cond, max_new_tokens, emb_masks=None, cfg_scale=1.0, cfg_interval=-1, **sampling_kwargs):
    """The reference's plain AR sampler (llamagen_solver.py:144-194: prefill, then decode_n_tokens one token at a time,
    CFG switched off after `cfg_interval` steps, every token drawn from the GLOBAL torch generator by sample(), :75-84)
    on the engine: window-1 forwards over the static KV cache.  Same signature, same return ([1, max_new_tokens])."""
    device = cond.device
    if device.type != "cuda":
        raise RuntimeError("the SJD engine needs the model and its inputs on a CUDA device (no CPU fallback)")
    if cond.shape[0] != 1:
        raise ValueError("one prompt per call")
    do_cfg = cfg_scale > 1.0
    if model.model_type == "c2i":
        cond_combined = torch.cat([cond, torch.ones_like(cond) * model.num_classes]) if do_cfg else cond
        T = 1
    elif model.model_type == "t2i":
        cond_combined = torch.cat([cond, torch.zeros_like(cond) + model.cls_embedding.uncond_embedding]) if do_cfg else cond
        T = cond.shape[1]
        if emb_masks is not None and not bool(emb_masks.bool().all()):
            raise NotImplementedError("caption padding masks (emb_masks with zeros) are not implemented on the SJD path")
    else:
        raise Exception("please check model type")
    cond_embeds = model.cls_embedding(cond_combined, train=False)[:, : model.cls_token_num]
    rows = 2 if do_cfg else 1
    model.setup_caches(max_batch_size=rows, max_seq_length=T + max_new_tokens, dtype=model.tok_embeddings.weight.dtype)
    cap = T + max_new_tokens + int(getattr(model, "max_num_new_tokens", 1)) + 8
    if not hasattr(model, "_sjd_stack"):
        raise RuntimeError("wrap the model with scheduler.jacobi_iteration_lumina_mgpt.renew_sampler first (it owns the engine)")
    stack = model._sjd_stack(rows, int(-(-cap // 64) * 64), device)
    pos = torch.arange(T, dtype=torch.int32, device=device).repeat(rows)
    logits = stack.forward(T, pos, pos, 0, [0] * rows, embeds=cond_embeds.to(torch.bfloat16).contiguous(), n_logit_tokens=1)
    tok = _first_token(logits[:, 0].float(), cfg_scale, **sampling_kwargs)
    out = [tok]
    for i in range(max_new_tokens - 1):
        scale = cfg_scale if not (cfg_interval > -1 and i > cfg_interval) else 1.0   # cfg_flag (:134-136)
        ids = tok.to(torch.int32).reshape(1).repeat(rows).contiguous()
        p1 = torch.full((rows,), T + i, dtype=torch.int32, device=device)
        logits = stack.forward(1, p1, p1, T + i, [0] * rows, ids=ids, n_logit_tokens=1)
        tok = _first_token(logits[:, 0].float(), scale, **sampling_kwargs)
        out.append(tok)
    return torch.cat(out, dim=1).to(torch.int)
