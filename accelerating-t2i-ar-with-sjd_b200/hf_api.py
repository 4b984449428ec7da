"""Host-side mirror of the reference's plugin interface for the SJD path — same names, same keyword
arguments, same error behaviour — backed by the sm_100a engine instead of PyTorch-eager code.

reference                                                       here
---------------------------------------------------------------------------------------------------------
scheduler/jacobi_iteration_lumina_mgpt.py:1340 renew_pipeline_sampler   renew_pipeline_sampler
                                          :598  renew_sampler            renew_sampler  (JacobiSampler._sample)
                                          :1253 renew_backbone           renew_backbone (mask is index math in-kernel)
                                          :432  renew_pipeline           renew_pipeline (create_logits_processor)
scheduler/logit_processor_3dim.py:45,158,355   3-D processors           descriptor classes, evaluated by sjd_verify
scheduler/logit_processor_3dim.py:207-353      Anole 3-D processors     descriptor classes -> engine.AnoleGrammarState
scheduler/jacobi_iteration_anhole.py:318       renew_pipeline_sampler   renew_pipeline_sampler_anole
llamagen/llamagen_solver.py:196,349            renew_llamagen, LlamaGenSolver   same names

`_sample(input_ids, logits_processor, stopping_criteria, generation_config, synced_gpus, streamer,
logits_warper=None, **model_kwargs) -> LongTensor [1, P+N]` keeps the HF-facing contract, so
`model.generate()` / `solver.generate()` callers (test_lumina_mgpt.py:130, test_llamagen.py:163) are unchanged.
The transformer weights are packed once (first call) into the engine's layout; models are mutated in place by
class swapping exactly like the reference.
"""
from __future__ import annotations

import logging
import math

import torch
from torch import nn

from . import engine as _engine
from .model import DeviceStack, StackShape

try:  # HF is plumbing for the boundary only
    from transformers.generation.logits_process import (LogitsProcessor, LogitsProcessorList, TemperatureLogitsWarper,
                                                        TopKLogitsWarper, TopPLogitsWarper)
except Exception:  # pragma: no cover
    LogitsProcessor = object
    LogitsProcessorList = list
    TopKLogitsWarper = TemperatureLogitsWarper = TopPLogitsWarper = None


# --------------------------------------------------------------------------------------------------------
# processors: parameter holders with the reference's constructor signatures.  Their arithmetic runs inside
# sjd_verify; calling them on tensors is not part of the product path.
# --------------------------------------------------------------------------------------------------------
class _DeviceEvaluated(LogitsProcessor):
    def __call__(self, input_ids, scores):
        raise RuntimeError(f"{type(self).__name__} is evaluated on the GPU by sjd_verify; there is no host fallback")


class MultiTokensVLLogitsProcessor(_DeviceEvaluated):
    """scheduler/logit_processor_3dim.py:45-155"""

    def __init__(self, image_start_token_id=None, image_end_token_id=None, image_next_line_token_id=None,
                 patch_size=None, voc_size=None, device="cpu"):
        self.image_start_token_id = image_start_token_id
        self.image_end_token_id = image_end_token_id
        self.image_next_line_token_id = image_next_line_token_id
        self.patch_size, self.voc_size = patch_size, voc_size
        self.h_latent_dim = self.w_latent_dim = None


class MultiTokensInterleavedTopKLogitsWarper(_DeviceEvaluated):
    """scheduler/logit_processor_3dim.py:158-204"""

    def __init__(self, image_top_k, text_top_k, image_start_token_id=None, image_end_token_id=None,
                 filter_value=-float("Inf"), min_tokens_to_keep=1):
        if not isinstance(text_top_k, int) or text_top_k <= 0:
            raise ValueError(f"`text_top_k` has to be a strictly positive integer, but is {text_top_k}")
        if not isinstance(image_top_k, int) or text_top_k <= 0:
            raise ValueError(f"`image_top_k` has to be a strictly positive integer, but is {image_top_k}")
        self.image_top_k = max(image_top_k, min_tokens_to_keep)
        self.text_top_k = max(text_top_k, min_tokens_to_keep)
        self.image_start_token_id, self.image_end_token_id = image_start_token_id, image_end_token_id


class TopPLogitsWarper3d(_DeviceEvaluated):
    """scheduler/logit_processor_3dim.py:355-419.  Evaluated after top-k inside sjd_verify (block_top_p, csrc/verify.cu);
    top_p = 1.0 (every shipped driver) removes nothing."""

    def __init__(self, top_p, filter_value=-float("Inf"), min_tokens_to_keep=1):
        top_p = float(top_p)
        if top_p < 0 or top_p > 1.0:
            raise ValueError(f"`top_p` has to be a float > 0 and < 1, but is {top_p}")
        if not isinstance(min_tokens_to_keep, int) or (min_tokens_to_keep < 1):
            raise ValueError(f"`min_tokens_to_keep` has to be a positive integer, but is {min_tokens_to_keep}")
        self.top_p = top_p
        self.filter_value, self.min_tokens_to_keep = filter_value, min_tokens_to_keep
        if min_tokens_to_keep != 1 or filter_value != -float("Inf"):
            raise NotImplementedError("TopPLogitsWarper3d on device keeps exactly one token at least and fills with -inf")


# ---- Anole: the 3-D variants of HF's Chameleon processors (scheduler/logit_processor_3dim.py:207-353) ----------------
def _id_list(x):
    return [int(t) for t in (x.flatten().tolist() if hasattr(x, "flatten") else x)]


class AllowOnlyTokensAtRelativeOffsetLogitsProcessor3d(_DeviceEvaluated):
    """:207-256"""

    def __init__(self, trigger_token_id, allowed_token_ids, offset, exclusive=False, device="cpu"):
        self.trigger_token_id, self.allowed_token_ids = int(trigger_token_id), _id_list(allowed_token_ids)
        self.offset, self.exclusive = int(offset), bool(exclusive)


class SuppressTokensInIndexRangeLogitsProcessor3d(_DeviceEvaluated):
    """:258-286"""

    def __init__(self, suppress_tokens, start_index, end_index=None, device="cpu"):
        self.suppress_tokens = _id_list(suppress_tokens)
        self.start_index = start_index
        self.end_index = end_index if end_index is not None else math.inf


class AllowOnlyTokensInRelativeWindowLogitsProcessor3d(_DeviceEvaluated):
    """:288-338"""

    def __init__(self, trigger_token_id, allowed_token_ids, window_width, exclusive=False, device="cpu"):
        self.trigger_token_id, self.allowed_token_ids = int(trigger_token_id), _id_list(allowed_token_ids)
        self.window_width, self.exclusive = int(window_width), bool(exclusive)


class SuppressTokensAtBeginLogitsProcessor3d(SuppressTokensInIndexRangeLogitsProcessor3d):
    """:340-349"""

    def __init__(self, begin_suppress_tokens, begin_index, device="cpu"):
        super().__init__(begin_suppress_tokens, begin_index, begin_index + 1, device=device)
        self.begin_index = begin_index

    def set_begin_index(self, begin_index):
        self.start_index, self.end_index, self.begin_index = begin_index, begin_index + 1, begin_index


class SuppressTokensLogitsProcessor3d(SuppressTokensInIndexRangeLogitsProcessor3d):
    """:351-353"""

    def __init__(self, suppress_tokens, device="cpu"):
        super().__init__(suppress_tokens, 0, device=device)


def _anole_grammar(procs, plain_k, vocab):
    """The processor sets renew_pipeline_anole.generate builds per multimodal_generation_mode
    (scheduler/jacobi_iteration_anhole.py:170-265) -> engine.AnoleGrammarState(mode):
        image-only              at-offset + in-window + no-late-image + suppress(everything else) + suppress-at-begin(eos)
        interleaved-text-image  at-offset + in-window + no-late-image
        text-only               suppress(image ids + boi + eoi)
    ("unrestricted" installs none of them and never reaches this function.)  Any other combination has no device
    implementation."""
    at = [p for p in procs if isinstance(p, AllowOnlyTokensAtRelativeOffsetLogitsProcessor3d)]
    win = [p for p in procs if isinstance(p, AllowOnlyTokensInRelativeWindowLogitsProcessor3d)]
    beg = [p for p in procs if isinstance(p, SuppressTokensAtBeginLogitsProcessor3d)]
    rng = [p for p in procs if type(p) is SuppressTokensInIndexRangeLogitsProcessor3d]
    sup = [p for p in procs if isinstance(p, SuppressTokensLogitsProcessor3d) or
           (type(p).__name__ == "SuppressTokensLogitsProcessor" and hasattr(p, "suppress_tokens"))]
    counts = (len(at), len(win), len(rng), len(sup), len(beg))
    if counts == (0, 0, 0, 1, 0):
        # text-only: the suppressed ids are one contiguous image-id run plus begin- and end-of-image
        ids = sorted(_id_list(sup[0].suppress_tokens))
        runs, start = [], 0
        for i in range(1, len(ids) + 1):
            if i == len(ids) or ids[i] != ids[i - 1] + 1:
                runs.append((ids[start], ids[i - 1] + 1))
                start = i
        big = max(runs, key=lambda r: r[1] - r[0])
        rest = [t for lo_, hi_ in runs if (lo_, hi_) != big for t in range(lo_, hi_)]
        if len(rest) > 2:
            raise NotImplementedError("text-only Anole grammar: more than two suppressed ids outside the image-id range")
        rest += [-1] * (2 - len(rest))
        return _engine.AnoleGrammarState(rest[0], rest[1], -1, big[0], big[1], 0, max_length=0, begin_index=0,
                                         top_k=plain_k, mode="text-only")
    if counts[:3] != (1, 1, 1) or counts[3:] not in ((1, 1), (0, 0)):
        raise NotImplementedError("this combination of 3-D Chameleon processors is none of the Anole adaptor's modes "
                                  "(scheduler/jacobi_iteration_anhole.py:170-265)")
    at, win, rng = at[0], win[0], rng[0]
    S = win.window_width
    lo, hi = _visual_range(win.allowed_token_ids)
    ok = (at.exclusive and win.exclusive and at.trigger_token_id == win.trigger_token_id and at.offset == S + 1
          and len(at.allowed_token_ids) == 1 and rng.suppress_tokens == [at.trigger_token_id]
          and rng.end_index == math.inf)
    boi, eoi = at.trigger_token_id, at.allowed_token_ids[0]
    if counts[3:] == (0, 0):
        if not ok:
            raise NotImplementedError("Anole processors are not in the 'interleaved-text-image' configuration")
        return _engine.AnoleGrammarState(boi, eoi, -1, lo, hi, S, max_length=int(rng.start_index) + S + 1, begin_index=0,
                                         top_k=plain_k, mode="interleaved-text-image")
    beg, sup = beg[0], sup[0]
    ok = ok and len(beg.suppress_tokens) == 1
    eos = beg.suppress_tokens[0]
    if ok and vocab is not None:
        kept = set(range(vocab)) - set(_id_list(sup.suppress_tokens))
        ok = kept == set(range(lo, hi)) | {eos, boi, eoi}
    if not ok:
        raise NotImplementedError("Anole processors are not in the 'image-only' configuration of "
                                  "scheduler/jacobi_iteration_anhole.py:200-240")
    return _engine.AnoleGrammarState(boi, eoi, eos, lo, hi, S, max_length=int(rng.start_index) + S + 1,
                                     begin_index=int(beg.begin_index), top_k=plain_k)


def renew_end_of_line_logit_processor_3d(model_class):
    """scheduler/jacobi_iteration_emu3.py:41-151: class-swaps Emu3PrefixConstrainedLogitsHelper (emu3/mllm/utils_emu3.py:
    19-62: height, width, img_token, eoi/eos/eol/eof/pad tokens, visual_tokens) into the 3-D processor.  Here it stays
    a parameter holder; its arithmetic runs in sjd_verify (engine.Emu3GrammarState)."""
    class EOLLogitProcessor3d(model_class, _DeviceEvaluated):
        _sjd_emu3_grammar = True
    return EOLLogitProcessor3d


def _visual_range(visual_tokens):
    v = torch.as_tensor(visual_tokens).flatten().tolist()
    lo, hi = min(v), max(v) + 1
    if hi - lo != len(set(v)):
        raise NotImplementedError("visual token ids must form one contiguous range for the device grammar")
    return lo, hi


def grammar_from_processors(processors, vocab=None):
    """Translate the reference's processor list into the engine's grammar state.  HF's own warpers, which
    `generate()` appends for `temperature != 1` / `top_p < 1` / `top_k`, are consumed too: the returned state carries
    `.temperature` (1.0 when no TemperatureLogitsWarper is in the list — the reference's _sample applies only what the
    list holds, never generation_config.temperature by itself)."""
    g = _grammar_from_processors(processors, vocab)
    g.temperature = 1.0
    for pr in processors or []:
        if TemperatureLogitsWarper is not None and isinstance(pr, TemperatureLogitsWarper):
            g.temperature *= float(pr.temperature)
    return g


def _grammar_from_processors(processors, vocab=None):
    vl = topk = emu = None
    plain_k, top_p, anole = 0, 1.0, []
    for pr in processors or []:
        if TemperatureLogitsWarper is not None and isinstance(pr, TemperatureLogitsWarper):
            continue      # a positive scale commutes with the grammar masks and the top-k selection: applied in sjd_verify
        if TopPLogitsWarper is not None and isinstance(pr, TopPLogitsWarper):
            # HF's 2-D nucleus warper == TopPLogitsWarper3d on every window row (logit_processor_3dim.py:355-419 is its port)
            if getattr(pr, "min_tokens_to_keep", 1) != 1 or getattr(pr, "filter_value", -float("inf")) != -float("inf"):
                raise NotImplementedError("TopPLogitsWarper on device keeps exactly one token at least and fills with -inf")
            top_p = float(pr.top_p)
            continue
        if getattr(pr, "_sjd_emu3_grammar", False):
            emu = pr
        elif isinstance(pr, MultiTokensVLLogitsProcessor):
            vl = pr
        elif isinstance(pr, MultiTokensInterleavedTopKLogitsWarper):
            topk = pr
        elif isinstance(pr, TopPLogitsWarper3d):
            top_p = pr.top_p   # sjd_verify applies it after top-k, the order every reference driver uses
        elif TopKLogitsWarper is not None and isinstance(pr, TopKLogitsWarper):
            if top_p != 1.0:
                raise NotImplementedError("top-k after top-p: sjd_verify applies top-k first")
            plain_k = int(pr.top_k)
        elif isinstance(pr, (AllowOnlyTokensAtRelativeOffsetLogitsProcessor3d, SuppressTokensInIndexRangeLogitsProcessor3d,
                             AllowOnlyTokensInRelativeWindowLogitsProcessor3d)) or \
                type(pr).__name__ == "SuppressTokensLogitsProcessor":
            anole.append(pr)
        else:
            raise NotImplementedError(f"logits processor {type(pr).__name__} has no device implementation")
    if top_p != 1.0 and (emu is not None or vl is not None or anole):
        raise NotImplementedError("top_p < 1 together with an image grammar")
    if anole:
        if emu is not None or vl is not None:
            raise NotImplementedError("Anole processors mixed with another image grammar")
        return _anole_grammar(anole, plain_k, vocab)
    if emu is not None:
        lo, hi = _visual_range(emu.visual_tokens)
        return _engine.Emu3GrammarState(int(emu.height), int(emu.width), int(emu.img_token), int(emu.eol_token),
                                        int(emu.eof_token), int(emu.eoi_token), int(emu.eos_token), int(emu.pad_token),
                                        lo, hi, top_k=plain_k)
    if vl is not None:
        return _engine.LuminaGrammarState(
            image_start=vl.image_start_token_id, image_end=vl.image_end_token_id, eol=vl.image_next_line_token_id,
            image_top_k=topk.image_top_k if topk else 0, text_top_k=topk.text_top_k if topk else 0)
    return _engine.PlainTopKState(top_k=plain_k, top_p=top_p)


# --------------------------------------------------------------------------------------------------------
# weight packing: HF Chameleon / Llama-like (Emu3) and LlamaGen module trees -> engine layout
# --------------------------------------------------------------------------------------------------------
def _rope_half(head_dim, n_pos, theta, round_bf16=True):
    from .families import rope_rotate_half
    return rope_rotate_half(head_dim, n_pos, theta, round_bf16)


def pack_hf_decoder(model, max_len: int, rows: int, device) -> DeviceStack:
    """ChameleonForConditionalGeneration / Emu3ForCausalLM-style trees: model.model.layers[i].self_attn.{q,k,v,o}_proj,
    (.q_norm/.k_norm), .mlp.{gate,up,down}_proj, .input_layernorm, .post_attention_layernorm; model.model.norm;
    model.lm_head (modeling_chameleon.py:235-369, :593-666)."""
    cfg = model.config
    if getattr(cfg, "swin_norm", False):
        raise NotImplementedError("swin_norm decoder layers (Chameleon-34B, modeling_chameleon.py:669-741) are not supported")
    core = model.model
    layers = core.layers
    H = cfg.num_attention_heads
    Hkv = getattr(cfg, "num_key_value_heads", H) or H
    d = cfg.hidden_size
    Dh = d // H
    qk_norm = hasattr(layers[0].self_attn, "q_norm")

    def norm_param(p, heads):  # Lumina keeps [model_parallel_size, Dh] and repeat-interleaves (:206-219)
        p = p.detach()
        if p.dim() == 1:
            p = p[None]
        return p.repeat_interleave(heads // p.shape[0], dim=0) if p.shape[0] != heads else p

    w = {"embed": core.embed_tokens.weight, "final_norm": core.norm.weight, "lm_head": model.lm_head.weight, "layers": []}
    for L in layers:
        a, m = L.self_attn, L.mlp
        e = {"attn_norm": L.input_layernorm.weight,
             "wqkv": torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0),
             "wo": a.o_proj.weight, "ffn_norm": L.post_attention_layernorm.weight,
             "w_gate_up": torch.cat([m.gate_proj.weight, m.up_proj.weight], 0), "w_down": m.down_proj.weight}
        if qk_norm:
            e.update(q_norm_w=norm_param(a.q_norm.weight, H), q_norm_b=norm_param(a.q_norm.bias, H),
                     k_norm_w=norm_param(a.k_norm.weight, Hkv), k_norm_b=norm_param(a.k_norm.bias, Hkv))
        w["layers"].append(e)
    shape = StackShape(len(layers), d, H, Hkv, Dh, cfg.intermediate_size, cfg.vocab_size,
                       float(getattr(cfg, "rms_norm_eps", 1e-5)), qk_norm, False)
    theta = getattr(cfg, "rope_theta", None)
    if theta is None:   # HF >= 5 keeps it in config.rope_parameters
        rp = getattr(cfg, "rope_parameters", None) or {}
        theta = rp.get("rope_theta", 10000.0) if isinstance(rp, dict) else 10000.0
    theta = float(theta)
    cos, sin = _rope_half(Dh, max_len, theta)
    return DeviceStack(shape, w, cos, sin, rows, max_len, device)


def pack_llamagen(model, max_len: int, rows: int, device) -> DeviceStack:
    """gpt-fast style tree of llamagen/llamagen.py:297-335 (fused wqkv, w1/w3/w2)."""
    from .families import rope_llamagen_2d
    c = model.config
    H = c.n_head
    Dh = c.dim // H
    w = {"embed": model.tok_embeddings.weight, "final_norm": model.norm.weight, "lm_head": model.output.weight,
         "layers": []}
    for L in model.layers:
        w["layers"].append({"attn_norm": L.attention_norm.weight, "wqkv": L.attention.wqkv.weight,
                            "wo": L.attention.wo.weight, "ffn_norm": L.ffn_norm.weight,
                            "w_gate_up": torch.cat([L.feed_forward.w1.weight, L.feed_forward.w3.weight], 0),
                            "w_down": L.feed_forward.w2.weight})
    d_ff = model.layers[0].feed_forward.w1.weight.shape[0]
    shape = StackShape(len(model.layers), c.dim, H, c.n_kv_head or H, Dh, d_ff, c.vocab_size, c.norm_eps, False, True)
    grid = int(math.isqrt(c.block_size))
    cos, sin = rope_llamagen_2d(grid, Dh, c.rope_base, c.cls_token_num)
    return DeviceStack(shape, w, cos, sin, rows, max(max_len, cos.shape[0]), device)


# --------------------------------------------------------------------------------------------------------
# renew_* : class swapping like the reference
# --------------------------------------------------------------------------------------------------------
def _criteria_limits(stopping_criteria, generation_config):
    eos, max_length, extra = [], None, []
    for c in stopping_criteria or []:
        if hasattr(c, "eos_token_id"):
            e = c.eos_token_id
            eos += [int(x) for x in (e.tolist() if hasattr(e, "tolist") else (e if isinstance(e, (list, tuple)) else [e]))]
        elif hasattr(c, "max_length"):
            max_length = int(c.max_length)
        else:
            extra.append(c)
    if max_length is None and generation_config is not None and getattr(generation_config, "max_length", None):
        max_length = int(generation_config.max_length)
    return eos, max_length, extra


def renew_sampler(model_class):
    class JacobiSampler(model_class, nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self._init_new_params()

        def _init_new_params(self, jacobi_loop_interval_l=1, jacobi_loop_interval_r=(768 // 16) ** 2 + 768 // 16,
                             max_num_new_tokens=16, guidance_scale=3.0, seed=42, multi_token_init_scheme="random",
                             do_cfg=True, prefix_token_sampler_scheme="speculative_jacobi",
                             use_chameleon_tokenizer=True, _init_doubled_attn_mask_cfg=False, **kwargs):
            """Same keywords/defaults as jacobi_iteration_lumina_mgpt.py:865-910.  The image-token vocabulary used
            for random drafts is ids 4..8195 (what VocabInfo.image_tokens yields for the Chameleon tokenizer, :879-886)
            unless `self.img_vocab` is set by the caller."""
            if use_chameleon_tokenizer:
                self.img_vocab = torch.arange(4, 8196, dtype=torch.long)
            elif not hasattr(self, "img_vocab"):
                self.img_vocab = None
            self.jacobi_loop_interval_l = jacobi_loop_interval_l
            self.jacobi_loop_interval_r = jacobi_loop_interval_r
            self.max_num_new_tokens = max_num_new_tokens
            self.max_jacobi_iter_num = min(200, self.max_num_new_tokens + 1)
            self.guidance_scale = guidance_scale
            self.seed = seed
            self.generator = None
            self.multi_token_init_scheme = multi_token_init_scheme
            self.do_cfg = do_cfg
            self.prefix_token_sampler_scheme = prefix_token_sampler_scheme
            self._init_doubled_attn_mask_cfg = _init_doubled_attn_mask_cfg

        # -- engine plumbing ----------------------------------------------------------------------------------
        def _sjd_params(self):
            return _engine.SJDParams(self.jacobi_loop_interval_l, self.jacobi_loop_interval_r, self.max_num_new_tokens,
                                     self.guidance_scale, self.seed, self.multi_token_init_scheme, self.do_cfg,
                                     self.prefix_token_sampler_scheme)

        def _sjd_fingerprint(self):
            """(data_ptr, version, dtype) of every parameter: changes on load_state_dict, in-place edits (LoRA merges),
            .to(dtype) and re-allocation — the engine holds a private re-laid-out copy of the weights."""
            return tuple((p.data_ptr(), p._version, p.dtype) for p in self.parameters())

        def sjd_invalidate(self):
            """Drop the engine's packed copy of the weights (and its KV cache); the next call re-packs."""
            st = getattr(self, "_sjd_stack_cache", None)
            if st is not None:
                st.close()
            object.__setattr__(self, "_sjd_stack_cache", None)
            object.__setattr__(self, "_sjd_engine_cache", None)

        def _sjd_stack(self, rows, max_len, device):
            """Weights are packed once per (rows, capacity); a larger cached context is reused as is, so a solver-side
            prefill and the following _sample share one KV cache.  A change of the module's parameters since the pack
            (fingerprint above) re-packs instead of silently decoding with stale weights."""
            st = getattr(self, "_sjd_stack_cache", None)
            fp = self._sjd_fingerprint()
            if st is not None and st.rows == rows and st.max_len >= max_len and st.ctx is not None and \
                    getattr(st, "_sjd_fp", None) == fp:
                return st
            self.sjd_invalidate()
            packer = pack_llamagen if hasattr(self, "tok_embeddings") else pack_hf_decoder
            st = packer(self, max_len, rows, device)
            st._sjd_fp = fp
            object.__setattr__(self, "_sjd_stack_cache", st)
            return st

        def _sjd_engine(self, stack, grammar):
            """One SJDEngine (two [W, V] fp32 probability buffers + pinned staging) per packed stack, reused across calls."""
            eng = getattr(self, "_sjd_engine_cache", None)
            vocab = self.img_vocab if self.img_vocab is not None else torch.arange(stack.shape.vocab)
            if eng is None or eng.stack is not stack:
                eng = _engine.SJDEngine(stack, self._sjd_params(), grammar, vocab)
                object.__setattr__(self, "_sjd_engine_cache", eng)
            else:
                eng.p, eng.grammar, eng.img_vocab = self._sjd_params(), grammar, vocab.cpu().long()
            return eng

        @torch.no_grad()
        def _sample(self, input_ids, logits_processor, stopping_criteria, generation_config, synced_gpus=False,
                    streamer=None, logits_warper=None, **model_kwargs):
            assert not getattr(generation_config, "return_dict_in_generate", False)
            if input_ids.shape[0] != 1:
                raise ValueError("the SJD sampler decodes one prompt per call (the reference's B>1 path is broken too)")
            if self.prefix_token_sampler_scheme not in ("speculative_jacobi", "jacobi"):
                raise ValueError(f"prefix_token_sampler_scheme: {self.prefix_token_sampler_scheme}")
            for name in ("pixel_values", "inputs_embeds"):
                if model_kwargs.get(name) is not None:
                    # the reference feeds them to its first forward (prepare_inputs_for_generation_jacobi, :715-724)
                    raise NotImplementedError(f"`{name}` (image-conditioned prompts) is not supported by the SJD engine: "
                                              "the prefill would run on raw placeholder ids")
            _engine.check_init_scheme(self.multi_token_init_scheme)
            device = input_ids.device
            do_cfg = bool(self.do_cfg) and self.guidance_scale != 1
            rows = 2 if do_cfg else 1
            eos, max_length, extra = _criteria_limits(stopping_criteria, generation_config)
            prompt = input_ids[0].tolist()
            kv_len0 = int(getattr(self, "_sjd_kv_len0", 0))   # tokens already cached by a solver-side prefill
            cap = (max_length or (len(prompt) + 4096)) + self.max_num_new_tokens + kv_len0 + 8
            stack = self._sjd_stack(rows, int(-(-cap // 64) * 64), device)
            grammar = grammar_from_processors(list(logits_processor or []) + list(logits_warper or []),
                                              vocab=stack.shape.vocab)
            eng = self._sjd_engine(stack, grammar)
            attn = model_kwargs.get("attention_mask")
            prefill_num = (attn.shape[1] - 1) if attn is not None else len(prompt) - 1
            kv_lo = [0, prefill_num] if (rows == 2 and not self._init_doubled_attn_mask_cfg
                                         and not hasattr(self, "tok_embeddings")) else [0] * rows
            uncond = None
            neg = model_kwargs.get("neg_input_ids")
            if neg is not None:
                # Emu3: the CFG-uncond row is a real negative prompt; both rows are left-padded to one length
                # (get_double_cfg_input_ids, logit_processor_3dim.py:422-440) and the padding is hidden by the mask
                if rows != 2:
                    raise ValueError("neg_input_ids given but CFG is disabled")
                pad = int(self.config.pad_token_id)
                P = max(len(prompt), neg.shape[1])
                uncond = [pad] * (P - neg.shape[1]) + neg[0].tolist()
                prompt = [pad] * (P - len(prompt)) + prompt
                kv_lo = [next((i for i, t in enumerate(r) if t != pad), len(r) - 1) for r in (prompt, uncond)]
                if attn is not None and attn.dim() == 2 and attn.shape[0] == 2:
                    kv_lo = [int((attn[b] == 0).long().cumprod(0).sum()) for b in range(2)]
            stop_fn = None
            if extra:
                def stop_fn(ids):
                    t = torch.tensor([ids])
                    return any(bool(torch.as_tensor(c(t, None)).any()) for c in extra)
            t1 = torch.cuda.Event(enable_timing=True)
            t2 = torch.cuda.Event(enable_timing=True)
            t1.record()
            ids = eng.generate(prompt, max_length=max_length or (len(prompt) + 4096), eos_token_ids=eos,
                               do_sample=bool(generation_config.do_sample), kv_len0=kv_len0, kv_lo=kv_lo,
                               temperature=float(grammar.temperature),
                               stop_fn=stop_fn, uncond_input_ids=uncond)
            t2.record()
            torch.cuda.synchronize()
            self.sjd_stats = eng.stats
            # the three lines drivers/logs rely on (jacobi_iteration_lumina_mgpt.py:1217-1220)
            t_inner = t1.elapsed_time(t2) / 1000
            print("Time elapsed inner: ", t_inner)
            print("gen loop num (NFE): ", eng.stats.nfe)
            print("tokens length: ", len(ids))
            logging.info(f"Time elapsed inner: {t_inner}")              # (:1221-1223: the same three lines through logging)
            logging.info(f"gen loop num (NFE): {eng.stats.nfe}")
            logging.info(f"tokens length: {len(ids)}")
            if streamer is not None:
                streamer.put(torch.tensor(ids[len(prompt):]))
                streamer.end()
            return torch.tensor([ids], dtype=input_ids.dtype, device=device)

    return JacobiSampler


def renew_sampler_forward(model_class):
    """scheduler/jacobi_iteration_emu3.py:153-368.  The reference wraps forward() to build the 4-D mask itself; here the
    mask is index math in the attention kernel, so only the CFG input preparation and the parameter defaults remain."""
    class JacobiModel(model_class):
        def _init_new_params(self, *args, use_chameleon_tokenizer=False, _init_doubled_attn_mask_cfg=True,
                             visual_tokens=None, **kwargs):
            keep = getattr(self, "img_vocab", None)
            super()._init_new_params(*args, use_chameleon_tokenizer=use_chameleon_tokenizer,
                                     _init_doubled_attn_mask_cfg=_init_doubled_attn_mask_cfg, **kwargs)
            if keep is not None:   # renew_sampler ran first with its Chameleon default (reference quirk)
                self.img_vocab = keep
            if getattr(self, "img_vocab", None) is None:
                self.img_vocab = torch.as_tensor(visual_tokens) if visual_tokens is not None else None

        def prepare_batch_cfg_model_inputs(self, input_ids, neg_input_ids=None, attention_mask=None):
            """:234-278 — left-pads positive and negative prompt to one length and builds the [2B, P] keep-mask."""
            B, P = input_ids.shape
            pad = self.config.pad_token_id
            out = {"input_ids": input_ids, "attention_mask": attention_mask}
            rows = 2 * B if self.do_cfg else B
            maxP = max(P, neg_input_ids.shape[1] if neg_input_ids is not None else P)
            keep = torch.zeros((rows, maxP), dtype=torch.bool, device=input_ids.device)
            keep[:B, -P:] = input_ids != pad
            if neg_input_ids is not None:
                both = torch.full((2 * B, maxP), pad, dtype=input_ids.dtype, device=input_ids.device)
                both[:B, -P:] = input_ids
                both[B:, -neg_input_ids.shape[1]:] = neg_input_ids
                out["input_ids"] = both
                out["pos_input_ids"] = both[:B]
                keep = both != pad
            if attention_mask is None:
                out["attention_mask"] = keep.to(torch.float32)
            elif attention_mask.shape[0] == B:
                raise NotImplementedError
            return out

    return JacobiModel


def renew_solver(model, processor, **jacobi_param_dict):
    """scheduler/jacobi_iteration_emu3.py:370-412: returns (model, LogitsProcessorList([constrained_fn]))."""
    h = jacobi_param_dict.pop("h", None)
    w = jacobi_param_dict.pop("w", None)
    jacobi_param_dict.pop("neg_inputs", None)
    jacobi_param_dict.pop("classifier_free_guidance", None)
    constrained_fn = processor.build_prefix_constrained_fn(h, w)
    constrained_fn.__class__ = renew_end_of_line_logit_processor_3d(constrained_fn.__class__)
    model.__class__ = renew_sampler(model.__class__)
    model._init_new_params(**jacobi_param_dict)
    model.__class__ = renew_sampler_forward(model.__class__)
    model._init_new_params(visual_tokens=constrained_fn.visual_tokens, **jacobi_param_dict)
    return model, LogitsProcessorList([constrained_fn])


def renew_backbone(model_class):
    """The reference overrides _update_causal_mask to understand 3-D window masks (:1253-1338); here the mask is
    `kv_lo <= j <= kv_len + i` evaluated inside the attention kernel, so the backbone class needs no change."""
    class JacobiBackbone(model_class):
        pass
    return JacobiBackbone


def renew_pipeline(model_class):
    class JacobiPipeline(model_class):
        def _init_new_params(self, guidance_scale=3.0, image_top_k=2000, text_top_k=10, **kwargs):
            self.cfg, self.image_top_k, self.text_top_k = guidance_scale, image_top_k, text_top_k

        def create_logits_processor(self, cfg=3.0, image_top_k=2000, text_top_k=10):
            image_top_k = getattr(self, "image_top_k", image_top_k)
            text_top_k = getattr(self, "text_top_k", text_top_k)
            ip = self.item_processor
            start = ip.token2id(ip.image_start_token)
            end = ip.token2id(ip.image_end_token)
            return LogitsProcessorList([
                MultiTokensVLLogitsProcessor(image_start_token_id=start, image_end_token_id=end,
                                             image_next_line_token_id=ip.token2id(ip.new_line_token), patch_size=32,
                                             voc_size=self.model.config.vocab_size, device=self.device),
                MultiTokensInterleavedTopKLogitsWarper(image_top_k=image_top_k, text_top_k=text_top_k,
                                                       image_start_token_id=start, image_end_token_id=end)])
    return JacobiPipeline


def renew_pipeline_sampler(pipe_line, **kwargs):
    """scheduler/jacobi_iteration_lumina_mgpt.py:1340-1346"""
    pipe_line.__class__ = renew_pipeline(pipe_line.__class__)
    pipe_line._init_new_params(**kwargs)
    pipe_line.model.__class__ = renew_sampler(pipe_line.model.__class__)
    pipe_line.model._init_new_params(**kwargs)
    pipe_line.model.model.__class__ = renew_backbone(pipe_line.model.model.__class__)
    return pipe_line


# --------------------------------------------------------------------------------------------------------
# Anole adaptor (scheduler/jacobi_iteration_anhole.py): HF ChameleonForConditionalGeneration as the pipeline
# --------------------------------------------------------------------------------------------------------
def renew_vocabulary_mapping(vocabulary_mapping):
    """:45-97 — adds the ids the adaptor needs to HF's ChameleonImageVocabularyMapping (which only knows names)."""
    from functools import cached_property

    class IndexVocabularyMapping(vocabulary_mapping):
        def _init_new_params(self, image_token_id=8711, boi_token_id=8197, eoi_token_id=8196):
            for name, val in (("image_token_id", image_token_id), ("boi_token_id", boi_token_id),
                              ("eoi_token_id", eoi_token_id)):
                if not hasattr(self, name):
                    setattr(self, name, val)

        @cached_property
        def image_token_ids(self):
            return sorted(v for k, v in self.vocab_map.items() if k.startswith("IMGIMG"))

    return IndexVocabularyMapping


def renew_pipeline_anole(model_class):
    """:99-285 — `generate(..., multimodal_generation_mode=...)` installs the 3-D Chameleon processors for the mode and
    hands over to HF's generate, which ends in the renewed `_sample`."""
    class JacobiPipeline(model_class):
        def _init_new_params(self, guidance_scale=3.0, image_top_k=2000, text_top_k=10, **kwargs):
            self.cfg, self.image_top_k, self.text_top_k = guidance_scale, image_top_k, text_top_k
            self.vocabulary_mapping = self.model.vocabulary_mapping
            self.vocabulary_mapping.__class__ = renew_vocabulary_mapping(self.vocabulary_mapping.__class__)
            self.vocabulary_mapping._init_new_params()

        @torch.no_grad()
        def generate(self, inputs=None, generation_config=None, logits_processor=None, multimodal_generation_mode=None,
                     **kwargs):
            vm, S = self.vocabulary_mapping, int(self.model.image_seq_length)
            mode = (multimodal_generation_mode or getattr(generation_config, "multimodal_generation_mode", None)
                    or "text-only")
            gc = generation_config if generation_config is not None else self.generation_config
            no_len = kwargs.get("max_length") is None and kwargs.get("max_new_tokens") is None
            if mode == "image-only" and no_len and generation_config is None:
                kwargs["max_new_tokens"] = S + 2            # boi + image + eoi (:115-126)
            input_ids = kwargs.get("input_ids", inputs)
            P = int(input_ids.shape[-1])
            if kwargs.get("max_new_tokens") is not None:
                max_length = P + int(kwargs["max_new_tokens"])
            elif kwargs.get("max_length") is not None:
                max_length = int(kwargs["max_length"])
            elif getattr(gc, "max_new_tokens", None) is not None:
                max_length = P + int(gc.max_new_tokens)
            else:
                max_length = int(gc.max_length)
            procs = LogitsProcessorList(list(logits_processor or []))
            dev = self.device
            image_ids = vm.image_token_ids
            at_offset = AllowOnlyTokensAtRelativeOffsetLogitsProcessor3d(
                trigger_token_id=vm.boi_token_id, allowed_token_ids=[vm.eoi_token_id], offset=S + 1, exclusive=True,
                device=dev)
            in_window = AllowOnlyTokensInRelativeWindowLogitsProcessor3d(
                trigger_token_id=vm.boi_token_id, allowed_token_ids=image_ids, window_width=S, exclusive=True, device=dev)
            no_late_image = SuppressTokensInIndexRangeLogitsProcessor3d(
                suppress_tokens=[vm.boi_token_id], start_index=max_length - S - 1, device=dev)
            if mode == "text-only":
                procs.append(SuppressTokensLogitsProcessor3d(
                    suppress_tokens=image_ids + [vm.boi_token_id, vm.eoi_token_id], device=dev))
            elif mode == "image-only":
                if max_length - P < S + 2:
                    import warnings
                    warnings.warn(f"image-only generation needs max_new_tokens >= {S + 2} (begin-of-image, {S} image "
                                  f"tokens, end-of-image); got {max_length - P}")
                allowed = set(image_ids) | {self.config.eos_token_id, vm.boi_token_id, vm.eoi_token_id}
                procs.extend([at_offset, in_window, no_late_image,
                              SuppressTokensLogitsProcessor3d(
                                  suppress_tokens=[t for t in range(self.vocab_size) if t not in allowed], device=dev),
                              SuppressTokensAtBeginLogitsProcessor3d(
                                  begin_suppress_tokens=[self.config.eos_token_id], begin_index=P, device=dev)])
            elif mode == "interleaved-text-image":
                procs.extend([at_offset, in_window, no_late_image])
            elif mode != "unrestricted":
                raise ValueError(f"Unknown multimodal generation mode: {mode}. Please choose one of 'unrestricted', "
                                 "'text-only', 'image-only', or 'interleaved-text-image'.")
            return super().generate(inputs=inputs, generation_config=generation_config, logits_processor=procs, **kwargs)

        @property
        def vocab_size(self):
            return int(self.config.vocab_size)

        def decode_image_tokens(self, bpe_tokens):
            return self.model.decode_image_tokens(bpe_tokens)

    return JacobiPipeline


def renew_backbone_adapt_anole(model_class):
    """:287-316 — a multi-token accept can overshoot the image: keep the first image_seq_length tokens before the
    VQ decoder (which is outside the SJD path and stays the model's own)."""
    class JacobiBackboneAdaptedAnole(model_class):
        def decode_image_tokens(self, bpe_tokens):
            if bpe_tokens.shape[1] != self.image_seq_length:
                bpe_tokens = bpe_tokens[:, : self.image_seq_length]
            return self.vqmodel.decode(self.convert_bpe2img_tokens(bpe_tokens))

    return JacobiBackboneAdaptedAnole


def renew_pipeline_sampler_anole(pipe_line, processor, **kwargs):
    """scheduler/jacobi_iteration_anhole.py:318-330 (exported there as renew_pipeline_sampler)."""
    pipe_line.model.__class__ = renew_backbone(pipe_line.model.__class__)
    pipe_line.model.__class__ = renew_backbone_adapt_anole(pipe_line.model.__class__)
    if not hasattr(pipe_line.model, "image_seq_length"):
        pipe_line.model.image_seq_length = processor.image_seq_length
    print(pipe_line.model.image_seq_length)
    pipe_line.__class__ = renew_pipeline_anole(pipe_line.__class__)
    pipe_line._init_new_params(**kwargs)
    pipe_line.__class__ = renew_sampler(pipe_line.__class__)
    pipe_line._init_new_params(**kwargs)
    return pipe_line


# --------------------------------------------------------------------------------------------------------
# synthetic-prompt solver used by bench.py / tests: the public call with host buffers on both sides
# --------------------------------------------------------------------------------------------------------
class SyntheticLuminaSolver:
    """generate(prompt_ids: pinned host LongTensor [1, P]) -> host LongTensor [1, P + N].  Mirrors the part of
    FlexARInferenceSolver.generate (lumina_mgpt/inference_solver.py:298-354) between tokenisation and VQ decode."""

    def __init__(self, eng: _engine.SJDEngine, max_length: int, eos_token_ids=(8710,)):
        self.engine, self.max_length, self.eos = eng, max_length, list(eos_token_ids)

    @torch.no_grad()
    def generate(self, prompt_ids: torch.Tensor, seed: int | None = None, do_sample=True, temperature=1.0):
        dev = self.engine.dev
        d_prompt = prompt_ids.to(dev, non_blocking=True)          # host -> device, like inference_solver.py:326
        prompt = d_prompt[0].tolist()
        P = len(prompt)
        if seed is not None:
            self.engine.p.seed = seed
        ids = self.engine.generate(prompt, max_length=self.max_length, eos_token_ids=self.eos, do_sample=do_sample,
                                   temperature=temperature, kv_lo=[0, P - 1])
        self.engine.stats.h2d_bytes += prompt_ids.numel() * 8
        out = torch.tensor([ids], dtype=torch.int64, device=dev).cpu()  # device -> host result (:350)
        self.engine.stats.d2h_bytes += out.numel() * 8
        return out
