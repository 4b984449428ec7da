"""ORACLE tooling (runs only in the build container, where /root/reference exists): drive the UNMODIFIED
reference scheduler (scheduler/jacobi_iteration_lumina_mgpt.py, scheduler/logit_processor_3dim.py) on CPU and
record what it does, as golden fixtures under tests/golden/.

  python oracle/mint_golden.py            # writes tests/golden/sjd_loop_*.json, forward_*.npz

The reference pins transformers 4.47.1; this image has 5.5.0, so the eight small compatibility shims of
SURVEY.md Appendix C are applied before importing it (aliases for removed names — no reference logic changes).
The transformer is replaced by oracle/fake_lm.py (exactly reproducible logits), so a fixture is just
{config, prompt, seed} -> {token sequence, per-iteration trace}.  Forward fixtures run the reference's own
model code (llamagen/llamagen.py, lumina_mgpt/model/chameleon/modeling_chameleon.py) in fp32 on seeded random
weights and store a few logit rows for oracle/ref_forward.py to be checked against.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("SJD_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))
sys.dont_write_bytecode = True

from oracle.fake_lm import fake_logits  # noqa: E402


def apply_shims():
    import transformers
    import transformers.generation.logits_process as LP
    from transformers.generation.utils import GenerationMixin
    LP.LogitsWarper = LP.LogitsProcessor
    _hus = GenerationMixin._has_unfinished_sequences
    GenerationMixin._has_unfinished_sequences = (
        lambda self, fin, synced, device, cur_len=None, max_length=None: _hus(self, fin, synced, device))
    GenerationMixin._extract_past_from_model_output = lambda self, outputs: ("past_key_values", outputs.past_key_values)

    class LegacyDynamicCache(transformers.cache_utils.Cache):
        def __init__(self):
            self.key_cache, self.value_cache = [], []

        def __len__(self):
            return len(self.key_cache)

        def get_seq_length(self, layer_idx=0):
            return 0 if len(self.key_cache) <= layer_idx else self.key_cache[layer_idx].shape[-2]

        def update(self, k, v, layer_idx, cache_kwargs=None):
            if len(self.key_cache) <= layer_idx:
                self.key_cache.append(k)
                self.value_cache.append(v)
            else:
                self.key_cache[layer_idx] = torch.cat([self.key_cache[layer_idx], k], -2)
                self.value_cache[layer_idx] = torch.cat([self.value_cache[layer_idx], v], -2)
            return self.key_cache[layer_idx], self.value_cache[layer_idx]

    transformers.DynamicCache = LegacyDynamicCache
    if not torch.cuda.is_available():
        import time

        class _Ev:
            def __init__(self, *a, **k):
                self.t = 0.0

            def record(self):
                self.t = time.time()

            def elapsed_time(self, other):
                return (other.t - self.t) * 1e3

        torch.cuda.Event = _Ev
        torch.cuda.synchronize = lambda *a, **k: None
        torch.cuda.manual_seed_all = lambda *a, **k: None
    return LegacyDynamicCache


def make_fake_lm(V: int, sharp: float):
    from transformers.generation.utils import GenerationMixin
    from types import SimpleNamespace as Out

    class Cfg:
        vocab_size = V
        is_encoder_decoder = False
        pad_token_id = 0

    class FakeLM(torch.nn.Module, GenerationMixin):
        def __init__(self):
            super().__init__()
            self.config = Cfg()

        def forward(self, input_ids=None, position_ids=None, cache_position=None, past_key_values=None,
                    use_cache=True, attention_mask=None, **kw):
            rows, W = input_ids.shape
            kv_len = int(cache_position[0])
            lg = fake_logits(input_ids.tolist(), kv_len, W, V, sharp).reshape(rows, W, V)
            dummy = torch.zeros(rows, 1, W, 1)
            past_key_values.update(dummy, dummy, 0)
            return Out(logits=torch.from_numpy(lg), past_key_values=past_key_values)

    return FakeLM()


def load_reference_scheduler():
    """Import the REFERENCE's scheduler modules by file path.  (`import scheduler` would resolve to this repo's
    drop-in package of the same name: a regular package beats the reference's namespace package on sys.path.)"""
    import importlib.util
    import types
    if "scheduler" in sys.modules and not str(getattr(sys.modules["scheduler"], "__path__", [""])[0]).startswith(str(REF)):
        for k in [k for k in sys.modules if k == "scheduler" or k.startswith("scheduler.")]:
            del sys.modules[k]
    pkg = types.ModuleType("scheduler")
    pkg.__path__ = [str(REF / "scheduler")]
    sys.modules["scheduler"] = pkg
    mods = {}
    for name in ("logit_processor_3dim", "jacobi_iteration_lumina_mgpt"):
        spec = importlib.util.spec_from_file_location(f"scheduler.{name}", REF / "scheduler" / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"scheduler.{name}"] = mod
        spec.loader.exec_module(mod)
        assert str(REF) in mod.__file__
        mods[name] = mod
    return mods["jacobi_iteration_lumina_mgpt"], mods["logit_processor_3dim"]


def run_reference_loop(case: dict, Cache) -> dict:
    sys.path.insert(0, str(REF))
    J, LP3 = load_reference_scheduler()
    MultiTokensInterleavedTopKLogitsWarper = LP3.MultiTokensInterleavedTopKLogitsWarper
    MultiTokensVLLogitsProcessor = LP3.MultiTokensVLLogitsProcessor
    TopPLogitsWarper3d = LP3.TopPLogitsWarper3d
    from transformers import GenerationConfig
    from transformers.generation.logits_process import LogitsProcessorList, TopKLogitsWarper
    from transformers.generation.stopping_criteria import (EosTokenCriteria, MaxLengthCriteria,
                                                           StoppingCriteriaList)

    V = case["V"]
    m = make_fake_lm(V, case["sharp"])
    m.__class__ = J.renew_sampler(m.__class__)
    m._init_new_params(use_chameleon_tokenizer=False, **case["jacobi"])
    lo, hi = case["img_vocab"]
    m.img_vocab = torch.arange(lo, hi)
    if case["grammar"] == "lumina":
        procs = LogitsProcessorList([
            MultiTokensVLLogitsProcessor(8197, 8196, 8803, 32, V),
            MultiTokensInterleavedTopKLogitsWarper(case["image_top_k"], case["text_top_k"], 8197, 8196)])
    elif case["grammar"] == "emu3":
        # Emu3PrefixConstrainedLogitsHelper (emu3/mllm/utils_emu3.py:19-62) class-swapped to the 3-D processor exactly
        # like renew_solver does (scheduler/jacobi_iteration_emu3.py:379-380); HF appends TopKLogitsWarper for
        # GenerationConfig(top_k=...) (test_emu3.py:83-90)
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_utils_emu3", REF / "emu3" / "mllm" / "utils_emu3.py")
        U = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(U)
        spec = importlib.util.spec_from_file_location("scheduler.jacobi_iteration_emu3", REF / "scheduler" / "jacobi_iteration_emu3.py")
        E = importlib.util.module_from_spec(spec)
        sys.modules["scheduler.jacobi_iteration_emu3"] = E
        spec.loader.exec_module(E)
        e = case["emu3"]
        fn = U.Emu3PrefixConstrainedLogitsHelper(e["height"], e["width"], e["img_token"], e["eoi"], e["eos"], e["eol"],
                                                 e["eof"], e["pad"], torch.arange(e["visual"][0], e["visual"][1]))
        fn.__class__ = E.renew_end_of_line_logit_processor_3d(fn.__class__)
        procs = LogitsProcessorList([fn, TopKLogitsWarper(top_k=case["image_top_k"])])
    elif case["grammar"] == "anole":
        # the processor list renew_pipeline_anole.generate builds for multimodal_generation_mode="image-only"
        # (scheduler/jacobi_iteration_anhole.py:200-240); HF appends TopKLogitsWarper (GenerationConfig.top_k)
        a = case["anole"]
        image_ids = list(range(a["image"][0], a["image"][1]))
        allowed = image_ids + [a["eos"], a["boi"], a["eoi"]]
        S = a["image_seq_length"]
        three = [
            LP3.AllowOnlyTokensAtRelativeOffsetLogitsProcessor3d(trigger_token_id=a["boi"], allowed_token_ids=[a["eoi"]],
                                                                 offset=S + 1, exclusive=True),
            LP3.AllowOnlyTokensInRelativeWindowLogitsProcessor3d(trigger_token_id=a["boi"], allowed_token_ids=image_ids,
                                                                 window_width=S, exclusive=True),
            LP3.SuppressTokensInIndexRangeLogitsProcessor3d(suppress_tokens=[a["boi"]],
                                                            start_index=case["max_length"] - S - 1)]
        mode = a.get("mode", "image-only")
        if mode == "text-only":            # jacobi_iteration_anhole.py:190-198
            plist = [LP3.SuppressTokensLogitsProcessor3d(suppress_tokens=image_ids + [a["boi"], a["eoi"]])]
        elif mode == "interleaved-text-image":   # :241-248
            plist = three
        else:                              # image-only, :200-240
            plist = three + [
                LP3.SuppressTokensLogitsProcessor3d(suppress_tokens=[t for t in range(V) if t not in set(allowed)]),
                LP3.SuppressTokensAtBeginLogitsProcessor3d(begin_suppress_tokens=[a["eos"]],
                                                           begin_index=len(case["prompt"]))]
        procs = LogitsProcessorList(plist + [TopKLogitsWarper(top_k=case["image_top_k"])])
    else:
        procs = LogitsProcessorList([TopKLogitsWarper(top_k=case["image_top_k"]),
                                     TopPLogitsWarper3d(top_p=case.get("top_p", 1.0))])
    gc = GenerationConfig(max_new_tokens=case["max_length"], max_length=case["max_length"], temperature=1.0,
                          top_k=None, do_sample=case["do_sample"], eos_token_id=case["eos"] or None)
    gc._pad_token_tensor = torch.tensor(0) if case["eos"] else None
    crit = [MaxLengthCriteria(case["max_length"])]
    if case["eos"]:
        crit.append(EosTokenCriteria(eos_token_id=case["eos"]))
    trace = []
    orig = J.prefix_matching_next_tokens

    def spy(*a, **k):
        r = orig(*a, **k)
        trace.append({"W": int(k["model_input_ids"].shape[1]), "n_new": int(r[1].shape[1]),
                      "tokens": [int(t) for t in r[1][0]]})
        return r

    J.prefix_matching_next_tokens = spy
    try:
        prompt = torch.tensor([case["prompt"]])
        import contextlib
        import io
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            out = m._sample(prompt, logits_processor=procs, stopping_criteria=StoppingCriteriaList(crit),
                            generation_config=gc, synced_gpus=False, streamer=None,
                            attention_mask=torch.ones_like(prompt), past_key_values=Cache(), use_cache=True)
    finally:
        J.prefix_matching_next_tokens = orig
    return {"ids": [int(t) for t in out[0]], "trace": trace}


_ANOLE = dict(boi=8197, eoi=8196, eos=2, image=[4, 8196], image_seq_length=24)
_ANOLE_TEXT = dict(_ANOLE, mode="text-only")
_ANOLE_MIX = dict(_ANOLE, mode="interleaved-text-image")
_EMU3 = dict(height=4, width=6, img_token=900, eol=901, eof=902, eoi=903, eos=904, pad=905, visual=[1000, 3048])
LOOP_CASES = {
    # Anole (a12): five 3-D Chameleon processors + TopK(50); boi forced first, 24 image tokens, eoi only if a trip
    # ends exactly on the offset (else the window jumps over it), max_new_tokens = S + 2 like model_loader.py:404-409
    "anole_spec_w8": dict(V=9216, sharp=14.0, grammar="anole", anole=_ANOLE, image_top_k=50, text_top_k=10,
                          prompt=[0, 300, 400, 500], img_vocab=[4, 8196], do_sample=True, eos=[2],
                          max_length=4 + 24 + 2,
                          jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=24 + 4, max_num_new_tokens=8,
                                      guidance_scale=3.0, seed=6, multi_token_init_scheme="random", do_cfg=True,
                                      prefix_token_sampler_scheme="speculative_jacobi")),
    "anole_spec_w4_hits_eoi": dict(V=9216, sharp=9.0, grammar="anole", anole=_ANOLE, image_top_k=50, text_top_k=10,
                                   prompt=[0, 300, 400], img_vocab=[4, 8196], do_sample=True, eos=[2],
                                   max_length=3 + 24 + 2,
                                   jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=20, max_num_new_tokens=4,
                                               guidance_scale=7.0, seed=11, multi_token_init_scheme="random", do_cfg=True,
                                               prefix_token_sampler_scheme="speculative_jacobi")),
    # Anole's other generation modes (a12, round 2): text-only = everything but image ids / boi / eoi; interleaved = image
    # ids for S tokens after a begin-of-image, end-of-image forced right after, text (no image ids, no eoi, boi only early)
    # elsewhere.  The interleaved prompt ends in begin-of-image so that the run crosses image -> eoi -> text.
    "anole_text_only_w8": dict(V=9216, sharp=12.0, grammar="anole", anole=_ANOLE_TEXT, image_top_k=50, text_top_k=10,
                               prompt=[0, 300, 400, 500], img_vocab=[4, 8196], do_sample=True, eos=[2],
                               max_length=4 + 40,
                               jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=30, max_num_new_tokens=8,
                                           guidance_scale=3.0, seed=8, multi_token_init_scheme="random", do_cfg=True,
                                           prefix_token_sampler_scheme="speculative_jacobi")),
    "anole_interleaved_w6": dict(V=9216, sharp=12.0, grammar="anole", anole=_ANOLE_MIX, image_top_k=50, text_top_k=10,
                                 prompt=[0, 300, 400, 8197], img_vocab=[4, 8196], do_sample=True, eos=[2],
                                 max_length=4 + 24 + 1 + 70,
                                 jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=24 + 1 + 9, max_num_new_tokens=6,
                                             guidance_scale=3.0, seed=12, multi_token_init_scheme="random", do_cfg=True,
                                             prefix_token_sampler_scheme="speculative_jacobi")),
    # LlamaGen processors with a real nucleus: TopK(100) then TopPLogitsWarper3d(0.8) (a10)
    "plain_topk_topp_spec_w16": dict(V=1024, sharp=12.0, grammar="plain", image_top_k=100, text_top_k=10, top_p=0.8,
                                     prompt=[207], img_vocab=[0, 1024], do_sample=True, eos=[], max_length=100,
                                     jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=90,
                                                 max_num_new_tokens=16, guidance_scale=4.0, seed=1,
                                                 multi_token_init_scheme="random", do_cfg=True,
                                                 prefix_token_sampler_scheme="speculative_jacobi")),
    "plain_topp_only_jacobi_w8": dict(V=1024, sharp=6.0, grammar="plain", image_top_k=1024, text_top_k=10, top_p=0.5,
                                      prompt=[3], img_vocab=[0, 1024], do_sample=True, eos=[], max_length=60,
                                      jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=50,
                                                  max_num_new_tokens=8, guidance_scale=2.0, seed=4,
                                                  multi_token_init_scheme="random", do_cfg=True,
                                                  prefix_token_sampler_scheme="jacobi")),
    # Emu3 grammar (a11): 4 x 6 grid -> EOL after every 6 visual tokens, then EOF, EOI, EOS (stops the run)
    "emu3_spec_w8": dict(V=4096, sharp=14.0, grammar="emu3", emu3=_EMU3, image_top_k=512, text_top_k=10,
                         prompt=[5, 17, 23, 900], img_vocab=[1000, 3048], do_sample=True, eos=[904],
                         max_length=4 + 7 * 4 + 3 + 4,
                         jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=200, max_num_new_tokens=8,
                                     guidance_scale=3.0, seed=2, multi_token_init_scheme="random", do_cfg=True,
                                     prefix_token_sampler_scheme="speculative_jacobi")),
    # ... and past EOS into the forced-PAD region (no EOS criterion), window crossing EOF/EOI/EOS
    "emu3_spec_w16_pad_tail": dict(V=4096, sharp=10.0, grammar="emu3", emu3=_EMU3, image_top_k=300, text_top_k=10,
                                   prompt=[900], img_vocab=[1000, 3048], do_sample=True, eos=[],
                                   max_length=1 + 7 * 4 + 3 + 6,
                                   jacobi=dict(jacobi_loop_interval_l=2, jacobi_loop_interval_r=200,
                                               max_num_new_tokens=16, guidance_scale=2.0, seed=9,
                                               multi_token_init_scheme="random", do_cfg=True,
                                               prefix_token_sampler_scheme="speculative_jacobi")),
    "lumina_spec_w8": dict(V=9216, sharp=16.0, grammar="lumina", image_top_k=2000, text_top_k=10,
                           prompt=[1, 100, 200, 8197, 8808, 8808], img_vocab=[4, 8196], do_sample=True, eos=[8710],
                           max_length=6 + 8 * 9 + 3,
                           jacobi=dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=8 * 8 + 8 - 10,
                                       max_num_new_tokens=8, guidance_scale=3.0, seed=0,
                                       multi_token_init_scheme="random", do_cfg=True,
                                       prefix_token_sampler_scheme="speculative_jacobi")),
    "lumina_spec_w16_cross_eoi": dict(V=9216, sharp=12.0, grammar="lumina", image_top_k=500, text_top_k=10,
                                      prompt=[7, 8197, 8807, 8809], img_vocab=[4, 8196], do_sample=True, eos=[8710],
                                      max_length=4 + 6 * 11 + 12,
                                      jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=200,
                                                  max_num_new_tokens=16, guidance_scale=3.0, seed=3,
                                                  multi_token_init_scheme="random", do_cfg=True,
                                                  prefix_token_sampler_scheme="speculative_jacobi")),
    "lumina_jacobi_greedy_w8": dict(V=9216, sharp=16.0, grammar="lumina", image_top_k=2000, text_top_k=10,
                                    prompt=[1, 100, 200, 8197, 8808, 8808], img_vocab=[4, 8196], do_sample=False,
                                    eos=[8710], max_length=6 + 8 * 9 + 3,
                                    jacobi=dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=8 * 8 + 8 - 10,
                                                max_num_new_tokens=8, guidance_scale=3.0, seed=0,
                                                multi_token_init_scheme="random", do_cfg=True,
                                                prefix_token_sampler_scheme="jacobi")),
    "lumina_jacobi_greedy_w1": dict(V=9216, sharp=16.0, grammar="lumina", image_top_k=2000, text_top_k=10,
                                    prompt=[1, 100, 200, 8197, 8808, 8808], img_vocab=[4, 8196], do_sample=False,
                                    eos=[8710], max_length=6 + 8 * 9 + 3,
                                    jacobi=dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=8 * 8 + 8 - 10,
                                                max_num_new_tokens=1, guidance_scale=3.0, seed=0,
                                                multi_token_init_scheme="random", do_cfg=True,
                                                prefix_token_sampler_scheme="jacobi")),
    "lumina_jacobi_sample_w8": dict(V=9216, sharp=16.0, grammar="lumina", image_top_k=2000, text_top_k=10,
                                    prompt=[1, 100, 200, 8197, 8808, 8808], img_vocab=[4, 8196], do_sample=True,
                                    eos=[8710], max_length=6 + 8 * 9 + 3,
                                    jacobi=dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=8 * 8 + 8 - 10,
                                                max_num_new_tokens=8, guidance_scale=3.0, seed=1,
                                                multi_token_init_scheme="random", do_cfg=True,
                                                prefix_token_sampler_scheme="jacobi")),
    "lumina_spec_greedy_w8": dict(V=9216, sharp=16.0, grammar="lumina", image_top_k=2000, text_top_k=10,
                                  prompt=[1, 100, 200, 8197, 8808, 8808], img_vocab=[4, 8196], do_sample=False,
                                  eos=[8710], max_length=6 + 8 * 9 + 3,
                                  jacobi=dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=8 * 8 + 8 - 10,
                                              max_num_new_tokens=8, guidance_scale=3.0, seed=2,
                                              multi_token_init_scheme="random", do_cfg=True,
                                              prefix_token_sampler_scheme="speculative_jacobi")),
    "lumina_spec_nocfg_w4": dict(V=9216, sharp=16.0, grammar="lumina", image_top_k=50, text_top_k=10,
                                 prompt=[1, 8197, 8806, 8806], img_vocab=[4, 8196], do_sample=True, eos=[8710],
                                 max_length=4 + 4 * 5 + 3,
                                 jacobi=dict(jacobi_loop_interval_l=2, jacobi_loop_interval_r=4 * 4 + 4 - 3,
                                             max_num_new_tokens=4, guidance_scale=1.0, seed=5,
                                             multi_token_init_scheme="random", do_cfg=True,
                                             prefix_token_sampler_scheme="speculative_jacobi")),
    "plain_topk_spec_w16": dict(V=1024, sharp=12.0, grammar="plain", image_top_k=100, text_top_k=10,
                                prompt=[207], img_vocab=[0, 1024], do_sample=True, eos=[], max_length=120,
                                jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=100,
                                            max_num_new_tokens=16, guidance_scale=4.0, seed=0,
                                            multi_token_init_scheme="repeat_horizon", do_cfg=True,
                                            prefix_token_sampler_scheme="speculative_jacobi")),
}


# Spatial draft initialisation with the Lumina grammar (the only grammar that knows the latent width).  The UNMODIFIED
# reference CRASHES on these (IndexError at jacobi_iteration_lumina_mgpt.py:577: it indexes the [B, 1, V] score tensor
# with sequence positions before it reaches the 'repeat' / 'sample' branch), which is why its own drivers use 'random'
# for Lumina (test_lumina_mgpt.py:59) and 'repeat_horizon' only for LlamaGen, where no width is known and the scheme
# degenerates to 'random' (golden plain_topk_spec_w16).  No golden can be minted; tests/ run engine vs oracle on them and
# tests/test_host_cpu.py shows the crash.  (python oracle/mint_golden.py crash  reproduces it.)
UNPINNED_CASES = {
    "lumina_spec_w8_repeat_horizon": dict(V=9216, sharp=16.0, grammar="lumina", image_top_k=2000, text_top_k=10,
                                          prompt=[1, 100, 200, 8197, 8808, 8808], img_vocab=[4, 8196], do_sample=True,
                                          eos=[8710], max_length=6 + 8 * 9 + 3,
                                          jacobi=dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=8 * 8 + 8 - 10,
                                                      max_num_new_tokens=8, guidance_scale=3.0, seed=0,
                                                      multi_token_init_scheme="repeat_horizon", do_cfg=True,
                                                      prefix_token_sampler_scheme="speculative_jacobi")),
    "lumina_jacobi_w16_repeat_horizon": dict(V=9216, sharp=10.0, grammar="lumina", image_top_k=500, text_top_k=10,
                                             prompt=[7, 8197, 8807, 8809], img_vocab=[4, 8196], do_sample=True, eos=[8710],
                                             max_length=4 + 6 * 11 + 12,
                                             jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=200,
                                                         max_num_new_tokens=16, guidance_scale=3.0, seed=3,
                                                         multi_token_init_scheme="repeat_horizon", do_cfg=True,
                                                         prefix_token_sampler_scheme="jacobi")),
}


def mint_loops(out_dir: Path, only=None):
    Cache = apply_shims()
    for name, case in LOOP_CASES.items():
        if only and name not in only:
            continue
        res = run_reference_loop(case, Cache)
        nfe = len(res["trace"])
        n_new = len(res["ids"]) - len(case["prompt"])
        print(f"{name}: {n_new} new tokens in {nfe} NFE ({n_new / nfe:.2f} tok/iter)")
        with open(out_dir / f"sjd_loop_{name}.json", "w") as f:
            json.dump({"case": case, "result": res,
                       "minted_with": {"torch": torch.__version__, "reference": "tyshiwo1/Accelerating-T2I-AR-with-SJD@b389cfb"}},
                      f, separators=(",", ":"))


if __name__ == "__main__":
    out = Path(os.environ.get("SJD_GOLDEN_OUT", REPO / "tests" / "golden"))
    out.mkdir(parents=True, exist_ok=True)
    which = sys.argv[1:] or ["loops", "forward"]
    if "loops" in which:
        mint_loops(out, only=[w for w in which if w in LOOP_CASES])
    if "crash" in which:
        Cache = apply_shims()
        for name, case in UNPINNED_CASES.items():
            try:
                run_reference_loop(case, Cache)
                print(name, "ran (unexpected)")
            except IndexError as e:
                print(name, "-> reference raises IndexError:", e)
    if "forward" in which:
        from oracle import mint_forward_golden as F
        F.mint_llamagen()
        F.mint_chameleon()
