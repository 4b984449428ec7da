"""ORACLE tooling (build container only, needs /root/reference): the forward-INCLUDED end-to-end golden.

  python oracle/mint_e2e_golden.py        # writes tests/golden/e2e_chameleon_greedy_jacobi_{cfg3_4x4,nocfg_6x6}.json

Runs the UNMODIFIED reference — vendored ChameleonForConditionalGeneration (lumina_mgpt/model/chameleon/
modeling_chameleon.py) with the reference's renewed backbone mask (scheduler/jacobi_iteration_lumina_mgpt.py:1253-1338)
and JacobiSampler._sample (:912-1249), Lumina 3-D processors (scheduler/logit_processor_3dim.py) — greedy,
prefix_token_sampler_scheme='jacobi', window 8, CFG 3.0, on the tiny decoder of oracle/e2e_case.py, on CPU, twice:
in fp32 and with the module cast to bf16 (the precision the reference ships with).  Both runs must give the same
tokens; the fixture stores them with the per-iteration trace and the smallest top-1 margin seen (in bf16 ulp of the
logit scale), plus the outcome of a sensitivity probe (attention output zeroed -> different tokens).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("SJD_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))
sys.dont_write_bytecode = True


def build_reference_model(case, J, dtype=torch.float32):
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF / "lumina_mgpt"))
    from model.chameleon.configuration_chameleon import ChameleonConfig
    ChameleonConfig.rope_scaling = None   # HF 5.5 auto-fills a dict the vendored code cannot parse
    from model.chameleon.modeling_chameleon import ChameleonForConditionalGeneration
    from oracle import e2e_case
    H = case["n_heads"]
    Dh = case["d_model"] // H
    cfg = ChameleonConfig(vocab_size=case["vocab"], hidden_size=case["d_model"], intermediate_size=case["d_ff"],
                          num_hidden_layers=case["n_layers"], num_attention_heads=H, num_key_value_heads=H,
                          max_position_embeddings=256, rms_norm_eps=case["rms_eps"], mask_image_logits=False,
                          attn_implementation="sdpa", rope_theta=case["rope_theta"], model_parallel_size=H,
                          vocabulary_map={"<image>": 3, "IMGIMGA": 4, "IMGIMGB": 5},
                          vq_config={"embed_dim": 8, "num_embeddings": 16, "resolution": 32, "channel_multiplier": [1, 1],
                                     "base_channels": 32, "num_res_blocks": 1, "latent_channels": 8})
    m = ChameleonForConditionalGeneration(cfg).float().eval()
    w = e2e_case.build_weights(case)
    ff = case["d_ff"]
    with torch.no_grad():
        m.model.embed_tokens.weight.copy_(w["embed"]); m.model.norm.weight.copy_(w["final_norm"])
        m.lm_head.weight.copy_(w["lm_head"])
        for L, wl in zip(m.model.layers, w["layers"]):
            a_, f_ = L.self_attn, L.mlp
            L.input_layernorm.weight.copy_(wl["attn_norm"]); L.post_attention_layernorm.weight.copy_(wl["ffn_norm"])
            hd = H * Dh
            a_.q_proj.weight.copy_(wl["wqkv"][:hd]); a_.k_proj.weight.copy_(wl["wqkv"][hd:2 * hd])
            a_.v_proj.weight.copy_(wl["wqkv"][2 * hd:]); a_.o_proj.weight.copy_(wl["wo"])
            f_.gate_proj.weight.copy_(wl["w_gate_up"][:ff]); f_.up_proj.weight.copy_(wl["w_gate_up"][ff:])
            f_.down_proj.weight.copy_(wl["w_down"])
            for mod, kw, kb in ((a_.q_norm, "q_norm_w", "q_norm_b"), (a_.k_norm, "k_norm_w", "k_norm_b")):
                assert tuple(mod.weight.shape) == (H, Dh), mod.weight.shape
                mod.weight.copy_(wl[kw]); mod.bias.copy_(wl[kb])
    m = m.to(dtype)
    from transformers.generation.utils import GenerationMixin
    if not isinstance(m, GenerationMixin):   # HF >= 4.50: PreTrainedModel no longer inherits it (compat shim)
        m.__class__ = type("ChameleonForConditionalGeneration", (m.__class__, GenerationMixin), {})
    m.model.__class__ = J.renew_backbone(m.model.__class__)
    m.__class__ = J.renew_sampler(m.__class__)
    m._init_new_params(use_chameleon_tokenizer=False, **case["jacobi"])
    lo, hi = case["img_vocab"]
    m.img_vocab = torch.arange(lo, hi)
    return m


def run(case, J, LP3, Cache, dtype=torch.float32, knock_out_attention=False):
    from transformers import GenerationConfig
    from transformers.generation.logits_process import LogitsProcessorList
    from transformers.generation.stopping_criteria import EosTokenCriteria, MaxLengthCriteria, StoppingCriteriaList
    m = build_reference_model(case, J, dtype)
    if knock_out_attention:
        for L in m.model.layers:
            L.self_attn.o_proj.weight.data.zero_()
    V = case["vocab"]
    procs = LogitsProcessorList([
        LP3.MultiTokensVLLogitsProcessor(8197, 8196, 8803, 32, V),
        LP3.MultiTokensInterleavedTopKLogitsWarper(case["image_top_k"], case["text_top_k"], 8197, 8196)])
    gc = GenerationConfig(max_new_tokens=case["max_length"], max_length=case["max_length"], temperature=1.0, top_k=None,
                          do_sample=False, eos_token_id=case["eos"])
    gc._pad_token_tensor = torch.tensor(0)
    crit = StoppingCriteriaList([MaxLengthCriteria(case["max_length"]), EosTokenCriteria(eos_token_id=case["eos"])])
    trace, margins = [], []
    orig_match, orig_samp = J.prefix_matching_next_tokens, J.sampling_logits2tokens

    def spy_match(*a, **k):
        r = orig_match(*a, **k)
        trace.append({"W": int(k["model_input_ids"].shape[1]), "n_new": int(r[1].shape[1])})
        return r

    def spy_samp(logits, *a, **k):
        toks, probs = orig_samp(logits, *a, **k)
        n = probs.shape[-2]
        top2 = torch.topk(probs.float().reshape(-1, probs.shape[-1]), 2, dim=-1)[0]
        gap = torch.log(top2[:, 0]) - torch.log(top2[:, 1].clamp(min=1e-38))     # logit-space margin of the argmax
        scale = float(logits[:, -n:].float().abs().max())
        ulp = 2.0 ** (torch.floor(torch.log2(torch.tensor(scale))).item() - 7)
        margins.append((gap / ulp).tolist())
        return toks, probs

    J.prefix_matching_next_tokens, J.sampling_logits2tokens = spy_match, spy_samp
    try:
        prompt = torch.tensor([case["prompt"]])
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            out = m._sample(prompt, logits_processor=procs, stopping_criteria=crit, generation_config=gc,
                            synced_gpus=False, streamer=None, attention_mask=torch.ones_like(prompt),
                            past_key_values=Cache(), use_cache=True)
    finally:
        J.prefix_matching_next_tokens, J.sampling_logits2tokens = orig_match, orig_samp
    # decisive argmaxes = the window positions whose outputs were accepted this iteration (they see the correct prefix);
    # the others only decide WHEN a token is accepted (the trace), not WHICH token (greedy Jacobi converges to the AR
    # greedy sequence, SURVEY §4 invariant (i))
    decisive = [x for ms, t in zip(margins, trace) for x in ms[:t["n_new"]] if x < 1e30]
    every = [x for ms in margins for x in ms if x < 1e30]
    return {"ids": [int(t) for t in out[0]], "trace": trace, "min_margin_ulp": min(decisive), "min_margin_all_ulp": min(every)}


def mint(name, target_ulp, max_seeds):
    from oracle.mint_golden import apply_shims, load_reference_scheduler
    from oracle import e2e_case
    Cache = apply_shims()
    J, LP3 = load_reference_scheduler()
    case = dict(e2e_case.CASES[name])
    best = None
    for head_seed in range(1, max_seeds):      # search the lm_head rescaling seed for the widest worst-case margin
        case["head_seed"] = head_seed
        r32 = run(case, J, LP3, Cache, torch.float32)
        if best is None or r32["min_margin_ulp"] > best[1]["min_margin_ulp"]:
            best = (head_seed, r32)
            print(f"{name}: head_seed {head_seed}: min decisive margin {r32['min_margin_ulp']:.1f} ulp, "
                  f"{len(r32['trace'])} NFE", flush=True)
        if r32["min_margin_ulp"] >= target_ulp:
            break
    head_seed, r32 = best
    case["head_seed"] = head_seed
    r16 = run(case, J, LP3, Cache, torch.bfloat16)
    ko = run(case, J, LP3, Cache, torch.float32, knock_out_attention=True)
    assert r16["ids"] == r32["ids"], "bf16 and fp32 reference runs disagree: margins are not robust"
    assert ko["ids"] != r32["ids"], "the token stream does not depend on the attention output"
    changed = sum(a != b for a, b in zip(ko["ids"], r32["ids"]))
    n_new = len(r32["ids"]) - len(case["prompt"])
    print(f"{name}: chosen head_seed {head_seed}: {n_new} tokens in {len(r32['trace'])} NFE, min decisive margin "
          f"{r32['min_margin_ulp']:.1f} ulp (fp32) / {r16['min_margin_ulp']:.1f} ulp (bf16); attention knock-out changes "
          f"{changed} tokens")
    out = {"case": case, "result": {"ids": r32["ids"], "trace": r32["trace"], "trace_bf16": r16["trace"],
                                    "min_margin_ulp_fp32": r32["min_margin_ulp"],
                                    "min_margin_all_ulp_fp32": r32["min_margin_all_ulp"],
                                    "min_margin_ulp_bf16": r16["min_margin_ulp"], "knockout_changed_tokens": changed},
           "minted_with": {"torch": torch.__version__, "reference": "tyshiwo1/Accelerating-T2I-AR-with-SJD@b389cfb"}}
    (REPO / "tests" / "golden" / f"e2e_chameleon_greedy_jacobi_{name}.json").write_text(json.dumps(out, separators=(",", ":")))


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3_4x4", "nocfg_6x6"]
    for name in which:
        mint(name, target_ulp=12.0 if name.startswith("cfg") else 6.0, max_seeds=600)
