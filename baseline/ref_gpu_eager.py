"""Bench tooling (not product code): time the UNMODIFIED reference's PyTorch-eager SJD — JacobiSampler._sample
(scheduler/jacobi_iteration_lumina_mgpt.py:912-1249) over the vendored ChameleonForConditionalGeneration
(lumina_mgpt/model/chameleon/modeling_chameleon.py:1494-1591) with the reference's renewed backbone mask (:1253-1338) and
3-D processors (scheduler/logit_processor_3dim.py) — on the GPU of this box, at the same shape, precision, window,
guidance and top-k as bench.py's own arm (BASELINE config 2: Lumina-mGPT-7B shape, random-init bf16 weights).
This is the denominator of BASELINE.json's target ">= 2x the reference's own single-GPU PyTorch SJD".

The reference is imported from baseline/_ref (scripts/install_reference.py; git-ignored, travels with gpurun) — never
from /root/reference at run time.  The image ships transformers 5.5 while the reference pins 4.47.1, so the eight
aliases of SURVEY.md Appendix C are applied first (oracle/mint_golden.apply_shims: removed names only, no reference
logic changes).  The time reported is the reference's OWN timer ("Time elapsed inner", :1050-1055,:1213-1223: CUDA
events around its while-loop, prefill included as its first iteration) divided by its own NFE counter.
"""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF_ROOT = REPO / "baseline" / "_ref"


def available() -> str | None:
    """None when the reference install is usable, else the reason (one line)."""
    if not (REF_ROOT / "scheduler" / "jacobi_iteration_lumina_mgpt.py").exists():
        return "baseline/_ref is missing (run scripts/install_reference.py in the build container)"
    return None


def _load():
    os.environ["SJD_REFERENCE"] = str(REF_ROOT)
    if str(REPO) not in sys.path:
        sys.path.insert(0, str(REPO))
    from oracle import mint_golden as MG     # the compat shims + by-path loader of the reference scheduler
    MG.REF = REF_ROOT
    Cache = MG.apply_shims()
    J, LP3 = MG.load_reference_scheduler()
    for p in (str(REF_ROOT), str(REF_ROOT / "lumina_mgpt")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from model.chameleon.configuration_chameleon import ChameleonConfig
    ChameleonConfig.rope_scaling = None      # HF 5.5 auto-fills a dict the vendored code cannot parse
    from model.chameleon.modeling_chameleon import ChameleonForConditionalGeneration
    return J, LP3, Cache, ChameleonConfig, ChameleonForConditionalGeneration


def build_model(device, *, n_layers=32, d_model=4096, n_heads=32, d_ff=11008, vocab=65536, max_pos=4096, seed=0,
                std=0.02, dtype=None):
    """Random-init Lumina-mGPT-7B-shaped ChameleonForConditionalGeneration, created directly on `device` in bf16."""
    import torch
    from transformers.generation.utils import GenerationMixin
    J, LP3, Cache, ChameleonConfig, Model = _load()
    dtype = dtype or torch.bfloat16
    cfg = ChameleonConfig(vocab_size=vocab, hidden_size=d_model, intermediate_size=d_ff, num_hidden_layers=n_layers,
                          num_attention_heads=n_heads, num_key_value_heads=n_heads, max_position_embeddings=max_pos,
                          rms_norm_eps=1e-5, mask_image_logits=False, attn_implementation="sdpa",
                          vocabulary_map={"<image>": 3, "IMGIMGA": 4, "IMGIMGB": 5},
                          vq_config={"embed_dim": 8, "num_embeddings": 16, "resolution": 32, "channel_multiplier": [1, 1],
                                     "base_channels": 32, "num_res_blocks": 1, "latent_channels": 8})
    # the reference's demo scripts switch the default initialisers off for speed (test_llamagen.py:11-12)
    keep = (torch.nn.Linear.reset_parameters, torch.nn.LayerNorm.reset_parameters, torch.nn.Embedding.reset_parameters)
    torch.nn.Linear.reset_parameters = torch.nn.LayerNorm.reset_parameters = torch.nn.Embedding.reset_parameters = \
        lambda self: None
    old = torch.get_default_dtype()
    try:
        torch.set_default_dtype(dtype)
        with torch.device(device):
            try:
                from transformers.modeling_utils import no_init_weights
                ctx = no_init_weights()
            except Exception:   # pragma: no cover
                ctx = contextlib.nullcontext()
            with ctx:
                m = Model(cfg)
    finally:
        torch.set_default_dtype(old)
        torch.nn.Linear.reset_parameters, torch.nn.LayerNorm.reset_parameters, torch.nn.Embedding.reset_parameters = keep
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() >= 2 and "norm" not in name:
                p.copy_((torch.randn(p.shape, generator=g, device=device, dtype=torch.float32) * std).to(p.dtype))
            elif name.endswith("bias"):
                p.zero_()
            else:
                p.fill_(1.0)
    m = m.to(dtype).eval()
    if not isinstance(m, GenerationMixin):   # HF >= 4.50: PreTrainedModel no longer inherits it (compat shim)
        m.__class__ = type("ChameleonForConditionalGeneration", (m.__class__, GenerationMixin), {})
    m.model.__class__ = J.renew_backbone(m.model.__class__)
    m.__class__ = J.renew_sampler(m.__class__)
    return m, (J, LP3, Cache)


def run(model, mods, *, prompt, max_length, window=32, guidance=3.0, image_top_k=2000, text_top_k=10, seed=0,
        grid=48, scheme="speculative_jacobi", do_sample=True, eos=(8710,)) -> dict:
    """One call of the reference's _sample on `prompt` (list of ids ending in <boi> h w) up to `max_length` tokens.
    Returns its own timer / NFE counter and the tokens."""
    import torch
    from transformers import GenerationConfig
    from transformers.generation.logits_process import LogitsProcessorList
    from transformers.generation.stopping_criteria import EosTokenCriteria, MaxLengthCriteria, StoppingCriteriaList
    J, LP3, Cache = mods
    dev = next(model.parameters()).device
    model._init_new_params(use_chameleon_tokenizer=False, jacobi_loop_interval_l=3,
                           jacobi_loop_interval_r=grid * grid + grid - 10, max_num_new_tokens=window,
                           guidance_scale=guidance, seed=seed, multi_token_init_scheme="random", do_cfg=True,
                           prefix_token_sampler_scheme=scheme)
    model.img_vocab = torch.arange(4, 8196)
    V = model.config.vocab_size
    procs = LogitsProcessorList([
        LP3.MultiTokensVLLogitsProcessor(8197, 8196, 8803, 32, V, device=dev) if "device" in
        LP3.MultiTokensVLLogitsProcessor.__init__.__code__.co_varnames else LP3.MultiTokensVLLogitsProcessor(8197, 8196, 8803, 32, V),
        LP3.MultiTokensInterleavedTopKLogitsWarper(image_top_k, text_top_k, 8197, 8196)])
    gc = GenerationConfig(max_new_tokens=max_length, max_length=max_length, temperature=1.0, top_k=None,
                          do_sample=do_sample, eos_token_id=list(eos))
    gc._pad_token_tensor = torch.tensor(0, device=dev)
    crit = StoppingCriteriaList([MaxLengthCriteria(max_length), EosTokenCriteria(eos_token_id=list(eos))])
    ids = torch.tensor([list(prompt)], device=dev)
    buf = io.StringIO()
    with torch.no_grad(), contextlib.redirect_stdout(buf):
        out = model._sample(ids, logits_processor=procs, stopping_criteria=crit, generation_config=gc,
                            synced_gpus=False, streamer=None, attention_mask=torch.ones_like(ids),
                            past_key_values=Cache(), use_cache=True)
    text = buf.getvalue()
    t = float(re.search(r"Time elapsed inner:\s*([0-9.eE+-]+)", text).group(1))
    nfe = int(re.search(r"gen loop num \(NFE\):\s*(\d+)", text).group(1))
    n_new = int(out.shape[1]) - len(prompt)
    return {"seconds": t, "nfe": nfe, "new_tokens": n_new, "ms_per_nfe": 1e3 * t / max(nfe, 1),
            "tokens_per_s": n_new / t if t > 0 else 0.0, "accepted_per_iter": n_new / max(nfe, 1),
            "ids": out[0].tolist()}


if __name__ == "__main__":   # tiny CPU/GPU self-check: python baseline/ref_gpu_eager.py
    import torch
    why = available()
    if why:
        print(why)
        sys.exit(0)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    m, mods = build_model(dev, n_layers=2, d_model=256, n_heads=2, d_ff=512, vocab=9216, max_pos=256,
                          dtype=torch.bfloat16 if dev == "cuda" else torch.float32)
    r = run(m, mods, prompt=[1, 100, 200, 8197, 8808, 8808], max_length=6 + 8 * 9 + 3, window=8, grid=8)
    print({k: v for k, v in r.items() if k != "ids"})
