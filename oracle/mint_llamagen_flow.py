"""ORACLE tooling (build container only, needs /root/reference): golden of the reference's COMPLETE LlamaGen SJD flow
(config 1 of BASELINE.json, the test_llamagen.py call sequence at test_llamagen.py:85-88,151-169):

    GPT (llamagen/llamagen.py) -> renew_llamagen -> renew_sampler -> LlamaGenSolver.generate
    = condition prefill + first-token sample (llamagen_solver.py:95-104, :75-84) + JacobiSampler._sample (:443)

run UNMODIFIED on CPU in fp32 (HF-5.5 compat shims only) on a small class-conditional GPT whose weights are
oracle.ref_forward.random_weights(seed) (so the test can regenerate them).  Writes tests/golden/llamagen_flow.json.
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

CASE = dict(dim=256, n_layer=2, n_head=4, vocab=1024, grid=8, num_classes=10, cls_token_num=1,
            weights_seed=303, weights_std=0.08, cls_seed=304, class_id=3, cfg_scale=4.0, temperature=1.0,
            top_k=100, top_p=1.0, global_seed=77,
            jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=8 * 8 - 8 - 2, max_num_new_tokens=8,
                        guidance_scale=4.0, seed=5, multi_token_init_scheme="random", do_cfg=True,
                        prefix_token_sampler_scheme="speculative_jacobi"))


# BASELINE config 1 at its stated size: class-conditional GPT-B (12 layers, d 768, 12 heads, 16 384 codes), 256 x 256 ->
# 16 x 16 = 256 image tokens, window 16, cfg 4, top-k 1000 (test_llamagen.py:27-50 uses window 16 / top-k 1000 as well)
CASE_GPTB = dict(dim=768, n_layer=12, n_head=12, vocab=16384, grid=16, num_classes=1000, cls_token_num=1,
                 weights_seed=311, weights_std=0.04, cls_seed=312, class_id=207, cfg_scale=4.0, temperature=1.0,
                 top_k=1000, top_p=1.0, global_seed=78,
                 jacobi=dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=16 * 16 - 16 - 2, max_num_new_tokens=16,
                             guidance_scale=4.0, seed=6, multi_token_init_scheme="random", do_cfg=True,
                             prefix_token_sampler_scheme="speculative_jacobi"))
CASES = {"toy": (CASE, "llamagen_flow.json"), "gptb": (CASE_GPTB, "llamagen_flow_gptb.json")}


@torch.no_grad()
def main(which="toy"):
    CASE, fname = CASES[which]
    from oracle.mint_golden import apply_shims, load_reference_scheduler
    apply_shims()
    J, _ = load_reference_scheduler()
    sys.path.insert(0, str(REF))
    from llamagen.llamagen import ModelArgs, Transformer
    from llamagen.llamagen_solver import LlamaGenSolver, renew_llamagen
    from oracle import ref_forward as RF
    c = CASE
    args = ModelArgs(dim=c["dim"], n_layer=c["n_layer"], n_head=c["n_head"], vocab_size=c["vocab"],
                     block_size=c["grid"] ** 2, cls_token_num=c["cls_token_num"], num_classes=c["num_classes"],
                     model_type="c2i", class_dropout_prob=0.1)
    m = Transformer(args).float().eval()
    ff = m.layers[0].feed_forward.w1.weight.shape[0]
    rcfg = RF.StackConfig(args.n_layer, args.dim, args.n_head, args.n_head, args.dim // args.n_head, ff, args.vocab_size,
                          args.norm_eps, rope_interleaved=True, family="llamagen")
    w = RF.random_weights(rcfg, seed=c["weights_seed"], std=c["weights_std"])
    m.tok_embeddings.weight.copy_(w["embed"]); m.norm.weight.copy_(w["final_norm"]); m.output.weight.copy_(w["lm_head"])
    for L, wl in zip(m.layers, w["layers"]):
        L.attention_norm.weight.copy_(wl["attn_norm"]); L.attention.wqkv.weight.copy_(wl["wqkv"])
        L.attention.wo.weight.copy_(wl["wo"]); L.ffn_norm.weight.copy_(wl["ffn_norm"])
        L.feed_forward.w1.weight.copy_(wl["w_gate_up"][:ff]); L.feed_forward.w3.weight.copy_(wl["w_gate_up"][ff:])
        L.feed_forward.w2.weight.copy_(wl["w_down"])
    g = torch.Generator().manual_seed(c["cls_seed"])
    m.cls_embedding.embedding_table.weight.copy_(
        RF.bf16r(torch.randn(m.cls_embedding.embedding_table.weight.shape, generator=g) * c["weights_std"]))
    m.__class__ = renew_llamagen(m.__class__)
    m._init_new_params(**c["jacobi"])
    m.__class__ = J.renew_sampler(m.__class__)
    m._init_new_params(use_chameleon_tokenizer=False, **c["jacobi"])
    m.img_vocab = torch.arange(c["vocab"])
    solver = LlamaGenSolver(m, c["top_k"], c["top_p"])
    torch.manual_seed(c["global_seed"])        # the first image token is drawn from the GLOBAL generator (:81)
    out = solver.generate(torch.tensor([c["class_id"]]), c["grid"] ** 2, None, cfg_scale=c["cfg_scale"],
                          temperature=c["temperature"], top_k=c["top_k"], top_p=c["top_p"], sample_logits=True)
    res = {"tokens": out[0].tolist(), "ff": ff, "norm_eps": args.norm_eps, "rope_base": args.rope_base}
    (REPO / "tests" / "golden" / fname).write_text(json.dumps({"case": c, "result": res}, separators=(",", ":")))
    print(fname + ":", len(res["tokens"]), "tokens", res["tokens"][:12])


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["toy", "gptb"]):
        main(name)
