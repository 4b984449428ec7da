"""ORACLE (test infrastructure, not product code): the forward-included end-to-end case — a tiny Chameleon-shaped
decoder whose greedy argmax is robust to bf16 rounding, shared by the minting script (oracle/mint_e2e_golden.py, which
runs the UNMODIFIED reference on it) and by the tests that rebuild the same weights for the oracle stack and the GPU engine.

Why the lm_head rows are rescaled: greedy + 'jacobi' is the one regime in which reference and engine must agree token
for token on real weights (SURVEY §4 invariant (i)), but two correct bf16 pipelines differ by a few bf16 ulp of the
logit scale, and with i.i.d. random weights the top-1 / top-2 gap over 8 192 image ids falls below that in a fifth of
all positions.  Giving the lm_head rows log-normal norms makes the logit distribution heavy-tailed, so the winner leads
by tens of ulp, while WHICH heavy row wins still depends on the direction of the hidden state (the minting script
checks that knocking out the attention output changes the token stream).
"""
from __future__ import annotations

import torch

from . import ref_forward as RF

def _case(grid: int, guidance: float, window: int = 8) -> dict:
    g2 = 8804 + grid // 2
    return dict(
        vocab=9216, d_model=256, n_layers=2, n_heads=2, d_ff=512, rms_eps=1e-5, rope_theta=10000.0,
        weights_seed=0, weights_std=0.08, head_seed=1, head_sigma=1.6,
        prompt=[1, 100, 200, 8197, g2, g2], img_vocab=[4, 8196], eos=[8710], image_top_k=2000, text_top_k=10,
        max_length=6 + grid * (grid + 1) + 3,
        jacobi=dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=grid * grid + grid - 4, max_num_new_tokens=window,
                    guidance_scale=guidance, seed=0, multi_token_init_scheme="random", do_cfg=True,
                    prefix_token_sampler_scheme="jacobi"))


# Two cases, because classifier-free guidance g*(c-u)+u multiplies the bf16 noise of the raw logits by up to 2g-1 = 5:
#   cfg3_4x4   CFG 3.0 (two rows, hidden prompt prefix) on a 4 x 4 latent image: few enough free decisions that a seed
#              with a >= 8 ulp worst-case margin exists;
#   nocfg_6x6  guidance 1.0 (the reference then decodes a single row, :1005) on a 6 x 6 image.
CASES = {"cfg3_4x4": _case(4, 3.0), "nocfg_6x6": _case(6, 1.0)}
CASE = CASES["cfg3_4x4"]


def stack_config(case=CASE) -> RF.StackConfig:
    H = case["n_heads"]
    return RF.StackConfig(case["n_layers"], case["d_model"], H, H, case["d_model"] // H, case["d_ff"], case["vocab"],
                          case["rms_eps"], qk_norm=True, rope_theta=case["rope_theta"])


def build_weights(case=CASE, device="cpu") -> dict:
    """random_weights(seed) with the lm_head rows rescaled by exp(N(0, head_sigma)) (values stay bf16-representable)."""
    w = RF.random_weights(stack_config(case), seed=case["weights_seed"], std=case["weights_std"], device="cpu")
    g = torch.Generator().manual_seed(case["head_seed"])
    scale = torch.exp(torch.randn(case["vocab"], generator=g) * case["head_sigma"]).clamp(max=40.0)
    w["lm_head"] = RF.bf16r(w["lm_head"] * scale[:, None])

    def to(x):
        return x.to(device) if torch.is_tensor(x) else x

    out = {k: to(v) for k, v in w.items() if k != "layers"}
    out["layers"] = [{k: to(v) for k, v in L.items()} for L in w["layers"]]
    return out


def oracle_decode(case, emulate_bf16: bool, device="cpu", trace=None):
    """sjd_oracle.decode driving RefStack on the case's weights (CFG-uncond row: prompt prefix hidden, RoPE position =
    slot - first visible key; scheduler/jacobi_iteration_lumina_mgpt.py:742-770).  Returns (ids, nfe)."""
    import numpy as np
    from . import sjd_oracle as O
    cfg = stack_config(case)
    w = build_weights(case, device)
    cos, sin = RF.rope_tables_rotate_half(cfg.head_dim, 256, case["rope_theta"], emulate_bf16)   # a bf16 model rounds the tables
    P = len(case["prompt"])
    do_cfg = case["jacobi"]["do_cfg"] and case["jacobi"]["guidance_scale"] != 1
    rows = 2 if do_cfg else 1
    kv_lo = [0, P - 1][:rows]
    stack = RF.RefStack(cfg, w, cos.to(device), sin.to(device), rows, 256, emulate_bf16=emulate_bf16)

    def logits_fn(rows_tokens, kv_len, n):
        ids = torch.tensor(rows_tokens, device=device)
        W = ids.shape[1]
        pos = torch.arange(kv_len, kv_len + W, device=device)[None].repeat(rows, 1)
        rope = torch.stack([(pos[b] - kv_lo[b]).clamp(min=0) for b in range(rows)])
        out = stack.forward(ids=ids, rope_pos=rope, kv_len=kv_len, kv_lo=kv_lo, cache_pos=pos, n_logit_tokens=n)
        return out.reshape(-1, cfg.vocab).cpu().numpy()

    return O.decode(logits_fn, case["prompt"], params=O.OracleParams(**case["jacobi"]),
                    grammar=O.LuminaGrammar(image_top_k=case["image_top_k"], text_top_k=case["text_top_k"]),
                    img_vocab=np.arange(*case["img_vocab"]), max_length=case["max_length"], eos_ids=case["eos"],
                    rows=rows, do_sample=False, trace=trace)
