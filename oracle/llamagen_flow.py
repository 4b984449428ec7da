"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's LlamaGen SJD flow
(llamagen/llamagen_solver.py:370-456): condition prefill (:95-104) -> CFG mix -> top-k / temperature -> first image
token from the GLOBAL torch generator (:75-84) -> JacobiSampler._sample on [first token] with the condition tokens
already cached (:436-443).  Pinned by tests/golden/llamagen_flow.json (oracle/mint_llamagen_flow.py)."""
from __future__ import annotations

import numpy as np
import torch

from . import ref_forward as RF
from . import sjd_oracle as O


def build_stack(case: dict, ff: int, norm_eps: float, rope_base: float, emulate_bf16=False, device="cpu"):
    H = case["n_head"]
    Dh = case["dim"] // H
    cfg = RF.StackConfig(case["n_layer"], case["dim"], H, H, Dh, ff, case["vocab"], norm_eps, rope_interleaved=True,
                         family="llamagen")
    w = RF.random_weights(cfg, seed=case["weights_seed"], std=case["weights_std"], device=device)
    g = torch.Generator().manual_seed(case["cls_seed"])
    cls_table = RF.bf16r(torch.randn(case["num_classes"] + 1, case["dim"], generator=g) * case["weights_std"]).to(device)
    cos, sin = RF.rope_tables_llamagen_2d(case["grid"], Dh, rope_base, case["cls_token_num"])
    return cfg, w, cls_table, cos, sin


def first_token(logits_2rows: np.ndarray, cfg_scale: float, temperature: float, top_k: int) -> int:
    """prefill() + sample() (llamagen_solver.py:95-104, :75-84) on the last condition position; consumes the global
    torch generator exactly like torch.multinomial there."""
    c, u = torch.from_numpy(logits_2rows[0:1]), torch.from_numpy(logits_2rows[1:2])
    lg = (u + (c - u) * cfg_scale) if cfg_scale > 1.0 else c
    lg = lg / max(temperature, 1e-5)
    if top_k > 0:
        k = min(top_k, lg.shape[-1])
        lg = lg.masked_fill(lg < torch.topk(lg, k)[0][..., -1, None], float("-inf"))
    probs = torch.softmax(lg, dim=-1)
    return int(torch.multinomial(probs, num_samples=1)[0, 0])


def generate(case: dict, ff: int, norm_eps: float, rope_base: float, trace=None, max_trips=None):
    cfg, w, cls_table, cos, sin = build_stack(case, ff, norm_eps, rope_base)
    T, n_new = case["cls_token_num"], case["grid"] ** 2
    stack = RF.RefStack(cfg, w, cos, sin, rows=2, max_len=T + n_new + case["jacobi"]["max_num_new_tokens"] + 2,
                        emulate_bf16=False)
    cond = cls_table[torch.tensor([case["class_id"], case["num_classes"]])][:, None, :]   # class, CFG null class
    pos = torch.arange(T)[None].repeat(2, 1)
    lg = stack.forward(embeds=cond, rope_pos=pos, kv_len=0, kv_lo=[0, 0], cache_pos=pos, n_logit_tokens=1)
    torch.manual_seed(case["global_seed"])
    tok0 = first_token(lg[:, 0].numpy(), case["cfg_scale"], case["temperature"], case["top_k"])

    def logits_fn(rows_tokens, kv_len, n):
        ids = torch.tensor(rows_tokens)
        W = ids.shape[1]
        p = torch.arange(kv_len, kv_len + W)[None].repeat(2, 1)
        out = stack.forward(ids=ids, rope_pos=p, kv_len=kv_len, kv_lo=[0, 0], cache_pos=p, n_logit_tokens=n)
        return out.reshape(-1, cfg.vocab).numpy()

    ids, nfe = O.decode(logits_fn, [tok0], params=O.OracleParams(**case["jacobi"]), grammar=O.PlainTopK(top_k=case["top_k"]),
                        img_vocab=np.arange(case["vocab"]), max_length=n_new, eos_ids=[], rows=2, do_sample=True,
                        temperature=1.0, kv_len0=T, trace=trace, max_trips=max_trips)
    return (ids[-n_new:] if max_trips is None else ids), nfe
