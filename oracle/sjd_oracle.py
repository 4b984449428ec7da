"""ORACLE (test infrastructure, not product code): CPU restatement in numpy / pure Python of the reference's
Speculative Jacobi Decoding scheduler.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product path (accelerating-t2i-ar-with-sjd_b200/) never does.

Each function cites the reference lines it follows (paths relative to the reference repo root).
Parity status: the reference ships no tests or golden vectors ("parity unpinned" upstream); this oracle is
pinned against the reference's own code executed on CPU in the build container — oracle/mint_golden.py
drives the unmodified scheduler/jacobi_iteration_lumina_mgpt.py + scheduler/logit_processor_3dim.py and
writes tests/golden/*.json, which tests/test_oracle_golden.py replays through this file.

All arithmetic is float32 like the reference (logits are `.float()`ed, modeling_chameleon.py:1560-1561).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

F32 = np.float32
NEG_INF = F32(-np.inf)


# ---------------------------------------------------------------------------------------------------
# grammar: scheduler/logit_processor_3dim.py:25-155 (MultiTokensVLLogitsProcessor) as index arithmetic
# ---------------------------------------------------------------------------------------------------
def eol_positions(tokenlen: int, new_len: int, line_len: int) -> list[int]:
    """Window positions forced to a line/image terminator: check_eol_in_multitokens + get_eol_in_multitokens
    (logit_processor_3dim.py:25-43).  tokenlen = tokens already generated after <boi,h,w>."""
    L, R = tokenlen + 1, tokenlen + new_len
    lo = L // line_len + 1 if L % line_len != 0 else L // line_len
    hi = R // line_len
    return [line_len * m - (tokenlen + 1) for m in range(lo, hi + 1)]


@dataclass
class LuminaGrammar:
    """State the reference keeps inside MultiTokensVLLogitsProcessor / the interleaved top-k warper."""
    image_start: int = 8197
    image_end: int = 8196
    eol: int = 8803
    grid_base: int = 8804
    allow_lo: int = 4
    allow_hi: int = 8196       # image tokens 4..8195  (logit_processor_3dim.py:65)
    image_top_k: int = 2000
    text_top_k: int = 10
    # cached like the reference caches h/w (they are reset when #start == #end)
    h: int | None = None
    w: int | None = None
    start_index: int | None = None

    def describe(self, ids: list[int], n: int) -> dict:
        """Decisions for an n-position window following the accepted prefix `ids`
        (logit_processor_3dim.py:84-155 and :190-204)."""
        n_start = sum(1 for t in ids if t == self.image_start)
        n_end = sum(1 for t in ids if t == self.image_end)
        d = {"in_image": n_start == n_end + 1, "allow": None, "forced": [-1] * n,
             "top_k": self.image_top_k if n_start == n_end + 1 else self.text_top_k,
             "no_cfg": n_start == n_end}
        if n_start == n_end:
            self.h = self.w = self.start_index = None
            return d
        if n_start != n_end + 1:
            return d
        if self.start_index is None:
            self.start_index = max(i for i, t in enumerate(ids) if t == self.image_start)
        new_token_num = len(ids) - (self.start_index + 1)
        if new_token_num < 2:
            return d
        if self.h is None or self.w is None:
            self.h = (ids[self.start_index + 1] - self.grid_base) * 2
            self.w = (ids[self.start_index + 2] - self.grid_base) * 2
        tokenlen = len(ids) - (self.start_index + 3)
        d["allow"] = (self.allow_lo, self.allow_hi)
        line = self.w + 1
        for pos in eol_positions(tokenlen, n, line):
            if 0 <= pos < n:
                d["forced"][pos] = self.eol
        for pos in eol_positions(tokenlen, n, line * self.h + 1):
            if 0 <= pos < n:
                d["forced"][pos] = self.image_end
        return d


@dataclass
class PlainTopK:
    """LlamaGen / Emu3-style processors without grammar: HF TopKLogitsWarper (+ TopPLogitsWarper3d with p = 1,
    a no-op on probabilities; llamagen/llamagen_solver.py:458-470)."""
    top_k: int = 0

    def describe(self, ids: list[int], n: int) -> dict:
        return {"in_image": True, "allow": None, "forced": [-1] * n, "top_k": self.top_k, "no_cfg": False}


# ---------------------------------------------------------------------------------------------------
# logits -> scores -> probabilities -> tokens   (jacobi_iteration_lumina_mgpt.py:82-132)
# ---------------------------------------------------------------------------------------------------
def cfg_mix(cond: np.ndarray, uncond: np.ndarray, g: float) -> np.ndarray:
    """guidance_scale * (c - u) + u in float32, each op rounded (:104)."""
    g = F32(g)
    return (g * (cond.astype(F32) - uncond.astype(F32))).astype(F32) + uncond.astype(F32)


def apply_grammar(scores: np.ndarray, desc: dict) -> np.ndarray:
    s = scores.astype(F32).copy()
    W, V = s.shape
    if desc["allow"] is not None:
        lo, hi = desc["allow"]
        s[:, :lo] = NEG_INF
        s[:, hi:] = NEG_INF
    for i, tok in enumerate(desc["forced"]):
        if tok >= 0:
            s[i, :] = NEG_INF
            s[i, tok] = F32(0)
    return s


def topk_filter(s: np.ndarray, k: int) -> np.ndarray:
    """scores < (k-th largest) -> -inf; ties kept (logit_processor_3dim.py:201-203; HF TopKLogitsWarper)."""
    if k <= 0:
        return s
    V = s.shape[-1]
    k = min(k, V)
    kth = np.partition(s, V - k, axis=-1)[..., V - k][..., None]
    out = s.copy()
    out[s < kth] = NEG_INF
    return out


def softmax(s: np.ndarray) -> np.ndarray:
    m = s.max(axis=-1, keepdims=True)
    e = np.exp((s - m).astype(F32)).astype(F32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)


def multinomial1(p: np.ndarray, noise_e: np.ndarray) -> np.ndarray:
    """torch.multinomial(p, 1) == argmax(p / Exp(1)) (ATen Distributions.cpp fast path); lowest index on ties."""
    return np.argmax((p / noise_e.astype(F32)).astype(F32), axis=-1)


def logits_to_probs(logits: np.ndarray, W: int, desc: dict, *, has_uncond: bool, apply_cfg: bool, guidance: float,
                    temperature: float = 1.0) -> np.ndarray:
    cond = logits[:W]
    s = cfg_mix(cond, logits[W:2 * W], guidance) if (has_uncond and apply_cfg) else cond.astype(F32)
    s = apply_grammar(s, desc)
    if temperature != 1.0:
        s = (s / F32(temperature)).astype(F32)
    s = topk_filter(s, desc["top_k"])
    return s


# ---------------------------------------------------------------------------------------------------
# verify: SpeculativeSampler.__call__ (:247-315), reject_sampling_single_token (:209-241),
#         find_first_misaligned_token_inds (:317-333), prefix_matching_next_tokens (:335-376)
# ---------------------------------------------------------------------------------------------------
@dataclass
class VerifyResult:
    matched: int
    rejected: bool
    tokens: np.ndarray      # [W] after accept / resample
    next_tokens: np.ndarray  # [W] raw samples from p
    p: np.ndarray           # [W, V]
    text_mode: bool = False


def verify(logits: np.ndarray, W: int, desc: dict, draft: np.ndarray, q_rows: list, *, has_uncond: bool,
           apply_cfg: bool, guidance: float, temperature: float = 1.0, do_sample: bool = True,
           scheme: str = "speculative_jacobi", noise_e1=None, noise_u=None, noise_e2=None,
           residual_desc_fn=None) -> VerifyResult:
    """One Jacobi iteration after the forward.  q_rows[i] is the [V] distribution draft[i] was drawn from, or None
    for a fresh random draft (one-hot at draft[i], jacobi_iteration_lumina_mgpt.py:511-514).
    residual_desc_fn(accepted_tokens) -> grammar desc for the single resampled position (the reference re-runs
    the processors on input_ids + accepted tokens, :297-306)."""
    s = logits_to_probs(logits, W, desc, has_uncond=has_uncond, apply_cfg=apply_cfg, guidance=guidance,
                        temperature=temperature)
    p = softmax(s)
    if do_sample:
        nxt = multinomial1(p, noise_e1[:W])
    else:
        nxt = np.argmax(s, axis=-1)
    nxt = nxt.astype(np.int64)
    tokens = nxt.copy()
    if W == 1:  # prefilling phase / AR step: the single new token is taken (:344-350)
        return VerifyResult(1, False, tokens, nxt, p)
    if scheme == "jacobi":
        first = W
        for i in range(1, W):
            if int(draft[i]) != int(nxt[i - 1]):
                first = i
                break
        return VerifyResult(first, False, tokens, nxt, p)
    first, rejected, text_mode = W, False, False
    for i in range(1, W):
        x = int(draft[i])
        px = p[i - 1, x]
        qx = F32(1.0) if q_rows[i] is None else F32(q_rows[i][x])
        ratio = np.minimum(F32(px) / qx, F32(1.0))
        if F32(noise_u[i]) < ratio:
            tokens[i - 1] = x
            continue
        first, rejected = i, True
        q = np.zeros(p.shape[1], F32)
        if q_rows[i] is None:
            q[x] = 1.0
        else:
            q = q_rows[i].astype(F32)
        with np.errstate(divide="ignore"):
            rl = np.log(np.maximum((p[i - 1] - q).astype(F32), F32(0))).astype(F32)
        rdesc = residual_desc_fn([int(t) for t in tokens[: i - 1]]) if residual_desc_fn else \
            {"allow": desc["allow"], "forced": [desc["forced"][i - 1]], "top_k": desc["top_k"]}
        text_mode = bool(rdesc.get("text_mode", False))
        rs = apply_grammar(rl[None, :], rdesc)
        if temperature != 1.0:
            rs = (rs / F32(temperature)).astype(F32)
        rs = topk_filter(rs, rdesc["top_k"])
        rp = softmax(rs)
        tokens[i - 1] = int(multinomial1(rp, noise_e2[None, :])[0])
        break
    return VerifyResult(first, rejected, tokens, nxt, p, text_mode)
