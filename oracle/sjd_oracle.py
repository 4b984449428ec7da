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
class Emu3Grammar:
    """EOLLogitProcessor3d (scheduler/jacobi_iteration_emu3.py:44-128) as index arithmetic: only visual ids
    [visual_lo, visual_hi) are allowed; window positions are forced to EOL every width+1 tokens after the image token,
    to EOF / EOI / EOS at (width+1)*height + 1 / 2 / 3, and — once the window reaches past that — the LAST rows of the
    window (python slice semantics of `batch_scores[start:, :]` with a possibly negative start, :120-125) to PAD.
    Non-visual ids get finfo.min there, which softmax / top-k treat like -inf.  HF appends TopKLogitsWarper(top_k)."""
    height: int = 90
    width: int = 90
    img_token: int = 0
    eol: int = 0
    eof: int = 0
    eoi: int = 0
    eos: int = 0
    pad: int = 0
    visual_lo: int = 0
    visual_hi: int = 0
    top_k: int = 2048

    def describe(self, ids: list[int], n: int) -> dict:
        offset = ids.index(self.img_token)            # first occurrence (offset_cache, :52-54)
        tokenlen = len(ids) - (offset + 1)
        d = {"in_image": True, "allow": (self.visual_lo, self.visual_hi), "forced": [-1] * n, "top_k": self.top_k,
             "no_cfg": False}
        line = self.width + 1
        for line_len, tok in ((line, self.eol), (line * self.height + 1, self.eof), (line * self.height + 2, self.eoi),
                              (line * self.height + 3, self.eos)):
            for pos in eol_positions(tokenlen, n, line_len):
                if 0 <= pos < n:
                    d["forced"][pos] = tok
        limit = line * self.height + 3
        if tokenlen + n > limit:
            for pos in list(range(n))[limit - tokenlen:]:
                d["forced"][pos] = self.pad
        return d


@dataclass
class PlainTopK:
    """LlamaGen / Emu3-style processors without grammar: HF TopKLogitsWarper (+ TopPLogitsWarper3d with p = 1,
    a no-op on probabilities; llamagen/llamagen_solver.py:458-470)."""
    top_k: int = 0
    top_p: float = 1.0

    def describe(self, ids: list[int], n: int) -> dict:
        return {"in_image": True, "allow": None, "forced": [-1] * n, "top_k": self.top_k, "no_cfg": False,
                "top_p": self.top_p}


@dataclass
class AnoleGrammar:
    """The five 3-D Chameleon processors the Anole adaptor installs for `multimodal_generation_mode="image-only"`
    (scheduler/jacobi_iteration_anhole.py:200-240) followed by HF's TopKLogitsWarper, restated as ONE disallowed-id mask
    per call: every processor looks at the ACCEPTED prefix only (its length / a token at a fixed distance from its end)
    and applies its decision to every window position alike (logit_processor_3dim.py:242-256, :280-286, :323-338).
    Disallowed ids are filled with finfo.min (not -inf).  A multi-token accept can therefore jump over the position
    where end-of-image would have been forced; the adaptor truncates to image_seq_length afterwards
    (jacobi_iteration_anhole.py:309-311)."""
    vocab: int = 65536
    boi: int = 8197
    eoi: int = 8196
    eos: int = 2
    image_lo: int = 4
    image_hi: int = 8196          # IMGIMG* ids of the Chameleon vocabulary: 4..8195
    image_seq_length: int = 1024
    max_length: int = 0           # generation_config.max_length (prompt + max_new_tokens)
    begin_index: int = 0          # prompt length (SuppressTokensAtBegin)
    top_k: int = 50               # HF default top_k when do_sample=True
    mode: str = "image-only"      # multimodal_generation_mode (jacobi_iteration_anhole.py:170-265)

    def disallowed(self, ids: list[int]) -> np.ndarray:
        V, S, cur = self.vocab, self.image_seq_length, len(ids)
        img = np.zeros(V, bool)
        img[self.image_lo:self.image_hi] = True
        dis = np.zeros(V, bool)
        if self.mode == "text-only":   # SuppressTokensLogitsProcessor3d(image ids + [boi, eoi])  (:190-198)
            dis |= img
            dis[[self.boi, self.eoi]] = True
            return dis
        if self.mode == "unrestricted":
            return dis
        # AllowOnlyTokensAtRelativeOffsetLogitsProcessor3d(boi, [eoi], offset=S+1, exclusive=True)  (:242-256)
        not_eoi = np.ones(V, bool)
        not_eoi[self.eoi] = False
        off = S + 1
        if cur < off:
            dis |= ~not_eoi
        else:
            dis |= ~(not_eoi ^ (ids[-off] == self.boi))
        # AllowOnlyTokensInRelativeWindowLogitsProcessor3d(boi, image ids, window_width=S, exclusive=True)  (:323-338)
        ww = min(S, cur)
        dis |= ~((~img) ^ (self.boi in ids[-ww:]))
        # SuppressTokensInIndexRangeLogitsProcessor3d([boi], start=max_length - S - 1)  (:280-286)
        if not (self.max_length - S - 1 > cur):
            dis[self.boi] = True
        if self.mode == "interleaved-text-image":   # only the three processors above (:241-248)
            return dis
        # SuppressTokensLogitsProcessor3d(everything but image ids, eos, boi, eoi)
        ok = img.copy()
        ok[[self.eos, self.boi, self.eoi]] = True
        dis |= ~ok
        # SuppressTokensAtBeginLogitsProcessor3d([eos], begin_index): active for begin <= cur <= begin + 1  (:283)
        if not (self.begin_index > cur or cur > self.begin_index + 1):
            dis[self.eos] = True
        return dis

    def describe(self, ids: list[int], n: int) -> dict:
        return {"in_image": True, "allow": None, "forced": [-1] * n, "top_k": self.top_k, "no_cfg": False,
                "masked": self.disallowed(ids)}


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
    if desc.get("masked") is not None:   # Anole processors: masked_fill(..., finfo.min)
        s[:, desc["masked"]] = np.finfo(F32).min
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


def topp_filter(s: np.ndarray, top_p: float) -> np.ndarray:
    """TopPLogitsWarper3d (logit_processor_3dim.py:406-419): ascending sort, softmax, running sum; entries whose
    running sum is <= 1 - top_p are removed, the largest is always kept.  The comparison runs in float32 (a Python
    scalar compared with a float32 tensor)."""
    if top_p >= 1.0:
        return s   # cumulative probability <= 0 only for entries that are -inf already
    order = np.argsort(s, axis=-1, kind="stable")
    srt = np.take_along_axis(s, order, axis=-1)
    cum = np.cumsum(softmax(srt), axis=-1, dtype=F32)
    rm_sorted = cum <= F32(1.0 - top_p)
    rm_sorted[..., -1:] = False
    rm = np.zeros_like(rm_sorted)
    np.put_along_axis(rm, order, rm_sorted, axis=-1)
    out = s.copy()
    out[rm] = NEG_INF
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
    s = topp_filter(s, desc.get("top_p", 1.0))
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
            {"allow": desc["allow"], "forced": [desc["forced"][i - 1]], "top_k": desc["top_k"],
             "top_p": desc.get("top_p", 1.0), "masked": desc.get("masked")}
        text_mode = bool(rdesc.get("text_mode", False))
        rs = apply_grammar(rl[None, :], rdesc)
        if temperature != 1.0:
            rs = (rs / F32(temperature)).astype(F32)
        rs = topk_filter(rs, rdesc["top_k"])
        rs = topp_filter(rs, rdesc.get("top_p", 1.0))
        rp = softmax(rs)
        e2 = noise_e2() if callable(noise_e2) else noise_e2   # callable: drawn only now, like the reference
        tokens[i - 1] = int(multinomial1(rp, np.asarray(e2, F32).reshape(1, -1))[0])
        break
    return VerifyResult(first, rejected, tokens, nxt, p, text_mode)


# ---------------------------------------------------------------------------------------------------
# the decode loop: JacobiSampler._sample (jacobi_iteration_lumina_mgpt.py:912-1249) restated
# ---------------------------------------------------------------------------------------------------
@dataclass
class OracleParams:
    """_init_new_params (:865-910)"""
    jacobi_loop_interval_l: int = 1
    jacobi_loop_interval_r: int = (768 // 16) ** 2 + 768 // 16
    max_num_new_tokens: int = 16
    guidance_scale: float = 3.0
    seed: int | None = 42
    multi_token_init_scheme: str = "random"
    do_cfg: bool = True
    prefix_token_sampler_scheme: str = "speculative_jacobi"


class TorchNoise:
    """Draws noise with torch exactly where the reference does:
    CPU-global randint for fresh drafts (:505-509); generator exponential_ inside torch.multinomial (:118);
    generator rand([B,W,V]) (:260); generator exponential_ for the residual multinomial only on rejection (:237)."""

    def __init__(self, seed, device="cpu"):
        import random
        import torch
        self.torch = torch
        self.g = None
        if seed is not None:  # set_seed (:36-45) + per-call generator (:1021-1023)
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
            self.g = torch.Generator(device).manual_seed(seed)
        self.device = device

    def randint(self, high, n):
        return self.torch.randint(0, high, (1, n))[0].numpy()

    def exponential(self, W, V):
        return self.torch.empty((W, V), dtype=self.torch.float32, device=self.device).exponential_(
            1.0, generator=self.g).cpu().numpy()

    def uniform(self, W, V):
        return self.torch.rand((1, W, V), dtype=self.torch.float32, device=self.device,
                               generator=self.g)[0].cpu().numpy()


def horizon_init(fresh, scheme, ids, carried, img_w, prefill_num):
    """get_multi_token_for_preparation's non-'random' branch (:516-594).  The fresh tokens have already been drawn
    uniformly (the reference draws them first in every scheme, :517-523).  When the grammar knows the latent width, a
    fresh draft at absolute index a whose column (a - origin) % (w + 1) is not 0 becomes a copy of the token at index
    a - 1 of [accepted ids | carried drafts] — clamped to the last known token, so every such draft of one window
    repeats the last known token ('repeat'; :565-586).  origin = prefill_num + 3 (:536)."""
    if scheme == "random" or not fresh:
        return fresh
    if img_w is None:
        return fresh
    width, origin = int(img_w) + 1, prefill_num + 3
    a0 = len(ids) + len(carried)
    if a0 < origin:
        return fresh
    if "horizon" not in scheme:
        raise AssertionError(f"multi_token_init_scheme should be 'horizon' or 'vertical', but got {scheme}")
    if "sample" in scheme:
        raise NotImplementedError("'sample_horizon' indexes a [B, 1, V] score tensor with sequence positions upstream "
                                  "and crashes; not restated")
    if "repeat" not in scheme:
        raise AssertionError(f"multi_token_init_scheme should be 'sample' or 'repeat', but got {scheme}")
    known = list(ids) + list(carried)
    out = list(fresh)
    for r in range(len(fresh)):
        a = a0 + r
        if (a - origin) % width - 1 >= 0:
            out[r] = int(known[min(a - 1, len(known) - 1)])
    return out


def decode(logits_fn, input_ids, *, params: OracleParams, grammar, img_vocab, max_length, eos_ids=(), rows=2,
           do_sample=True, temperature=1.0, noise=None, max_trips=None, trace=None, stop_fn=None, kv_len0=0):
    """Run the SJD loop.  logits_fn(row_tokens: list[list[int]], kv_len: int, n_logit: int) -> float32
    [rows * n_logit, V] is the model forward for one window (tokens at cache slots kv_len..); the loop handles
    window construction (:606-740), draft bookkeeping (:378-430), the window-size rule (:1142-1144) and stopping
    (:1200-1203).  Returns (ids, nfe)."""
    p = params
    ids = [int(t) for t in input_ids]
    cur_len = len(ids)
    do_cfg = bool(p.do_cfg) and p.guidance_scale != 1 and rows == 2
    noise = noise or TorchNoise(p.seed)
    lr = (cur_len + p.jacobi_loop_interval_l, cur_len + p.jacobi_loop_interval_r)
    prefill_num = cur_len - 1      # attention_mask.shape[1] - 1 at entry (:1000)
    out_W = 1
    carried_tokens: list[int] = []
    carried_q: list = []          # distributions the carried drafts were sampled from
    kv_len = kv_len0   # keys cached before the loop (LlamaGen: the condition tokens, llamagen_solver.py:417)
    nfe = 0
    first = True
    img_vocab = np.asarray(img_vocab)
    while True:
        if first:
            window, q_rows = list(ids), [None] * len(ids)
        else:
            n_fill = out_W - 1
            keep_t, keep_q = carried_tokens[:n_fill], carried_q[:n_fill]
            n_rand = max(n_fill - len(carried_tokens), 0)
            fresh = [int(img_vocab[j]) for j in noise.randint(len(img_vocab), n_rand)] if n_rand > 0 else []
            fresh = horizon_init(fresh, p.multi_token_init_scheme, ids, carried_tokens, getattr(grammar, "w", None),
                                 prefill_num)
            window = [ids[-1]] + keep_t + fresh
            q_rows = [None] + keep_q + [None] * n_rand
        W = len(window)
        desc = grammar.describe(ids, out_W)
        logits = logits_fn([window] * rows, kv_len, out_W)
        V = logits.shape[-1]
        wv, qv = window[-out_W:], q_rows[-out_W:]
        e1 = noise.exponential(out_W, V) if do_sample else None
        u = e2 = None
        if out_W > 1 and p.prefix_token_sampler_scheme == "speculative_jacobi":
            rs = noise.uniform(out_W, V)
            u = rs[np.arange(out_W), np.asarray(wv)]

        def lazy_e2():  # the residual noise is drawn only if a rejection happens (:237-240)
            return noise.exponential(1, V)[0]

        def resid_desc(accepted):
            d = grammar.describe(ids + accepted, 1)
            d["text_mode"] = not d["in_image"]
            return d
        res = verify(logits, out_W, desc, np.asarray(wv), qv, has_uncond=(rows == 2),
                     apply_cfg=do_cfg and not desc["no_cfg"], guidance=p.guidance_scale, temperature=temperature,
                     do_sample=do_sample, scheme=p.prefix_token_sampler_scheme, noise_e1=e1, noise_u=u,
                     noise_e2=lazy_e2, residual_desc_fn=resid_desc)
        if first or out_W <= 1:
            new, n_cached = [int(res.tokens[-1])], W
            carried_tokens, carried_q = [], []
        else:
            m = res.matched
            new, n_cached = [int(t) for t in res.tokens[:m]], m
            carried_tokens = [int(t) for t in res.tokens[m:]]
            carried_q = [res.p[j] for j in range(m, out_W)]
        next_W = min(p.max_num_new_tokens, lr[1] - cur_len) if (lr[0] <= cur_len < lr[1]) else 1
        if trace is not None:
            trace.append({"W": W, "n_new": len(new), "rejected": int(res.rejected), "tokens": list(new)})
        ids += new
        kv_len += n_cached
        out_W = next_W
        cur_len = len(ids)
        nfe += 1
        first = False
        if ids[-1] in set(eos_ids) or cur_len >= max_length or (stop_fn and stop_fn(ids)):
            break
        if max_trips is not None and nfe >= max_trips:
            break
    return ids, nfe
