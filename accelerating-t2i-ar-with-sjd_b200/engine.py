"""Host side of the SJD loop.  (verify_call: thin wrapper over sjd_verify; the decode loop follows below.)"""
from __future__ import annotations

import os
import ctypes as C

import numpy as _np
import torch

from . import _lib

_scratch: dict = {}


def _buf(key, shape, dtype, device):
    k = (key, tuple(shape), dtype, str(device))
    t = _scratch.get(k)
    if t is None:
        t = torch.empty(shape, dtype=dtype, device=device)
        _scratch[k] = t
    return t


def verify_call(logits: torch.Tensor, W: int, V: int, desc: dict, draft: torch.Tensor, q_row: torch.Tensor | None,
                p_prev: torch.Tensor | None, *, has_uncond: bool, apply_cfg: bool, guidance: float,
                temperature: float, do_sample: bool, scheme: int, noise_e1=None, noise_u=None, noise_e2=None,
                eoi_token: int = -1, text_top_k: int = 0, p_cur: torch.Tensor | None = None,
                sync: bool = True, rng: "PhiloxNoise | None" = None) -> dict:
    """Run the device verify step on fp32 logits [(2|1)*W, V].  `desc` carries the grammar decision for this
    window: {'allow': (lo, hi) | None, 'forced': [W ints], 'top_k': int}.  Returns device tensors and, when
    `sync`, the host ints `matched` / `rejected`."""
    dev = logits.device
    L = _lib.lib()
    forced = torch.tensor(desc["forced"], dtype=torch.int32, device=dev) if any(t >= 0 for t in desc["forced"]) else None
    if p_cur is None:
        p_cur = torch.empty(W, V, dtype=torch.float32, device=dev)
    a = _lib.VerifyArgs()
    a.logits = logits.data_ptr()
    a.W, a.V = W, V
    a.has_uncond, a.apply_cfg = int(has_uncond), int(apply_cfg)
    a.guidance, a.temperature = float(guidance), float(temperature)
    a.allow_lo, a.allow_hi = desc["allow"] if desc.get("allow") else (0, 0)
    a.allow_mode = int(desc.get("allow_mode", 0))
    a.ban[0], a.ban[1] = desc.get("ban", (-1, -1))
    a.forced = forced.data_ptr() if forced is not None else None
    a.top_k, a.do_sample, a.scheme = int(desc["top_k"]), int(do_sample), int(scheme)
    a.top_p_thresh = top_p_threshold(desc.get("top_p", 1.0))
    a.draft = draft.data_ptr()
    a.q_row = q_row.data_ptr() if q_row is not None else None
    a.p_prev = p_prev.data_ptr() if p_prev is not None else None
    a.p_cur = p_cur.data_ptr()
    a.noise_e1 = noise_e1.data_ptr() if noise_e1 is not None else None
    a.noise_u = noise_u.data_ptr() if noise_u is not None else None
    a.noise_e2 = noise_e2.data_ptr() if noise_e2 is not None else None
    if rng is not None:   # device-side noise: the three draws of one trip, in the reference's order
        a.rng_mode, a.rng_seed = 1, rng.seed
        a.rng_off[0], a.rng_span[0] = rng.draw(W * V)
        a.rng_off[1], a.rng_span[1] = rng.draw(W * V)
        a.rng_off[2], a.rng_span[2] = rng.draw(V)
    a.eoi_token, a.text_top_k = int(eoi_token), int(text_top_k)
    resid = _buf("resid", (V,), torch.float32, dev)
    nxt = torch.empty(W, dtype=torch.int32, device=dev)
    out_tok = torch.empty(W, dtype=torch.int32, device=dev)
    info = torch.empty(4, dtype=torch.int32, device=dev)
    a.resid, a.next_tokens, a.out_tokens, a.out_info = resid.data_ptr(), nxt.data_ptr(), out_tok.data_ptr(), info.data_ptr()
    k = ("sync_ws", str(dev))
    if k not in _scratch:
        _scratch[k] = torch.zeros(1, dtype=torch.int32, device=dev)   # zero once; the kernel leaves it zero
    a.sync_ws = _scratch[k].data_ptr()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(L.sjd_verify(C.byref(a), C.c_void_p(stream)), "sjd_verify")
    res = {"tokens": out_tok, "next_tokens": nxt, "info": info, "p": p_cur, "_keep": (forced,)}
    if sync:
        h = info.cpu()
        res["matched"], res["rejected"] = int(h[0]), bool(h[1])
    return res


def top_p_threshold(top_p: float) -> float:
    """float32(1 - top_p), the value TopPLogitsWarper3d compares the ascending running probability with
    (scheduler/logit_processor_3dim.py:411); 0 switches the filter off (top_p = 1 removes nothing)."""
    top_p = float(top_p)
    if top_p < 0 or top_p > 1.0:
        raise ValueError(f"`top_p` has to be a float > 0 and < 1, but is {top_p}")
    return float(_np.float32(1.0 - top_p)) if top_p < 1.0 else 0.0


# =====================================================================================================
# The SJD decode loop (host control only: integers, a few pinned-memory copies and two C calls per
# Jacobi iteration).  Mirrors JacobiSampler._sample (scheduler/jacobi_iteration_lumina_mgpt.py:912-1249):
#   window = [last accepted | carried-over drafts | fresh random drafts]   (:606-740, :470-596)
#   forward over the static KV cache, CFG rows batched                        (:1086-1107)
#   verify on device (sjd_verify)                                             (:1113-1140)
#   next window size rule                                                     (:1142-1144)
#   append matched tokens, keep unmatched as drafts, roll the cache back      (:1146-1160, :378-430)
# =====================================================================================================
import random as _random
import time as _time
from dataclasses import dataclass, field

from .model import DeviceStack


class LuminaGrammarState:
    """Incremental host-side state of the Chameleon/Lumina image grammar that the reference recomputes from
    input_ids every iteration (scheduler/logit_processor_3dim.py:84-155, :190-204): image tokens 4..8195 only,
    forced end-of-line every (w+1) tokens, forced end-of-image after h rows; CFG off outside an image
    (check_is_force_no_cfg, scheduler/jacobi_iteration_lumina_mgpt.py:70-80)."""

    def __init__(self, image_start=8197, image_end=8196, eol=8803, grid_base=8804, allow=(4, 8196),
                 image_top_k=2000, text_top_k=10):
        self.image_start, self.image_end, self.eol, self.grid_base = image_start, image_end, eol, grid_base
        self.allow, self.image_top_k, self.text_top_k = allow, image_top_k, text_top_k
        self.eoi_token = image_end
        self.reset()

    def reset(self):
        self.n_start = self.n_end = 0
        self.since_start: list[int] = []   # tokens after the last image-start token
        self.h = self.w = None

    def observe(self, tokens):
        for t in tokens:
            if t == self.image_start:
                self.n_start += 1
                self.since_start = []
                continue
            if t == self.image_end:
                self.n_end += 1
            self.since_start.append(t)

    @property
    def no_cfg(self) -> bool:
        return self.n_start == self.n_end

    def note_residual_call(self, accepted):
        """On a rejection the reference re-runs its processors on the prefix plus the accepted drafts
        (reject_sampling_single_token, :209-241); if those close the image the processor forgets h and w right there
        (logit_processor_3dim.py:89-92), which the NEXT window's draft initialisation reads (:1069) before any other
        processor call."""
        n_start = self.n_start + sum(1 for t in accepted if t == self.image_start)
        n_end = self.n_end + sum(1 for t in accepted if t == self.image_end)
        if n_start == n_end:
            self.h = self.w = None

    def describe(self, n: int) -> dict:
        in_img = self.n_start == self.n_end + 1
        d = {"allow": None, "forced": [-1] * n, "top_k": self.image_top_k if in_img else self.text_top_k}
        if self.n_start == self.n_end:
            self.h = self.w = None
            return d
        if not in_img or len(self.since_start) < 2:
            return d
        if self.h is None:
            self.h = (self.since_start[0] - self.grid_base) * 2
            self.w = (self.since_start[1] - self.grid_base) * 2
        tokenlen = len(self.since_start) - 2
        d["allow"] = self.allow
        for line_len, tok in ((self.w + 1, self.eol), ((self.w + 1) * self.h + 1, self.image_end)):
            lo_, hi_ = tokenlen + 1, tokenlen + n
            first = -(-lo_ // line_len)  # ceil
            for mult in range(first, hi_ // line_len + 1):
                pos = line_len * mult - (tokenlen + 1)
                if 0 <= pos < n:
                    d["forced"][pos] = tok
        return d


class Emu3GrammarState:
    """Host-side state of the Emu3 image grammar (reference: EOLLogitProcessor3d, scheduler/jacobi_iteration_emu3.py:44-128,
    built on Emu3PrefixConstrainedLogitsHelper, emu3/mllm/utils_emu3.py:19-62): only visual ids are allowed; positions
    after the image token are forced to EOL every width+1 tokens, to EOF / EOI / EOS at (width+1)*height + 1 / 2 / 3 and —
    once a window reaches past that — its LAST rows to PAD (the reference slices `batch_scores[start:, :]` with a start
    that goes negative, :120-125; reproduced as is).  CFG is never switched off for Emu3 (the processor has no
    image_start_token_id, jacobi_iteration_lumina_mgpt.py:1086-1096).  HF appends TopKLogitsWarper(top_k)."""
    eoi_token = -1      # no text-mode switch: the grammar is purely positional
    text_top_k = 0
    no_cfg = False

    def __init__(self, height, width, img_token, eol, eof, eoi, eos, pad, visual_lo, visual_hi, top_k=2048):
        self.height, self.width, self.img_token = height, width, img_token
        self.eol, self.eof, self.eoi, self.eos, self.pad = eol, eof, eoi, eos, pad
        self.allow, self.top_k = (visual_lo, visual_hi), top_k
        self.reset()

    def reset(self):
        self.tokenlen = None   # tokens after the first image token; None until it has been seen

    def observe(self, tokens):
        for t in tokens:
            if self.tokenlen is None:
                if t == self.img_token:
                    self.tokenlen = 0
            else:
                self.tokenlen += 1

    def describe(self, n: int) -> dict:
        if self.tokenlen is None:
            raise ValueError("Emu3 grammar: the prompt does not contain the image token")
        d = {"allow": self.allow, "forced": [-1] * n, "top_k": self.top_k}
        line = self.width + 1
        for line_len, tok in ((line, self.eol), (line * self.height + 1, self.eof), (line * self.height + 2, self.eoi),
                              (line * self.height + 3, self.eos)):
            lo_, hi_ = self.tokenlen + 1, self.tokenlen + n
            for mult in range(-(-lo_ // line_len), hi_ // line_len + 1):
                pos = line_len * mult - (self.tokenlen + 1)
                if 0 <= pos < n:
                    d["forced"][pos] = tok
        limit = line * self.height + 3
        if self.tokenlen + n > limit:
            for pos in list(range(n))[limit - self.tokenlen:]:
                d["forced"][pos] = self.pad
        return d

    def describe_residual(self, n: int) -> list:
        """Forced id of the residual distribution at each reject position j: the reference re-runs the processor on a
        ONE-token window after the j accepted tokens (reject_sampling_single_token,
        jacobi_iteration_lumina_mgpt.py:209-241), and this grammar depends on the window length in its PAD tail."""
        out, keep = [], self.tokenlen
        for j in range(n):
            self.tokenlen = keep + j
            out.append(self.describe(1)["forced"][0])
        self.tokenlen = keep
        return out


class PlainTopKState:
    """No grammar: HF TopKLogitsWarper(k) then TopPLogitsWarper3d(p) — LlamaGen (llamagen/llamagen_solver.py:458-470)."""
    eoi_token = -1
    text_top_k = 0
    no_cfg = False

    def __init__(self, top_k=0, top_p=1.0):
        self.top_k, self.top_p = top_k, top_p

    def reset(self):
        pass

    def observe(self, tokens):
        pass

    def describe(self, n: int) -> dict:
        return {"allow": None, "forced": [-1] * n, "top_k": self.top_k, "top_p": self.top_p}


class AnoleGrammarState:
    """Host-side state of the Anole / HF-Chameleon grammar: the 3-D Chameleon processors that
    renew_pipeline_anole.generate installs per `multimodal_generation_mode` (scheduler/jacobi_iteration_anhole.py:170-265;
    classes at scheduler/logit_processor_3dim.py:207-353) plus HF's TopKLogitsWarper.  Every one of them inspects the
    ACCEPTED prefix only (its length, the token S+1 places from its end, whether begin-of-image is among its last S
    tokens) and applies one decision to every window position, so a trip's grammar is one candidate set:

      image-only              image ids [image_lo, image_hi)  while begin-of-image is within the last S accepted tokens,
                              {eoi}                            when begin-of-image sits exactly S+1 places back,
                              a subset of {eos, boi}           otherwise (boi until max_length - S - 1, eos except at the start)
      interleaved-text-image  image ids / {eoi} as above; otherwise everything EXCEPT image ids, eoi and — from index
                              max_length - S - 1 on — boi
      text-only               everything except image ids, boi and eoi
      unrestricted            everything (hf_api maps it to PlainTopKState)

    A window that straddles the end of the image keeps "image ids" for all its positions — the reference behaves the
    same and truncates to image_seq_length afterwards (:309-311).  CFG is never switched off (the first processor has no
    image_start_token_id, jacobi_iteration_lumina_mgpt.py:1086-1096)."""
    eoi_token = -1
    text_top_k = 0
    no_cfg = False

    def __init__(self, boi, eoi, eos, image_lo, image_hi, image_seq_length, max_length, begin_index, top_k=50,
                 mode="image-only"):
        if mode not in ("image-only", "interleaved-text-image", "text-only"):
            raise ValueError(f"Unknown multimodal generation mode: {mode}")
        self.boi, self.eoi, self.eos = int(boi), int(eoi), int(eos)
        self.allow = (int(image_lo), int(image_hi))
        self.S, self.max_length, self.begin_index, self.top_k = int(image_seq_length), int(max_length), int(begin_index), int(top_k)
        self.mode = mode
        self.reset()

    def reset(self):
        self.n = 0          # accepted tokens so far
        self.boi_at = []    # their positions holding begin-of-image

    def observe(self, tokens):
        for t in tokens:
            if t == self.boi:
                self.boi_at.append(self.n)
            self.n += 1

    def _decide(self, n: int):
        """Candidate set after an accepted prefix of length n (positions of boi as observed):
        ('image',) | ('forced', id) | ('pair', id, id) | ('text', ban0, ban1)."""
        S = self.S
        if self.mode == "text-only":
            return ("text", self.boi, self.eoi)
        at_offset = n >= S + 1 and (n - (S + 1)) in self.boi_at          # input_ids[-(S+1)] == boi
        in_window = any(n - min(S, n) <= p < n for p in self.boi_at)      # boi in input_ids[-min(S, n):]
        late = not (self.max_length - S - 1 > n)                          # begin-of-image suppressed from here on
        if in_window:
            if at_offset:
                raise NotImplementedError("Anole grammar: two begin-of-image tokens S+1 apart leave no allowed id")
            return ("image",)
        if at_offset:
            return ("forced", self.eoi)
        if self.mode == "interleaved-text-image":
            return ("text", self.eoi, self.boi if late else -1)
        left = []
        if not (self.begin_index <= n <= self.begin_index + 1):
            left.append(self.eos)
        if not late:
            left.append(self.boi)
        if len(left) == 1:
            return ("forced", left[0])
        if len(left) == 2:
            return ("pair", left[0], left[1])
        raise NotImplementedError(f"Anole grammar: no id is allowed after {n} tokens (max_length too small for an image)")

    def _desc(self, d, n):
        if d[0] == "image":
            return {"allow": self.allow, "forced": [-1] * n, "top_k": self.top_k, "allow_mode": 1, "ban": [-1, -1]}
        if d[0] == "forced":
            return {"allow": None, "forced": [d[1]] * n, "top_k": self.top_k, "allow_mode": 0, "ban": [-1, -1]}
        if d[0] == "pair":
            return {"allow": None, "forced": [-1] * n, "top_k": self.top_k, "allow_mode": 3, "ban": [d[1], d[2]]}
        return {"allow": self.allow, "forced": [-1] * n, "top_k": self.top_k, "allow_mode": 2, "ban": [d[1], d[2]]}

    def describe(self, n: int) -> dict:
        return self._desc(self._decide(self.n), n)

    def describe_residual(self, n: int, drafts=None) -> list:
        """Forced id (encoded -2 - id, see below) of the residual at reject position j: the processors are re-run on the
        prefix plus the j drafts accepted before it (jacobi_iteration_lumina_mgpt.py:297-306).  `drafts` = the window's
        tokens ([0] = last accepted token); only begin-of-image among them moves the triggers.  Positions whose residual
        set is the window's return -1; a single id returns the "mask-forced" code; ONE other multi-id set per window is
        handed to the kernel through `self.resid_desc` (sjd_verify_args.resid_*): the positions after a forced
        begin-of-image (image ids) or after a forced end-of-image (text).  Two different other sets in one window are
        refused."""
        out, other, other_from = [], None, 0
        d0 = self._decide(self.n)
        keep = list(self.boi_at)
        try:
            for j in range(n):
                if j > 0:
                    tok = drafts[j] if drafts is not None else (d0[1] if d0[0] == "forced" else None)
                    if tok == self.boi:
                        self.boi_at.append(self.n + j - 1)
                d = self._decide(self.n + j)
                if d[0] == "forced":
                    out.append(-2 - d[1])   # "mask-forced" code of sjd_verify (include/sjd_b200.h, forced_resid): these
                    # processors fill the other ids with finfo.min rather than writing a one-hot row, which matters when
                    # the forced id has no residual mass
                elif d == d0 and other is None:
                    out.append(-1)
                else:
                    if other is None:
                        other, other_from = d, j
                    elif other != d:   # (also: back to the window's set after another one)
                        raise NotImplementedError("Anole grammar: the residual candidate set changes twice inside one "
                                                  "window (e.g. it straddles index max_length - image_seq_length - 1 right "
                                                  "after an image); use a smaller window there")
                    out.append(-1)
        finally:
            self.boi_at = keep
        self.resid_desc = dict(self._desc(other, 1), resid_from=other_from) if other is not None else None
        return out


@dataclass
class SJDParams:
    """Same names and defaults as the reference's _init_new_params (jacobi_iteration_lumina_mgpt.py:865-910)."""
    jacobi_loop_interval_l: int = 1
    jacobi_loop_interval_r: int = (768 // 16) ** 2 + 768 // 16
    max_num_new_tokens: int = 16
    guidance_scale: float = 3.0
    seed: int | None = 42
    multi_token_init_scheme: str = "random"
    do_cfg: bool = True
    prefix_token_sampler_scheme: str = "speculative_jacobi"


@dataclass
class SJDStats:
    nfe: int = 0
    new_tokens: int = 0
    t_inner: float = 0.0
    trace: list = field(default_factory=list)   # per trip (W, matched, rejected)
    h2d_bytes: int = 0                          # pinned-host -> device bytes copied by the loop
    d2h_bytes: int = 0                          # device -> pinned-host bytes (accepted tokens, counters)
    kv_read_tokens: int = 0                     # sum over forwards of the keys each CFG row's window attended to


def check_init_scheme(scheme: str):
    """multi_token_init_scheme as get_multi_token_for_preparation reads it (jacobi_iteration_lumina_mgpt.py:502-594):
    'random', or a name containing 'horizon' and 'repeat'.  'vertical' asserts upstream (:560); 'sample_horizon' (like
    'repeat_horizon') dies upstream with an IndexError as soon as the Lumina grammar knows the image width (:577) — here
    'repeat_horizon' runs with the semantics the code states (see horizon_init) and 'sample_horizon' is refused."""
    if scheme == "random":
        return
    if "horizon" not in scheme:
        raise ValueError(f"multi_token_init_scheme should be 'horizon' or 'vertical', but got {scheme}")
    if "sample" in scheme:
        raise NotImplementedError("multi_token_init_scheme 'sample_horizon' crashes in the reference (IndexError, "
                                  "jacobi_iteration_lumina_mgpt.py:577) and has no implementation here; use "
                                  "'repeat_horizon' or 'random'")
    if "repeat" not in scheme:
        raise ValueError(f"multi_token_init_scheme should be 'sample' or 'repeat', but got {scheme}")


def horizon_init(fresh, scheme, ids, carried, img_w, prefill_num):
    """Spatial draft initialisation, get_multi_token_for_preparation's non-'random' branch (:516-594).  The fresh tokens
    are drawn uniformly first in every scheme (:517-523, same RNG consumption).  Once the grammar knows the latent
    width w, a fresh draft at absolute sequence index a whose column (a - origin) % (w + 1) is not the first of its row
    becomes a copy of its left neighbour, the token at index a - 1 of [accepted ids | carried drafts] — clamped to the
    last known token (:570-574), so the fresh drafts of one window repeat the last known token up to the next row
    start.  origin = prefill_num + 3 (:536).  Grammars without a width (LlamaGen, Emu3, Anole) keep the random drafts,
    like the reference (img_width = 0, :533-537)."""
    if scheme == "random" or not fresh or img_w is None:
        return fresh
    width, origin = int(img_w) + 1, prefill_num + 3
    a0 = len(ids) + len(carried)
    if a0 < origin:
        return fresh
    out = list(fresh)
    n_known = len(ids) + len(carried)
    for r in range(len(fresh)):
        a = a0 + r
        if (a - origin) % width - 1 >= 0:
            j = min(a - 1, n_known - 1)
            out[r] = int(ids[j]) if j < len(ids) else int(carried[j - len(ids)])
    return out


def set_seed(seed: int):
    """jacobi_iteration_lumina_mgpt.py:36-45"""
    _random.seed(seed)
    _np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class NoiseSource:
    """Noise in the exact order/shape the reference draws it from its torch.Generator
    (SURVEY Appendix B): exponential_ [W,V] for multinomial, rand [1,W,V] for the accept test and, only when a
    draft was rejected, exponential_ [1,V] for the residual multinomial."""

    def __init__(self, seed: int | None, device, gen_device=None):
        """gen_device: where the torch.Generator lives (default: the compute device, like the reference's
        torch.Generator(input_ids.device), :1023).  'cpu' reproduces a CPU run of the reference bit for bit."""
        self.device = torch.device(device)
        self.gen_device = torch.device(gen_device) if gen_device is not None else self.device
        self.g = torch.Generator(self.gen_device).manual_seed(seed) if seed is not None else None

    def multinomial_noise(self, W, V):
        e = torch.empty((W, V), dtype=torch.float32, device=self.gen_device).exponential_(1.0, generator=self.g)
        return e.to(self.device)

    def accept_noise(self, W, V, draft_dev):
        rs = torch.rand((1, W, V), dtype=torch.float32, device=self.gen_device, generator=self.g)
        u = rs[0].gather(1, draft_dev.to(self.gen_device).long()[:, None]).squeeze(1)
        return u.to(self.device).contiguous()

    def residual_noise_speculative(self, V):
        """Draw the residual noise but remember how to un-draw it when no rejection happened."""
        gen = self.g
        if gen is None:
            gen = (torch.cuda.default_generators[self.gen_device.index or 0] if self.gen_device.type == "cuda"
                   else torch.default_generator)
        state = gen.get_state()
        e2 = torch.empty((1, V), dtype=torch.float32, device=self.gen_device).exponential_(1.0, generator=self.g)
        return e2.to(self.device), (gen, state)

    @staticmethod
    def undo(token):
        gen, state = token
        gen.set_state(state)


class PhiloxNoise:
    """The same three draws as NoiseSource on a torch.Generator(device='cuda').manual_seed(seed), but nothing is
    generated on the host side: sjd_verify (rng_mode = 1) computes the consumed elements itself from (seed, philox
    offset, torch's launch geometry).  This object only does torch's bookkeeping: every draw of `numel` elements moves
    the generator's offset by ((numel - 1) // (span * 4) + 1) * 4 with span = 256 * min(SMs * maxThreadsPerSM // 256,
    ceil(numel / 256)) (calc_execution_policy, ATen/native/cuda/DistributionTemplates.h).  Bit parity with torch is
    tested on the GPU (tests/test_gpu_parity.py::test_device_philox_matches_torch_generator)."""

    def __init__(self, seed: int, device):
        self.seed, self.offset = int(seed) & 0xFFFFFFFFFFFFFFFF, 0
        pr = torch.cuda.get_device_properties(device)
        self.max_grid = pr.multi_processor_count * (pr.max_threads_per_multi_processor // 256)

    def span(self, numel: int) -> int:
        return 256 * min(self.max_grid, -(-numel // 256))

    def peek(self, numel: int):
        return self.offset, self.span(numel)

    def draw(self, numel: int):
        off, span = self.offset, self.span(numel)
        self.offset += ((numel - 1) // (span * 4) + 1) * 4
        return off, span


def default_noise(seed, device):
    """Device-side Philox noise when the reference would use a seeded CUDA generator; the torch-tensor path otherwise
    (seed None = torch's global generator, or SJD_HOST_NOISE=1)."""
    import os
    dev = torch.device(device)
    if seed is not None and dev.type == "cuda" and os.environ.get("SJD_HOST_NOISE", "0") != "1":
        return PhiloxNoise(seed, dev)
    return NoiseSource(seed, dev)


class SJDEngine:
    """One prompt, `rows` CFG rows (cond [+ uncond]) — the reference's effective batch."""

    def __init__(self, stack: DeviceStack, params: SJDParams, grammar, img_vocab: torch.Tensor,
                 noise_factory=None):
        self.stack, self.p, self.grammar = stack, params, grammar
        self.img_vocab = img_vocab.cpu().long()
        self.dev = stack.device
        self.V = stack.shape.vocab
        self.rows = stack.rows
        Wmax = _lib.SJD_MAX_TOKENS // self.rows
        self.Wmax = Wmax
        self.pbuf = [torch.zeros(Wmax, self.V, dtype=torch.float32, device=self.dev) for _ in range(2)]
        n_i32 = 3 * _lib.SJD_MAX_TOKENS + 4 * Wmax + 16
        self.h_stage = torch.empty(n_i32, dtype=torch.int32).pin_memory()
        self.h_np = self.h_stage.numpy()   # same pinned memory: filled with numpy (cheaper than torch.tensor per field)
        self._stage_copied = None          # event: the last H2D copy out of h_stage[:3M] has completed
        self.d_stage = torch.empty(n_i32, dtype=torch.int32, device=self.dev)
        self.d_out = torch.empty(4 + Wmax, dtype=torch.int32, device=self.dev)
        # result of a verify step: {matched, rejected, first, text_mode, tokens[Wmax], done flag}.  Pinned host memory the
        # verify kernel writes DIRECTLY (mapped under unified addressing) and flags with a sequence number behind a
        # system-scope fence: the loop polls the flag instead of copying device -> host and synchronising the stream
        # (SJD_ZERO_COPY=0 restores copy + synchronize)
        self.h_out = torch.zeros(5 + Wmax, dtype=torch.int32).pin_memory()
        self.h_out_np = self.h_out.numpy()
        self._flag_idx = 4 + Wmax
        self._seq = 0
        self._zero_copy = os.environ.get("SJD_ZERO_COPY", "1") != "0"
        self._views = {}                    # cached slices of the staging buffers per token-row count
        self._stream = None                 # torch's current stream, looked up once per generate() (~8 us per look-up)
        # SJD_NVTX=1: NVTX ranges "sjd.forward" / "sjd.verify" around the two C calls of an iteration (a timeline aid for
        # nsys / ncu --nvtx; off by default: a push / pop pair costs host time on the per-iteration path)
        self._nvtx = os.environ.get("SJD_NVTX", "0") == "1"
        self.d_nxt = torch.empty(Wmax, dtype=torch.int32, device=self.dev)
        self.resid = torch.empty(self.V, dtype=torch.float32, device=self.dev)
        self.d_sync = torch.zeros(1, dtype=torch.int32, device=self.dev)   # sjd_verify's last-CTA counter (stays zero)
        self.noise_factory = noise_factory or default_noise
        self.lib = _lib.lib()
        self.stats = SJDStats()

    def _img_vocab_np(self):
        iv = self.img_vocab
        if getattr(self, "_ivn_src", None) is not iv:      # hf_api swaps img_vocab between calls
            self._ivn, self._ivn_src = iv.numpy(), iv
        return self._ivn

    # -- one forward over `tokens` per row starting at cache slot kv_len -----------------------------------
    def _forward(self, row_tokens, kv_len, kv_lo, n_logit, embeds=None):
        rows, W = self.rows, len(row_tokens[0])
        M = rows * W
        # the pinned staging buffer is reused by every call: its previous async copy must have left it (a chunked
        # prefill issues several forwards without a stream sync in between; in the decode loop the event is long done)
        if self._stage_copied is not None:
            self._stage_copied.synchronize()
        hn = self.h_np
        pos = _np.arange(kv_len, kv_len + W, dtype=_np.int32)
        for b in range(rows):
            hn[b * W:(b + 1) * W] = row_tokens[b]
            _np.maximum(pos - kv_lo[b], 0, out=hn[M + b * W:M + (b + 1) * W])   # RoPE position = slot - first visible key
            hn[2 * M + b * W:2 * M + (b + 1) * W] = pos
        v = self._views.get(M)
        if v is None:
            ds = self.d_stage
            v = self._views[M] = (ds[:3 * M], self.h_stage[:3 * M], ds[:M], ds[M:2 * M], ds[2 * M:3 * M])
        v[0].copy_(v[1], non_blocking=True)
        if self._stage_copied is None:
            self._stage_copied = torch.cuda.Event()
        stream = self._stream if self._stream is not None else torch.cuda.current_stream(self.dev)
        self._stage_copied.record(stream)
        self.stats.h2d_bytes += 12 * M
        if self._stream is not None and getattr(self.stack, "ctx", None) is not None:   # (test doubles of the stack take no handle)
            return self.stack.forward(W, v[3], v[4], kv_len, kv_lo, ids=None if embeds is not None else v[2], embeds=embeds,
                                      n_logit_tokens=n_logit, stream_handle=stream.cuda_stream)
        return self.stack.forward(W, v[3], v[4], kv_len, kv_lo, ids=None if embeds is not None else v[2], embeds=embeds,
                                  n_logit_tokens=n_logit)

    @torch.no_grad()
    def generate(self, input_ids, *, max_length: int, eos_token_ids=(), do_sample=True, temperature=1.0,
                 kv_len0: int = 0, kv_lo=None, uncond_input_ids=None, stop_fn=None, has_eos_criteria=None,
                 collect_trace=False):
        """input_ids: list[int] tokens not yet in the cache (whole prompt, or LlamaGen's first image token with the
        condition already cached at slots [0, kv_len0)).  uncond_input_ids: tokens fed to the CFG-uncond row on
        the first trip (same length; default = input_ids).  Returns the full token list (prompt + generated)."""
        p, rows, V, dev = self.p, self.rows, self.V, self.dev
        ids = [int(t) for t in input_ids]
        cur_len = len(ids)
        do_cfg = bool(p.do_cfg) and (p.guidance_scale != 1) and rows == 2
        if rows == 2 and not do_cfg:
            raise ValueError("stack was built with 2 CFG rows but CFG is disabled")
        kv_lo = list(kv_lo) if kv_lo is not None else [0] * rows
        scheme = {"speculative_jacobi": 0, "jacobi": 1}.get(p.prefix_token_sampler_scheme)
        if scheme is None:
            raise ValueError(f"prefix_token_sampler_scheme: {p.prefix_token_sampler_scheme}")
        check_init_scheme(p.multi_token_init_scheme)
        prefill_num = cur_len - 1   # attention_mask.shape[1] - 1 at entry (:1000)
        if p.seed is not None:
            set_seed(p.seed)   # python / numpy / torch global RNGs: the CPU one drives the fresh-draft randint
        noise = self.noise_factory(p.seed, dev)
        lr = (cur_len + p.jacobi_loop_interval_l, cur_len + p.jacobi_loop_interval_r)
        grammar = self.grammar
        grammar.reset()
        grammar.observe(ids)
        eos = set(int(e) for e in eos_token_ids)
        stats = SJDStats()
        self.stats = stats

        out_W = 1                 # output_token_num
        carried: list[int] = []   # unmatched next tokens (drafts for the next trip)
        carried_row0 = 0          # row of pbuf[prev] holding carried[0]'s distribution
        cur = 0                   # pbuf index written by the current trip
        kv_len = kv_len0
        first_trip = True
        finished = False
        torch.cuda.synchronize(dev)
        self._stream = torch.cuda.current_stream(dev)
        t0 = _time.perf_counter()
        while not finished:
            # ---- window -------------------------------------------------------------------------------
            if first_trip:
                window = list(ids)
                q_row = [-1] * len(window)
            else:
                n_fill = out_W - 1
                keep = carried[:n_fill] if len(carried) > n_fill else carried
                n_rand = max(n_fill - len(carried), 0)
                fresh = []
                if n_rand > 0:
                    r = torch.randint(0, len(self.img_vocab), (1, n_rand))   # CPU global RNG, like :505-509
                    fresh = self._img_vocab_np()[r.numpy()[0]].tolist()
                    fresh = horizon_init(fresh, p.multi_token_init_scheme, ids, carried, getattr(grammar, "w", None),
                                         prefill_num)
                window = [ids[-1]] + keep + fresh
                q_row = [-1] + [carried_row0 + j for j in range(len(keep))] + [-1] * n_rand
            W = len(window)
            if kv_len + W > self.stack.max_len:
                raise RuntimeError(f"KV cache too small: need {kv_len + W}, have {self.stack.max_len}")
            n_rope = getattr(self.stack, "n_rope_pos", None)
            if n_rope is not None and kv_len + W - min(kv_lo) > n_rope:
                # the reference fails loudly on freqs_cis[input_pos] (llamagen/llamagen.py:386); reading past the table
                # on the device would accept garbage drafts instead
                raise RuntimeError(f"RoPE table too small: position {kv_len + W - min(kv_lo) - 1} needed, table has "
                                   f"{n_rope} rows (jacobi_loop_interval_r / max_num_new_tokens reach past the image)")
            no_cfg = bool(grammar.no_cfg)
            # ---- forward (prefill may be chunked; only the last `n_out` positions need logits) ----------
            n_out = out_W
            if first_trip:
                un = [int(t) for t in uncond_input_ids] if uncond_input_ids is not None else window
                chunk = max(1, _lib.SJD_MAX_TOKENS // rows)
                s = 0
                while s < W:
                    e = min(W, s + chunk)
                    rt = [window[s:e]] + ([un[s:e]] if rows == 2 else [])
                    logits = self._forward(rt, kv_len + s, kv_lo, n_out if e == W else 1)
                    s = e
            else:
                rt = [window] * rows
                if self._nvtx:
                    torch.cuda.nvtx.range_push("sjd.forward")
                logits = self._forward(rt, kv_len, kv_lo, n_out)
                if self._nvtx:
                    torch.cuda.nvtx.range_pop()
                    torch.cuda.nvtx.range_push("sjd.verify")
            # ---- verify ---------------------------------------------------------------------------------
            Wv = n_out
            desc = grammar.describe(Wv)
            hs, ds, hn = self.h_stage, self.d_stage, self.h_np
            base = 3 * _lib.SJD_MAX_TOKENS
            hn[base:base + Wv] = window[-Wv:]
            hn[base + self.Wmax: base + self.Wmax + Wv] = q_row[-Wv:]
            hn[base + 2 * self.Wmax: base + 2 * self.Wmax + Wv] = desc["forced"]
            resid_forced = None
            if hasattr(grammar, "describe_residual"):
                try:
                    resid_forced = grammar.describe_residual(Wv, window[-Wv:])
                except TypeError:
                    resid_forced = grammar.describe_residual(Wv)
            if resid_forced is not None:
                hn[base + 3 * self.Wmax: base + 3 * self.Wmax + Wv] = resid_forced
            ds[base: base + 4 * self.Wmax].copy_(hs[base: base + 4 * self.Wmax], non_blocking=True)
            stats.h2d_bytes += 16 * self.Wmax
            d_draft = ds[base: base + Wv]
            a = _lib.VerifyArgs()
            a.logits = logits.data_ptr()
            a.W, a.V = Wv, V
            a.has_uncond, a.apply_cfg = int(rows == 2), int(do_cfg and not no_cfg)
            a.guidance, a.temperature = float(p.guidance_scale), float(temperature)
            a.allow_lo, a.allow_hi = desc["allow"] if desc["allow"] else (0, 0)
            a.allow_mode = int(desc.get("allow_mode", 0))
            a.ban[0], a.ban[1] = desc.get("ban", (-1, -1))
            a.forced = ds[base + 2 * self.Wmax:].data_ptr()
            a.forced_resid = ds[base + 3 * self.Wmax:].data_ptr() if resid_forced is not None else None
            rd = getattr(grammar, "resid_desc", None) if resid_forced is not None else None
            if rd is not None:   # the residual's own candidate set (Anole: after a forced begin- / end-of-image)
                a.resid_set, a.resid_allow_mode = 1, int(rd.get("allow_mode", 0))
                a.resid_allow_lo, a.resid_allow_hi = rd["allow"] if rd["allow"] else (0, 0)
                a.resid_ban[0], a.resid_ban[1] = rd.get("ban", (-1, -1))
                a.resid_from = int(rd.get("resid_from", 0))
            a.top_k, a.do_sample, a.scheme = int(desc["top_k"]), int(do_sample), scheme
            a.top_p_thresh = top_p_threshold(desc.get("top_p", 1.0))
            a.draft = d_draft.data_ptr()
            a.q_row = ds[base + self.Wmax:].data_ptr()
            a.p_prev, a.p_cur = self.pbuf[1 - cur].data_ptr(), self.pbuf[cur].data_ptr()
            keep_alive = []
            undo = None
            pending_e2 = False
            if isinstance(noise, PhiloxNoise):
                a.rng_mode, a.rng_seed = 1, noise.seed
                for k in range(3):
                    a.rng_off[k], a.rng_span[k] = 0, 256
                if do_sample:
                    a.rng_off[0], a.rng_span[0] = noise.draw(Wv * V)
                if scheme == 0 and Wv > 1:
                    a.rng_off[1], a.rng_span[1] = noise.draw(Wv * V)
                    a.rng_off[2], a.rng_span[2] = noise.peek(V)    # consumed (and counted) only on a rejection
                    pending_e2 = True
            else:
                if do_sample:
                    e1 = noise.multinomial_noise(Wv, V)
                    keep_alive.append(e1)
                    a.noise_e1 = e1.data_ptr()
                if scheme == 0 and Wv > 1:
                    u = noise.accept_noise(Wv, V, d_draft)
                    e2, undo = noise.residual_noise_speculative(V)
                    keep_alive += [u, e2]
                    a.noise_u, a.noise_e2 = u.data_ptr(), e2.data_ptr()
            a.eoi_token, a.text_top_k = int(grammar.eoi_token), int(grammar.text_top_k)
            a.resid, a.next_tokens = self.resid.data_ptr(), self.d_nxt.data_ptr()
            a.sync_ws = self.d_sync.data_ptr()
            stream = self._stream
            if self._zero_copy:
                self._seq = (self._seq % 0x3FFFFFFF) + 1
                a.out_info, a.out_tokens = self.h_out.data_ptr(), self.h_out[4:].data_ptr()
                a.done_flag, a.done_seq = self.h_out[self._flag_idx:].data_ptr(), self._seq
                _lib.check(self.lib.sjd_verify(C.byref(a), C.c_void_p(stream.cuda_stream)), "sjd_verify")
                stats.d2h_bytes += 4 * (5 + Wv)
                flag, fi, seq = self.h_out_np, self._flag_idx, self._seq
                spins = 0
                while flag[fi] != seq:
                    spins += 1
                    if spins > 2_000_000:            # ~ seconds: something is wrong on the device; surface it
                        stream.synchronize()
                        if flag[fi] != seq:
                            raise RuntimeError("sjd_verify finished without publishing its result")
            else:
                a.out_info, a.out_tokens = self.d_out.data_ptr(), self.d_out[4:].data_ptr()
                _lib.check(self.lib.sjd_verify(C.byref(a), C.c_void_p(stream.cuda_stream)), "sjd_verify")
                self.h_out[:4 + Wv].copy_(self.d_out[:4 + Wv], non_blocking=True)
                stats.d2h_bytes += 4 * (4 + Wv)
                stream.synchronize()
            if self._nvtx and not first_trip:
                torch.cuda.nvtx.range_pop()
            res = self.h_out_np[:4 + Wv].tolist()
            matched, rejected = res[0], bool(res[1])
            toks = res[4:4 + Wv]
            if undo is not None and not rejected:
                noise.undo(undo)
            if pending_e2 and rejected:
                noise.draw(V)
            # ---- bookkeeping ----------------------------------------------------------------------------
            if first_trip or out_W <= 1:
                n_cached = W            # every fed token is now a valid cache entry
                new = [toks[-1]]
                carried, carried_row0 = [], 0
            else:
                n_cached = matched
                new = toks[:matched]
                carried = toks[matched:]
                carried_row0 = matched
            next_W = (min(p.max_num_new_tokens, lr[1] - cur_len)
                      if (cur_len >= lr[0] and cur_len < lr[1]) else 1)
            ids += new
            if rejected and hasattr(grammar, "note_residual_call"):
                grammar.note_residual_call(new[:-1])
            grammar.observe(new)
            stats.kv_read_tokens += sum(kv_len + W - lo_ for lo_ in kv_lo)
            kv_len += n_cached
            stats.nfe += 1
            if collect_trace:
                stats.trace.append((W, len(new), int(rejected)))
            out_W = next_W
            cur = 1 - cur
            cur_len = len(ids)
            first_trip = False
            if (ids[-1] in eos) or cur_len >= max_length or (stop_fn is not None and stop_fn(ids)):
                finished = True
        self._stream = None
        torch.cuda.synchronize(dev)
        stats.t_inner = _time.perf_counter() - t0
        stats.new_tokens = len(ids) - len(input_ids)
        return ids
