"""Host side of the SJD loop.  (verify_call: thin wrapper over sjd_verify; the decode loop follows below.)"""
from __future__ import annotations

import ctypes as C

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
                sync: bool = True) -> dict:
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
    a.forced = forced.data_ptr() if forced is not None else None
    a.top_k, a.do_sample, a.scheme = int(desc["top_k"]), int(do_sample), int(scheme)
    a.draft = draft.data_ptr()
    a.q_row = q_row.data_ptr() if q_row is not None else None
    a.p_prev = p_prev.data_ptr() if p_prev is not None else None
    a.p_cur = p_cur.data_ptr()
    a.noise_e1 = noise_e1.data_ptr() if noise_e1 is not None else None
    a.noise_u = noise_u.data_ptr() if noise_u is not None else None
    a.noise_e2 = noise_e2.data_ptr() if noise_e2 is not None else None
    a.eoi_token, a.text_top_k = int(eoi_token), int(text_top_k)
    resid = _buf("resid", (V,), torch.float32, dev)
    nxt = torch.empty(W, dtype=torch.int32, device=dev)
    out_tok = torch.empty(W, dtype=torch.int32, device=dev)
    info = torch.empty(4, dtype=torch.int32, device=dev)
    a.resid, a.next_tokens, a.out_tokens, a.out_info = resid.data_ptr(), nxt.data_ptr(), out_tok.data_ptr(), info.data_ptr()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(L.sjd_verify(C.byref(a), C.c_void_p(stream)), "sjd_verify")
    res = {"tokens": out_tok, "next_tokens": nxt, "info": info, "p": p_cur, "_keep": (forced,)}
    if sync:
        h = info.cpu()
        res["matched"], res["rejected"] = int(h[0]), bool(h[1])
    return res
