"""ORACLE helper (test infrastructure): a deterministic, exactly reproducible stand-in for the transformer.

logits[row, i, v] depend only on (token fed at window position i, its cache position, CFG row, v) through
32-bit integer hashing, and every value is a dyadic rational that float32 represents exactly — so the
reference scheduler (driven through this fake model in oracle/mint_golden.py), the numpy oracle and the CUDA
verify kernel all see bit-identical logits on any machine, with nothing but the seed to store.

Tokens in the same residue class mod 4 get similar (not identical) distributions, which makes carried-over
Jacobi drafts face acceptance ratios p/q both below and above 1, exercising accept, reject and resample.
"""
from __future__ import annotations

import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    x = x & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x2C1B3C6D)) & M32
    x ^= x >> np.uint64(12)
    x = (x * np.uint64(0x297A2D39)) & M32
    x ^= x >> np.uint64(15)
    return x


def _unit(tok, pos, row, v, salt):
    x = (np.uint64(tok) * np.uint64(73856093) + np.uint64(pos) * np.uint64(19349663)
         + np.uint64(row) * np.uint64(83492791) + v * np.uint64(40503) + np.uint64(salt))
    h16 = _mix(x) >> np.uint64(16)
    return h16.astype(np.float32) / np.float32(65536.0) - np.float32(0.5)   # exact in fp32


def fake_logits_row(tok: int, pos: int, row: int, V: int, sharp: float = 16.0) -> np.ndarray:
    v = np.arange(V, dtype=np.uint64)
    base = _unit(0, pos, row, v, 5)            # position-only term: shared by every draft at this slot
    coarse = _unit(tok % 4, pos, row, v, 17)
    fine = _unit(tok, pos, row, v, 99)
    return (base * np.float32(sharp) + coarse * np.float32(sharp / 16)
            + fine * np.float32(sharp / 32)).astype(np.float32)


def fake_logits(row_tokens, kv_len: int, n_logit: int, V: int, sharp: float = 16.0) -> np.ndarray:
    """Same contract as oracle.sjd_oracle.decode's logits_fn: [rows * n_logit, V] for the last n_logit window
    positions of each row; position = cache slot of the token."""
    out = []
    for row, toks in enumerate(row_tokens):
        W = len(toks)
        for i in range(W - n_logit, W):
            out.append(fake_logits_row(int(toks[i]), kv_len + i, row, V, sharp))
    return np.stack(out)
