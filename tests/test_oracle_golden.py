"""CPU: the numpy oracle must reproduce what the UNMODIFIED reference scheduler did (fixtures minted by
oracle/mint_golden.py from /root/reference run on CPU): identical token sequences and per-iteration traces."""
import numpy as np
import pytest

from conftest import load_loop_goldens
from oracle import sjd_oracle as O
from oracle.fake_lm import fake_logits

GOLD = load_loop_goldens()


def run_oracle(case, max_trips=None):
    V = case["V"]

    def logits_fn(rows_tokens, kv_len, n):
        return fake_logits(rows_tokens, kv_len, n, V, case["sharp"])

    j = case["jacobi"]
    params = O.OracleParams(**j)
    if case["grammar"] == "lumina":
        grammar = O.LuminaGrammar(image_top_k=case["image_top_k"], text_top_k=case["text_top_k"])
    else:
        grammar = O.PlainTopK(top_k=case["image_top_k"])
    trace = []
    do_cfg = j["do_cfg"] and j["guidance_scale"] != 1
    ids, nfe = O.decode(logits_fn, case["prompt"], params=params, grammar=grammar,
                        img_vocab=np.arange(*case["img_vocab"]), max_length=case["max_length"],
                        eos_ids=case["eos"], rows=2 if do_cfg else 1, do_sample=case["do_sample"],
                        trace=trace, max_trips=max_trips)
    return ids, nfe, trace


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_reference(name):
    g = GOLD[name]
    ids, nfe, trace = run_oracle(g["case"])
    ref = g["result"]
    assert [t["n_new"] for t in trace] == [t["n_new"] for t in ref["trace"]], "accepted-count trace differs"
    assert [t["W"] for t in trace] == [t["W"] for t in ref["trace"]], "window-size trace differs"
    assert ids == ref["ids"], "token sequence differs from the reference"
    assert nfe == len(ref["trace"])


def test_greedy_jacobi_equals_ar(loop_goldens):
    """SURVEY §4 invariant (i): greedy + 'jacobi' with window W emits exactly the window-1 (AR) tokens."""
    a = loop_goldens["lumina_jacobi_greedy_w8"]["result"]
    b = loop_goldens["lumina_jacobi_greedy_w1"]["result"]
    assert a["ids"] == b["ids"]
    assert len(a["trace"]) < len(b["trace"])


def test_eol_positions_match_reference_formula():
    # logit_processor_3dim.py:25-43, line_len 9: EOL where (tokenlen + j + 1) % 9 == 0
    for tokenlen in range(0, 40):
        for n in (1, 4, 8, 16, 32):
            got = O.eol_positions(tokenlen, n, 9)
            want = [j for j in range(n) if (tokenlen + j + 1) % 9 == 0]
            assert got == want


def test_topk_keeps_ties():
    s = np.array([[1.0, 3.0, 3.0, 2.0, -np.inf, 0.5]], np.float32)
    out = O.topk_filter(s, 2)
    assert np.isfinite(out[0]).tolist() == [False, True, True, False, False, False]
    out = O.topk_filter(s, 3)
    assert np.isfinite(out[0]).tolist() == [False, True, True, True, False, False]
    # k larger than the number of finite entries: threshold is -inf, nothing more is removed
    out = O.topk_filter(s, 6)
    assert np.isfinite(out[0]).tolist() == [True, True, True, True, False, True]
