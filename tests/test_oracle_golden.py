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
    elif case["grammar"] == "emu3":
        e = case["emu3"]
        grammar = O.Emu3Grammar(e["height"], e["width"], e["img_token"], e["eol"], e["eof"], e["eoi"], e["eos"], e["pad"],
                                e["visual"][0], e["visual"][1], top_k=case["image_top_k"])
    elif case["grammar"] == "anole":
        a = case["anole"]
        grammar = O.AnoleGrammar(vocab=V, boi=a["boi"], eoi=a["eoi"], eos=a["eos"], image_lo=a["image"][0],
                                 image_hi=a["image"][1], image_seq_length=a["image_seq_length"],
                                 max_length=case["max_length"], begin_index=len(case["prompt"]), top_k=case["image_top_k"],
                                 mode=a.get("mode", "image-only"))
    else:
        grammar = O.PlainTopK(top_k=case["image_top_k"], top_p=case.get("top_p", 1.0))
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


# ------------------------------------------------------------------------------------------------------------
# forward oracle (oracle/ref_forward.RefStack) vs the reference's own model code (fixtures minted by
# oracle/mint_forward_golden.py from llamagen/llamagen.py and lumina_mgpt/model/chameleon/modeling_chameleon.py, fp32 CPU)
# ------------------------------------------------------------------------------------------------------------
def _load_forward(name):
    from conftest import GOLDEN
    return np.load(GOLDEN / f"forward_{name}.npz")


def test_forward_oracle_matches_reference_llamagen():
    import torch
    from oracle import ref_forward as RF
    z = _load_forward("llamagen")
    L, d, H, Hkv, Dh, ff, V = (int(x) for x in z["cfg"])
    cfg = RF.StackConfig(L, d, H, Hkv, Dh, ff, V, float(z["eps"][0]), rope_interleaved=True, family="llamagen")
    w = RF.random_weights(cfg, seed=int(z["seed"][0]), std=0.05)   # what the minting script copied into the reference
    cos, sin = RF.rope_tables_llamagen_2d(int(z["grid"][0]), Dh, float(z["rope_base"][0]), 1)
    ref = RF.RefStack(cfg, w, cos, sin, rows=2, max_len=40, emulate_bf16=False)
    ids_i = 0
    for ci, (kind, kv_len, W) in enumerate(z["calls"]):
        pos = torch.arange(kv_len, kv_len + W)[None].repeat(2, 1)
        if kind == 0:   # condition-token prefill: the class embedding comes in as `embeds`
            lg = ref.forward(embeds=torch.from_numpy(z["cond_embeds"]), rope_pos=pos, kv_len=int(kv_len), kv_lo=[0, 0],
                             cache_pos=pos)
        else:
            lg = ref.forward(ids=torch.from_numpy(z[f"ids{ids_i}"]), rope_pos=pos, kv_len=int(kv_len), kv_lo=[0, 0],
                             cache_pos=pos)
            ids_i += 1
        want = z[f"logits{ci}"]
        assert lg.shape == want.shape
        assert np.abs(lg.numpy() - want).max() <= 2e-4, f"call {ci}: {np.abs(lg.numpy() - want).max()}"


def test_forward_oracle_matches_reference_chameleon():
    """Lumina-mGPT backbone incl. per-head QK-LayerNorm, rotate-half RoPE, the renewed 3-D-mask-aware
    _update_causal_mask, a KV roll-back and the CFG-uncond row whose prompt prefix is hidden."""
    import torch
    from oracle import ref_forward as RF
    z = _load_forward("chameleon")
    L, d, H, Hkv, Dh, ff, V = (int(x) for x in z["cfg"])
    P = int(z["P"][0])
    cfg = RF.StackConfig(L, d, H, Hkv, Dh, ff, V, float(z["eps"][0]), qk_norm=True, rope_theta=float(z["theta"][0]))
    w = RF.random_weights(cfg, seed=int(z["seed"][0]), std=0.05)   # what the minting script copied into the reference
    r = int(z["qk_norm_rows"][0])
    if r != H:   # the vendored ChameleonLayerNorm shares rows across heads (repeat_interleave, :206-219)
        for Lw in w["layers"]:
            for k in ("q_norm_w", "q_norm_b", "k_norm_w", "k_norm_b"):
                Lw[k] = Lw[k][:r].repeat_interleave(H // r, dim=0)
    cos, sin = RF.rope_tables_rotate_half(Dh, 128, float(z["theta"][0]), False)
    ref = RF.RefStack(cfg, w, cos, sin, rows=2, max_len=64, emulate_bf16=False)
    kv_lo = [0, P - 1]
    for ci, (kv_len, W) in enumerate(z["calls"]):
        pos = torch.arange(kv_len, kv_len + W)[None].repeat(2, 1)
        rope = torch.stack([(pos[b] - kv_lo[b]).clamp(min=0) for b in range(2)])
        lg = ref.forward(ids=torch.from_numpy(z[f"ids{ci}"]), rope_pos=rope, kv_len=int(kv_len), kv_lo=kv_lo, cache_pos=pos)
        want = z[f"logits{ci}"]
        got = lg.numpy()
        if ci == 0:
            # uncond row, prompt prefix positions: every key is hidden for them -> the reference attends uniformly
            # (min_dtype masks), the engine outputs zeros; those rows are never consumed (only the last token's
            # logits are, sampling_logits2tokens :97).  Compare what is consumed.
            got, want = got[:, -1], want[:, -1]
        assert np.abs(got - want).max() <= 2e-4, f"call {ci}: {np.abs(got - want).max()}"


def test_llamagen_flow_oracle_matches_reference():
    """BASELINE config 1 in miniature: the reference's complete test_llamagen.py call sequence (GPT -> renew_llamagen ->
    renew_sampler -> LlamaGenSolver.generate, fp32 on CPU) vs the oracle restatement of that flow driving the forward
    oracle: identical image-token sequence."""
    import json
    from conftest import GOLDEN
    from oracle import llamagen_flow
    g = json.loads((GOLDEN / "llamagen_flow.json").read_text())
    case, ref = g["case"], g["result"]
    tokens, nfe = llamagen_flow.generate(case, ref["ff"], ref["norm_eps"], ref["rope_base"])
    assert tokens == ref["tokens"]
    assert nfe < len(tokens)


def test_llamagen_flow_oracle_matches_reference_at_gptb_size():
    """BASELINE config 1 at its stated size (class-conditional GPT-B: 12 layers, d 768, 16 384 codes; 16 x 16 tokens,
    window 16, cfg 4, top-k 1000): tests/golden/llamagen_flow_gptb.json is the UNMODIFIED reference's test_llamagen.py
    flow on CPU (oracle/mint_llamagen_flow.py gptb, 243 forwards).  The whole flow takes minutes on CPU, so the oracle
    flow is checked on its first 10 forwards here (condition prefill, first token, nine Jacobi windows); the GPU test
    replays the full length."""
    import json
    from conftest import GOLDEN
    from oracle import llamagen_flow
    g = json.loads((GOLDEN / "llamagen_flow_gptb.json").read_text())
    case, ref = g["case"], g["result"]
    ids, nfe = llamagen_flow.generate(case, ref["ff"], ref["norm_eps"], ref["rope_base"], max_trips=9)
    assert nfe == 9 and len(ids) >= 9
    assert ids == ref["tokens"][:len(ids)]


@pytest.mark.parametrize("emulate_bf16", [False, True])
@pytest.mark.parametrize("name", ["cfg3_4x4", "nocfg_6x6"])
def test_forward_included_end_to_end_golden(name, emulate_bf16):
    """tests/golden/e2e_chameleon_greedy_jacobi_*.json (oracle/mint_e2e_golden.py): the UNMODIFIED reference — vendored
    Chameleon forward + renewed mask + JacobiSampler._sample, greedy, 'jacobi', window 8 — run on a tiny decoder, with
    CFG 3 (two rows, hidden prompt prefix) and without.  The loop oracle driving the forward oracle must emit the same
    tokens and, in fp32, the same accepted-count trace; with the bf16 rounding points emulated the tokens must still be
    the same (every decisive argmax leads by several bf16 ulp of the logit scale — that is what lets the GPU test demand
    token equality with the forward included)."""
    import json
    from conftest import GOLDEN
    from oracle import e2e_case
    g = json.loads((GOLDEN / f"e2e_chameleon_greedy_jacobi_{name}.json").read_text())
    trace = []
    ids, nfe = e2e_case.oracle_decode(g["case"], emulate_bf16, trace=trace)
    assert ids == g["result"]["ids"]
    assert g["result"]["min_margin_ulp_fp32"] >= (8.0 if name.startswith("cfg") else 3.5)
    if not emulate_bf16:
        assert [t["n_new"] for t in trace] == [t["n_new"] for t in g["result"]["trace"]]
        assert nfe == len(g["result"]["trace"])
