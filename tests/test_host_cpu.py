"""CPU tests (no GPU): the C-ABI library loads and exports every symbol of include/sjd_b200.h, host-side
logic of the product (grammar state, window bookkeeping helpers, replica sharding over gloo) agrees with the
oracle restatement, and the product never imports the oracle."""
import ctypes
import os
import random
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "accelerating-t2i-ar-with-sjd_b200"


@pytest.fixture(scope="module")
def lib():
    import sjd_b200  # noqa: F401
    from sjd_b200 import _lib
    _lib.build()
    return _lib


def test_library_exports_every_header_symbol(lib):
    names = lib.header_symbols()
    assert len(names) >= 14
    dll = ctypes.CDLL(str(lib.LIB_PATH))
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/sjd_b200.h but not exported"
    assert lib.lib().sjd_version() >= 100


def test_library_is_built_for_sm100a_with_tcgen05_and_tma(lib):
    sass = subprocess.run(["cuobjdump", "-sass", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"


def test_stream_k_workspace_is_host_computable(lib):
    L = lib.lib()
    # 2 KB of counters + 32 B of {arrived, done} per tile + per-tile row statistics + two segments of tpu fp32 partial
    # tiles per CTA (tpu = weight tiles per stream-K unit, 1 unless SJD_GEMM_TPU says otherwise)
    tpu = int(os.environ.get("SJD_GEMM_TPU", "1"))
    assert L.sjd_gemm_workspace_bytes(4096, 4096, 64, 148) == 2048 + 32 * 32 + 32 * 64 * 4 + 2 * tpu * 148 * 64 * 128 * 4
    assert L.sjd_gemm_workspace_bytes(184622, 4096, 128, 148) == 2048 + 1443 * 32 + 1443 * 128 * 4 + 2 * tpu * 148 * 128 * 128 * 4


def test_no_compute_without_gpu_fails_loudly(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sjd_b200.model import DeviceStack, StackShape
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceStack(StackShape(1, 64, 1, 1, 64, 64, 128), {}, torch.zeros(4, 32), torch.zeros(4, 32), 1, 16)


def test_product_never_imports_oracle():
    for p in list(PKG.rglob("*.py")) + [ROOT / "sjd_b200.py"] + list((ROOT / "scheduler").glob("*.py")) + \
            list((ROOT / "llamagen").glob("*.py")) + list((ROOT / "lumina_mgpt").glob("*.py")):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{p} imports the oracle"


def test_grammar_state_matches_oracle_grammar(lib):
    """Product-side incremental grammar == oracle's from-scratch restatement on random token streams."""
    from oracle import sjd_oracle as O
    from sjd_b200.engine import LuminaGrammarState
    rnd = random.Random(0)
    for trial in range(30):
        h, w = rnd.randint(1, 4), rnd.randint(1, 5)
        ids = [rnd.randint(8900, 9000) for _ in range(rnd.randint(1, 5))] + [8197, 8804 + h, 8804 + w]
        og, pg = O.LuminaGrammar(image_top_k=77, text_top_k=5), LuminaGrammarState(image_top_k=77, text_top_k=5)
        pg.reset()
        pg.observe(ids)
        total = (2 * w + 1) * 2 * h + 1
        produced = 0
        while produced < total + 6:
            n = rnd.randint(1, 9)
            d_o, d_p = og.describe(ids, n), pg.describe(n)
            assert d_o["forced"] == d_p["forced"], (trial, ids)
            assert (d_o["allow"] is None) == (d_p["allow"] is None)
            assert d_o["top_k"] == d_p["top_k"]
            assert d_o["no_cfg"] == pg.no_cfg
            m = rnd.randint(1, n)   # accept m tokens, honouring forced ones
            new = [d_o["forced"][j] if d_o["forced"][j] >= 0 else rnd.randint(4, 8195) for j in range(m)]
            ids = ids + new
            pg.observe(new)
            produced += m


def test_emu3_grammar_state_matches_oracle_grammar(lib):
    """Emu3 grammar: incremental product-side state == oracle restatement (itself pinned to the reference's
    EOLLogitProcessor3d by tests/golden/sjd_loop_emu3_*.json), through EOL / EOF / EOI / EOS and into the PAD tail."""
    from oracle import sjd_oracle as O
    from sjd_b200.engine import Emu3GrammarState
    rnd = random.Random(1)
    for trial in range(30):
        h, w = rnd.randint(1, 4), rnd.randint(1, 6)
        args = (h, w, 900, 901, 902, 903, 904, 905, 1000, 3048)
        og, pg = O.Emu3Grammar(*args, top_k=64), Emu3GrammarState(*args, top_k=64)
        ids = [rnd.randint(1, 800) for _ in range(rnd.randint(0, 4))] + [900]
        pg.reset()
        pg.observe(ids)
        produced, total = 0, (w + 1) * h + 3
        while produced < total + 12:
            n = rnd.randint(1, 11)
            d_o, d_p = og.describe(ids, n), pg.describe(n)
            assert d_o["forced"] == d_p["forced"], (trial, ids, n)
            assert tuple(d_o["allow"]) == tuple(d_p["allow"]) and d_o["top_k"] == d_p["top_k"]
            m = rnd.randint(1, n)
            new = [d_o["forced"][j] if d_o["forced"][j] >= 0 else rnd.randint(1000, 3047) for j in range(m)]
            ids = ids + new
            pg.observe(new)
            produced += m


def test_prompt_sharding_covers_all_prompts(lib):
    from sjd_b200.replicas import shard_prompts
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in shard_prompts(8, r, world))
        assert got == list(range(8))
        assert all(len(shard_prompts(8, r, world)) == 8 // world for r in range(world))


_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, sjd_b200
from sjd_b200 import replicas
rank, world, local = replicas.init_process_group("gloo")
mine = replicas.shard_prompts(8, rank, world)
tok, nfe = sum(100 + i for i in mine), sum(10 + i for i in mine)
replicas.barrier()
g = replicas.gather_counters(tok, nfe, 1)
t = replicas.max_over_ranks(1.0 + rank)
if rank == 0:
    print(json.dumps(dict(world=world, total_tokens=int(g[:, 0].sum()), total_nfe=int(g[:, 1].sum()),
                          done=int(g[:, 2].sum()), tmax=t)))
"""


def test_two_rank_gloo_counter_gather(lib, tmp_path):
    """world_size-2 CPU run of the only exchange step on the path (accepted-token counters)."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out == dict(world=2, total_tokens=sum(100 + i for i in range(8)), total_nfe=sum(10 + i for i in range(8)),
                       done=2, tmax=2.0)


# ------------------------------------------------------------------------------------------- Anole grammar (a12)
def _anole_pair(S=12, P=5, extra=2, top_k=50, V=9216):
    from oracle import sjd_oracle as O
    from sjd_b200 import engine
    max_length = P + S + extra
    o = O.AnoleGrammar(vocab=V, boi=8197, eoi=8196, eos=2, image_lo=4, image_hi=8196, image_seq_length=S,
                       max_length=max_length, begin_index=P, top_k=top_k)
    e = engine.AnoleGrammarState(8197, 8196, 2, 4, 8196, S, max_length, P, top_k=top_k)
    return o, e


def _allowed_from_desc(d, V):
    """allowed-id set of window position 0 as the verify kernel will see it (sjd_verify_args: allow range, allow_mode, ban)"""
    if d["forced"][0] >= 0:
        return {d["forced"][0]}
    mode, ban = d.get("allow_mode", 0), [b for b in d.get("ban", (-1, -1))]
    if mode == 3:
        return set(ban)
    lo, hi = d["allow"] if d["allow"] else (0, V)
    if mode == 2:
        return set(range(V)) - set(range(lo, hi)) - set(ban)
    return set(range(lo, hi)) - (set(ban) if mode in (1, 4) else set())


def test_anole_grammar_state_matches_oracle_masks():
    """engine.AnoleGrammarState (what the kernel is told) == the five reference processors restated as masks in the
    oracle, for every prefix length of an image-only generation — including prefixes a multi-token accept produced by
    jumping over the end-of-image offset."""
    rng = random.Random(0)
    V, S, P = 9216, 12, 5
    o, e = _anole_pair(S=S, P=P)
    for overshoot in (False, True):
        ids = [0, 300, 400, 500, 600]
        e.reset()
        e.observe(ids)
        while len(ids) < P + S + 2:
            d = e.describe(4)
            allowed = set(np.flatnonzero(~o.disallowed(ids)).tolist())
            assert _allowed_from_desc(d, V) == allowed, (len(ids), d["forced"][:1], sorted(allowed)[:4])
            assert d["forced"] == [d["forced"][0]] * 4, "one decision for every window position"
            # (the trip that forces begin-of-image is the window-1 trip after the prompt, jacobi_loop_interval_l >= 1)
            resid = e.describe_residual(1 if allowed == {8197} else 3)
            for j, f in enumerate(resid):   # residual at reject position j: prefix + j accepted drafts
                acc = [7] * j if len(allowed) > 1 else [next(iter(allowed))] * j
                want = set(np.flatnonzero(~o.disallowed(ids + acc)).tolist())
                assert (want == {-2 - f}) if f <= -2 else (f == -1 and want == set(range(4, 8196))), (len(ids), j, f)
            if len(allowed) == 1:
                new = [next(iter(allowed))]
            else:
                n_new = rng.randint(1, 4) if overshoot else 1
                new = [rng.randrange(4, 8196) for _ in range(n_new)]
            ids += new
            e.observe(new)
        assert ids[P] == 8197
        if not overshoot:
            assert ids[P + S + 1] == 8196, "with single-token accepts end-of-image is forced at the offset"


def test_anole_two_id_sets_are_expressible_now():
    """Round 1 refused the {eos, boi} set that appears when max_new_tokens leaves room for a second image; the verify
    kernel now takes a one-or-two-id candidate set (allow_mode 3)."""
    o, e = _anole_pair(S=4, P=2, extra=12)
    ids = [0, 1, 8197, 10, 11, 12, 13, 8196]
    e.observe(ids)
    d = e.describe(1)
    assert d["allow_mode"] == 3 and _allowed_from_desc(d, 9216) == set(np.flatnonzero(~o.disallowed(ids)).tolist()) == {2, 8197}


@pytest.mark.parametrize("mode", ["text-only", "interleaved-text-image"])
def test_anole_other_modes_match_oracle_masks(mode):
    """a12 (round 2): the candidate set engine.AnoleGrammarState(mode) hands to sjd_verify == the reference's processors
    for that multimodal_generation_mode restated as masks (oracle.AnoleGrammar(mode), itself pinned to the reference by
    the goldens anole_text_only_w8 / anole_interleaved_w6), for every prefix of a run that crosses text -> begin-of-image
    -> S image tokens -> forced end-of-image -> text -> the index from which begin-of-image is suppressed."""
    from oracle import sjd_oracle as O
    from sjd_b200 import engine
    V, S, P = 9216, 6, 3
    max_length = P + 30
    o = O.AnoleGrammar(vocab=V, boi=8197, eoi=8196, eos=2, image_lo=4, image_hi=8196, image_seq_length=S,
                       max_length=max_length, begin_index=P, top_k=50, mode=mode)
    e = engine.AnoleGrammarState(8197, 8196, 2, 4, 8196, S, max_length, P, top_k=50, mode=mode)
    ids = [0, 300, 400]
    e.observe(ids)
    script = [9000, 9001, 8197] + [10, 11, 12, 13, 14, 15] + [8196] + [9100 + i for i in range(16)]
    for tok in script:
        d = e.describe(3)
        want = set(np.flatnonzero(~o.disallowed(ids)).tolist())
        assert _allowed_from_desc(d, V) == want, (mode, len(ids), d["allow_mode"], d["ban"])
        if mode == "interleaved-text-image" or tok not in (8197, 8196) and not (4 <= tok < 8196):
            assert tok in want or mode == "text-only", (mode, len(ids), tok)
        ids.append(tok)
        e.observe([tok])
    # residual decisions inside a text window: same set unless a draft is begin-of-image (then the kernel cannot express it)
    e2 = engine.AnoleGrammarState(8197, 8196, 2, 4, 8196, S, max_length, P, top_k=50, mode=mode)
    e2.observe([0, 300, 400])
    assert e2.describe_residual(3, [400, 9000, 9001]) == [-1, -1, -1] and e2.resid_desc is None
    if mode == "interleaved-text-image":
        # an accepted begin-of-image draft switches the residual of the positions behind it to image ids: handed to the
        # kernel as its second candidate set (sjd_verify_args.resid_*, from position 1 on)
        assert e2.describe_residual(3, [400, 8197, 9001]) == [-1, -1, -1]
        assert e2.resid_desc["allow"] == (4, 8196) and e2.resid_desc["allow_mode"] == 1 and e2.resid_desc["resid_from"] == 1
        want = set(np.flatnonzero(~o.disallowed([0, 300, 400, 8197, 9001][:4] + [])).tolist())
        assert want == set(range(4, 8196))


def test_anole_processors_translate_to_grammar_state():
    from sjd_b200 import engine, hf_api as H
    from transformers.generation.logits_process import TopKLogitsWarper
    V, S, P = 9216, 24, 4
    image = list(range(4, 8196))
    allowed = set(image) | {2, 8197, 8196}
    procs = [H.AllowOnlyTokensAtRelativeOffsetLogitsProcessor3d(8197, [8196], offset=S + 1, exclusive=True),
             H.AllowOnlyTokensInRelativeWindowLogitsProcessor3d(8197, image, window_width=S, exclusive=True),
             H.SuppressTokensInIndexRangeLogitsProcessor3d([8197], start_index=P + S + 2 - S - 1),
             H.SuppressTokensLogitsProcessor3d([t for t in range(V) if t not in allowed]),
             H.SuppressTokensAtBeginLogitsProcessor3d([2], begin_index=P), TopKLogitsWarper(top_k=50)]
    g = H.grammar_from_processors(procs, vocab=V)
    assert isinstance(g, engine.AnoleGrammarState)
    assert (g.boi, g.eoi, g.eos, g.allow, g.S, g.max_length, g.begin_index, g.top_k) == \
        (8197, 8196, 2, (4, 8196), S, P + S + 2, P, 50)
    with pytest.raises(RuntimeError, match="no host fallback"):
        procs[0](torch.zeros(1, 3, dtype=torch.long), torch.zeros(1, 2, V))
    with pytest.raises(NotImplementedError):
        H.grammar_from_processors(procs[:2] + procs[3:], vocab=V)      # one processor of the set missing
    # the other modes of renew_pipeline_anole.generate (a12, round 2)
    gi = H.grammar_from_processors(procs[:3] + procs[5:], vocab=V)     # interleaved-text-image: the first three + TopK
    assert gi.mode == "interleaved-text-image" and (gi.boi, gi.eoi, gi.allow, gi.S) == (8197, 8196, (4, 8196), S)
    gt = H.grammar_from_processors([H.SuppressTokensLogitsProcessor3d(image + [8197, 8196]), procs[5]], vocab=V)
    assert gt.mode == "text-only" and gt.describe(2)["allow_mode"] == 2
    assert _allowed_from_desc(gt.describe(2), V) == set(range(V)) - set(image) - {8197, 8196}   # ([4, 8198) is one run here)
    procs[1] = H.AllowOnlyTokensInRelativeWindowLogitsProcessor3d(8197, image, window_width=S, exclusive=False)
    with pytest.raises(NotImplementedError):
        H.grammar_from_processors(procs, vocab=V)


def test_top_p_processor_translates_and_validates():
    from sjd_b200 import engine, hf_api as H
    from transformers.generation.logits_process import TopKLogitsWarper
    g = H.grammar_from_processors([TopKLogitsWarper(top_k=100), H.TopPLogitsWarper3d(top_p=0.8)])
    assert isinstance(g, engine.PlainTopKState) and g.top_k == 100 and g.top_p == 0.8
    assert g.describe(3)["top_p"] == 0.8
    assert engine.top_p_threshold(1.0) == 0.0
    assert engine.top_p_threshold(0.9) == float(np.float32(1.0 - 0.9))
    with pytest.raises(ValueError):
        H.TopPLogitsWarper3d(top_p=1.5)
    with pytest.raises(NotImplementedError):
        H.grammar_from_processors([H.TopPLogitsWarper3d(top_p=0.8), TopKLogitsWarper(top_k=100)])


def test_top_p_oracle_follows_reference_formula():
    """oracle.topp_filter against the reference arithmetic written out with torch (sort ascending, softmax, cumsum,
    `<= 1 - top_p`, keep the last) — scheduler/logit_processor_3dim.py:406-419."""
    from oracle import sjd_oracle as O
    g = torch.Generator().manual_seed(3)
    for top_p in (0.3, 0.8, 0.95):
        s = torch.randn(4, 777, generator=g) * 3
        s[0, 5:200] = -float("inf")
        srt, idx = torch.sort(s, descending=False)
        rm = srt.softmax(-1).cumsum(-1) <= (1 - top_p)
        rm[..., -1:] = 0
        want = s.masked_fill(rm.scatter(-1, idx, rm), -float("inf")).numpy()
        got = O.topp_filter(s.numpy(), top_p)
        assert np.array_equal(np.isinf(got), np.isinf(want))
    assert O.topp_filter(s.numpy(), 1.0) is not None


def test_anole_adaptor_generate_reaches_sample_with_the_image_only_grammar():
    """scheduler.jacobi_iteration_anhole.renew_pipeline_sampler on a tiny HF ChameleonForConditionalGeneration: HF's own
    generate() plumbing (this image: transformers 5.x) must end in the renewed `_sample` with the five 3-D processors
    (+ HF's TopKLogitsWarper), which translate to the Anole grammar state.  `_sample` itself needs the GPU."""
    from transformers import ChameleonConfig, ChameleonForConditionalGeneration
    from scheduler.jacobi_iteration_anhole import renew_pipeline_sampler
    from sjd_b200 import engine, hf_api as H
    names = {f"IMGIMG{chr(65 + i // 10)}{chr(65 + i % 10)}Z": 4 + i for i in range(60)}
    cfg = ChameleonConfig(vocab_size=128, hidden_size=128, intermediate_size=128, num_hidden_layers=1,
                          num_attention_heads=1, num_key_value_heads=1, max_position_embeddings=64,
                          vocabulary_map={"<image>": 3, **names}, eos_token_id=2, bos_token_id=0, pad_token_id=1,
                          vq_config={"embed_dim": 8, "num_embeddings": 16, "resolution": 32, "channel_multiplier": [1, 1],
                                     "base_channels": 32, "num_res_blocks": 1, "latent_channels": 8})
    m = ChameleonForConditionalGeneration(cfg).eval()
    m.model.vocabulary_mapping.boi_token_id = 70
    m.model.vocabulary_mapping.eoi_token_id = 71

    class Proc:
        image_seq_length = 9

    m = renew_pipeline_sampler(m, Proc(), jacobi_loop_interval_l=1, jacobi_loop_interval_r=20, max_num_new_tokens=4,
                               guidance_scale=3.0, seed=0, multi_token_init_scheme="random", do_cfg=True,
                               image_top_k=50, text_top_k=10, prefix_token_sampler_scheme="speculative_jacobi")
    assert m.model.image_seq_length == 9 and m.vocabulary_mapping.image_token_ids == list(range(4, 64))
    seen = {}

    def fake_sample(self, *args, **kw):
        import inspect
        b = inspect.signature(orig).bind(self, *args, **kw)   # HF must be able to call the REAL _sample like this
        b.apply_defaults()
        input_ids, logits_processor = b.arguments["input_ids"], b.arguments["logits_processor"]
        stopping_criteria, generation_config = b.arguments["stopping_criteria"], b.arguments["generation_config"]
        seen["procs"], seen["gc"], seen["crit"] = logits_processor, generation_config, stopping_criteria
        return input_ids

    cls = type(m)
    orig = cls._sample
    cls._sample = fake_sample
    try:
        ids = torch.tensor([[0, 80, 81, 82]])
        m.generate(input_ids=ids, attention_mask=torch.ones_like(ids), multimodal_generation_mode="image-only",
                   do_sample=True, top_k=7)
    finally:
        cls._sample = orig
    g = H.grammar_from_processors(list(seen["procs"]), vocab=cfg.vocab_size)
    assert isinstance(g, engine.AnoleGrammarState)
    assert (g.boi, g.eoi, g.eos, g.allow, g.S, g.max_length, g.begin_index, g.top_k) == (70, 71, 2, (4, 64), 9, 4 + 11, 4, 7)
    assert int(seen["gc"].max_length) == 4 + 11


# ------------------------------------------------------------------------ round 2: draft initialisation schemes (f2)
def _ref_root():
    for p in (ROOT / "baseline" / "_ref", Path("/root/reference")):
        if (p / "scheduler" / "jacobi_iteration_lumina_mgpt.py").exists():
            return p
    return None


def test_multi_token_init_scheme_is_validated_not_ignored(lib):
    """every non-'random' scheme used to mean 'random' silently.  Now: 'repeat_horizon' is implemented (engine ==
    oracle restatement of jacobi_iteration_lumina_mgpt.py:516-594), anything the reference asserts on raises ValueError, and
    'sample_horizon' (an IndexError upstream) is refused."""
    from sjd_b200 import engine
    from oracle import sjd_oracle as O
    engine.check_init_scheme("random")
    engine.check_init_scheme("repeat_horizon")
    for bad in ("vertical", "repeat_vertical", "horizon"):
        with pytest.raises(ValueError):
            engine.check_init_scheme(bad)
    with pytest.raises(NotImplementedError):
        engine.check_init_scheme("sample_horizon")
    # width 8 (+1 for the end-of-line slot), prompt of 6 tokens -> origin = (6 - 1) + 3 = 8
    ids = list(range(100, 120))            # 20 accepted tokens
    carried = [7001, 7002]
    fresh = [1, 2, 3, 4, 5, 6]
    out = engine.horizon_init(fresh, "repeat_horizon", ids, carried, 8, 5)
    # absolute indices 22..27 -> columns (a - 8) % 9 = 5, 6, 7, 8, 0, 1: copies of the last known token except at column 0
    assert out == [7002, 7002, 7002, 7002, 5, 7002]
    assert engine.horizon_init(fresh, "random", ids, carried, 8, 5) == fresh
    assert engine.horizon_init(fresh, "repeat_horizon", ids, carried, None, 5) == fresh      # no width (LlamaGen, Emu3)
    rng = random.Random(0)
    for _ in range(200):
        n_ids, n_c, n_f = rng.randint(1, 60), rng.randint(0, 9), rng.randint(0, 9)
        a = [rng.randint(4, 8195) for _ in range(n_ids)]
        c = [rng.randint(4, 8195) for _ in range(n_c)]
        f = [rng.randint(4, 8195) for _ in range(n_f)]
        w, pre = rng.choice([None, 4, 6, 48]), rng.randint(0, 12)
        assert engine.horizon_init(f, "repeat_horizon", a, c, w, pre) == O.horizon_init(f, "repeat_horizon", a, c, w, pre)


def test_reference_itself_crashes_on_lumina_horizon_schemes():
    """Why 'repeat_horizon' + the Lumina grammar has no reference-minted golden: the unmodified reference raises
    IndexError at jacobi_iteration_lumina_mgpt.py:577 as soon as a window inside the image needs fresh drafts."""
    ref = _ref_root()
    if ref is None:
        pytest.skip("no reference checkout / baseline/_ref here")
    code = ("import os, sys; sys.dont_write_bytecode = True; sys.path.insert(0, %r); os.environ['SJD_REFERENCE'] = %r\n"
            "from oracle import mint_golden as M\nM.REF = __import__('pathlib').Path(%r)\nC = M.apply_shims()\n"
            "case = M.UNPINNED_CASES['lumina_spec_w8_repeat_horizon']\n"
            "try:\n    M.run_reference_loop(case, C); print('RAN')\nexcept IndexError as e:\n    print('INDEXERROR', e)\n"
            % (str(ROOT), str(ref), str(ref)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "INDEXERROR" in r.stdout, (r.stdout[-500:], r.stderr[-800:])


def test_oracle_repeat_horizon_changes_the_drafts_only_inside_the_image():
    """Oracle loop with the spatial initialisation on the two unpinned cases: it must differ from 'random' (the scheme
    is live), keep the grammar intact, and leave the greedy-Jacobi fixed point untouched."""
    from oracle import mint_golden as M
    from test_oracle_golden import run_oracle
    differs = 0
    for name, case in M.UNPINNED_CASES.items():
        ids_h, nfe_h, tr_h = run_oracle(case)
        rnd = dict(case, jacobi=dict(case["jacobi"], multi_token_init_scheme="random"))
        ids_r, nfe_r, tr_r = run_oracle(rnd)
        differs += int([t["n_new"] for t in tr_h] != [t["n_new"] for t in tr_r] or ids_h != ids_r)
        P = len(case["prompt"])
        h, w = (case["prompt"][-2] - 8804) * 2, (case["prompt"][-1] - 8804) * 2
        img = ids_h[P:]
        assert all(img[i] == 8803 for i in range(w, min(len(img), h * (w + 1)), w + 1)), name
    assert differs >= 1, "the spatial initialisation never changed a run: the scheme is not live"


# ------------------------------------------------------------------------ round 2: the drop-in boundary, for real
_IMPORT_BLOCKS = {
    # the import lines of the three reference demo scripts that touch modules this repository shadows or must not shadow
    "test_llamagen.py:16-20": ["from llamagen.tokenizer.tokenizer_image.vq_model import VQ_models",
                               "from llamagen.language.t5 import T5Embedder",
                               "from llamagen.llamagen import GPT_models",
                               "from llamagen.llamagen_solver import LlamaGenSolver, renew_llamagen, generate",
                               "from scheduler.jacobi_iteration_lumina_mgpt import renew_sampler"],
    "test_lumina_mgpt.py:10,101": ["from lumina_mgpt.inference_solver import FlexARInferenceSolver",
                                   "from scheduler.jacobi_iteration_lumina_mgpt import renew_pipeline_sampler"],
    "test_emu3.py:16,145": ["from emu3.mllm.processing_emu3 import Emu3Processor",
                            "from scheduler.jacobi_iteration_emu3 import renew_solver"],
}


def test_reference_demo_scripts_import_blocks_resolve_with_this_repo_first():
    """README's drop-in claim: with this repository AHEAD of a reference checkout on sys.path (the reference's scripts
    append ./ and ./lumina_mgpt/ themselves, test_lumina_mgpt.py:4-5), every import of test_llamagen.py / test_lumina_mgpt.py
    / test_emu3.py resolves — scheduler.* / llamagen.llamagen / llamagen.llamagen_solver to THIS repository, the VQ
    decoder, T5 embedder, FlexARInferenceSolver and Emu3Processor to the reference's own files (round 1 shadowed
    llamagen.tokenizer / llamagen.language with a regular package).  HF-5.5 name shims of SURVEY App. C applied first;
    a missing third-party dependency of the reference (ftfy for its T5 text cleaner) is reported, not a failure."""
    ref = _ref_root()
    if ref is None:
        pytest.skip("no reference checkout / baseline/_ref here")
    lines = [l for block in _IMPORT_BLOCKS.values() for l in block]
    code = ("import sys, json; sys.dont_write_bytecode = True\n"
            "sys.path[:0] = [%r, %r, %r]\n"
            "from oracle.mint_golden import apply_shims; apply_shims()\n"
            "out = {}\n"
            "for l in %r:\n"
            "    try:\n"
            "        exec(l, {}); out[l] = sys.modules[l.split()[1]].__file__\n"
            "    except ModuleNotFoundError as e:\n"
            "        out[l] = 'MISSING:' + str(e.name)\n"
            "    except Exception as e:\n"
            "        out[l] = 'ERROR:' + repr(e)[:200]\n"
            "print('RESULT' + json.dumps(out))\n" % (str(ROOT), str(ref), str(ref / "lumina_mgpt"), lines))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    import json
    res = json.loads(r.stdout.split("RESULT", 1)[1])
    ours = ("scheduler.", "llamagen.llamagen ", "llamagen.llamagen_solver ")
    for line, where in res.items():
        if where.startswith("MISSING:"):
            # only third-party packages the reference itself needs may be missing, never a module of either repository
            assert where.split(":", 1)[1].split(".")[0] not in ("llamagen", "scheduler", "lumina_mgpt", "emu3", "model",
                                                                  "xllmx", "data"), (line, where)
            continue
        assert not where.startswith("ERROR:"), (line, where)
        mine = any(line.split()[1].startswith(o.strip()) and (o.endswith(".") or line.split()[1] == o.strip()) for o in ours)
        if mine:
            assert where.startswith(str(ROOT)) and "baseline" not in where, (line, where)
        else:
            assert where.startswith(str(ref)), (line, where)


def test_llamagen_renew_calls_of_the_demo_script_on_a_stub_checkpoint():
    """test_llamagen.py:72-88 on CPU: GPT_models[...]() -> renew_llamagen -> renew_sampler -> _init_new_params(**dict with the
    script's extra keys) -> load_state_dict of a stub checkpoint -> LlamaGenSolver(...).  (The forward needs the GPU.)"""
    from llamagen.llamagen import GPT_models
    from llamagen.llamagen_solver import LlamaGenSolver, renew_llamagen
    from scheduler.jacobi_iteration_lumina_mgpt import renew_sampler
    gpt = GPT_models["GPT-B"](block_size=16 ** 2, cls_token_num=1, model_type="c2i", num_classes=10, vocab_size=512)
    jd = dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=(256 // 16) ** 2 - 16 - 2, max_num_new_tokens=16,
              guidance_scale=7.5, seed=None, multi_token_init_scheme="repeat_horizon", do_cfg=True, image_top_k=1000,
              text_top_k=10, prefix_token_sampler_scheme="speculative_jacobi")
    gpt.__class__ = renew_llamagen(gpt.__class__)
    gpt._init_new_params(**jd)
    gpt.__class__ = renew_sampler(gpt.__class__)
    gpt._init_new_params(**jd)
    sd = {k: v.clone() for k, v in gpt.state_dict().items()}
    missing = gpt.load_state_dict(sd, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys
    solver = LlamaGenSolver(model=gpt, image_top_k=jd["image_top_k"], image_top_p=1.0)
    assert [type(p).__name__ for p in solver.create_logits_processor()] == ["TopKLogitsWarper", "TopPLogitsWarper3d"]
    assert gpt.multi_token_init_scheme == "repeat_horizon" and gpt.max_num_new_tokens == 16
    with pytest.raises(RuntimeError, match="CUDA"):
        solver.generate(torch.tensor([3]), 256, None, cfg_scale=7.5, temperature=1.0, top_k=1000, top_p=1.0)


def test_hf_warpers_are_consumed_by_the_grammar_translation():
    """HF generate() appends TemperatureLogitsWarper / TopPLogitsWarper / TopKLogitsWarper for non-default
    sampling settings; they must reach sjd_verify instead of raising, and the kernel temperature must come from the list."""
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper
    from sjd_b200 import engine, hf_api
    g = hf_api.grammar_from_processors([TopKLogitsWarper(50), TemperatureLogitsWarper(0.7), TopPLogitsWarper(0.9)])
    assert isinstance(g, engine.PlainTopKState) and g.top_k == 50 and abs(g.top_p - 0.9) < 1e-12
    assert abs(g.temperature - 0.7) < 1e-12
    g = hf_api.grammar_from_processors([hf_api.MultiTokensVLLogitsProcessor(8197, 8196, 8803, 32, 65536),
                                        hf_api.MultiTokensInterleavedTopKLogitsWarper(2000, 10, 8197, 8196)])
    assert isinstance(g, engine.LuminaGrammarState) and g.temperature == 1.0


# ------------------------------------------------------------------------ round 2: loader + multi-GPU prompt runner (f1)
_LAUNCH_WORKER = '''
import sys, json
sys.path.insert(0, {root!r})
import torch
import sjd_b200
from sjd_b200 import launcher

def loader(model_name, device=None, seed=None, **kw):
    class M:
        class engine:
            class stats:
                new_tokens, nfe = 0, 0
    def fwd(prompt):
        M.engine.stats.new_tokens, M.engine.stats.nfe = len(prompt), 3
        return torch.tensor([len(prompt)])
    return M, fwd

prompts = ["p" * (i + 1) for i in range(7)]
res = launcher.run_prompts("stub", prompts, output_dir={out!r}, seed=0, loader=loader, backend="gloo")
res2 = launcher.run_prompts("stub", prompts, output_dir={out!r}, seed=0, loader=loader, backend="gloo")   # resume: all skipped
import torch.distributed as dist
if dist.get_rank() == 0:
    print(json.dumps(dict(done=res[:, 0].tolist(), tok=res[:, 1].tolist(), nfe=res[:, 2].tolist(), again=res2[:, 0].tolist())))
'''


def test_two_rank_gloo_prompt_runner(lib, tmp_path):
    """sjd_b200.launcher over gloo, world size 2: prompt i runs on rank i mod 2 (multi_gpu_dataframe_split.py:31-63), tensor
    results land in <idx>.pt (multi_gpu_infer_with_prompt.py:58-61), counters are all-gathered, a second run skips
    every finished prompt (:56-57)."""
    import json
    out = tmp_path / "work"
    script = tmp_path / "l.py"
    script.write_text(_LAUNCH_WORKER.format(root=str(ROOT), out=str(out)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29579", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["done"] == [4, 3] and r["again"] == [0, 0]
    assert r["tok"] == [1 + 3 + 5 + 7, 2 + 4 + 6] and r["nfe"] == [12, 9]
    assert sorted(p.name for p in out.iterdir()) == [f"{i}.pt" for i in range(7)]
    assert int(torch.load(out / "4.pt")[0]) == 5


def test_model_loader_dispatch_and_names():
    """model_wrappers.model_loader exposes the reference's entry points (model_loader.py:347-360, :564-574) and dispatches
    on the model name like it; unknown names raise NotImplementedError like upstream."""
    from model_wrappers import model_loader as ML
    assert str(ROOT) in ML.__file__
    for fn in ("load_pretrained_model", "get_forward_func", "load_lumina_mgpt", "load_anole", "load_emu3", "load_llamagen",
               "get_lumina_mgpt_forward_func", "get_anole_forward_func", "get_emu3_forward_func", "get_llamagen_forward_func"):
        assert callable(getattr(ML, fn)), fn
    with pytest.raises(NotImplementedError):
        ML.load_pretrained_model("some/other-model")
    with pytest.raises(NotImplementedError):
        ML.get_forward_func("some/other-model", None)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ML.load_pretrained_model("synthetic/llamagen-gpt-b", device="cpu", n_layers=1)


# ------------------------------------------------------------------------ f3: VQ decoders (token ids -> pixels)
@pytest.mark.parametrize("family", ["llamagen", "chameleon"])
def test_vq_decoder_plan_and_conv_stack_match_reference_golden_on_cpu(family):
    """The decoder's structure is read from the state dict's keys (two naming schemes); its convolution stack (library
    calls) on the latents the reference formula gives must reproduce the pixels the UNMODIFIED reference module produced
    (tests/golden/vq_decode_*.json, oracle/mint_vq_golden.py).  The token-side kernel (sjd_vq_lookup) is checked on the
    GPU; constructing the decoder without a GPU must fail loudly."""
    import json
    import torch
    import torch.nn.functional as F
    from sjd_b200 import vq_decode
    from oracle.vq_case import fill_state
    g = json.loads((ROOT / "tests" / "golden" / f"vq_decode_{family}.json").read_text())
    sd = fill_state(g["shapes"], g["seed"])
    with pytest.raises(RuntimeError):
        vq_decode.VQDecoder(sd, "cpu")
    dec = vq_decode.VQDecoder.__new__(vq_decode.VQDecoder)
    dec.sd = {k: v.float() for k, v in sd.items()}
    dec.layout = "llamagen" if any(k.startswith("decoder.conv_blocks.") for k in sd) else "chameleon"
    assert dec.layout == family
    dec.program = dec._plan()
    ops = [op for op, _ in dec.program]
    assert ops.count("up") == 2 and ops.count("attn") >= 1 and ops[:3] == ["res", "attn", "res"]
    cb = sd["quantize.embedding.weight"]
    if g["l2_norm"]:
        cb = F.normalize(cb, p=2, dim=-1)
    codes = torch.tensor(g["codes"])
    B, h, w = g["batch"], g["h"], g["w"]
    zq = cb[codes].reshape(B, h, w, -1).permute(0, 3, 1, 2).contiguous()
    lat = F.conv2d(zq, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    with torch.no_grad():
        px = dec.decode_latents(lat)
    ref = torch.tensor(g["pixels"]).reshape(g["out_shape"])
    assert list(px.shape) == g["out_shape"]
    assert (px - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def test_emu3_vq_decoder_stack_matches_reference_golden_on_cpu():
    """Emu3's vision-tokenizer decoder for images: temporal stack (inference BatchNorm3d, causal 3-D convolutions, x2 up-sampling
    in time) + latent-conditioned 2-D decoder, driven from the latents the reference formula gives, must reproduce the pixels
    of the unmodified Emu3VisionVQModel.decode (tests/golden/vq_decode_emu3.json).  For one frame the causal (3, 1, 1)
    post_quant_conv reduces to its last temporal tap — asserted here against F.conv3d with the reference's padding."""
    import json
    import torch
    import torch.nn.functional as F
    from sjd_b200 import vq_decode
    from oracle.vq_case import fill_state
    g = json.loads((ROOT / "tests" / "golden" / "vq_decode_emu3.json").read_text())
    sd = fill_state(g["shapes"], g["seed"])
    with pytest.raises(RuntimeError):
        vq_decode.Emu3VQDecoder(sd, "cpu")
    dec = vq_decode.Emu3VQDecoder.__new__(vq_decode.Emu3VQDecoder)
    dec.sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    dec.program = dec._plan()
    assert dec.n_time_res == 1 and dec.n_time_up == 2
    B, h, w = g["batch"], g["h"], g["w"]
    codes = torch.tensor(g["codes"])
    zq = sd["quantize.embedding.weight"][codes].reshape(B, h, w, -1).permute(0, 3, 1, 2).contiguous()
    pq_w, pq_b = sd["post_quant_conv.conv.weight"], sd["post_quant_conv.conv.bias"]
    z_ref = F.conv3d(F.pad(zq.unsqueeze(2), (0, 0, 0, 0, 2, 0)), pq_w, pq_b)[:, :, 0]
    z = F.conv2d(zq, pq_w[:, :, -1], pq_b)
    assert (z - z_ref).abs().max().item() < 1e-6
    px = dec.decode_latents(z, zq)
    ref = torch.tensor(g["pixels"]).reshape(g["out_shape"])
    assert list(px.shape) == g["out_shape"]
    assert (px - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


# ------------------------------------------------------------------------ f4: T5 encoder (LlamaGen text-to-image captions)
def _t5_case():
    import torch
    from transformers import T5Config, T5EncoderModel
    torch.manual_seed(0)
    cfg = T5Config(d_model=128, d_kv=64, num_heads=2, d_ff=256, num_layers=2, vocab_size=100,
                   feed_forward_proj="gated-gelu", dropout_rate=0.0)
    m = T5EncoderModel(cfg).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(p.bfloat16().float())          # bf16-representable weights: only activation rounding can differ
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 100, (2, 24), generator=g)
    mask = torch.ones(2, 24, dtype=torch.long)
    mask[1, 15:] = 0                               # a padded caption
    with torch.no_grad():
        ref = m(input_ids=ids, attention_mask=mask)["last_hidden_state"]
    return m, ids, mask, ref


def test_t5_encoder_structure_matches_hf_on_cpu(monkeypatch):
    """llamagen/language/t5.py:78-83 calls HF's T5EncoderModel; sjd_b200.t5_encoder restates its stack (T5LayerNorm, unscaled
    attention with bucketed relative position bias shared from block 0, padding mask, gated gelu_new FFN, final norm) around
    ONE primitive, y = x W^T.  With that primitive in fp32 torch the restatement must equal HF to rounding; with bf16 operands
    (what sjd_gemm_bf16 computes, checked on the GPU) it must stay within bf16 activation error.  No GPU -> construction raises."""
    import torch
    from sjd_b200 import t5_encoder
    m, ids, mask, ref = _t5_case()
    with pytest.raises(RuntimeError):
        t5_encoder.T5EncoderB200.from_module(m, device="cpu")
    monkeypatch.setattr(t5_encoder.T5EncoderB200, "__init__", _t5_cpu_init)
    enc = t5_encoder.T5EncoderB200.from_module(m, device="cpu")
    enc._linear = lambda x, w: x @ w.float().T
    out = enc.forward(ids, mask, graph=False)
    assert (out - ref).abs().max().item() < 2e-5
    enc._linear = lambda x, w: x.bfloat16().float() @ w.float().T
    out = enc.forward(ids, mask, graph=False)
    d = (out - ref).abs()
    assert d.max().item() < 3e-2 and d.mean().item() < 5e-3, (d.max().item(), d.mean().item())


def _t5_cpu_init(self, state_dict, *, num_heads, d_kv, relative_attention_num_buckets=32, relative_attention_max_distance=128,
                 layer_norm_epsilon=1e-6, gated=None, act="gelu_new", device="cuda"):
    """Test-only constructor: the product's __init__ minus the CUDA-only parts (workspace, staging buffer, device check)."""
    import torch
    if str(device) != "cpu":
        raise AssertionError("test helper")
    sd = {k[len("encoder."):] if k.startswith("encoder.") else k: v for k, v in state_dict.items()}
    self.device, self.H, self.dkv = torch.device("cpu"), num_heads, d_kv
    self.n_buckets, self.max_dist, self.eps, self.act = relative_attention_num_buckets, relative_attention_max_distance, layer_norm_epsilon, act
    self.embed = sd["embed_tokens.weight" if "embed_tokens.weight" in sd else "shared.weight"].float()
    self.d = self.embed.shape[1]
    self.gated = bool(gated)
    self.rel_bias = sd["block.0.layer.0.SelfAttention.relative_attention_bias.weight"].float()
    self.layers = []
    n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("block."))
    for i in range(n_layers):
        a, f = f"block.{i}.layer.0.", f"block.{i}.layer.1."
        self.layers.append(dict(
            ln1=sd[a + "layer_norm.weight"].float(),
            qkv=torch.cat([sd[a + "SelfAttention.q.weight"], sd[a + "SelfAttention.k.weight"], sd[a + "SelfAttention.v.weight"]], 0).bfloat16(),
            o=sd[a + "SelfAttention.o.weight"].bfloat16(), ln2=sd[f + "layer_norm.weight"].float(),
            wi=torch.cat([sd[f + "DenseReluDense.wi_0.weight"], sd[f + "DenseReluDense.wi_1.weight"]], 0).bfloat16(),
            wo=sd[f + "DenseReluDense.wo.weight"].bfloat16()))
    self.final_ln = sd["final_layer_norm.weight"].float()
    self.inner, self.d_ff = self.H * self.dkv, self.layers[0]["wo"].shape[1]


# ------------------------------------------------------------------------ attention_sw.cu: host-side work split
@pytest.mark.parametrize("W,H,Hkv,rows,kv_len,kv_lo", [
    (32, 32, 32, 2, 1200, [0, 66]),      # the bench shape: 64 heads x 10 key tiles
    (32, 32, 32, 2, 2400, [0, 66]),
    (64, 32, 8, 2, 4096, [0, 5]),        # Emu3-Gen window 64: four row tiles per kv head
    (16, 32, 8, 2, 700, [0, 300]),       # GQA stacked into one unit; row 1 hides two whole key tiles
    (1, 32, 32, 1, 90, [0]),             # a plain AR step, one key tile
    (8, 2, 2, 2, 300, [0, 36]),          # toy shape
])
def test_attention_sw_work_split_covers_every_unit_once(lib, W, H, Hkv, rows, kv_len, kv_lo):
    """The unit table attention_sw.cu's CTAs read (pure host arithmetic, sjd_debug_attn_sw_split): ranges are contiguous, cover
    [0, units) exactly once; in cluster mode CTA r*k + j holds slice j of run r's key tiles, every CTA has a tile and the grid
    is one wave; in the segment form (grid cap) the cost-balanced ranges differ by at most two units unless they cross heads."""
    import ctypes as C
    L = lib.lib()
    kl = (C.c_int32 * len(kv_lo))(*kv_lo)
    for sms, max_cluster, grid_cap in ((148, 4, 0), (148, 0, 0), (148, 4, 7), (132, 2, 0)):
        ub = (C.c_uint16 * 150)()
        info = (C.c_int32 * 8)()
        assert L.sjd_debug_attn_sw_split(W, H, Hkv, rows, kv_len, kl, sms, max_cluster, grid_cap, ub, info) == 0
        grid, k, ncols, n_units, n_chunks, mtiles, hpc, nv = list(info)
        Wp = (W + 7) // 8 * 8
        assert n_chunks == (kv_len + W + 127) // 128 and hpc * Wp <= 64 and ncols in (32, 64) and nv in (2, 3)
        assert n_units == n_chunks * Hkv * mtiles * rows and hpc * mtiles >= H // Hkv
        b = list(ub[:grid + 1])
        assert b[0] == 0 and b[grid] == n_units and all(b[i] <= b[i + 1] for i in range(grid)), (b, n_units)
        if grid_cap:
            assert k == 0 and grid <= grid_cap
        if k:
            runs = Hkv * mtiles * rows
            assert k in (1, 2, 4) and k <= max_cluster and grid == runs * k <= sms and k <= n_chunks
            for c in range(grid):
                r, j = divmod(c, k)
                assert b[c] == r * n_chunks + j * n_chunks // k and b[c + 1] > b[c]
        else:
            assert grid <= sms
            if grid == sms and n_units >= 4 * grid:
                sizes = [b[i + 1] - b[i] for i in range(grid)]
                assert max(sizes) - min(s for s in sizes if s) <= max(4, n_chunks // 2), sizes
