"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libsjd_b200.so; the oracle (oracle/) is only the checker.

  * GEMM (tcgen05 stream-K)            vs torch fp32 matmul of the same bf16 operands      (tolerance: fp32 sum order)
  * verify kernels                      vs oracle.sjd_oracle.verify                         (tokens / counts bit-exact)
  * SJD loop (engine + verify kernels)  vs fixtures minted from the UNMODIFIED reference    (token-exact, trace-exact)
  * window forward                      vs oracle.ref_forward.RefStack (bf16-emulating)     (<= 2 bf16 ulp of logits)
  * full engine on a tiny real stack    vs oracle loop replaying the engine's own logits    (token-exact)
"""
import numpy as np
import pytest
import torch

from conftest import ATTN_MODES, load_loop_goldens, set_attn

pytestmark = pytest.mark.gpu

GOLD = load_loop_goldens()


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sjd_b200  # noqa: F401
    from sjd_b200 import _lib, engine, families, model
    from oracle import fake_lm, ref_forward, sjd_oracle
    return dict(lib=_lib.lib(), _lib=_lib, engine=engine, model=model, families=families, O=sjd_oracle,
                RF=ref_forward, fake=fake_lm, dev=torch.device("cuda:0"))


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("N,K,M", [(128, 64, 16), (256, 128, 1), (384, 256, 32), (1000, 256, 48), (2304, 768, 32),
                                   (4096, 4096, 64), (4096, 4096, 256), (12288, 4096, 64), (22016, 4096, 64),
                                   (4096, 11008, 64), (16384, 768, 32), (5000, 512, 80)])
def test_gemm_matches_fp32_reference(env, N, K, M):
    L, _lib, dev = env["lib"], env["_lib"], env["dev"]
    torch.manual_seed(N * 7 + K + M)
    m_tile = (M + 15) // 16 * 16
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    x = torch.zeros(m_tile, K, device=dev, dtype=torch.bfloat16)
    x[:M] = torch.randn(M, K, device=dev).bfloat16()
    ws = torch.zeros(L.sjd_gemm_workspace_bytes(N, K, m_tile, 0), device=dev, dtype=torch.uint8)
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):   # second launch proves the in-kernel counters were left re-armed
        out.fill_(float("nan"))
        _lib.check(L.sjd_gemm_bf16(w.data_ptr(), N, K, x.data_ptr(), m_tile, M, out.data_ptr(), 1, 0, ws.data_ptr(), 0, st),
                   "gemm")
    torch.cuda.synchronize()
    ref = x[:M].double() @ w.double().T
    err = (out.double() - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()) * (K / 64) ** 0.5, f"max abs err {err}"


def test_gemm_stream_k_grid_independence(env):
    """The stream-K split must not change results beyond fp32 summation order, for any CTA count."""
    L, _lib, dev = env["lib"], env["_lib"], env["dev"]
    N, K, M, m_tile = 1536, 1024, 40, 48
    torch.manual_seed(3)
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    x = torch.zeros(m_tile, K, device=dev, dtype=torch.bfloat16)
    x[:M] = torch.randn(M, K, device=dev).bfloat16()
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for grid in (1, 7, 37, 148, 0):
        ws = torch.zeros(L.sjd_gemm_workspace_bytes(N, K, m_tile, grid), device=dev, dtype=torch.uint8)
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        _lib.check(L.sjd_gemm_bf16(w.data_ptr(), N, K, x.data_ptr(), m_tile, M, out.data_ptr(), 1, 0, ws.data_ptr(), grid, st),
                   "gemm")
        outs.append(out)
    torch.cuda.synchronize()
    ref = x[:M].float() @ w.float().T
    for o in outs:
        assert (o - ref).abs().max().item() < 1e-4
    # same grid twice -> bit identical (deterministic reduction order)
    ws = torch.zeros(L.sjd_gemm_workspace_bytes(N, K, m_tile, 0), device=dev, dtype=torch.uint8)
    out2 = torch.empty(M, N, device=dev, dtype=torch.float32)
    _lib.check(L.sjd_gemm_bf16(w.data_ptr(), N, K, x.data_ptr(), m_tile, M, out2.data_ptr(), 1, 0, ws.data_ptr(), 0, st),
               "gemm")
    torch.cuda.synchronize()
    assert torch.equal(outs[-1], out2)
    # bf16 output epilogue
    out3 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _lib.check(L.sjd_gemm_bf16(w.data_ptr(), N, K, x.data_ptr(), m_tile, M, out3.data_ptr(), 0, 0, ws.data_ptr(), 0, st),
               "gemm")
    torch.cuda.synchronize()
    assert torch.equal(out3, out2.bfloat16())


# ---------------------------------------------------------------------------------------------- verify
def _verify_case(env, seed, *, W=8, V=9216, scheme="speculative_jacobi", do_sample=True, guidance=3.0,
                 has_uncond=True, apply_cfg=True, temperature=1.0, top_k=2000, grammar=True, u_scale=1.0, top_p=1.0,
                 allow=None):
    O, dev, engine = env["O"], env["dev"], env["engine"]
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal(((2 if has_uncond else 1) * W, V)) * 1.5).astype(np.float32)
    if grammar:
        g = O.LuminaGrammar(image_top_k=top_k)
        ids = [1, 100, 8197, 8808, 8808] + [int(x) for x in rng.integers(4, 8196, size=seed % 11)]
        desc = g.describe(ids, W)
    else:
        desc = {"allow": None, "forced": [-1] * W, "top_k": top_k, "in_image": True, "no_cfg": False}
    desc["top_p"] = top_p
    if allow is not None:   # a wide candidate range (Emu3's 32 768 visual ids): the strided, range-restricted path
        desc["allow"] = tuple(allow)
    # distributions of this trip, to build plausible drafts: draft[i] ~ p[i-1], q = a perturbed p[i-1]
    s = O.logits_to_probs(logits, W, desc, has_uncond=has_uncond, apply_cfg=apply_cfg, guidance=guidance,
                          temperature=temperature)
    p = O.softmax(s)
    d_lo, d_hi = allow if allow is not None else (4, min(8196, V))
    draft = rng.integers(d_lo, d_hi, size=W).astype(np.int64)
    p_prev = np.zeros((W, V), np.float32)
    q_rows, q_idx = [None] * W, [-1] * W
    for i in range(1, W):
        if rng.random() < 0.75:
            pert = s[i - 1] + (rng.standard_normal(V) * 0.7).astype(np.float32)
            p_prev[i] = O.softmax(pert[None])[0]
            draft[i] = int(np.argmax(p_prev[i] / rng.exponential(size=V)))
            q_rows[i], q_idx[i] = p_prev[i], i
    e1 = rng.exponential(size=(W, V)).astype(np.float32)
    u = (rng.random(W) * u_scale).astype(np.float32)
    e2 = rng.exponential(size=V).astype(np.float32)
    ref = O.verify(logits, W, desc, draft, q_rows, has_uncond=has_uncond, apply_cfg=apply_cfg, guidance=guidance,
                   temperature=temperature, do_sample=do_sample, scheme=scheme, noise_e1=e1, noise_u=u, noise_e2=e2)
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev) if dt is None else \
        torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
    out = engine.verify_call(t(logits), W, V, desc, t(draft, torch.int32), torch.tensor(q_idx, dtype=torch.int32, device=dev),
                             t(p_prev), has_uncond=has_uncond, apply_cfg=apply_cfg, guidance=guidance,
                             temperature=temperature, do_sample=do_sample, scheme=0 if scheme == "speculative_jacobi" else 1,
                             noise_e1=t(e1), noise_u=t(u), noise_e2=t(e2), eoi_token=8196, text_top_k=10)
    torch.cuda.synchronize()
    return ref, out


@pytest.mark.parametrize("seed", range(12))
def test_verify_speculative_matches_oracle(env, seed):
    ref, out = _verify_case(env, seed, u_scale=0.6 if seed % 3 else 1.0)
    assert out["matched"] == ref.matched
    assert out["rejected"] == ref.rejected
    assert out["tokens"].cpu().numpy().tolist() == ref.tokens.tolist()
    assert out["next_tokens"].cpu().numpy().tolist() == ref.next_tokens.tolist()
    p = out["p"].cpu().numpy()
    assert np.array_equal(p > 0, ref.p > 0), "top-k / grammar support differs"
    assert np.abs(p - ref.p).max() <= 1e-6


def test_verify_covers_accepts_and_rejects(env):
    """The seeded cases above must exercise both outcomes, else the parity claim is hollow."""
    ms = [_verify_case(env, s, u_scale=0.6 if s % 3 else 1.0)[0].matched for s in range(12)]
    assert max(ms) >= 3 and min(ms) == 1, ms


@pytest.mark.parametrize("kw", [
    dict(scheme="jacobi", do_sample=False), dict(scheme="jacobi", do_sample=True),
    dict(do_sample=False), dict(apply_cfg=False), dict(has_uncond=False, apply_cfg=False),
    dict(temperature=0.7), dict(top_k=0), dict(top_k=1), dict(top_k=50, grammar=False, V=1024),
    dict(W=1), dict(W=32, V=16384, grammar=False, top_k=1000), dict(W=16, V=65536), dict(u_scale=0.05),
    dict(W=8, V=60000, grammar=False, allow=(20011, 52779), top_k=2048),            # range wider than the register path
    dict(W=4, V=60000, grammar=False, allow=(1024, 40000), top_k=2048, top_p=0.9),  # ... with a nucleus on top
])
def test_verify_variants_match_oracle(env, kw):
    for seed in (1, 2, 3):
        ref, out = _verify_case(env, 100 + seed, **kw)
        assert out["matched"] == ref.matched, kw
        assert out["tokens"].cpu().numpy().tolist() == ref.tokens.tolist(), kw
        assert np.abs(out["p"].cpu().numpy() - ref.p).max() <= 1e-6


@pytest.mark.parametrize("kw", [
    dict(top_p=0.9), dict(top_p=0.5, top_k=0), dict(top_p=0.8, grammar=False, V=16384, top_k=1000, W=16),
    dict(top_p=0.95, scheme="jacobi"), dict(top_p=0.7, do_sample=False), dict(top_p=0.6, V=65536, grammar=False, top_k=0),
    dict(top_p=0.85, temperature=0.8, u_scale=0.1),
])
def test_verify_top_p_matches_oracle(env, kw):
    """TopPLogitsWarper3d after top-k (a10).  The reference's removed set is decided by an fp32 running sum in sorted
    order; ours by an exact fixed-point sum, so an entry whose running sum lies within 2e-6 of 1 - top_p may fall on
    either side.  Seeds whose boundary is that close are skipped (never more than one in three); everything else must
    agree exactly: support, tokens, accepted count, probabilities to 1e-6."""
    O = env["O"]
    checked = 0
    for seed in (1, 2, 3):
        ref, out = _verify_case(env, 200 + seed, **kw)
        p = out["p"].cpu().numpy()
        if not np.array_equal(p > 0, ref.p > 0):
            # is every disagreement a boundary case?  mass of the disputed entries relative to the threshold
            diff = (p > 0) != (ref.p > 0)
            rows = np.flatnonzero(diff.any(1))
            for r in rows:
                assert diff[r].sum() <= 1, "more than one disputed entry in a row"
                kept = ref.p[r] > 0
                srt = np.sort(np.where(kept | diff[r], np.maximum(ref.p[r], p[r]), 0.0))
                assert np.abs(np.cumsum(srt[srt > 0]) / srt.sum() - (1 - kw["top_p"])).min() < 2e-6
            continue
        checked += 1
        assert out["matched"] == ref.matched, kw
        assert out["tokens"].cpu().numpy().tolist() == ref.tokens.tolist(), kw
        assert np.abs(p - ref.p).max() <= 1e-6
        n_allowed = 8192 if kw.get("grammar", True) else kw.get("V", 9216)
        topk_only = min(kw.get("top_k", 2000) or n_allowed, n_allowed)
        assert ((ref.p > 0).sum(1) < topk_only).all(), "top-p removed nothing: hollow case"
    assert checked >= 2


def test_verify_top_p_ties_removed_lowest_id_first(env):
    """Equal probabilities straddling the threshold: a stable ascending sort removes the lowest ids first, as many as
    fit under 1 - top_p; the largest entry always survives, even for top_p = 0."""
    O, dev, engine = env["O"], env["dev"], env["engine"]
    W, V = 3, 640
    logits = np.full((W, V), -np.inf, np.float32)
    logits[0, [3, 50, 200, 400, 639]] = [0.0, 0.0, 0.0, 0.0, np.log(4.0)]    # p = 1/8 x4, 1/2: top_p 0.7 removes ids 3, 50
    logits[1, [7, 9]] = [1.0, 1.0]                                          # two equal entries, top_p 0.7: remove id 7? 0.5 > 0.3: none
    logits[2, 100:110] = 0.0                                                # ten equal entries of 0.1: three removed (0.3 <= 0.3000000119)
    e1 = np.ones((W, V), np.float32)
    for top_p, want in ((0.7, [[200, 400, 639], [7, 9], list(range(103, 110))]), (0.0, [[639], [9], [109]])):
        desc = {"allow": None, "forced": [-1] * W, "top_k": 0, "top_p": top_p}
        ref = O.verify(logits, W, desc, np.array([0, 5, 5]), [None] * W, has_uncond=False, apply_cfg=False, guidance=1.0,
                       do_sample=True, scheme="jacobi", noise_e1=e1)
        out = engine.verify_call(torch.from_numpy(logits).to(dev), W, V, desc, torch.tensor([0, 5, 5], dtype=torch.int32, device=dev),
                                 None, None, has_uncond=False, apply_cfg=False, guidance=1.0, temperature=1.0,
                                 do_sample=True, scheme=1, noise_e1=torch.from_numpy(e1).to(dev))
        p = out["p"].cpu().numpy()
        for r in range(W):
            assert np.flatnonzero(ref.p[r] > 0).tolist() == want[r], (top_p, r, np.flatnonzero(ref.p[r] > 0))
            assert np.flatnonzero(p[r] > 0).tolist() == want[r], (top_p, r, np.flatnonzero(p[r] > 0))
        assert np.abs(p - ref.p).max() <= 1e-6
        assert out["tokens"].cpu().numpy().tolist() == ref.tokens.tolist()


def test_verify_topk_ties_and_small_support(env):
    """Ties with the k-th largest score are kept; k beyond the finite support removes nothing."""
    O, dev, engine = env["O"], env["dev"], env["engine"]
    W, V = 2, 512
    logits = np.full((W, V), -3.0, np.float32)
    logits[0, [5, 9, 17, 33]] = [2.0, 1.0, 1.0, 1.0]      # k = 2 -> the three tied 1.0 stay
    logits[1, :] = -np.inf
    logits[1, [7, 8]] = [0.5, 0.25]                        # two finite entries, k = 2
    desc = {"allow": None, "forced": [-1, -1], "top_k": 2}
    e1 = np.ones((W, V), np.float32)
    ref = O.verify(logits, W, desc, np.array([0, 5]), [None, None], has_uncond=False, apply_cfg=False, guidance=1.0,
                   do_sample=True, scheme="jacobi", noise_e1=e1)
    out = engine.verify_call(torch.from_numpy(logits).to(dev), W, V, desc, torch.tensor([0, 5], dtype=torch.int32, device=dev),
                             None, None, has_uncond=False, apply_cfg=False, guidance=1.0, temperature=1.0,
                             do_sample=True, scheme=1, noise_e1=torch.from_numpy(e1).to(dev))
    p = out["p"].cpu().numpy()
    assert (p[0] > 0).sum() == 4 and (p[1] > 0).sum() == 2
    assert np.abs(p - ref.p).max() <= 1e-6
    assert out["tokens"].cpu().numpy().tolist() == ref.tokens.tolist()


# ---------------------------------------------------------------------------- SJD loop vs the reference
class _FakeStack:
    def __init__(self, V, rows, dev, max_len=4096):
        from types import SimpleNamespace
        self.shape = SimpleNamespace(vocab=V)
        self.rows, self.device, self.max_len = rows, dev, max_len


def _engine_for_case(env, case, noise_dev="cpu"):
    engine, fake, dev = env["engine"], env["fake"], env["dev"]
    j = case["jacobi"]
    do_cfg = j["do_cfg"] and j["guidance_scale"] != 1
    rows = 2 if do_cfg else 1
    if case["grammar"] == "lumina":
        grammar = engine.LuminaGrammarState(image_top_k=case["image_top_k"], text_top_k=case["text_top_k"])
    elif case["grammar"] == "emu3":
        e = case["emu3"]
        grammar = engine.Emu3GrammarState(e["height"], e["width"], e["img_token"], e["eol"], e["eof"], e["eoi"], e["eos"],
                                          e["pad"], e["visual"][0], e["visual"][1], top_k=case["image_top_k"])
    elif case["grammar"] == "anole":
        a = case["anole"]
        grammar = engine.AnoleGrammarState(a["boi"], a["eoi"], a["eos"], a["image"][0], a["image"][1], a["image_seq_length"],
                                           case["max_length"], len(case["prompt"]), top_k=case["image_top_k"],
                                           mode=a.get("mode", "image-only"))
    else:
        grammar = engine.PlainTopKState(top_k=case["image_top_k"], top_p=case.get("top_p", 1.0))

    class Eng(engine.SJDEngine):
        def _forward(self, row_tokens, kv_len, kv_lo, n_logit, embeds=None):
            lg = fake.fake_logits(row_tokens, kv_len, n_logit, case["V"], case["sharp"])
            return torch.from_numpy(lg).to(dev).view(len(row_tokens), n_logit, case["V"])

    params = engine.SJDParams(**j)
    return Eng(_FakeStack(case["V"], rows, dev), params, grammar, torch.arange(*case["img_vocab"]),
               noise_factory=lambda seed, d: engine.NoiseSource(seed, d, gen_device=noise_dev))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_engine_loop_reproduces_reference_tokens(env, name):
    """Host loop + CUDA verify kernels, fed the reference's logits and the reference's CPU noise stream, must emit
    the reference's exact token sequence and accepted-count trace."""
    g = GOLD[name]
    case, ref = g["case"], g["result"]
    eng = _engine_for_case(env, case)
    ids = eng.generate(case["prompt"], max_length=case["max_length"], eos_token_ids=case["eos"],
                       do_sample=case["do_sample"], collect_trace=True)
    assert [t[1] for t in eng.stats.trace] == [t["n_new"] for t in ref["trace"]]
    assert [t[0] for t in eng.stats.trace] == [t["W"] for t in ref["trace"]]
    assert ids == ref["ids"]
    assert eng.stats.nfe == len(ref["trace"])


# ------------------------------------------------------------------------------------------ forward
def _family(env, name):
    RF = env["RF"]
    if name == "chameleon":
        return RF.StackConfig(2, 256, 2, 2, 128, 512, 9216, 1e-5, qk_norm=True), \
            RF.rope_tables_rotate_half(128, 512, 10000.0, True), [0, 36]
    if name == "llamagen":
        return RF.StackConfig(3, 256, 4, 4, 64, 768, 1024, 1e-5, rope_interleaved=True, family="llamagen"), \
            RF.rope_tables_llamagen_2d(24, 64, 10000, 1), [0, 0]
    return RF.StackConfig(2, 512, 4, 1, 128, 1024, 5000, 1e-5, family="emu3", rope_theta=1e6), \
        RF.rope_tables_rotate_half(128, 512, 1e6, True), [0, 5]


@pytest.mark.parametrize("attn", ATTN_MODES)
@pytest.mark.parametrize("family", ["chameleon", "llamagen", "emu3"])
def test_window_forward_matches_reference_stack(env, family, attn, monkeypatch):
    """Prefill, an AR step, two Jacobi windows with a 9-token roll-back in between, and a short window —
    logits within 4 bf16 ulp (mean well below one ulp) of the bf16-emulating reference, and no further from the exact
    fp32 forward than that bf16 reference itself is."""
    RF, model, dev = env["RF"], env["model"], env["dev"]
    # every attention kernel on every shape it supports (the context reads SJD_ATTN when it is created): "tc" = tcgen05 +
    # TMEM (attention_tc.cu), "tct" = its transposed small-window variant (attention_tct.cu; head dim 128, windows <= 64,
    # else the next choice), "mma" = mma.sync (attention.cu), "auto" = the per-window choice the product makes
    set_attn(monkeypatch, attn)
    cfg, (cos, sin), kv_lo = _family(env, family)
    w = RF.random_weights(cfg, seed=1, device=dev)
    rows, max_len = 2, 320
    shape = model.StackShape(cfg.n_layers, cfg.d_model, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.d_ff,
                             cfg.vocab, cfg.rms_eps, cfg.qk_norm, cfg.rope_interleaved)
    ds = model.DeviceStack(shape, w, cos, sin, rows, max_len, dev)
    ref = RF.RefStack(cfg, w, cos.to(dev), sin.to(dev), rows, max_len, emulate_bf16=True)
    ref32 = RF.RefStack(cfg, w, cos.to(dev), sin.to(dev), rows, max_len, emulate_bf16=False)
    g = torch.Generator().manual_seed(5)
    kv_len = 0
    for step, W in enumerate([37, 1, 16, 16, 5, 128]):
        if step == 3:
            kv_len -= 9
        ids = torch.randint(0, cfg.vocab, (rows, W), generator=g).to(dev)
        pos = torch.arange(kv_len, kv_len + W, device=dev)[None].repeat(rows, 1)
        rope_pos = torch.stack([(pos[b] - kv_lo[b]).clamp(min=0) for b in range(rows)])
        n = 1 if step == 0 else W
        lg = ds.forward(W, rope_pos.int().flatten().contiguous(), pos.int().flatten().contiguous(), kv_len, kv_lo,
                        ids=ids.int().flatten().contiguous(), n_logit_tokens=n).clone()
        lr = ref.forward(ids=ids, rope_pos=rope_pos, kv_len=kv_len, kv_lo=kv_lo, cache_pos=pos, n_logit_tokens=n)
        torch.cuda.synchronize()
        l32 = ref32.forward(ids=ids, rope_pos=rope_pos, kv_len=kv_len, kv_lo=kv_lo, cache_pos=pos, n_logit_tokens=n)
        torch.cuda.synchronize()
        # (a) against the bf16-emulating reference: two bf16 pipelines that differ only in fp32 summation order
        #     disagree by a few bf16 ulp OF THE LOGIT SCALE at worst (a logit is a sum of O(1) terms, so its error
        #     does not shrink with its own magnitude) and by a fraction of that ulp on average
        ulp = 2.0 ** (torch.floor(torch.log2(lr.abs().max())).item() - 7)
        assert (lg - lr).abs().max().item() <= 2.5 * ulp, f"{family} step {step}: max {(lg - lr).abs().max().item()}"
        assert (lg - lr).abs().mean().item() < 0.5 * ulp, (lg - lr).abs().mean().item()
        # (b) against the exact fp32 forward: our error must not exceed the reference's own bf16 error
        e_ours, e_ref = (lg - l32).abs(), (lr - l32).abs()
        assert e_ours.max().item() <= 1.5 * e_ref.max().item() + 1e-6, (e_ours.max().item(), e_ref.max().item())
        assert e_ours.mean().item() <= 1.25 * e_ref.mean().item() + 1e-6, (e_ours.mean().item(), e_ref.mean().item())
        kv_len += W
    ds.close()


def test_engine_on_real_stack_matches_oracle_replay(env):
    """Tiny Chameleon-shaped stack, real kernels end to end.  The oracle loop is replayed on the logits the
    engine's forward produced (captured per trip) with the same CPU noise stream: tokens must be identical."""
    RF, model, engine, O, dev = env["RF"], env["model"], env["engine"], env["O"], env["dev"]
    cfg, (cos, sin), _ = _family(env, "chameleon")
    w = RF.random_weights(cfg, seed=4, std=0.08, device=dev)
    shape = model.StackShape(cfg.n_layers, cfg.d_model, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.d_ff,
                             cfg.vocab, cfg.rms_eps, cfg.qk_norm, cfg.rope_interleaved)
    ds = model.DeviceStack(shape, w, cos, sin, 2, 256, dev)
    captured = []

    class Eng(engine.SJDEngine):
        def _forward(self, row_tokens, kv_len, kv_lo, n_logit, embeds=None):
            lg = super()._forward(row_tokens, kv_len, kv_lo, n_logit, embeds)
            captured.append((len(row_tokens[0]), kv_len, n_logit, lg.detach().cpu().numpy().reshape(-1, cfg.vocab).copy()))
            return lg

    prompt = [1, 100, 200, 300, 8197, 8808, 8808]
    kw = dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=8 * 8 + 8 - 10, max_num_new_tokens=8, guidance_scale=3.0,
              seed=0, multi_token_init_scheme="random", do_cfg=True, prefix_token_sampler_scheme="speculative_jacobi")
    eng = Eng(ds, engine.SJDParams(**kw), engine.LuminaGrammarState(), torch.arange(4, 8196),
              noise_factory=lambda seed, d: engine.NoiseSource(seed, d, gen_device="cpu"))
    max_len = len(prompt) + 8 * 9 + 2
    ids = eng.generate(prompt, max_length=max_len, eos_token_ids=[8710], kv_lo=[0, len(prompt) - 1], collect_trace=True)
    it = iter(captured)

    def replay(rows_tokens, kv_len, n):
        W, kv, nn, lg = next(it)
        assert (W, kv, nn) == (len(rows_tokens[0]), kv_len, n), "oracle and engine disagree on the window schedule"
        return lg

    ids_o, nfe_o = O.decode(replay, prompt, params=O.OracleParams(**kw), grammar=O.LuminaGrammar(),
                            img_vocab=np.arange(4, 8196), max_length=max_len, eos_ids=[8710], rows=2)
    assert ids == ids_o and eng.stats.nfe == nfe_o
    # grammar sanity on real kernels: EOL every 9th image token, EOI after 8 rows
    img = ids[len(prompt):]
    assert all(img[i] == 8803 for i in range(8, 72, 9)) and img[72] == 8196
    assert eng.stats.nfe < len(img), "Jacobi decoding must need fewer forwards than tokens"
    ds.close()


# ------------------------------------------------------------------------------- LlamaGen boundary (config 1)
@pytest.mark.parametrize("golden", ["llamagen_flow.json", "llamagen_flow_gptb.json"])
def test_llamagen_solver_flow_on_gpu(env, golden):
    """(llamagen_flow_gptb.json = BASELINE config 1 at its stated size: GPT-B, 256 tokens, window 16, cfg 4, top-k 1000.)
    The reference's test_llamagen.py call sequence against this repo's drop-in modules, on the GPU:
    llamagen.llamagen.Transformer -> renew_llamagen -> renew_sampler -> LlamaGenSolver.generate.  Every forward's
    logits are captured and the oracle flow (pinned to the reference by tests/golden/llamagen_flow.json) is replayed on
    them with the same generators: identical image tokens."""
    import json
    from conftest import GOLDEN
    from llamagen.llamagen import ModelArgs, Transformer
    from llamagen.llamagen_solver import LlamaGenSolver, renew_llamagen
    from scheduler.jacobi_iteration_lumina_mgpt import renew_sampler
    from oracle import llamagen_flow
    RF, O, model_mod, dev = env["RF"], env["O"], env["model"], env["dev"]
    g = json.loads((GOLDEN / golden).read_text())
    case, ref = g["case"], g["result"]
    args = ModelArgs(dim=case["dim"], n_layer=case["n_layer"], n_head=case["n_head"], vocab_size=case["vocab"],
                     block_size=case["grid"] ** 2, cls_token_num=case["cls_token_num"], num_classes=case["num_classes"],
                     model_type="c2i", class_dropout_prob=0.1)
    m = Transformer(args)
    cfg, w, cls_table, _, _ = llamagen_flow.build_stack(case, ref["ff"], ref["norm_eps"], ref["rope_base"])
    ff = ref["ff"]
    with torch.no_grad():
        m.tok_embeddings.weight.copy_(w["embed"]); m.norm.weight.copy_(w["final_norm"]); m.output.weight.copy_(w["lm_head"])
        m.cls_embedding.embedding_table.weight.copy_(cls_table)
        for L, wl in zip(m.layers, w["layers"]):
            L.attention_norm.weight.copy_(wl["attn_norm"]); L.attention.wqkv.weight.copy_(wl["wqkv"])
            L.attention.wo.weight.copy_(wl["wo"]); L.ffn_norm.weight.copy_(wl["ffn_norm"])
            L.feed_forward.w1.weight.copy_(wl["w_gate_up"][:ff]); L.feed_forward.w3.weight.copy_(wl["w_gate_up"][ff:])
            L.feed_forward.w2.weight.copy_(wl["w_down"])
    m = m.to(dev, torch.bfloat16).eval()
    m.__class__ = renew_llamagen(m.__class__)
    m._init_new_params(**case["jacobi"])
    m.__class__ = renew_sampler(m.__class__)
    m._init_new_params(use_chameleon_tokenizer=False, **case["jacobi"])
    m.img_vocab = torch.arange(case["vocab"])
    captured = []
    orig_forward = model_mod.DeviceStack.forward

    def spy(self, *a, **k):
        lg = orig_forward(self, *a, **k)
        captured.append(lg.detach().float().cpu().numpy().reshape(-1, case["vocab"]).copy())
        return lg

    model_mod.DeviceStack.forward = spy
    try:
        solver = LlamaGenSolver(m, case["top_k"], case["top_p"])
        torch.manual_seed(case["global_seed"])
        # the engine's default noise generator lives on the GPU; the oracle replay below uses the same device generator
        out = solver.generate(torch.tensor([case["class_id"]], device=dev), case["grid"] ** 2, None,
                              cfg_scale=case["cfg_scale"], temperature=case["temperature"], top_k=case["top_k"],
                              top_p=case["top_p"], sample_logits=True)
    finally:
        model_mod.DeviceStack.forward = orig_forward
    tokens = out[0].tolist()
    assert len(tokens) == case["grid"] ** 2 and all(0 <= t < case["vocab"] for t in tokens)
    assert m.sjd_stats.nfe < len(tokens), "Jacobi decoding must need fewer forwards than tokens"
    # ---- oracle replay on the captured logits (first token: global generator on the device, like the solver) ----
    it = iter(captured)
    torch.manual_seed(case["global_seed"])
    lg0 = torch.from_numpy(next(it)).to(dev)
    from llamagen.llamagen_solver import _first_token
    tok0 = int(_first_token(lg0, case["cfg_scale"], temperature=case["temperature"], top_k=case["top_k"],
                            top_p=case["top_p"])[0, 0])
    assert tok0 == tokens[0]
    ids_o, nfe_o = O.decode(lambda r, k, n: next(it), [tok0], params=O.OracleParams(**case["jacobi"]),
                            grammar=O.PlainTopK(top_k=case["top_k"]), img_vocab=np.arange(case["vocab"]),
                            max_length=case["grid"] ** 2, eos_ids=[], rows=2, do_sample=True, temperature=1.0,
                            kv_len0=case["cls_token_num"], noise=O.TorchNoise(case["jacobi"]["seed"], device=str(dev)))
    assert ids_o[-len(tokens):] == tokens and nfe_o == m.sjd_stats.nfe


# ------------------------------------------------------------------------------- Emu3 boundary (config 4 family)
def test_emu3_adaptor_flow_on_gpu(env):
    """The Emu3 entry points of the plugin API (scheduler.jacobi_iteration_emu3: renew_end_of_line_logit_processor_3d,
    renew_sampler_forward, prepare_batch_cfg_model_inputs, `_sample(..., neg_input_ids=...)`) on a tiny Llama-style GQA
    model (what Emu3's LM is): left-padded negative prompt as the CFG-uncond row, positional EOL/EOF/EOI/EOS grammar,
    HF TopKLogitsWarper.  Engine output == oracle replay on the captured logits; grammar positions checked."""
    from transformers import GenerationConfig, LlamaConfig, LlamaForCausalLM
    from transformers.generation.logits_process import LogitsProcessorList, TopKLogitsWarper
    from transformers.generation.stopping_criteria import EosTokenCriteria, MaxLengthCriteria, StoppingCriteriaList
    from scheduler.jacobi_iteration_emu3 import renew_end_of_line_logit_processor_3d, renew_sampler_forward
    from scheduler.jacobi_iteration_lumina_mgpt import renew_sampler
    O, model_mod, dev = env["O"], env["model"], env["dev"]
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=4096, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=1, rms_norm_eps=1e-5, max_position_embeddings=256,
                      pad_token_id=0, tie_word_embeddings=False)
    m = LlamaForCausalLM(cfg)
    with torch.no_grad():
        for p_ in m.parameters():
            if p_.dim() >= 2:
                p_.normal_(0.0, 0.08)
    m = m.to(dev, torch.bfloat16).eval()

    class Helper:   # stand-in for Emu3PrefixConstrainedLogitsHelper (emu3/mllm/utils_emu3.py:19-41): attributes only
        def __init__(self):
            self.height, self.width, self.img_token = 4, 6, 900
            self.eol_token, self.eof_token, self.eoi_token, self.eos_token, self.pad_token = 901, 902, 903, 904, 905
            self.visual_tokens = torch.arange(1000, 3048)
            self.offset_cache = {}

    fn = Helper()
    fn.__class__ = renew_end_of_line_logit_processor_3d(fn.__class__)
    jac = dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=200, max_num_new_tokens=8, guidance_scale=3.0, seed=4,
               multi_token_init_scheme="random", do_cfg=True, prefix_token_sampler_scheme="speculative_jacobi")
    m.__class__ = renew_sampler(m.__class__)
    m._init_new_params(use_chameleon_tokenizer=False, **jac)
    m.__class__ = renew_sampler_forward(m.__class__)
    m._init_new_params(visual_tokens=fn.visual_tokens, **jac)
    assert m.img_vocab is not None and int(m.img_vocab[0]) == 1000
    pos = torch.tensor([[5, 17, 23, 11, 900]], device=dev)
    neg = torch.tensor([[7, 9, 900]], device=dev)
    mi = m.prepare_batch_cfg_model_inputs(pos, neg)
    assert mi["input_ids"].tolist() == [[5, 17, 23, 11, 900], [0, 0, 7, 9, 900]]
    assert mi["attention_mask"].tolist() == [[1, 1, 1, 1, 1], [0, 0, 1, 1, 1]]
    P, n_img = 5, 7 * 4 + 3
    crit = StoppingCriteriaList([MaxLengthCriteria(P + n_img + 4), EosTokenCriteria(eos_token_id=[904])])
    gc = GenerationConfig(max_length=P + n_img + 4, do_sample=True, temperature=1.0, top_k=None)
    captured = []
    orig_forward = model_mod.DeviceStack.forward

    def spy(self, *a, **k):
        lg = orig_forward(self, *a, **k)
        captured.append(lg.detach().float().cpu().numpy().reshape(-1, cfg.vocab_size).copy())
        return lg

    model_mod.DeviceStack.forward = spy
    try:
        out = m._sample(mi["pos_input_ids"], logits_processor=LogitsProcessorList([fn, TopKLogitsWarper(top_k=512)]),
                        stopping_criteria=crit, generation_config=gc, synced_gpus=False, streamer=None,
                        attention_mask=mi["attention_mask"], neg_input_ids=neg)
    finally:
        model_mod.DeviceStack.forward = orig_forward
    ids = out[0].tolist()
    img = ids[P:]
    assert all(img[i] == 901 for i in range(6, 28, 7)), img          # EOL closes every row of 6 visual tokens
    assert img[28:31] == [902, 903, 904], img                          # EOF, EOI, EOS
    assert all(1000 <= t < 3048 for i, t in enumerate(img[:28]) if i % 7 != 6)
    it = iter(captured)
    g = O.Emu3Grammar(4, 6, 900, 901, 902, 903, 904, 905, 1000, 3048, top_k=512)
    ids_o, nfe_o = O.decode(lambda r, k, n: next(it), [5, 17, 23, 11, 900], params=O.OracleParams(**jac), grammar=g,
                            img_vocab=np.arange(1000, 3048), max_length=P + n_img + 4, eos_ids=[904], rows=2,
                            do_sample=True, noise=O.TorchNoise(jac["seed"], device=str(dev)))
    assert ids_o == ids and nfe_o == m.sjd_stats.nfe


# ------------------------------------------------------- independent check: HF's own Chameleon forward (Anole's model)
def test_forward_matches_hf_chameleon_bf16(env):
    """The module tree Anole runs (transformers.ChameleonForConditionalGeneration, model_wrappers/model_loader.py:12-13)
    packed by hf_api.pack_hf_decoder and run by the engine vs that model's OWN PyTorch bf16 forward on the GPU:
    prefill and window logits within 3 bf16 ulp of the logit scale, mean within half an ulp (both pipelines round every linear / norm output to bf16)."""
    from transformers import ChameleonConfig, ChameleonForConditionalGeneration
    from sjd_b200 import hf_api
    dev = env["dev"]
    torch.manual_seed(1)
    cfg = ChameleonConfig(vocab_size=2048, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                          num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-5,
                          vocabulary_map={"<image>": 3, "IMGIMGA": 4, "IMGIMGB": 5},
                          vq_config={"embed_dim": 8, "num_embeddings": 16, "resolution": 32, "channel_multiplier": [1, 1],
                                     "base_channels": 32, "num_res_blocks": 1, "latent_channels": 8})
    m = ChameleonForConditionalGeneration(cfg)
    with torch.no_grad():
        for n_, p_ in m.named_parameters():
            if "vqmodel" in n_:
                continue
            if p_.dim() >= 2 and "norm" not in n_:
                p_.normal_(0.0, 0.02)
            elif "norm" in n_ and n_.endswith("weight"):
                p_.normal_(1.0, 0.1)
            elif "norm" in n_ and n_.endswith("bias"):
                p_.normal_(0.0, 0.1)
    m = m.to(dev, torch.bfloat16).eval()
    P = 23
    ids = torch.randint(6, cfg.vocab_size, (1, P), device=dev)
    with torch.no_grad():
        ref = m(input_ids=ids, use_cache=False).logits.float()[0]          # [P, V]
    stack = hf_api.pack_hf_decoder(m, max_len=128, rows=1, device=dev)
    pos = torch.arange(P, dtype=torch.int32, device=dev)
    ours = stack.forward(P, pos, pos, 0, [0], ids=ids[0].int().contiguous(), n_logit_tokens=P)[0].clone()
    torch.cuda.synchronize()
    keep = (ref > -1e30).all(0)     # HF blanks the image-token columns of the logits (finfo.min); compare the others
    assert int(keep.sum()) >= cfg.vocab_size - 8
    ref, ours = ref[:, keep], ours[:, keep]
    ulp = 2.0 ** (torch.floor(torch.log2(ref.abs().max())).item() - 7)
    err = (ours - ref).abs()
    print('hf-chameleon prefill: max err', err.max().item(), 'mean', err.mean().item(), 'ulp', ulp, 'absmax', ref.abs().max().item())
    assert err.max().item() <= 3.0 * ulp, (err.max().item(), ulp)
    assert err.mean().item() <= 0.5 * ulp
    # a window step over the cache as well: tokens P..P+7 with the first P cached by the call above
    W = 8
    ids2 = torch.randint(6, cfg.vocab_size, (1, W), device=dev)
    with torch.no_grad():
        ref2 = m(input_ids=torch.cat([ids, ids2], 1), use_cache=False).logits.float()[0, P:]
    pos2 = torch.arange(P, P + W, dtype=torch.int32, device=dev)
    ours2 = stack.forward(W, pos2, pos2, P, [0], ids=ids2[0].int().contiguous(), n_logit_tokens=W)[0].clone()
    torch.cuda.synchronize()
    err2 = (ours2[:, keep] - ref2[:, keep]).abs()
    assert err2.max().item() <= 3.0 * ulp, (err2.max().item(), ulp)
    stack.close()


# ------------------------------------------------------------------------------- Anole boundary (config 5 family)
def test_anole_adaptor_flow_on_gpu(env):
    """scheduler.jacobi_iteration_anhole.renew_pipeline_sampler on a tiny HF ChameleonForConditionalGeneration (the module
    tree Anole runs), through HF's own generate(): the five 3-D Chameleon processors + TopK evaluated by sjd_verify, CFG
    with the hidden-prefix uncond row, real kernels end to end.  Engine output == oracle replay on the captured logits
    (the oracle applies the reference processors as masks); begin-of-image first, then image ids only."""
    from transformers import ChameleonConfig, ChameleonForConditionalGeneration
    from scheduler.jacobi_iteration_anhole import renew_pipeline_sampler
    O, model_mod, dev = env["O"], env["model"], env["dev"]
    torch.manual_seed(2)
    n_img, S, V = 200, 20, 512
    names = {f"IMGIMG{chr(65 + i // 100)}{chr(65 + (i // 10) % 10)}{chr(65 + i % 10)}Z": 4 + i for i in range(n_img)}
    cfg = ChameleonConfig(vocab_size=V, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                          num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-5,
                          vocabulary_map={"<image>": 3, **names}, eos_token_id=2, bos_token_id=0, pad_token_id=1,
                          vq_config={"embed_dim": 8, "num_embeddings": 16, "resolution": 32, "channel_multiplier": [1, 1],
                                     "base_channels": 32, "num_res_blocks": 1, "latent_channels": 8})
    m = ChameleonForConditionalGeneration(cfg)
    with torch.no_grad():
        for n_, p_ in m.named_parameters():
            if "vqmodel" not in n_ and p_.dim() >= 2 and "norm" not in n_:
                p_.normal_(0.0, 0.08)
    m = m.to(dev, torch.bfloat16).eval()
    boi, eoi = 300, 301
    m.model.vocabulary_mapping.boi_token_id = boi
    m.model.vocabulary_mapping.eoi_token_id = eoi

    class Proc:
        image_seq_length = S

    jac = dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=S + 4, max_num_new_tokens=6, guidance_scale=3.0, seed=3,
               multi_token_init_scheme="random", do_cfg=True, prefix_token_sampler_scheme="speculative_jacobi")
    m = renew_pipeline_sampler(m, Proc(), image_top_k=50, text_top_k=10, **jac)
    m.img_vocab = torch.arange(4, 4 + n_img)     # random drafts from this model's image ids
    captured = []
    orig_forward = model_mod.DeviceStack.forward

    def spy(self, *a, **k):
        lg = orig_forward(self, *a, **k)
        captured.append(lg.detach().float().cpu().numpy().reshape(-1, V).copy())
        return lg

    prompt = [0, 400, 401, 402, 403]
    ids_in = torch.tensor([prompt], device=dev)
    model_mod.DeviceStack.forward = spy
    try:
        out = m.generate(input_ids=ids_in, attention_mask=torch.ones_like(ids_in),
                         multimodal_generation_mode="image-only", max_new_tokens=S + 2, do_sample=True, top_k=40)
    finally:
        model_mod.DeviceStack.forward = orig_forward
    ids = out[0].tolist()
    P = len(prompt)
    assert len(ids) == P + S + 2 and ids[P] == boi
    assert all(4 <= t < 4 + n_img or t == eoi for t in ids[P + 1:]), ids
    assert all(4 <= t < 4 + n_img for t in ids[P + 1:P + 1 + S]), ids
    assert m.sjd_stats.nfe < S + 2, "Jacobi decoding must need fewer forwards than tokens"
    it = iter(captured)
    g = O.AnoleGrammar(vocab=V, boi=boi, eoi=eoi, eos=2, image_lo=4, image_hi=4 + n_img, image_seq_length=S,
                       max_length=P + S + 2, begin_index=P, top_k=40)
    ids_o, nfe_o = O.decode(lambda r, k, n: next(it), prompt, params=O.OracleParams(**jac), grammar=g,
                            img_vocab=np.arange(4, 4 + n_img), max_length=P + S + 2, eos_ids=[2], rows=2,
                            do_sample=True, noise=O.TorchNoise(jac["seed"], device=str(dev)))
    assert ids_o == ids and nfe_o == m.sjd_stats.nfe


# ------------------------------------------------------------- size-independent properties on the real kernels
def _lumina_slice(env, n_layers, max_len):
    """Lumina-mGPT-7B width (d 4096, 32 heads, ff 11008, V 65536), a few layers deep."""
    from sjd_b200 import families
    model, dev = env["model"], env["dev"]
    shape = families.lumina_7b()
    shape.n_layers = n_layers
    w = families.random_weights(shape, seed=3, std=0.03, device=dev)
    cos, sin = families.rope_rotate_half(128, max_len, 10000.0, True)
    ds = model.DeviceStack(shape, w, cos, sin, 2, max_len, dev)
    del w
    return ds


def test_full_width_decode_is_deterministic_and_grammatical(env):
    """BASELINE config 2 at full width (2 layers deep): a 16 x 16 latent image with window 32, cfg 3, top-k 2000 through
    engine + real kernels.  Same seed twice -> identical tokens and identical accepted-count trace (every split-K
    fix-up, attention merge and verify reduction runs in a fixed order); end-of-line every 17th image token,
    end-of-image after 16 rows; fewer forwards than tokens."""
    engine, dev = env["engine"], env["dev"]
    g = 16
    prompt = [1] + list(range(9000, 9040)) + [8197, 8804 + g // 2, 8804 + g // 2]
    ds = _lumina_slice(env, 2, 448)
    kw = dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=g * g + g - 10, max_num_new_tokens=32, guidance_scale=3.0,
              seed=11, multi_token_init_scheme="random", do_cfg=True, prefix_token_sampler_scheme="speculative_jacobi")
    runs = []
    for _ in range(2):
        eng = engine.SJDEngine(ds, engine.SJDParams(**kw), engine.LuminaGrammarState(), torch.arange(4, 8196))
        ids = eng.generate(prompt, max_length=len(prompt) + g * (g + 1) + 2, eos_token_ids=[8710],
                           kv_lo=[0, len(prompt) - 1], collect_trace=True)
        runs.append((ids, list(eng.stats.trace), eng.stats.nfe))
    assert runs[0] == runs[1], "same seed, different result: some reduction order is not fixed"
    img = runs[0][0][len(prompt):]
    assert all(img[i] == 8803 for i in range(g, g * (g + 1), g + 1)) and img[g * (g + 1)] == 8196
    assert all(4 <= t < 8196 for i, t in enumerate(img[:g * (g + 1)]) if i % (g + 1) != g)
    assert runs[0][2] < len(img)
    ds.close()


def test_full_width_window_logits_do_not_depend_on_the_window(env):
    """The logits of a token must not depend on how many draft tokens share its forward (the GEMM's per-element
    accumulation order is fixed by the stream-K partition, not by the number of token rows; the attention of a query
    only sees keys up to itself): position i of a 32-token window == the same token fed in a 1-token window after the
    first i were cached, bit for bit in the GEMMs and to fp32 rounding of the attention merge (key splits differ)."""
    dev = env["dev"]
    ds = _lumina_slice(env, 2, 256)
    V = ds.shape.vocab
    gen = torch.Generator().manual_seed(4)
    P, W = 40, 32
    pre = torch.randint(4, 8196, (2, P), generator=gen).int().to(dev)
    win = torch.randint(4, 8196, (1, W), generator=gen).int().repeat(2, 1).to(dev)
    kv_lo = [0, 7]

    def fwd(ids2, kv_len):
        Wn = ids2.shape[1]
        pos = torch.arange(kv_len, kv_len + Wn, dtype=torch.int32, device=dev)
        rope = torch.cat([(pos - kv_lo[b]).clamp(min=0) for b in range(2)]).int().contiguous()
        return ds.forward(Wn, rope, pos.repeat(2).contiguous(), kv_len, kv_lo, ids=ids2.flatten().contiguous(),
                          n_logit_tokens=Wn).clone()

    fwd(pre, 0)
    big = fwd(win, P)                                   # [2, W, V]
    fwd(pre, 0)
    for i in range(W):
        one = fwd(win[:, i:i + 1].contiguous(), P + i)  # caches token i, so the next step sees it
        d = (one[:, 0] - big[:, i]).abs().max().item()
        assert d <= 2.0 ** -6, (i, d)                   # bf16 logits of magnitude ~4: one ulp is 2^-6 .. 2^-5
    # BASELINE config 5's window sweep at full width: the first W positions of a 128-token draft row must not care whether
    # the window is 8, 16, 32, 64 or 128 wide (attention_sw.cu up to 64, attention_tc.cu at 128; 16 .. 256 GEMM rows)
    win128 = torch.randint(4, 8196, (1, 128), generator=gen).int().repeat(2, 1).to(dev)
    big128 = fwd(win128, P)
    ulp = 2.0 ** (torch.floor(torch.log2(big128.abs().max())).item() - 7)
    for Ws in (8, 16, 32, 64):
        part = fwd(win128[:, :Ws].contiguous(), P)      # roll-back is just the smaller kv_len
        d = (part - big128[:, :Ws]).abs()
        # different attention kernels round P to bf16 against different references (the window-128 kernel against the exact
        # running row maximum, attention_sw.cu against an integer bound of an 8-column group): two valid bf16 pipelines
        # with independent roundings, a few ulp apart at worst, a fraction of an ulp on average (measured 0.26 ulp; both
        # stay inside the 0.5 ulp mean bound against the bf16-emulating oracle, tests/test_gpu_baseline_sizes.py)
        assert d.max().item() <= 4.0 * ulp and d.mean().item() <= 0.35 * ulp, (Ws, d.max().item(), d.mean().item(), ulp)
    torch.cuda.synchronize()
    ds.close()


# ------------------------------------------------------------------------- edges: long prompts, capacity, bad arguments
def test_chunked_prefill_equals_one_pass_reference(env):
    """A prompt longer than one forward call holds (rows * tokens <= 256) is prefilled in chunks through the same
    kernels; the engine's first generated token and the cache it leaves behind must be what a single pass over the
    whole prompt gives: last-position logits vs the oracle stack's one-pass logits, then one more window on top."""
    RF, model, engine, dev = env["RF"], env["model"], env["engine"], env["dev"]
    cfg, (cos, sin), _ = _family(env, "chameleon")
    w = RF.random_weights(cfg, seed=7, device=dev)
    shape = model.StackShape(cfg.n_layers, cfg.d_model, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.d_ff,
                             cfg.vocab, cfg.rms_eps, cfg.qk_norm, cfg.rope_interleaved)
    P, rows, max_len = 300, 2, 448
    ds = model.DeviceStack(shape, w, cos, sin, rows, max_len, dev)
    ref = RF.RefStack(cfg, w, cos.to(dev), sin.to(dev), rows, max_len, emulate_bf16=True)
    g = torch.Generator().manual_seed(9)
    prompt = torch.randint(8900, 9200, (P,), generator=g).tolist()
    kv_lo = [0, P - 1]
    captured = []

    class Eng(engine.SJDEngine):
        def _forward(self, row_tokens, kv_len, kv_lo_, n_logit, embeds=None):
            lg = super()._forward(row_tokens, kv_len, kv_lo_, n_logit, embeds)
            captured.append((len(row_tokens[0]), kv_len, lg.detach().float().clone()))
            return lg

    kw = dict(jacobi_loop_interval_l=1, jacobi_loop_interval_r=64, max_num_new_tokens=8, guidance_scale=3.0, seed=0,
              multi_token_init_scheme="random", do_cfg=True, prefix_token_sampler_scheme="speculative_jacobi")
    eng = Eng(ds, engine.SJDParams(**kw), engine.PlainTopKState(top_k=50), torch.arange(4, 8196))
    eng.generate(prompt, max_length=P + 10, kv_lo=kv_lo)
    chunks = [c for c in captured if c[1] < P]
    assert [c[0] for c in chunks] == [128, 128, 44] and [c[1] for c in chunks] == [0, 128, 256]
    ids = torch.tensor([prompt, prompt], device=dev)
    pos = torch.arange(P, device=dev)[None].repeat(rows, 1)
    rope = torch.stack([(pos[b] - kv_lo[b]).clamp(min=0) for b in range(rows)])
    lr = ref.forward(ids=ids, rope_pos=rope, kv_len=0, kv_lo=kv_lo, cache_pos=pos, n_logit_tokens=1)
    ulp = 2.0 ** (torch.floor(torch.log2(lr.abs().max())).item() - 7)
    assert (chunks[-1][2].view(rows, -1) - lr.view(rows, -1)).abs().max().item() <= 2.5 * ulp
    # the first Jacobi window after the prefill reads the whole chunk-built cache
    Wn, kv, lg = captured[len(chunks)]
    assert kv == P
    ds.close()


def test_capacity_and_argument_errors_are_loud(env):
    """KV cache overflow, too many token rows, missing noise: negative status from the C ABI -> RuntimeError in Python,
    never a silent truncation."""
    RF, model, engine, dev, lib = env["RF"], env["model"], env["engine"], env["dev"], env["lib"]
    cfg, (cos, sin), _ = _family(env, "chameleon")
    w = RF.random_weights(cfg, seed=2, device=dev)
    shape = model.StackShape(cfg.n_layers, cfg.d_model, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.d_ff,
                             cfg.vocab, cfg.rms_eps, cfg.qk_norm, cfg.rope_interleaved)
    ds = model.DeviceStack(shape, w, cos, sin, 2, 64, dev)
    W = 16
    ids = torch.zeros(2 * W, dtype=torch.int32, device=dev)
    pos = torch.arange(W, dtype=torch.int32, device=dev).repeat(2)
    with pytest.raises(RuntimeError, match="overflow"):
        ds.forward(W, pos, pos, 56, [0, 0], ids=ids, n_logit_tokens=W)          # 56 + 16 > 64 slots
    big = torch.zeros(2 * 129, dtype=torch.int32, device=dev)
    with pytest.raises(RuntimeError, match="out of range"):
        ds.forward(129, big, big, 0, [0, 0], ids=big, n_logit_tokens=1)         # 258 token rows > 256
    eng = engine.SJDEngine(ds, engine.SJDParams(max_num_new_tokens=8, jacobi_loop_interval_r=40, seed=0),
                           engine.PlainTopKState(top_k=10), torch.arange(4, 8196))
    with pytest.raises(RuntimeError, match="KV cache too small"):
        eng.generate([1, 2, 3], max_length=200, kv_lo=[0, 2])
    logits = torch.zeros(2 * 4, 128, device=dev)
    desc = {"allow": None, "forced": [-1] * 4, "top_k": 0}
    with pytest.raises(RuntimeError, match="noise_e1"):
        engine.verify_call(logits, 4, 128, desc, torch.zeros(4, dtype=torch.int32, device=dev), None, None,
                           has_uncond=True, apply_cfg=True, guidance=3.0, temperature=1.0, do_sample=True, scheme=1)
    ds.close()


def test_full_size_config2_image_is_deterministic_and_grammatical(env):
    """BASELINE config 2 at FULL size: Lumina-mGPT-7B shape (32 layers, 13.5 GB of weights), one 768 x 768 image
    (48 rows of 48 tokens + end-of-line, end-of-image), window 32, cfg 3.0, top-k 2000, fixed seed — decoded twice.
    Identical token streams and accepted-count traces (bit-reproducible kernels), the grammar exactly in place over all
    2 353 image-region tokens, accepted tokens per forward > 1, and the KV cache never rewound below the accepted prefix."""
    from sjd_b200 import families
    engine, model, dev = env["engine"], env["model"], env["dev"]
    shape = families.lumina_7b()
    w = families.random_weights(shape, seed=0, device=dev)
    cos, sin = families.rope_rotate_half(128, 2560, 10000.0, True)
    ds = model.DeviceStack(shape, w, cos, sin, 2, 2560, dev)
    del w
    torch.cuda.empty_cache()
    g = 48
    gen = torch.Generator().manual_seed(1000)
    prompt = torch.randint(8900, 65000, (64,), generator=gen).tolist() + [8197, 8804 + g // 2, 8804 + g // 2]
    kw = dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=g * g + g - 10, max_num_new_tokens=32, guidance_scale=3.0,
              seed=0, multi_token_init_scheme="random", do_cfg=True, prefix_token_sampler_scheme="speculative_jacobi")
    runs = []
    for _ in range(2):
        eng = engine.SJDEngine(ds, engine.SJDParams(**kw), engine.LuminaGrammarState(image_top_k=2000), torch.arange(4, 8196))
        ids = eng.generate(prompt, max_length=len(prompt) + g * (g + 1) + 2, eos_token_ids=[8710],
                           kv_lo=[0, len(prompt) - 1], collect_trace=True)
        runs.append((ids, list(eng.stats.trace), eng.stats.nfe))
    assert runs[0] == runs[1]
    ids, trace, nfe = runs[0]
    img = ids[len(prompt):]
    n = g * (g + 1)
    assert len(img) >= n + 1 and img[n] == 8196
    assert all(img[i] == 8803 for i in range(g, n, g + 1))
    assert all(4 <= t < 8196 for i, t in enumerate(img[:n]) if i % (g + 1) != g)
    assert sum(t[1] for t in trace) == len(img) and nfe < len(img)
    assert all(1 <= t[1] <= t[0] for t in trace), "accepted count outside [1, window]"
    ds.close()


def test_config4_width_long_cache_kernels_agree(env):
    """BASELINE config 4 at full width (Emu3-Gen: 32 query / 8 kv heads, d_ff 14 336, V = 184 622; 2 layers deep) over a
    4 096-token cache: the same 64 draft tokens fed as one window of 64 (mma.sync attention, 256 stacked rows per kv
    head), as two windows of 32 (tcgen05 attention, 128 rows) and one by one (window 1) must give the same logits to
    bf16 rounding — three attention kernels / split plans, two GEMM row counts, one answer.  Then the wide verify:
    tokens restricted to the 32 768 visual ids."""
    from sjd_b200 import families
    model, engine, dev = env["model"], env["engine"], env["dev"]
    shape = families.emu3_gen()
    shape.n_layers = 2
    w = families.random_weights(shape, seed=5, std=0.03, device=dev)
    max_len = 4352
    cos, sin = families.rope_rotate_half(128, max_len, 1e6, True)
    ds = model.DeviceStack(shape, w, cos, sin, 2, max_len, dev)
    del w
    V, L, kv_lo = shape.vocab, 4096, [0, 16]
    gen = torch.Generator().manual_seed(8)

    def fwd(ids2, kv_len):
        Wn = ids2.shape[1]
        pos = torch.arange(kv_len, kv_len + Wn, dtype=torch.int32, device=dev)
        rope = torch.cat([(pos - kv_lo[b]).clamp(min=0) for b in range(2)]).int().contiguous()
        return ds.forward(Wn, rope, pos.repeat(2).contiguous(), kv_len, kv_lo, ids=ids2.flatten().contiguous(),
                          n_logit_tokens=Wn).clone()

    for s in range(0, L, 128):   # fill the cache
        fwd(torch.randint(0, V, (2, 128), generator=gen).int().to(dev), s)
    win = torch.randint(151854, 151854 + 32768, (1, 64), generator=gen).int().repeat(2, 1).to(dev)
    big = fwd(win, L)                                                     # [2, 64, V], window 64
    halves = torch.cat([fwd(win[:, :32].contiguous(), L), fwd(win[:, 32:].contiguous(), L + 32)], 1)
    ulp = 2.0 ** (torch.floor(torch.log2(big.abs().max())).item() - 7)
    assert (halves - big).abs().max().item() <= 2.0 * ulp, ((halves - big).abs().max().item(), ulp)
    for i in (0, 1, 31, 32, 63):                                          # the cache holds win[:, :64] from the calls above
        one = fwd(win[:, i:i + 1].contiguous(), L + i)
        assert (one[:, 0] - big[:, i]).abs().max().item() <= 2.0 * ulp, i
    # verify over the 184 622-wide vocabulary with Emu3's 32 768-id visual range (strided, range-restricted path)
    W = 64
    desc = {"allow": (151854, 151854 + 32768), "forced": [-1] * W, "top_k": 2048}
    e1 = torch.empty(W, V, device=dev).exponential_(generator=torch.Generator(dev).manual_seed(1))
    out = engine.verify_call(big.view(-1, V), W, V, desc, win[0].contiguous(), None, None, has_uncond=True, apply_cfg=True,
                             guidance=3.0, temperature=1.0, do_sample=True, scheme=1, noise_e1=e1)
    toks = out["tokens"].cpu()
    p = out["p"]
    assert bool(((toks >= 151854) & (toks < 151854 + 32768)).all())   # (ties with the k-th logit are kept: bf16 logits tie often)
    assert bool(((p > 0).sum(1) <= 2048 + 128).all()) and torch.allclose(p.sum(1), torch.ones(W, device=dev), atol=1e-4)
    assert float(p[:, :151854].sum()) == 0.0 and float(p[:, 151854 + 32768:].sum()) == 0.0
    ds.close()


# ------------------------------------------------------------------------------------------ device-side noise (round 2)
@pytest.mark.parametrize("shape", [(32, 65536), (1, 65536), (8, 9216), (64, 184622), (1, 184622), (3, 1000), (16, 16384)])
def test_device_philox_matches_torch_generator(env, shape):
    """sjd_verify's rng_mode = 1 must draw, element for element and bit for bit, what the reference's seeded
    torch.Generator(device='cuda') writes into the noise tensors: exponential_ [W, V] (torch.multinomial), rand [1, W, V]
    (accept test), exponential_ [1, V] (residual multinomial) — in that order from one generator, with the offset
    bookkeeping engine.PhiloxNoise does on the host."""
    L, _lib, engine, dev = env["lib"], env["_lib"], env["engine"], env["dev"]
    W, V = shape
    seed = 1234 + W
    g = torch.Generator(dev).manual_seed(seed)
    ph = engine.PhiloxNoise(seed, dev)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(2):   # two trips: the offsets must keep tracking torch's generator
        for kind, numel, make in ((0, W * V, lambda: torch.empty(W, V, device=dev).exponential_(1.0, generator=g)),
                                  (1, W * V, lambda: torch.rand(1, W, V, device=dev, generator=g)),
                                  (0, V, lambda: torch.empty(1, V, device=dev).exponential_(1.0, generator=g))):
            ref = make().flatten()
            off, span = ph.draw(numel)
            out = torch.empty(numel, device=dev)
            _lib.check(L.sjd_debug_philox(out.data_ptr(), numel, seed, off, span, kind, st), "philox")
            torch.cuda.synchronize()
            if not torch.equal(out, ref):
                bad = (out != ref).nonzero().flatten()
                # which transform variant WOULD match (developer aid; see philox_fill_kernel)
                hits = []
                for var in ([16, 17, 18, 24, 25, 26] if kind == 0 else [32, 40]):
                    o2 = torch.empty(numel, device=dev)
                    L.sjd_debug_philox(o2.data_ptr(), numel, seed, off, span, var, st)
                    torch.cuda.synchronize()
                    hits.append((var, int((o2 != ref).sum())))
                raise AssertionError(f"kind {kind} numel {numel} rep {rep}: {bad.numel()} of {numel} differ, first at "
                                     f"{int(bad[0])}: {float(out[bad[0]])!r} vs {float(ref[bad[0]])!r}; mismatches per "
                                     f"variant {hits}")
        assert ph.offset == g.get_offset(), (ph.offset, g.get_offset())
