"""Developer GPU check (not a test): quick numerics + timing of each kernel against torch / the oracle."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sjd_b200  # noqa
from sjd_b200 import _lib
from sjd_b200.model import DeviceStack, StackShape
from oracle import ref_forward as RF
from oracle import sjd_oracle as O

dev = torch.device("cuda:0")
L = _lib.lib()
print("lib version", L.sjd_version(), "SMs", L.sjd_device_sm_count(), flush=True)
which = sys.argv[1:] or ["gemm", "verify", "forward"]


def gemm_case(N, K, M, time_it=False, grid=0):
    torch.manual_seed(N + K + M)
    m_tile = (M + 15) // 16 * 16
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    x = torch.zeros(max(m_tile, 16), K, device=dev, dtype=torch.bfloat16)
    x[:M] = (torch.randn(M, K, device=dev)).bfloat16()
    wsb = L.sjd_gemm_workspace_bytes(N, K, m_tile, grid)
    ws = torch.zeros(wsb, device=dev, dtype=torch.uint8)
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.sjd_gemm_bf16(w.data_ptr(), N, K, x.data_ptr(), x.shape[0], M, out.data_ptr(), 1, 0, ws.data_ptr(), grid, st), "gemm")
    torch.cuda.synchronize()
    ref = x[:M].float() @ w.float().T
    err = (out - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    msg = f"gemm N={N} K={K} M={M}: max_abs_err={err:.3e} rel={rel:.3e}"
    if time_it:
        # rotate over several weight copies so that HBM (not L2) is measured
        ncopy = max(2, int(400e6 // (N * K * 2)) + 1)
        ws_list = [(torch.randn(N, K, device=dev) * 0.05).bfloat16() for _ in range(ncopy)]
        for i in range(3):
            L.sjd_gemm_bf16(ws_list[i % ncopy].data_ptr(), N, K, x.data_ptr(), x.shape[0], M, out.data_ptr(), 1, 0, ws.data_ptr(), grid, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for i in range(iters):
            L.sjd_gemm_bf16(ws_list[i % ncopy].data_ptr(), N, K, x.data_ptr(), x.shape[0], M, out.data_ptr(), 1, 0, ws.data_ptr(), grid, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        gbs = N * K * 2 / ms / 1e6
        msg += f"  time={ms*1e3:.1f}us  weight-stream={gbs:.0f} GB/s"
        # cuBLAS for comparison
        xx = x[:M]
        for i in range(3):
            torch.matmul(xx, ws_list[i % ncopy].T)
        e0.record()
        for i in range(iters):
            torch.matmul(xx, ws_list[i % ncopy].T)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / iters
        msg += f"  | cuBLAS {ms2*1e3:.1f}us {N*K*2/ms2/1e6:.0f} GB/s"
    print(msg, flush=True)
    return rel


if "gemm" in which:
    for (N, K, M) in [(128, 64, 16), (256, 128, 16), (384, 256, 32), (1000, 256, 48), (2304, 768, 32),
                      (4096, 4096, 64), (4096, 4096, 256), (4096, 4096, 2)]:
        gemm_case(N, K, M)
    for (N, K, M) in [(12288, 4096, 64), (4096, 4096, 64), (22016, 4096, 64), (4096, 11008, 64), (65536, 4096, 64),
                      (22016, 4096, 32), (22016, 4096, 128), (22016, 4096, 256)]:
        gemm_case(N, K, M, time_it=True)

if "verify" in which:
    from sjd_b200.engine import verify_call  # thin wrapper used by the engine
    rng = np.random.default_rng(0)
    for case in range(6):
        W, V = 8, 9216
        logits = (rng.standard_normal((2 * W, V)) * 2).astype(np.float32)
        g = O.LuminaGrammar()
        ids = [1, 100, 200, 8197, 8808, 8808] + [int(x) for x in rng.integers(4, 8196, size=case * 3 + 2)]
        desc = g.describe(ids, W)
        draft = rng.integers(4, 8196, size=W).astype(np.int64)
        p_prev = O.softmax(O.topk_filter(O.apply_grammar((rng.standard_normal((W, V)) * 2).astype(np.float32),
                                                         {"allow": (4, 8196), "forced": [-1] * W}), 2000))
        q_rows = [None] * W
        q_row_idx = [-1] * W
        for i in range(1, W):
            if rng.random() < 0.6:
                q_rows[i] = p_prev[i]
                q_row_idx[i] = i
                # make the draft plausible under q so that some are accepted
                draft[i] = int(np.argmax(p_prev[i] / rng.exponential(size=V)))
        e1 = rng.exponential(size=(W, V)).astype(np.float32)
        u = rng.random(W).astype(np.float32) * (0.3 if case % 2 else 1.0)
        e2 = rng.exponential(size=V).astype(np.float32)
        ref = O.verify(logits, W, desc, draft, q_rows, has_uncond=True, apply_cfg=True, guidance=3.0,
                       noise_e1=e1, noise_u=u, noise_e2=e2)
        out = verify_call(torch.from_numpy(logits).to(dev), W, V, desc, torch.from_numpy(draft).int().to(dev),
                          torch.tensor(q_row_idx, dtype=torch.int32, device=dev), torch.from_numpy(p_prev).to(dev),
                          has_uncond=True, apply_cfg=True, guidance=3.0, temperature=1.0, do_sample=True, scheme=0,
                          noise_e1=torch.from_numpy(e1).to(dev), noise_u=torch.from_numpy(u).to(dev),
                          noise_e2=torch.from_numpy(e2).to(dev), eoi_token=8196, text_top_k=10)
        torch.cuda.synchronize()
        tok = out["tokens"].cpu().numpy()
        ok = (out["matched"] == ref.matched) and (tok == ref.tokens).all()
        perr = np.abs(out["p"].cpu().numpy() - ref.p).max()
        print(f"verify case {case}: matched gpu={out['matched']} ref={ref.matched} rejected={out['rejected']}/{ref.rejected} "
              f"tokens_equal={bool((tok == ref.tokens).all())} p_err={perr:.2e} forced={desc['forced']} {'OK' if ok else 'MISMATCH'}",
              flush=True)
        if not ok:
            print("  gpu", tok, "\n  ref", ref.tokens, "\n  nxt", ref.next_tokens, out["next_tokens"].cpu().numpy())

if "forward" in which:
    def run_family(name, cfg, rope, rows, kv_lo_fn):
        torch.manual_seed(0)
        w = RF.random_weights(cfg, seed=1, device=dev)
        cos, sin = rope
        max_len = 320
        shape = StackShape(cfg.n_layers, cfg.d_model, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.d_ff, cfg.vocab,
                           cfg.rms_eps, cfg.qk_norm, cfg.rope_interleaved)
        ds = DeviceStack(shape, w, cos, sin, rows, max_len, dev)
        ref = RF.RefStack(cfg, w, cos.to(dev), sin.to(dev), rows, max_len, emulate_bf16=True)
        ref32 = RF.RefStack(cfg, w, cos.to(dev), sin.to(dev), rows, max_len, emulate_bf16=False)
        g = torch.Generator().manual_seed(5)
        kv_len = 0
        for step, W in enumerate([37, 1, 16, 16, 5]):
            if step == 3:
                kv_len -= 9  # roll back 9 rejected drafts
            ids = torch.randint(0, cfg.vocab, (rows, W), generator=g).to(dev)
            kv_lo = kv_lo_fn(kv_len) if step else kv_lo_fn(0)
            pos = torch.arange(kv_len, kv_len + W, device=dev)[None].repeat(rows, 1)
            rope_pos = pos.clone()
            for b in range(rows):
                rope_pos[b] = (pos[b] - kv_lo[b]).clamp(min=0)
            n = W if step else 1
            lg = ds.forward(W, rope_pos.int().flatten().contiguous(), pos.int().flatten().contiguous(), kv_len, kv_lo,
                            ids=ids.int().flatten().contiguous(), n_logit_tokens=n).clone()
            lr = ref.forward(ids=ids, rope_pos=rope_pos, kv_len=kv_len, kv_lo=kv_lo, cache_pos=pos, n_logit_tokens=n)
            l32 = ref32.forward(ids=ids, rope_pos=rope_pos, kv_len=kv_len, kv_lo=kv_lo, cache_pos=pos, n_logit_tokens=n)
            torch.cuda.synchronize()
            e_emul = (lg - lr).abs().max().item()
            e_32 = (lg - l32).abs().max().item()
            e_ref = (lr - l32).abs().max().item()
            print(f"{name} step{step} W={W} kv_len={kv_len} kv_lo={kv_lo}: |gpu-ref_bf16emul|={e_emul:.3e} |gpu-ref_fp32|={e_32:.3e} "
                  f"|ref_bf16emul-ref_fp32|={e_ref:.3e} logits_absmax={l32.abs().max().item():.3f}", flush=True)
            kv_len += W
        ds.close()

    cham = RF.StackConfig(2, 256, 2, 2, 128, 512, 9216, 1e-5, qk_norm=True)
    run_family("chameleon", cham, RF.rope_tables_rotate_half(128, 512, 10000.0, True), 2, lambda kv: [0, 36])
    lg = RF.StackConfig(3, 256, 4, 4, 64, 768, 1024, 1e-5, rope_interleaved=True, family="llamagen")
    run_family("llamagen", lg, RF.rope_tables_llamagen_2d(24, 64, 10000, 1), 2, lambda kv: [0, 0])
    emu = RF.StackConfig(2, 512, 4, 1, 128, 1024, 5000, 1e-5, family="emu3", rope_theta=1e6)
    run_family("emu3", emu, RF.rope_tables_rotate_half(128, 512, 1e6, True), 2, lambda kv: [0, 5])
print("done")
