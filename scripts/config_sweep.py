"""Per-trip latency of the window forward (+ verify) on the shapes of BASELINE.json's other configs, at fixed cache
length — the weight-independent number: Lumina-7B W=32 at several L, Emu3-Gen W=64 at 4 096 cached image
tokens (config 4), Chameleon-7B window sweep 8..128 (config 5).  Prints one JSON line per point with the HBM
roofline time of the same trip (weights + K/V read + K/V write, MEASURED_PEAKS.json bandwidth)."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import sjd_b200  # noqa: E402,F401
from sjd_b200 import _lib, engine, families, model  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = float(peaks.get("hbm_gbs", 6457.4)) * 1e9
TC = float(peaks.get("bf16_tflops", 1721.9)) * 1e12


def ev_time(fn, n=8, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, shape, points, max_len, prompt=67, allow=(4, 8196), top_k=2000, rope=None):
    w = families.random_weights(shape, seed=0, device=dev)
    cos, sin = rope if rope is not None else families.rope_rotate_half(shape.head_dim, max_len, 10000.0, True)
    st = model.DeviceStack(shape, w, cos, sin, rows=2, max_len=max_len, device=dev)
    del w
    torch.cuda.empty_cache()
    n_w = shape.n_layers * ((shape.n_heads + 2 * shape.n_kv_heads) * shape.head_dim * shape.d_model
                            + shape.n_heads * shape.head_dim * shape.d_model + 3 * shape.d_model * shape.d_ff) \
        + shape.d_model * shape.vocab
    for W, L in points:
        M = 2 * W
        ids = torch.randint(allow[0], allow[1], (M,), dtype=torch.int32).to(dev)
        pos = torch.arange(L, L + W, dtype=torch.int32)
        rope = torch.cat([pos, pos - (prompt - 1)]).to(dev)
        cpos = torch.cat([pos, pos]).to(dev)
        logits = None

        def fwd():
            nonlocal logits
            logits = st.forward(W, rope, cpos, L, [0, prompt - 1], ids=ids, n_logit_tokens=W)

        t_f = ev_time(fwd)
        V = shape.vocab
        desc = {"allow": allow, "forced": [-1] * W, "top_k": top_k}
        draft = torch.randint(allow[0], allow[1], (W,), dtype=torch.int32, device=dev)
        q_row = torch.full((W,), -1, dtype=torch.int32, device=dev)
        e1 = torch.empty(W, V, device=dev).exponential_()
        u = torch.rand(W, device=dev)
        e2 = torch.empty(V, device=dev).exponential_()
        pc = torch.empty(W, V, device=dev)
        lg = logits.view(-1, V)

        def ver():
            engine.verify_call(lg, W, V, desc, draft, q_row, pc, has_uncond=True, apply_cfg=True, guidance=3.0,
                               temperature=1.0, do_sample=True, scheme=0, noise_e1=e1, noise_u=u, noise_e2=e2,
                               p_cur=pc, sync=False)

        t_v = ev_time(ver)
        kv = shape.n_layers * 2 * 2 * shape.n_kv_heads * shape.head_dim * 2 * (L + W)     # K and V, 2 rows, bf16
        bytes_ = 2 * n_w + kv + M * V * 4
        flops = 2 * n_w * M + 4 * shape.n_layers * M * (L + W / 2) * shape.n_heads * shape.head_dim
        t_roof = max(bytes_ / HBM, flops / TC) * 1e3
        print(json.dumps({"config": name, "window": W, "kv_len": L, "rows": M, "ms_forward": round(t_f, 3),
                          "ms_verify": round(t_v, 3), "ms_roofline": round(t_roof, 3),
                          "frac_of_roofline": round(t_roof / t_f, 3), "bound": "hbm" if bytes_ / HBM > flops / TC else "tensor",
                          "GB_per_trip": round(bytes_ / 1e9, 2)}), flush=True)
    st.close()
    del st
    torch.cuda.empty_cache()


which = sys.argv[1:] or ["lumina", "emu3"]
if "lumina" in which:
    run("lumina7b/chameleon7b", families.lumina_7b(),
        [(32, 256), (32, 1200), (32, 2400), (8, 600), (16, 600), (32, 600), (64, 600), (128, 600)], 2688)
if "emu3" in which:
    run("emu3-gen", families.emu3_gen(), [(64, 4096), (64, 8000), (32, 4096)], 8320, prompt=48,
        allow=(151854, 151854 + 32768), top_k=2048)
if "llamagen" in which:   # config 1: LlamaGen GPT-B class-conditional 256 x 256 (16 x 16 latent grid + 1 condition token), window 16
    run("llamagen-gpt-b", families.llamagen("GPT-B"), [(16, 64), (16, 200), (1, 200)], 320, prompt=1, allow=(0, 16384),
        top_k=1000, rope=families.rope_llamagen_2d(16, 64, 10000, 1))
