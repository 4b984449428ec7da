"""Developer check of the tcgen05 attention: window forwards of the three tiny families against the oracle stack,
printing after every step (run with SJD_ATTN=tc under `timeout`)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import sjd_b200  # noqa
from sjd_b200 import model
from oracle import ref_forward as RF

dev = torch.device("cuda:0")
fams = {
    "chameleon": (RF.StackConfig(2, 256, 2, 2, 128, 512, 9216, 1e-5, qk_norm=True), RF.rope_tables_rotate_half(128, 512, 10000.0, True), [0, 36]),
    "llamagen": (RF.StackConfig(3, 256, 4, 4, 64, 768, 1024, 1e-5, rope_interleaved=True, family="llamagen"), RF.rope_tables_llamagen_2d(24, 64, 10000, 1), [0, 0]),
    "emu3": (RF.StackConfig(2, 512, 4, 1, 128, 1024, 5000, 1e-5, family="emu3", rope_theta=1e6), RF.rope_tables_rotate_half(128, 512, 1e6, True), [0, 5]),
}
which = sys.argv[1:] or list(fams)
for name in which:
    cfg, (cos, sin), kv_lo = fams[name]
    w = RF.random_weights(cfg, seed=1, device=dev)
    rows, max_len = 2, 320
    shape = model.StackShape(cfg.n_layers, cfg.d_model, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.d_ff, cfg.vocab,
                             cfg.rms_eps, cfg.qk_norm, cfg.rope_interleaved)
    ds = model.DeviceStack(shape, w, cos, sin, rows, max_len, dev)
    ref = RF.RefStack(cfg, w, cos.to(dev), sin.to(dev), rows, max_len, emulate_bf16=True)
    g = torch.Generator().manual_seed(5)
    kv_len = 0
    for step, W in enumerate([37, 1, 16, 16, 5, 128]):
        if step == 3:
            kv_len -= 9
        ids = torch.randint(0, cfg.vocab, (rows, W), generator=g).to(dev)
        pos = torch.arange(kv_len, kv_len + W, device=dev)[None].repeat(rows, 1)
        rope_pos = torch.stack([(pos[b] - kv_lo[b]).clamp(min=0) for b in range(rows)])
        n = 1 if step == 0 else W
        print(f"{name} step {step} W={W} kv_len={kv_len} ...", flush=True)
        lg = ds.forward(W, rope_pos.int().flatten().contiguous(), pos.int().flatten().contiguous(), kv_len, kv_lo,
                        ids=ids.int().flatten().contiguous(), n_logit_tokens=n).clone()
        torch.cuda.synchronize()
        lr = ref.forward(ids=ids, rope_pos=rope_pos, kv_len=kv_len, kv_lo=kv_lo, cache_pos=pos, n_logit_tokens=n)
        ulp = 2.0 ** (torch.floor(torch.log2(lr.abs().max())).item() - 7)
        d = (lg - lr).abs()
        print(f"   max err {d.max().item():.4f} mean {d.mean().item():.5f}  (ulp {ulp:.4f})  nan={bool(torch.isnan(lg).any())}", flush=True)
        kv_len += W
    ds.close()
