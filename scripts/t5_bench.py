"""Measurement for SURVEY §8 f4: the flan-t5-xl-shaped text encoder (24 layers, d 2048, 32 heads x 64, gated d_ff 5120; 2
captions x 120 tokens = 240 token rows, random-init bf16 weights) — HF's T5EncoderModel (eager, bf16, cuBLAS) against
sjd_b200.t5_encoder.T5EncoderB200 (every linear layer on sjd_gemm_bf16), CUDA events."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sjd_b200  # noqa
from sjd_b200 import t5_encoder
from transformers import T5Config, T5EncoderModel

dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = T5Config(d_model=2048, d_kv=64, num_heads=32, d_ff=5120, num_layers=24, vocab_size=32128,
               feed_forward_proj="gated-gelu", dropout_rate=0.0)
m = T5EncoderModel(cfg).eval().to(dev).to(torch.bfloat16)
ids = torch.randint(0, cfg.vocab_size, (2, 120), device=dev)
mask = torch.ones(2, 120, dtype=torch.long, device=dev); mask[1, 40:] = 0
enc = t5_encoder.T5EncoderB200.from_module(m)


def ev_time(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    t_hf = ev_time(lambda: m(input_ids=ids, attention_mask=mask))
    t_eager = ev_time(lambda: enc.forward(ids, mask, graph=False))
    t_us = ev_time(lambda: enc.forward(ids, mask))
    ref = m(input_ids=ids, attention_mask=mask)["last_hidden_state"].float()
    out = enc.forward(ids, mask)
wbytes = 24 * (4 * 2048 * 2048 + 3 * 5120 * 2048) * 2
d = (out - ref).abs()[mask.bool()]
print(f"flan-t5-xl-shaped encoder, 240 token rows: HF eager bf16 {t_hf:.2f} ms | T5EncoderB200 launch by launch {t_eager:.2f} ms, as a CUDA graph {t_us:.2f} ms "
      f"({wbytes / t_us / 1e6:.0f} GB/s of weights; HBM roofline {wbytes / 6457.4e6:.2f} ms) | "
      f"max |diff| vs HF bf16 {d.max().item():.3f}, mean {d.mean().item():.4f} (output scale {ref.abs().max().item():.1f})")
