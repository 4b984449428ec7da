"""Short profiling target (ncu-friendly): N Jacobi trips of the bench workload — Lumina-7B shapes, window W,
KV cache pre-set to L keys — through the same C-ABI calls the engine makes (sjd_ctx_forward + sjd_verify).
    python scripts/profile_trips.py [--trips 3] [--window 32] [--kv-len 1200] [--time]
With --time it prints CUDA-event timings per trip (forward / verify / noise) instead of being a profile target.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sjd_b200  # noqa
from sjd_b200 import _lib, engine, families, model

ap = argparse.ArgumentParser()
ap.add_argument("--trips", type=int, default=3)
ap.add_argument("--window", type=int, default=32)
ap.add_argument("--kv-len", type=int, default=1200)
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--time", action="store_true")
ap.add_argument("--host-noise", action="store_true", help="round-1 form: torch generator kernels fill the noise tensors")
args = ap.parse_args()

dev = torch.device("cuda:0")
shape = families.lumina_7b()
shape.n_layers = args.layers
W, V, L = args.window, shape.vocab, args.kv_len
weights = families.random_weights(shape, seed=0, device=dev)
cos, sin = families.rope_rotate_half(shape.head_dim, 2560, 10000.0, True)
stack = model.DeviceStack(shape, weights, cos, sin, rows=2, max_len=2560, device=dev)
del weights
P = 67
g = torch.Generator().manual_seed(0)
ids = torch.randint(4, 8196, (2 * W,), generator=g, dtype=torch.int32).to(dev)
pos = torch.arange(L, L + W, dtype=torch.int32)
rope = torch.cat([pos, pos - (P - 1)]).to(dev)
cpos = torch.cat([pos, pos]).to(dev)
desc = {"allow": (4, 8196), "forced": [-1] * W, "top_k": 2000}
desc["forced"][W // 2] = 8803
p_prev = torch.softmax(torch.randn(W, V, device=dev), -1)
draft = ids[:W].clone()
q_row = torch.tensor([-1] + list(range(1, W // 2)) + [-1] * (W - W // 2), dtype=torch.int32, device=dev)
gen = torch.Generator(dev).manual_seed(0)
philox = engine.PhiloxNoise(0, dev)


def trip(ev=None):
    if ev: ev[0].record()
    logits = stack.forward(W, rope, cpos, L, [0, P - 1], ids=ids, n_logit_tokens=W)
    if ev: ev[1].record()
    e1 = u = e2 = None
    if args.host_noise:
        e1 = torch.empty((W, V), device=dev).exponential_(1.0, generator=gen)
        u = torch.rand((1, W, V), device=dev, generator=gen)[0].gather(1, draft.long()[:, None]).squeeze(1).contiguous()
        e2 = torch.empty((1, V), device=dev).exponential_(1.0, generator=gen)
    if ev: ev[2].record()
    out = engine.verify_call(logits, W, V, desc, draft, q_row, p_prev, has_uncond=True, apply_cfg=True, guidance=3.0,
                             temperature=1.0, do_sample=True, scheme=0, noise_e1=e1, noise_u=u, noise_e2=e2,
                             eoi_token=8196, text_top_k=10, sync=True, rng=None if args.host_noise else philox)
    if ev: ev[3].record()
    return out


if args.time:
    for _ in range(3):
        trip()
    torch.cuda.synchronize()
    tot = [0.0, 0.0, 0.0]
    n = max(args.trips, 10)
    for _ in range(n):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        trip(ev)
        torch.cuda.synchronize()
        for i in range(3):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
    print(f"W={W} L={L} layers={args.layers}: forward {tot[0]/n:.3f} ms  noise {tot[1]/n:.3f} ms  verify {tot[2]/n:.3f} ms")
else:
    for _ in range(args.trips):
        trip()
    torch.cuda.synchronize()
print("done")
