"""Developer timing: where a Jacobi iteration's time goes on the DEVICE timeline of the real engine loop (one 768x768
image, bench workload): forward kernels, verify kernel, and the gap between verify's end and the next forward's first
kernel (result read-back, host window preparation, staging copy, launch latency)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa
import sjd_b200  # noqa
from sjd_b200 import engine, families, model

dev = torch.device("cuda:0")
shape = families.lumina_7b()
P = bench.PROMPT_TEXT + 3
max_length = P + bench.IMG_TOKENS + 2
w = families.random_weights(shape, seed=0, device=dev)
cos, sin = families.rope_rotate_half(shape.head_dim, 2560, 10000.0, True)
stack = model.DeviceStack(shape, w, cos, sin, rows=2, max_len=2560, device=dev)
del w
grammar = engine.LuminaGrammarState(image_top_k=bench.TOP_K, text_top_k=10)
ev = []


class Eng(engine.SJDEngine):
    def _forward(self, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = super()._forward(*a, **k)
        e1.record()
        ev.append((e0, e1))
        return out


eng = Eng(stack, engine.SJDParams(**bench.sjd_params(bench.WINDOW, 0)), grammar, torch.arange(4, 8196))
for rep in range(2):
    ev.clear()
    eng.p.seed = rep
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids = eng.generate(bench.synthetic_prompt(rep), max_length=max_length, eos_token_ids=[8710], kv_lo=[0, P - 1])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
fwd = sum(a.elapsed_time(b) for a, b in ev)
between = sum(ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1))   # end of forward i -> start of forward i+1
n = len(ev)
print(f"{n} forwards, wall {wall * 1e3 / n:.3f} ms per iteration: forward (incl. its staging copy) {fwd / n:.3f} ms, "
      f"verify + read-back + host + next staging {between / (n - 1):.3f} ms")
