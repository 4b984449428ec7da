"""Developer timing of the forward chain on Lumina-7B shapes: GEMM-only chain vs full forward, per window size."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sjd_b200  # noqa
from sjd_b200 import _lib, families, model

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
shape = families.lumina_7b(); shape.n_layers = layers
w = families.random_weights(shape, seed=0, device=dev)
cos, sin = families.rope_rotate_half(128, 2560, 10000.0, True)
st = model.DeviceStack(shape, w, cos, sin, rows=2, max_len=2560, device=dev)
del w
lib = _lib.lib()
s = torch.cuda.current_stream().cuda_stream
def ev_time(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
Ws = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [32, 16, 64]
for W in Ws:
    L, P = int(os.environ.get("SJD_BENCH_L", "1200")), 67
    ids = torch.randint(4, 8196, (2 * W,), dtype=torch.int32).to(dev)
    pos = torch.arange(L, L + W, dtype=torch.int32)
    rope = torch.cat([pos, pos - (P - 1)]).to(dev); cpos = torch.cat([pos, pos]).to(dev)
    tg = ev_time(lambda: lib.sjd_ctx_gemm_only(st.ctx, W, s))
    tf = ev_time(lambda: st.forward(W, rope, cpos, L, [0, P - 1], ids=ids, n_logit_tokens=W))
    ng = 4 * layers + 1
    wb = (shape.n_layers * (4 * 4096 * 4096 + 3 * 4096 * 11008) + 4096 * 65536) * 2
    print(f"W={W} layers={layers}: gemm-only chain {tg:.3f} ms ({tg/ng*1e3:.1f} us/gemm, {wb/tg/1e6:.0f} GB/s weights) | full forward {tf:.3f} ms "
          f"(attention+rest {(tf-tg)/layers*1e3:.1f} us/layer)", flush=True)
