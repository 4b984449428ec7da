"""Developer timing (round 2): are the chain kernel's stragglers the SAME CTAs every time?  Per op kind, the per-CTA
streaming time (A - t0: op entered -> last accumulator ready) of consecutive layers is correlated, and the slowest /
fastest CTAs are listed.  A stable pattern would justify a weighted stream-K partition."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sjd_b200  # noqa
from sjd_b200 import _lib, families, model
dev = torch.device("cuda:0")
shape = families.lumina_7b(); shape.n_layers = 6
w = families.random_weights(shape, seed=0, device=dev)
cos, sin = families.rope_rotate_half(128, 2560, 10000.0, True)
st = model.DeviceStack(shape, w, cos, sin, rows=2, max_len=2560, device=dev)
lib = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for _ in range(3): lib.sjd_ctx_gemm_only(st.ctx, W, s)
n = 4 * shape.n_layers + 1
runs = []
for rep in range(2):
    buf = torch.zeros(n, 256, 16, dtype=torch.int64, device=dev)
    lib.sjd_debug_gemm_stamps(buf.data_ptr(), n)
    lib.sjd_ctx_gemm_only(st.ctx, W, s)
    torch.cuda.synchronize()
    lib.sjd_debug_gemm_stamps(None, 0)
    runs.append(buf.cpu().double()[:, :148] / 1.965e3)
names = ["qkv", "o", "gate_up", "down"]
for k, nm in enumerate(names):
    for what, (a, c) in (("stream A-t0", (2, 0)), ("op E-t0", (6, 0))):
        vecs = []
        for b in runs:
            for l in range(1, shape.n_layers):
                t = b[4 * l + k]
                vecs.append(t[:, a] - t[:, c])
        V = torch.stack(vecs)                      # [runs * layers, 148]
        Vc = V - V.mean(1, keepdim=True)
        C = (Vc @ Vc.T) / (Vc.norm(dim=1)[:, None] * Vc.norm(dim=1)[None, :])
        off = C[~torch.eye(len(vecs), dtype=torch.bool)]
        m = V.mean(0)
        order = torch.argsort(m)
        print(f"{nm:8s} {what}: mean {V.mean():.1f} us, per-launch spread (max-mean) {float((V.max(1).values - V.mean(1)).mean()):.1f} us, "
              f"correlation of the per-CTA pattern across layers/runs mean {off.mean():.2f} min {off.min():.2f}; "
              f"CTA means: slowest {[(int(i), round(float(m[i]), 1)) for i in order[-5:]]} fastest {[(int(i), round(float(m[i]), 1)) for i in order[:5]]}")
