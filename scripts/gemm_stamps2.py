"""Developer timing (round 2): per-op timeline of the chain kernel from the epilogue warps' clock64 stamps.
Per CTA and op: t0 = the epilogue warps enter the op, A = accumulator of the LAST segment ready (the MMA is done),
P = all segments drained / parked, S = every partial of the op parked (row owners only), X = rows written (row owners),
E = the epilogue warps leave the op.  Prints mean / max over CTAs of A-t0, E-A (the CTA's tail) and E-t0."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sjd_b200  # noqa
from sjd_b200 import _lib, families, model
dev = torch.device("cuda:0")
shape = families.lumina_7b(); shape.n_layers = 4
w = families.random_weights(shape, seed=0, device=dev)
cos, sin = families.rope_rotate_half(128, 2560, 10000.0, True)
st = model.DeviceStack(shape, w, cos, sin, rows=2, max_len=2560, device=dev)
lib = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for _ in range(3): lib.sjd_ctx_gemm_only(st.ctx, W, s)
n = 4 * shape.n_layers + 1
buf = torch.zeros(n, 256, 16, dtype=torch.int64, device=dev)
lib.sjd_debug_gemm_stamps(buf.data_ptr(), n)
lib.sjd_ctx_gemm_only(st.ctx, W, s)
torch.cuda.synchronize()
lib.sjd_debug_gemm_stamps(None, 0)
b = buf.cpu().double() / 1.965e3   # us at 1965 MHz
names = ["qkv", "o", "gate_up", "down"]
for i in range(n):
    t = b[i, :148]
    def d(a, c):
        ok = (t[:, a] > 0) & (t[:, c] > 0)
        v = (t[:, a] - t[:, c])[ok]
        return f"{v.mean():6.1f}/{v.max():6.1f}" if len(v) else "   -  /   -  "
    nm = names[i % 4] if i < n - 1 else "lm_head"
    print(f"{nm:8s} mean/max us: MMA done (A-t0) {d(2,0)} | parked (P-A) {d(3,2)} | all parked seen (S-P) {d(5,3)} | rows written (X-S) {d(10,5)} "
          f"| tail E-A {d(6,2)} | op E-t0 {d(6,0)}")
# op-to-op gap inside a CTA: t0 of op i+1 minus E of op i is zero by construction (same warps); the wait shows up as A-t0 of
# the next op exceeding its streaming time
