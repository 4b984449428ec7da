"""Developer timing: where does a fused GEMM spend its time?  clock64 stamps of the epilogue stages per CTA."""
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
    t0 = t[:, 0]
    def col(j):
        v = t[:, j] - t0
        v = v[t[:, j] > 0]
        return (f"{v.mean():6.1f}/{v.max():6.1f}" if len(v) else "   -  /   -  ")
    nm = names[i % 4] if i < n - 1 else "lm_head"
    print(f"{nm:8s} mean/max us: last-acc {col(2)} parked {col(3)} | t0-arrived {col(8)} t0-share-done {col(9)} | fixup-done {col(4)} stats-in {col(5)} xn-written {col(10)} end {col(6)} | share: setup {col(7)} loads-done {col(14)} apply-done {col(15)} | probe L2 load cycles mean/max {b[i,:148,13].mean()*1.965e3:.0f}/{b[i,:148,13].max()*1.965e3:.0f}")
