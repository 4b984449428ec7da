"""Developer timing of the tcgen05 attention pipeline (CTA 0, clock64 stamps per unit; run with SJD_ATTN=tc)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sjd_b200  # noqa
from sjd_b200 import _lib, families, model

dev = torch.device("cuda:0")
shape = families.lumina_7b(); shape.n_layers = 2
w = families.random_weights(shape, seed=0, device=dev)
cos, sin = families.rope_rotate_half(128, 2560, 10000.0, True)
st = model.DeviceStack(shape, w, cos, sin, rows=2, max_len=2560, device=dev)
lib = _lib.lib()
W, L, P = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 1200, 67
ids = torch.randint(4, 8196, (2 * W,), dtype=torch.int32).to(dev)
pos = torch.arange(L, L + W, dtype=torch.int32)
rope = torch.cat([pos, pos - (P - 1)]).to(dev); cpos = torch.cat([pos, pos]).to(dev)
for _ in range(3):
    st.forward(W, rope, cpos, L, [0, P - 1], ids=ids, n_logit_tokens=W)
buf = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
lib.sjd_debug_attn_stamps(buf.data_ptr())
st.forward(W, rope, cpos, L, [0, P - 1], ids=ids, n_logit_tokens=W)
torch.cuda.synchronize()
lib.sjd_debug_attn_stamps(None)
b = buf.cpu().view(8, 16)
t0 = int(b[0, 15])
names = ["loads issued", "QK landed", "S issued", "PV inputs ready", "PV issued", "S seen", "max done", "P arrived", "O seen", "epi done"]
for n in range(8):
    if int(b[n, 0]) == 0:
        break
    print(f"unit {n}: " + "  ".join(f"{names[k]} {(int(b[n, k]) - t0) / 1.9e3:6.2f}" for k in range(10)))
