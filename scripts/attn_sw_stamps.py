"""Developer timing of attention_sw.cu (CTA 0, clock64 stamps per unit / segment; run with SJD_ATTN=sw)."""
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
W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
P = 67
ids = torch.randint(4, 8196, (2 * W,), dtype=torch.int32).to(dev)
pos = torch.arange(L, L + W, dtype=torch.int32)
rope = torch.cat([pos, pos - (P - 1)]).to(dev); cpos = torch.cat([pos, pos]).to(dev)
for _ in range(3):
    st.forward(W, rope, cpos, L, [0, P - 1], ids=ids, n_logit_tokens=W)
buf = torch.zeros(8 * 16 + 4 * 160, dtype=torch.int64, device=dev)
lib.sjd_debug_attn_stamps(buf.data_ptr())
st.forward(W, rope, cpos, L, [0, P - 1], ids=ids, n_logit_tokens=W)
torch.cuda.synchronize()
lib.sjd_debug_attn_stamps(None)
full = buf.cpu()
b = full[:128].view(8, 16)
t0 = int(b[0, 15])
names = ["K issued", "K,Q landed", "S issued", "PV ready", "PV issued", "S seen", "max done", "P arrived", "seg O seen", "seg epi done"]
print(f"W={W} L={L} (us after the dependency wait of CTA 0, last layer; rows 8/9 are indexed by SEGMENT)")
for n in range(8):
    print(f"unit {n}: " + "  ".join(f"{names[k]} {((int(b[n, k]) - t0) / 1.9e3 if int(b[n, k]) else float('nan')):6.2f}" for k in range(10)))

print("producer (us): after wait %.2f, early-Q done %.2f, cursors ready %.2f, first post-wait Q issued %.2f" % tuple(((int(b[0, k]) - t0) / 1.9e3 if int(b[0, k]) else float('nan')) for k in (10, 11, 12, 13)))
print("cluster tail of CTA 0 (us): tail entered %.2f, first cluster barrier passed %.2f, merged %.2f, rows stored %.2f, second barrier passed %.2f" % tuple(((int(b[1, k]) - t0) / 1.9e3 if int(b[1, k]) else float('nan')) for k in (10, 11, 12, 13, 14)))
c = full[128:].view(160, 4)
c = c[c[:, 0] > 0]
c = c[c[:, 2] > 0]
if len(c):
    s0 = int(c[:, 0].min())
    st_, en_ = (c[:, 0] - s0).float() / 1e3, (c[:, 1] - s0).float() / 1e3
    print(f"per-CTA wall clock (us after the first CTA's dependency wait returned): {len(c)} CTAs, "
          f"wait returned min/median/max {st_.min():.2f}/{st_.median():.2f}/{st_.max():.2f}, "
          f"end min/median/max {en_.min():.2f}/{en_.median():.2f}/{en_.max():.2f}")
    for nu in sorted(set(c[:, 2].tolist())):
        m = c[:, 2] == nu
        print(f"  CTAs with {nu} units: {int(m.sum())}, busy (end - own wait) median {(en_[m] - st_[m]).median():.2f} max {(en_[m] - st_[m]).max():.2f} us")

# what makes a CTA slow?  classify by the units it owns (contiguous split of runs x key tiles, key tiles fastest)
if len(c) and os.environ.get("SJD_STAMPS_CLASSES"):
    n_chunks = (L + W + 127) // 128
    n_units = n_chunks * 32 * 2
    G = len(c)
    rows = []
    for cta in range(G):
        u0 = int(c[cta, 3]); u1 = u0 + int(c[cta, 2])
        if u1 == u0:
            continue
        kts = [u % n_chunks for u in range(u0, u1)]
        runs = len({u // n_chunks for u in range(u0, u1)})
        rows.append((u1 - u0, int((n_chunks - 1) in kts), int(any(kt == 0 and (u // n_chunks) >= 32 for kt, u in zip(kts, range(u0, u1)))), runs,
                     kts[0] == n_chunks - 1, float(en_[cta] - st_[cta])))
    import collections
    agg = collections.defaultdict(list)
    for r in rows:
        agg[r[:5]].append(r[5])
    print("class (units, has window tile, has hidden-prefix tile, runs, STARTS with the window tile): n, median busy us, max")
    for k in sorted(agg):
        v = sorted(agg[k])
        print(f"  {k}: n={len(v)} median {v[len(v) // 2]:.2f} max {v[-1]:.2f}")
