"""Developer: cProfile of the engine's host loop over one bench image (what the host does per Jacobi iteration)."""
import cProfile, io, os, pstats, sys
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
eng = engine.SJDEngine(stack, engine.SJDParams(**bench.sjd_params(bench.WINDOW, 0)),
                       engine.LuminaGrammarState(image_top_k=bench.TOP_K, text_top_k=10), torch.arange(4, 8196))
eng.generate(bench.synthetic_prompt(0), max_length=max_length, eos_token_ids=[8710], kv_lo=[0, P - 1])
pr = cProfile.Profile()
pr.enable()
eng.generate(bench.synthetic_prompt(1), max_length=max_length, eos_token_ids=[8710], kv_lo=[0, P - 1])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
print("nfe", eng.stats.nfe)
