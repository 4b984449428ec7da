#!/usr/bin/env python
"""torchrun entry of the multi-GPU prompt runner (sjd_b200.launcher; reference: eval_model.py -> run_caption_gen)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sjd_b200  # noqa: E402,F401
from sjd_b200 import launcher  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="synthetic/lumina-mgpt-7b-768")
    ap.add_argument("--prompts", help="text file, one prompt per line (default: 8 built-in prompts)")
    ap.add_argument("--output-dir", default="./workdir")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max_num_new_tokens", type=int, default=32)
    ap.add_argument("--multi_token_init_scheme", default="random")
    ap.add_argument("--guidance_scale", type=float, default=3.0)
    ap.add_argument("--target_size", type=int, default=None)
    ap.add_argument("--n_layers", type=int, default=None, help="synthetic families: truncate the stack (smoke runs)")
    a = ap.parse_args()
    if a.prompts:
        prompts = [l.strip() for l in open(a.prompts) if l.strip()]
    else:
        prompts = [f"synthetic prompt {i}" for i in range(8)]
    kw = dict(max_num_new_tokens=a.max_num_new_tokens, multi_token_init_scheme=a.multi_token_init_scheme,
              guidance_scale=a.guidance_scale)
    if a.target_size:
        kw["target_size"] = a.target_size
    if a.n_layers:
        kw["n_layers"] = a.n_layers
    launcher.run_prompts(a.model, prompts, output_dir=a.output_dir, seed=a.seed, **kw)


if __name__ == "__main__":
    main()
