#!/usr/bin/env python
"""bench.py — image-tokens/sec & accepted-tokens/iter of the SJD hot path on Lumina-mGPT-7B shapes, 768x768.

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # reference's algorithm on the host cores (oracle port)

One "step" = one full 768x768 image: synthetic 64-token prompt + <boi,h,w>, 2 352 image-region tokens
(48 rows x (48 + EOL)) + EOI + 1, decoded by Speculative Jacobi Decoding with window 32, cfg 3.0, top-k 2000,
temperature 1 — BASELINE.json configs[1] on random-init weights of the exact 7B shape (no checkpoints offline).
Prints ONE JSON line (see DESIGN.md §6 for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

WINDOW = 32
GRID = 48                      # 768 / 16 latent rows / cols
PROMPT_TEXT = 64
IMG_TOKENS = GRID * (GRID + 1)  # 2352 image-region tokens incl. EOL
CFG, TOP_K = 3.0, 2000
METRIC = "image_tokens_per_sec"
UNIT = "tokens/s"


def bench_config(n_gpus, window):
    return {"workload": "Lumina-mGPT-7B shape (Chameleon 7B: 32L d4096 H32 ff11008 V65536), 768x768 (48x49 image tokens), "
                        f"SJD window {window}, cfg {CFG}, top-k {TOP_K}, temp 1.0, 1 prompt per GPU per step",
            "window": window, "prompts_per_gpu_per_step": 1, "parallelism": f"replicas x{n_gpus}",
            "weights": "random-init N(0,0.02) bf16", "l2": "weights 13.5 GB per trip >> 126 MB L2 (no flush needed)"}


def synthetic_prompt(seed: int):
    g = torch.Generator().manual_seed(1000 + seed)
    text = torch.randint(8900, 65000, (PROMPT_TEXT,), generator=g).tolist()
    return text + [8197, 8804 + GRID // 2, 8804 + GRID // 2]   # <boi>, h-grid, w-grid (item_processor.py:103-104)


def sjd_params(window, seed):
    return dict(jacobi_loop_interval_l=3, jacobi_loop_interval_r=GRID * GRID + GRID - 10, max_num_new_tokens=window,
                guidance_scale=CFG, seed=seed, multi_token_init_scheme="random", do_cfg=True,
                prefix_token_sampler_scheme="speculative_jacobi")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def stack_bytes(shape, rows):
    """(weight bytes of one forward incl. lm_head, KV bytes per attended key token summed over layers and K+V)"""
    d, ff, V, hd = shape.d_model, shape.d_ff, shape.vocab, shape.n_heads * shape.head_dim
    qkv_n = (shape.n_heads + 2 * shape.n_kv_heads) * shape.head_dim
    w = shape.n_layers * (qkv_n * d + d * hd + 2 * ff * d + d * ff) * 2 + V * d * 2
    kv_per_token = shape.n_layers * 2 * shape.n_kv_heads * shape.head_dim * 2
    return w, kv_per_token


def whole_trip_roofline(shape, rows, nfe, kv_read_tokens, seconds):
    """Algorithmic HBM bytes of the whole decode (every forward streams all weights once and reads the visible K/V of
    every CFG row once) over the device time, against the measured HBM peak."""
    w, kvt = stack_bytes(shape, rows)
    alg = nfe * w + kv_read_tokens * kvt
    peak, src = measured_peaks()
    ach = alg / seconds / 1e9
    return {"bound": "hbm", "alg_bytes_per_nfe": int(alg / max(nfe, 1)), "weights_bytes": int(w),
            "kv_bytes_per_nfe": int(kv_read_tokens * kvt / max(nfe, 1)), "achieved": round(ach, 1), "peak": peak,
            "unit": "GB/s", "frac": round(ach / peak, 4), "peak_source": src}


def projected(ms_per_nfe):
    """tokens/s at the acceptance the reference publishes for real checkpoints (README: 2.1-2.4 accepted tokens per
    forward with window 16-32): random-init weights accept ~1.06, so the measured tokens/s is an AR-rate number."""
    return {f"{a:.1f}": round(a / ms_per_nfe * 1e3, 1) for a in (2.1, 2.4)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            if d.get("hbm_gbs"):
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ======================================================================================= our arm
def run_ours(args):
    import sjd_b200  # noqa: F401
    from sjd_b200 import _lib, engine, families, model, replicas

    rank, world, local = replicas.init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.lib()
    shape = families.lumina_7b()
    P = PROMPT_TEXT + 3
    max_length = P + IMG_TOKENS + 2
    weights = families.random_weights(shape, seed=0, device=dev)
    demo = args.logit_scale != 1.0 or args.top_k != TOP_K
    if args.logit_scale != 1.0:   # demonstration of the multi-token path at scale, NOT the headline workload
        weights["lm_head"] = (weights["lm_head"].float() * args.logit_scale).to(torch.bfloat16)
    cos, sin = families.rope_rotate_half(shape.head_dim, 2560, 10000.0, True)
    stack = model.DeviceStack(shape, weights, cos, sin, rows=2, max_len=2560, device=dev)
    del weights
    grammar = engine.LuminaGrammarState(image_top_k=args.top_k, text_top_k=10)
    eng = engine.SJDEngine(stack, engine.SJDParams(**sjd_params(args.window, 0)), grammar, torch.arange(4, 8196))

    def one_image(idx):
        eng.p.seed = idx
        prompt = synthetic_prompt(idx)
        ids = eng.generate(prompt, max_length=max_length, eos_token_ids=[8710], kv_lo=[0, P - 1])
        return len(ids) - P, eng.stats.nfe, eng.stats

    # ---- warm-up ------------------------------------------------------------------------------------
    for i in range(args.warmup):
        one_image(10_000 + rank + i * world)
    # ---- timed: device timeline (CUDA events on the launching stream), max over ranks -----------------
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = lib.sjd_launch_count()
    replicas.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tok = nfe = 0
    h2d = d2h = kvr = 0
    for i in range(args.steps):
        t, n, st = one_image(rank + i * world)
        tok, nfe = tok + t, nfe + n
        h2d, d2h, kvr = h2d + st.h2d_bytes, d2h + st.d2h_bytes, kvr + st.kv_read_tokens
    e1.record()
    torch.cuda.synchronize()
    replicas.barrier()
    launches = lib.sjd_launch_count() - launches0
    t_dev = replicas.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)
    clk = clocks.stop()
    counters = replicas.gather_counters(tok, nfe, 1, dev)
    tot_tok, tot_nfe = int(counters[:, 0].sum()), int(counters[:, 1].sum())

    # ---- e2e: the public call a user makes, host buffers in / host tokens out, wall clock ----------------
    from sjd_b200 import hf_api
    solver = hf_api.SyntheticLuminaSolver(eng, max_length=max_length)
    replicas.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e_tok = 0
    e_h2d = e_d2h = 0
    for i in range(args.steps):
        prompt = torch.tensor([synthetic_prompt(rank + i * world)], dtype=torch.int64).pin_memory()
        out = solver.generate(prompt, seed=rank + i * world)        # returns host LongTensor [1, P + new]
        e_tok += out.shape[1] - P
        e_h2d += eng.stats.h2d_bytes
        e_d2h += eng.stats.d2h_bytes
    torch.cuda.synchronize()
    t_e2e = replicas.max_over_ranks(time.perf_counter() - w0, dev)
    replicas.barrier()
    e_tot = int(replicas.gather_counters(e_tok, 0, 1, dev)[:, 0].sum())

    # ---- roofline of the dominant kernel (gemm_chain_kernel): every GEMM of one trip, timed alone with CUDA events ----
    # sjd_ctx_gemm_only launches the same persistent chain kernels as the forward (1 + n_layers launches carrying
    # 4*n_layers + 1 GEMMs with their fused epilogues), attention skipped.
    roof = None
    cpu_base = None
    gpu_ref = None

    def one_image_bounded(window, n_img_tokens, seed):
        """our engine on the first n_img_tokens of an image with the given window (the span the eager reference leg times)"""
        e2 = engine.SJDEngine(stack, engine.SJDParams(**sjd_params(window, seed)), grammar, torch.arange(4, 8196))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        ids = e2.generate(synthetic_prompt(seed), max_length=P + n_img_tokens, eos_token_ids=[8710], kv_lo=[0, P - 1])
        b.record()
        torch.cuda.synchronize()
        return {"seconds": a.elapsed_time(b) / 1e3, "nfe": e2.stats.nfe, "new_tokens": len(ids) - P}

    if rank == 0:
        M = 2 * args.window
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            _lib.check(lib.sjd_ctx_gemm_only(stack.ctx, args.window, st), "gemm_only")
        reps = 10
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(reps):
            lib.sjd_ctx_gemm_only(stack.ctx, args.window, st)
        g1.record()
        torch.cuda.synchronize()
        n_launch = shape.n_layers + 1
        n_gemm = 4 * shape.n_layers + 1
        t_launch = g0.elapsed_time(g1) / 1e3 / reps / n_launch
        d, ff, V, hd = shape.d_model, shape.d_ff, shape.vocab, shape.n_heads * shape.head_dim
        qkv_n = (shape.n_heads + 2 * shape.n_kv_heads) * shape.head_dim
        gemms = [(qkv_n, d), (d, hd), (2 * ff, d), (d, ff)] * shape.n_layers + [(V, d)]
        # algorithmic bytes: weights once + activations in (bf16) + results out (bf16; fp32 for the logits)
        alg = (sum(N * K * 2 + M * K * 2 + M * N * 2 for N, K in gemms) + M * V * 2) / n_launch
        peak, peak_src = measured_peaks()
        achieved = alg / t_launch / 1e9
        traffic = None
        tp = ROOT / "profiles" / "r02_gemm_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"kernel": "gemm_chain_kernel (persistent; tcgen05 + TMA; o_proj, gate_up, down, next qkv | lm_head per launch)",
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": int(alg), "avg_launch_us": round(t_launch * 1e6, 2),
                "launches_timed": n_launch * reps, "gemms_per_trip": n_gemm,
                "us_per_gemm": round(t_launch * n_launch / n_gemm * 1e6, 2)}
        if world == 1 and not args.no_cpu_baseline and not demo:
            cpu_base = cpu_reference(args, budget_s=args.cpu_budget)
        if world == 1 and not args.no_gpu_reference and not demo:
            gpu_ref = gpu_eager_reference(args, dev, eng, one_image_bounded)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(tot_tok / t_dev, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(t_dev / args.steps * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(bench_config(world, args.window), **({"DEMONSTRATION_not_the_headline_workload":
                           f"lm_head scaled by {args.logit_scale}, image top-k {args.top_k}: flat logits make the "
                           "speculative test accept most drafts, which shows the multi-token path and what tokens/s does "
                           "when acceptance is high"} if demo else {})),
            "accepted_tokens_per_iter": round(tot_tok / tot_nfe, 3), "nfe_per_image": round(tot_nfe / (args.steps * world), 1),
            "ms_per_nfe": round(t_dev / (tot_nfe / world) * 1e3, 3),
            "whole_trip_roofline": whole_trip_roofline(shape, 2, nfe, kvr, e0.elapsed_time(e1) / 1e3),
            "projected_tokens_per_sec_per_gpu_at_published_acceptance": projected(t_dev / (tot_nfe / world) * 1e3),
            "vs_reference_note": "the --impl reference arm is ONE host's CPU cores whatever N is: its ratio is meaningful at "
                                 "N=1 only; the GPU-vs-GPU number is gpu_eager_reference",
            "clocks": clk,
            "e2e": {"value": round(e_tot / t_e2e, 2), "unit": UNIT, "h2d_bytes_per_step": e_h2d // args.steps,
                    "d2h_bytes_per_step": e_d2h // args.steps, "api": "SyntheticLuminaSolver.generate (host ids in, host ids out)"},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu_base, "gpu_eager_reference": gpu_ref,
        }
        print(json.dumps(line), flush=True)
    stack.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# ===================================================== the reference's own PyTorch-eager SJD on this GPU (north_star's denominator)
def gpu_eager_reference(args, dev, eng, ours_bounded):
    """BASELINE north_star: ">= 2x the reference's own single-GPU PyTorch SJD image-tokens/sec on Lumina-mGPT-7B 768x768".
    Runs the UNMODIFIED reference (baseline/_ref: scheduler/jacobi_iteration_lumina_mgpt.py + vendored Chameleon, HF-5.5
    name shims only) on this GPU at the same 7B shape / bf16 / cfg / top-k, on the first `--ref-tokens` image tokens of one
    768x768 image (a bounded sample: the whole image takes the reference ~2 minutes), for windows 32 and 16, timed by
    the reference's own CUDA-event timer; then our engine on the SAME span.  `speedup` = reference ms/NFE over ours
    (acceptance is a property of the weights, identical for both)."""
    sys.path.insert(0, str(ROOT / "baseline"))
    try:
        import ref_gpu_eager as RG
    except Exception as ex:   # pragma: no cover
        return {"unavailable": f"baseline/ref_gpu_eager.py import failed: {ex!r}"[:200]}
    why = RG.available()
    if why:
        return {"unavailable": why}
    out = {"impl": "unmodified reference (JacobiSampler._sample + vendored ChameleonForConditionalGeneration, sdpa), "
                   "PyTorch eager, bf16, same GPU", "torch": torch.__version__, "sample":
           f"first {args.ref_tokens} image tokens of one 768x768 image per window (prefill + AR steps + Jacobi windows)",
           "timer": "reference's own CUDA events (jacobi_iteration_lumina_mgpt.py:1050-1055,1213-1223)", "windows": {}}
    try:
        t0 = time.perf_counter()
        model, mods = RG.build_model(dev, seed=0)
        out["build_s"] = round(time.perf_counter() - t0, 1)
        P = PROMPT_TEXT + 3
        for window in (args.window, 16) if args.window != 16 else (16,):
            RG.run(model, mods, prompt=synthetic_prompt(77), max_length=P + 24, window=window, guidance=CFG,
                   image_top_k=TOP_K, seed=77, grid=GRID)                      # warm-up (cuBLAS heuristics, allocator)
            r = RG.run(model, mods, prompt=synthetic_prompt(0), max_length=P + args.ref_tokens, window=window,
                       guidance=CFG, image_top_k=TOP_K, seed=0, grid=GRID)
            ours_bounded(window, 24, 77)
            o = ours_bounded(window, args.ref_tokens, 0)
            out["windows"][str(window)] = {
                "reference_ms_per_nfe": round(r["ms_per_nfe"], 3), "reference_nfe": r["nfe"],
                "reference_new_tokens": r["new_tokens"], "reference_tokens_per_s": round(r["tokens_per_s"], 2),
                "ours_ms_per_nfe": round(o["seconds"] / o["nfe"] * 1e3, 3), "ours_nfe": o["nfe"],
                "ours_new_tokens": o["new_tokens"], "ours_tokens_per_s": round(o["new_tokens"] / o["seconds"], 2),
                "speedup_ms_per_nfe": round(r["ms_per_nfe"] / (o["seconds"] / o["nfe"] * 1e3), 3),
                "speedup_tokens_per_s": round((o["new_tokens"] / o["seconds"]) / r["tokens_per_s"], 3)}
        del model
        torch.cuda.empty_cache()
    except Exception as ex:
        import traceback
        out["error"] = f"{ex!r}"[:300]
        out["traceback_tail"] = traceback.format_exc()[-600:]
    return out


# ================================================================= reference arm / cpu baseline (oracle port)
def cpu_reference(args, budget_s=20.0, steps=1, warmup=0):
    """The reference's algorithm on the host cores: oracle/sjd_oracle.decode driving oracle/ref_forward.RefStack
    (fp32 PyTorch eager, all cores) at the SAME shapes/config, on a bounded sample: the first `n_trips` Jacobi
    iterations of one image (prefill + 3 AR steps + windows).  Throughput = new tokens / wall time."""
    import numpy as np
    from oracle import ref_forward as RF, sjd_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    cfg = RF.StackConfig(32, 4096, 32, 32, 128, 11008, 65536, 1e-5, qk_norm=True)
    # one set of random layer tensors, cloned per layer: distinct memory (same DRAM traffic as 32 different layers)
    # without spending minutes on 6.7e9 randn's; values are irrelevant to timing.
    one = RF.StackConfig(1, 4096, 32, 32, 128, 11008, 65536, 1e-5, qk_norm=True)
    w1 = RF.random_weights(one, seed=0)
    w = {"embed": w1["embed"], "final_norm": w1["final_norm"], "lm_head": w1["lm_head"],
         "layers": [{k: v.clone() for k, v in w1["layers"][0].items()} for _ in range(cfg.n_layers)]}
    cos, sin = RF.rope_tables_rotate_half(128, 2560, 10000.0, True)
    P = PROMPT_TEXT + 3
    results = []
    for it in range(warmup + steps):
        ref = RF.RefStack(cfg, w, cos, sin, rows=2, max_len=2560, emulate_bf16=False)
        kv_lo = [0, P - 1]

        def logits_fn(rows_tokens, kv_len, n):
            ids = torch.tensor(rows_tokens)
            W = ids.shape[1]
            pos = torch.arange(kv_len, kv_len + W)[None].repeat(2, 1)
            rope = torch.stack([(pos[b] - kv_lo[b]).clamp(min=0) for b in range(2)])
            lg = ref.forward(ids=ids, rope_pos=rope, kv_len=kv_len, kv_lo=kv_lo, cache_pos=pos, n_logit_tokens=n)
            return lg.reshape(-1, cfg.vocab).numpy()

        t_start = time.perf_counter()
        state = {"trips": 0}

        def stop(ids):
            state["trips"] += 1
            return (time.perf_counter() - t_start) > budget_s

        ids, nfe = O.decode(logits_fn, synthetic_prompt(it), params=O.OracleParams(**sjd_params(args.window, it)),
                            grammar=O.LuminaGrammar(image_top_k=TOP_K), img_vocab=np.arange(4, 8196),
                            max_length=P + IMG_TOKENS + 2, eos_ids=[8710], rows=2, stop_fn=stop)
        dt = time.perf_counter() - t_start
        if it >= warmup:
            results.append((len(ids) - P, nfe, dt))
    tok = sum(r[0] for r in results)
    nfe = sum(r[1] for r in results)
    dt = sum(r[2] for r in results)
    return {"value": round(tok / dt, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {nfe // max(1, len(results))} Jacobi iterations of one 768x768 image per step "
                      f"(~{budget_s:.0f}s of CPU work), fp32 PyTorch-eager oracle port of the reference loop, same config",
            "nfe": nfe, "new_tokens": tok, "seconds": round(dt, 2), "ms_per_nfe": round(dt / max(nfe, 1) * 1e3, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference(args, budget_s=args.cpu_budget, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": round(res["seconds"] / args.steps * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.gpus, args.window), "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ======================================================================== BASELINE configs 1 / 4 / 5 (own lines)
CONFIGS = {
    1: dict(model="synthetic/llamagen-gpt-b", window=16, kw=dict(guidance_scale=4.0, image_top_k=1000, target_size=256),
            name="LlamaGen GPT-B (12L d768), 256x256 (256 tokens), window 16, cfg 4, top-k 1000"),
    4: dict(model="synthetic/emu3-gen", window=64, kw=dict(guidance_scale=3.0, image_top_k=2048, target_size=720),
            name="Emu3-Gen shape (32L d4096 GQA 32:8 ff14336 V184622), 720x720 (90x91 tokens), window 64, cfg 3, top-k 2048"),
    5: dict(model="synthetic/anole-7b-512", window=32, kw=dict(guidance_scale=3.0, image_top_k=2000, target_size=512),
            name="Anole / Chameleon-7B shape, 512x512 (1024 image tokens, Anole grammar), cfg 3, top-k 2000",
            sweep=(8, 16, 32, 64, 128)),
}


def run_config(args):
    """One JSON line for BASELINE config 1, 4 or 5 through the public loader API (model_wrappers.model_loader:
    load_pretrained_model / get_forward_func on the synthetic family; host prompt string in, host token tensor out, so
    `value` and `e2e` are the same measurement).  Config 5 sweeps the window."""
    import sjd_b200  # noqa: F401
    from sjd_b200 import _lib, replicas
    from model_wrappers.model_loader import get_forward_func, load_pretrained_model
    rank, world, local = replicas.init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    C = CONFIGS[args.config]
    lib = _lib.lib()
    windows = C.get("sweep", (C["window"],)) if args.window == WINDOW else (args.window,)
    rows_out = []
    clk = None
    for window in windows:
        solver = load_pretrained_model(C["model"], device=dev, seed=0, max_num_new_tokens=window, **C["kw"])
        fwd = get_forward_func(C["model"], solver)
        for i in range(args.warmup):
            solver.engine.p.seed = 10_000 + i
            fwd(f"warm-up prompt {rank} {i}")
        clocks = ClockSampler(local)
        clocks.start()
        l0 = lib.sjd_launch_count()
        replicas.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tok = nfe = kvr = h2d = d2h = 0
        for i in range(args.steps):
            solver.engine.p.seed = rank + i * world
            out = fwd(f"prompt {rank + i * world}")
            st = solver.engine.stats
            tok, nfe, kvr = tok + int(out.numel()), nfe + st.nfe, kvr + st.kv_read_tokens
            h2d, d2h = h2d + st.h2d_bytes, d2h + st.d2h_bytes
        e1.record()
        torch.cuda.synchronize()
        replicas.barrier()
        t = replicas.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)
        clk = clocks.stop()
        cnt = replicas.gather_counters(tok, nfe, 1, dev)
        T, N = int(cnt[:, 0].sum()), int(cnt[:, 1].sum())
        rows_out.append({"window": window, "tokens_per_s": round(T / t, 2), "accepted_tokens_per_iter": round(T / N, 3),
                         "ms_per_nfe": round(t / (N / world) * 1e3, 3), "nfe_per_image": round(N / (args.steps * world), 1),
                         "whole_trip_roofline": whole_trip_roofline(solver.stack.shape, 2, nfe, kvr, e0.elapsed_time(e1) / 1e3),
                         "projected_tokens_per_sec_per_gpu_at_published_acceptance": projected(t / (N / world) * 1e3),
                         "gpu_launches": int(lib.sjd_launch_count() - l0), "seconds": round(t, 3),
                         "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps})
        solver.stack.close()
        del solver, fwd
        torch.cuda.empty_cache()
    if rank == 0:
        head = next((r for r in rows_out if r["window"] == C["window"]), rows_out[0])
        line = {"metric": METRIC, "value": head["tokens_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(head["seconds"] / args.steps * 1e3, 2), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"BASELINE config {args.config}: {C['name']}; 1 prompt per GPU per step",
                           "window": head["window"], "parallelism": f"replicas x{world}",
                           "weights": "random-init N(0,0.02) bf16", "l2": "weights per trip >> 126 MB L2 (no flush needed)"
                           if args.config != 1 else "GPT-B weights (0.2 GB) exceed the 126 MB L2; no flush"},
                "accepted_tokens_per_iter": head["accepted_tokens_per_iter"], "ms_per_nfe": head["ms_per_nfe"],
                "nfe_per_image": head["nfe_per_image"], "clocks": clk,
                "e2e": {"value": head["tokens_per_s"], "unit": UNIT, "h2d_bytes_per_step": head["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": head["d2h_bytes_per_step"],
                        "api": "model_wrappers.model_loader.get_forward_func(...)(prompt: str) -> LongTensor (host)"},
                "gpu_launches": head["gpu_launches"], "roofline": head["whole_trip_roofline"], "cpu_baseline": None,
                "window_sweep": rows_out if len(rows_out) > 1 else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--window", type=int, default=WINDOW)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work per reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference's eager SJD on this GPU")
    ap.add_argument("--ref-tokens", type=int, default=192, help="image tokens decoded by the eager-reference leg per window")
    ap.add_argument("--logit-scale", type=float, default=1.0,
                    help="demonstration only: scale the random lm_head by this factor (flat logits => drafts are accepted)")
    ap.add_argument("--top-k", type=int, default=TOP_K, help="demonstration only: image top-k (the headline uses 2000)")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE config: 2 (default; 3 is the same workload at --gpus 8), 1 = LlamaGen GPT-B, 4 = Emu3-Gen, "
                         "5 = Anole window sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config in (1, 4, 5):
        run_config(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
