"""CPU tests (no GPU): the C-ABI library loads and exports every symbol of include/sjd_b200.h, host-side
logic of the product (grammar state, window bookkeeping helpers, replica sharding over gloo) agrees with the
oracle restatement, and the product never imports the oracle."""
import ctypes
import os
import random
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "accelerating-t2i-ar-with-sjd_b200"


@pytest.fixture(scope="module")
def lib():
    import sjd_b200  # noqa: F401
    from sjd_b200 import _lib
    _lib.build()
    return _lib


def test_library_exports_every_header_symbol(lib):
    names = lib.header_symbols()
    assert len(names) >= 14
    dll = ctypes.CDLL(str(lib.LIB_PATH))
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/sjd_b200.h but not exported"
    assert lib.lib().sjd_version() >= 100


def test_library_is_built_for_sm100a_with_tcgen05_and_tma(lib):
    sass = subprocess.run(["cuobjdump", "-sass", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"


def test_stream_k_workspace_is_host_computable(lib):
    L = lib.lib()
    # 2 KB of counters + 32 B of {arrived, done} per tile + per-tile row statistics + two fp32 partial slots per CTA
    assert L.sjd_gemm_workspace_bytes(4096, 4096, 64, 148) == 2048 + 32 * 32 + 32 * 64 * 4 + 2 * 148 * 64 * 128 * 4
    assert L.sjd_gemm_workspace_bytes(184622, 4096, 128, 148) == 2048 + 1443 * 32 + 1443 * 128 * 4 + 2 * 148 * 128 * 128 * 4


def test_no_compute_without_gpu_fails_loudly(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sjd_b200.model import DeviceStack, StackShape
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceStack(StackShape(1, 64, 1, 1, 64, 64, 128), {}, torch.zeros(4, 32), torch.zeros(4, 32), 1, 16)


def test_product_never_imports_oracle():
    for p in list(PKG.rglob("*.py")) + [ROOT / "sjd_b200.py"] + list((ROOT / "scheduler").glob("*.py")) + \
            list((ROOT / "llamagen").glob("*.py")) + list((ROOT / "lumina_mgpt").glob("*.py")):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{p} imports the oracle"


def test_grammar_state_matches_oracle_grammar(lib):
    """Product-side incremental grammar == oracle's from-scratch restatement on random token streams."""
    from oracle import sjd_oracle as O
    from sjd_b200.engine import LuminaGrammarState
    rnd = random.Random(0)
    for trial in range(30):
        h, w = rnd.randint(1, 4), rnd.randint(1, 5)
        ids = [rnd.randint(8900, 9000) for _ in range(rnd.randint(1, 5))] + [8197, 8804 + h, 8804 + w]
        og, pg = O.LuminaGrammar(image_top_k=77, text_top_k=5), LuminaGrammarState(image_top_k=77, text_top_k=5)
        pg.reset()
        pg.observe(ids)
        total = (2 * w + 1) * 2 * h + 1
        produced = 0
        while produced < total + 6:
            n = rnd.randint(1, 9)
            d_o, d_p = og.describe(ids, n), pg.describe(n)
            assert d_o["forced"] == d_p["forced"], (trial, ids)
            assert (d_o["allow"] is None) == (d_p["allow"] is None)
            assert d_o["top_k"] == d_p["top_k"]
            assert d_o["no_cfg"] == pg.no_cfg
            m = rnd.randint(1, n)   # accept m tokens, honouring forced ones
            new = [d_o["forced"][j] if d_o["forced"][j] >= 0 else rnd.randint(4, 8195) for j in range(m)]
            ids = ids + new
            pg.observe(new)
            produced += m


def test_emu3_grammar_state_matches_oracle_grammar(lib):
    """Emu3 grammar: incremental product-side state == oracle restatement (itself pinned to the reference's
    EOLLogitProcessor3d by tests/golden/sjd_loop_emu3_*.json), through EOL / EOF / EOI / EOS and into the PAD tail."""
    from oracle import sjd_oracle as O
    from sjd_b200.engine import Emu3GrammarState
    rnd = random.Random(1)
    for trial in range(30):
        h, w = rnd.randint(1, 4), rnd.randint(1, 6)
        args = (h, w, 900, 901, 902, 903, 904, 905, 1000, 3048)
        og, pg = O.Emu3Grammar(*args, top_k=64), Emu3GrammarState(*args, top_k=64)
        ids = [rnd.randint(1, 800) for _ in range(rnd.randint(0, 4))] + [900]
        pg.reset()
        pg.observe(ids)
        produced, total = 0, (w + 1) * h + 3
        while produced < total + 12:
            n = rnd.randint(1, 11)
            d_o, d_p = og.describe(ids, n), pg.describe(n)
            assert d_o["forced"] == d_p["forced"], (trial, ids, n)
            assert tuple(d_o["allow"]) == tuple(d_p["allow"]) and d_o["top_k"] == d_p["top_k"]
            m = rnd.randint(1, n)
            new = [d_o["forced"][j] if d_o["forced"][j] >= 0 else rnd.randint(1000, 3047) for j in range(m)]
            ids = ids + new
            pg.observe(new)
            produced += m


def test_prompt_sharding_covers_all_prompts(lib):
    from sjd_b200.replicas import shard_prompts
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in shard_prompts(8, r, world))
        assert got == list(range(8))
        assert all(len(shard_prompts(8, r, world)) == 8 // world for r in range(world))


_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, sjd_b200
from sjd_b200 import replicas
rank, world, local = replicas.init_process_group("gloo")
mine = replicas.shard_prompts(8, rank, world)
tok, nfe = sum(100 + i for i in mine), sum(10 + i for i in mine)
replicas.barrier()
g = replicas.gather_counters(tok, nfe, 1)
t = replicas.max_over_ranks(1.0 + rank)
if rank == 0:
    print(json.dumps(dict(world=world, total_tokens=int(g[:, 0].sum()), total_nfe=int(g[:, 1].sum()),
                          done=int(g[:, 2].sum()), tmax=t)))
"""


def test_two_rank_gloo_counter_gather(lib, tmp_path):
    """world_size-2 CPU run of the only exchange step on the path (accepted-token counters)."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out == dict(world=2, total_tokens=sum(100 + i for i in range(8)), total_nfe=sum(10 + i for i in range(8)),
                       done=2, tmax=2.0)
