import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_loop_goldens():
    out = {}
    for p in sorted(GOLDEN.glob("sjd_loop_*.json")):
        out[p.stem.replace("sjd_loop_", "")] = json.loads(p.read_text())
    return out


@pytest.fixture(scope="session")
def loop_goldens():
    return load_loop_goldens()


# Attention kernels a forward test can force (the context reads these variables when it is created).  "sw" = the
# segment-accumulating small-window tcgen05 kernel (attention_sw.cu); its suffixes turn the developer knobs that make
# SMALL test shapes walk what the bench shape walks: g<N> caps the grid at N CTAs (one CTA then runs several heads, wraps
# its K / V rings, rotates its Q buffers and accumulates several key tiles per segment), grow<X> ends a segment whenever a
# tile's maximum exceeds the reference by X (0: nearly every tile -> accumulator rotation), n64 forces the 64-column
# instantiation, c<K> caps the cluster size of the in-kernel merge (c0: partial slots + merge pre-op in the next chain kernel;
# default: clusters of up to 4 CTAs per head — with grow0 the accumulators are rescaled in place at nearly every tile).
ATTN_MODES = ["auto", "tc", "tct", "mma", "sw", "sw:grow0", "sw:c2:n64:grow0", "sw:c0", "sw:g1", "sw:g3:grow0", "sw:g2:n64"]
_ATTN_KNOBS = ("SJD_ATTN", "SJD_ATTN_SW_GRID", "SJD_ATTN_SW_GROW", "SJD_ATTN_SW_NCOLS", "SJD_ATTN_SW_CLUSTER")


def set_attn(monkeypatch, attn):
    for k in _ATTN_KNOBS:
        monkeypatch.delenv(k, raising=False)
    if attn == "auto":
        return
    parts = attn.split(":")
    monkeypatch.setenv("SJD_ATTN", parts[0])
    for kn in parts[1:]:
        if kn.startswith("grow"):
            monkeypatch.setenv("SJD_ATTN_SW_GROW", kn[4:])
        elif kn.startswith("c"):
            monkeypatch.setenv("SJD_ATTN_SW_CLUSTER", kn[1:])
        elif kn.startswith("g"):
            monkeypatch.setenv("SJD_ATTN_SW_GRID", kn[1:])
        elif kn.startswith("n"):
            monkeypatch.setenv("SJD_ATTN_SW_NCOLS", kn[1:])
        else:
            raise ValueError(attn)
