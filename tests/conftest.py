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
