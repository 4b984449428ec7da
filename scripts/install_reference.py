"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with gpurun) so that
bench.py's `gpu_eager_reference` leg can time the reference's own PyTorch-eager SJD on the B200 (BASELINE north_star:
">= 2x the reference's own single-GPU PyTorch SJD").

  python scripts/install_reference.py            # needs /root/reference (the build container)

The reference's setup.py packages only its training toolkit (`xllmx`); the SJD path (scheduler/, the vendored Chameleon
under lumina_mgpt/model/, llamagen/, emu3/) is run "from the repository root" upstream (test_lumina_mgpt.py:3-4 appends
./ and ./lumina_mgpt/ to sys.path).  So the install is: pip install of setup.py (--no-deps, offline) plus a verbatim copy
of those module directories next to it.  Nothing under baseline/_ref is part of this repository's history or product.
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("SJD_REFERENCE", "/root/reference"))
DST = REPO / "baseline" / "_ref"
NEEDED = ["xllmx", "scheduler", "lumina_mgpt/model", "lumina_mgpt/inference_solver.py", "lumina_mgpt/data", "llamagen", "emu3",
          "model_wrappers", "utils.py"]


def main() -> int:
    if not REF.exists():
        print(f"{REF} not found: nothing installed (the GPU box uses the copy made in the build container)")
        return 0
    DST.mkdir(parents=True, exist_ok=True)
    tmp = Path("/tmp/sjd_refcopy")
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp, ignore=shutil.ignore_patterns("*.png", "*.jpg", "*.pth", "*.pt", "assets", ".git"))
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "-q",
                        "--find-links", "/opt/wheelhouse", "--upgrade", "--target", str(DST), str(tmp)],
                       capture_output=True, text=True)
    print("pip install (xllmx):", "ok" if r.returncode == 0 else r.stderr[-400:])
    for rel in NEEDED:
        src, dst = REF / rel, DST / rel
        if not src.exists():
            continue
        if src.is_dir():
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.png", "*.jpg", "*.pth", "*.pt",
                                                                    "*.ckpt", "*.safetensors"))
        else:
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copy2(src, dst)
    n = sum(1 for _ in DST.rglob("*.py"))
    print(f"installed reference into {DST} ({n} python files)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
