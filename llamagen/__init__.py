"""Drop-in module path of the reference's llamagen package (its test_llamagen.py does `from llamagen... import ...`).
This directory holds what the SJD hot path replaces: the GPT parameter tree (llamagen.llamagen, so reference
checkpoints load) and the solver (llamagen.llamagen_solver); the PyTorch forward is replaced by the sm_100a engine.

Everything else of the reference's package — `llamagen.tokenizer.tokenizer_image.vq_model` (VQ decoder),
`llamagen.language.t5` (T5 embedder), which test_llamagen.py:17-18 imports — stays the reference's own: this package
extends its `__path__` over every other `llamagen/` directory on sys.path, so with this repository AHEAD of the
reference checkout on sys.path the reference's sub-packages still resolve (submodules present here win)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
