"""Drop-in for the reference module of the same path: the [B, W, V] logits processors, as parameter holders whose
arithmetic runs inside sjd_verify on the GPU."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sjd_b200  # noqa: E402,F401
from sjd_b200.hf_api import (  # noqa: E402,F401
    AllowOnlyTokensAtRelativeOffsetLogitsProcessor3d, AllowOnlyTokensInRelativeWindowLogitsProcessor3d,
    MultiTokensInterleavedTopKLogitsWarper, MultiTokensVLLogitsProcessor, SuppressTokensAtBeginLogitsProcessor3d,
    SuppressTokensInIndexRangeLogitsProcessor3d, SuppressTokensLogitsProcessor3d, TopPLogitsWarper3d)
