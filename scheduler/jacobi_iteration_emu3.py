"""Drop-in for the reference module of the same path (Emu3 adaptor of the SJD plugin API), served by the sm_100a
engine — see accelerating-t2i-ar-with-sjd_b200/hf_api.py for the mapping to reference lines."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sjd_b200  # noqa: E402,F401
from sjd_b200.hf_api import (  # noqa: E402,F401
    renew_end_of_line_logit_processor_3d, renew_sampler_forward, renew_solver)
