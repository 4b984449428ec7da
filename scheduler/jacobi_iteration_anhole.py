"""Drop-in for the reference module of the same path (Anole adaptor of the SJD plugin API: HF
ChameleonForConditionalGeneration as the pipeline), served by the sm_100a engine — see
accelerating-t2i-ar-with-sjd_b200/hf_api.py for the mapping to reference lines."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sjd_b200  # noqa: E402,F401
from sjd_b200.hf_api import (  # noqa: E402,F401
    renew_backbone, renew_backbone_adapt_anole, renew_pipeline_anole, renew_sampler, renew_vocabulary_mapping)
from sjd_b200.hf_api import renew_pipeline_sampler_anole as renew_pipeline_sampler  # noqa: E402,F401
