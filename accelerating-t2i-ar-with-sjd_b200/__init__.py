"""B200-native Speculative Jacobi Decoding hot path (importable as ``sjd_b200``, see /sjd_b200.py).

Only what the hot path needs lives here: ``csrc/`` (sm_100a kernels + the C ABI of include/sjd_b200.h),
``_lib`` (ctypes binding), ``model`` (weight packing + window forward), ``engine`` (the SJD loop) and
``hf_api`` (mirror of the reference's renew_* / solver interface).
"""
__version__ = "0.1.0"
