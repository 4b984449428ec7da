"""ctypes binding of libsjd_b200.so (C ABI declared in include/sjd_b200.h).

The product path has no fallback: if the shared library is missing or a symbol is absent, importing
the engine raises.  PyTorch is used only for device memory and streams; pointers cross the boundary
as integers (``tensor.data_ptr()``).
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libsjd_b200.so"
HEADER = REPO_ROOT / "include" / "sjd_b200.h"
CSRC = PKG_DIR / "csrc"

SJD_MAX_ROWS = 8
SJD_MAX_TOKENS = 256

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/sjd_b200.cu for sm_100a into libsjd_b200.so (in-tree, travels to the GPU box)."""
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [HEADER]
    if LIB_PATH.exists() and not force:
        newest = max(p.stat().st_mtime for p in srcs)
        if LIB_PATH.stat().st_mtime >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC / "sjd_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


def header_symbols() -> list[str]:
    """Every function name declared in include/sjd_b200.h (used by the export test)."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sjd_[a-z0-9_]+)\s*\(", text)))


class VerifyArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("W", C.c_int32), ("V", C.c_int32),
        ("has_uncond", C.c_int32), ("apply_cfg", C.c_int32),
        ("guidance", C.c_float), ("temperature", C.c_float),
        ("allow_lo", C.c_int32), ("allow_hi", C.c_int32),
        ("forced", C.c_void_p), ("forced_resid", C.c_void_p), ("top_k", C.c_int32), ("top_p_thresh", C.c_float), ("do_sample", C.c_int32), ("scheme", C.c_int32),
        ("draft", C.c_void_p), ("q_row", C.c_void_p), ("p_prev", C.c_void_p), ("p_cur", C.c_void_p),
        ("noise_e1", C.c_void_p), ("noise_u", C.c_void_p), ("noise_e2", C.c_void_p),
        ("eoi_token", C.c_int32), ("text_top_k", C.c_int32),
        ("resid", C.c_void_p), ("next_tokens", C.c_void_p), ("out_tokens", C.c_void_p), ("out_info", C.c_void_p),
        ("sync_ws", C.c_void_p),
        ("rng_mode", C.c_int32), ("rng_seed", C.c_uint64), ("rng_off", C.c_uint64 * 3), ("rng_span", C.c_uint32 * 3),
        ("allow_mode", C.c_int32), ("ban", C.c_int32 * 2),
        ("resid_set", C.c_int32), ("resid_allow_mode", C.c_int32), ("resid_allow_lo", C.c_int32),
        ("resid_allow_hi", C.c_int32), ("resid_ban", C.c_int32 * 2), ("resid_from", C.c_int32),
        ("done_flag", C.c_void_p), ("done_seq", C.c_int32),
    ]

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.ban[0] = self.ban[1] = self.resid_ban[0] = self.resid_ban[1] = -1      # "no extra removed id"


class ModelCfg(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32), ("d_model", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("d_ff", C.c_int32), ("vocab", C.c_int32), ("rms_eps", C.c_float),
        ("qk_norm", C.c_int32), ("rope_interleaved", C.c_int32), ("rows", C.c_int32), ("max_len", C.c_int32),
        ("n_rope_pos", C.c_int32), ("logits_round_bf16", C.c_int32),
    ]


class LayerWeights(C.Structure):
    _fields_ = [
        ("attn_norm", C.c_void_p), ("wqkv", C.c_void_p),
        ("q_norm_w", C.c_void_p), ("q_norm_b", C.c_void_p), ("k_norm_w", C.c_void_p), ("k_norm_b", C.c_void_p),
        ("wo", C.c_void_p), ("ffn_norm", C.c_void_p), ("w_gate_up", C.c_void_p), ("w_down", C.c_void_p),
    ]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("W", C.c_int32), ("ids", C.c_void_p), ("embeds", C.c_void_p),
        ("rope_pos", C.c_void_p), ("cache_pos", C.c_void_p),
        ("kv_len", C.c_int32), ("kv_lo", C.c_int32 * SJD_MAX_ROWS),
        ("n_logit_tokens", C.c_int32), ("logits", C.c_void_p),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the shared library, declaring argument types.  Raises if it is missing (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found — run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "the SJD hot path has no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    L.sjd_version.restype = C.c_int
    L.sjd_last_error.restype = C.c_char_p
    L.sjd_device_sm_count.restype = C.c_int
    L.sjd_launch_count.restype = C.c_uint64
    L.sjd_debug_attn_stamps.restype = None
    L.sjd_debug_attn_stamps.argtypes = [C.c_void_p]
    L.sjd_debug_gemm_stamps.restype = None
    L.sjd_debug_gemm_stamps.argtypes = [C.c_void_p, C.c_int]
    L.sjd_gemm_workspace_bytes.restype = C.c_size_t
    L.sjd_gemm_workspace_bytes.argtypes = [C.c_int] * 4
    L.sjd_gemm_bf16.restype = C.c_int
    L.sjd_gemm_bf16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.sjd_verify.restype = C.c_int
    L.sjd_verify.argtypes = [C.POINTER(VerifyArgs), C.c_void_p]
    L.sjd_debug_philox.restype = C.c_int
    L.sjd_debug_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
    L.sjd_debug_attn_sw_split.restype = C.c_int
    L.sjd_debug_attn_sw_split.argtypes = [C.c_int] * 5 + [C.c_void_p] + [C.c_int] * 3 + [C.c_void_p, C.c_void_p]
    L.sjd_vq_lookup.restype = C.c_int
    L.sjd_vq_lookup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_void_p, C.c_void_p]
    L.sjd_ctx_create.restype = C.c_int
    L.sjd_ctx_create.argtypes = [C.POINTER(ModelCfg), C.POINTER(C.c_void_p)]
    L.sjd_ctx_destroy.restype = None
    L.sjd_ctx_destroy.argtypes = [C.c_void_p]
    L.sjd_ctx_device_bytes.restype = C.c_size_t
    L.sjd_ctx_device_bytes.argtypes = [C.c_void_p]
    L.sjd_ctx_set_layer.restype = C.c_int
    L.sjd_ctx_set_layer.argtypes = [C.c_void_p, C.c_int, C.POINTER(LayerWeights)]
    L.sjd_ctx_set_globals.restype = C.c_int
    L.sjd_ctx_set_globals.argtypes = [C.c_void_p] * 6
    L.sjd_ctx_forward.restype = C.c_int
    L.sjd_ctx_forward.argtypes = [C.c_void_p, C.POINTER(ForwardArgs), C.c_void_p]
    L.sjd_ctx_gemm_only.restype = C.c_int
    L.sjd_ctx_gemm_only.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().sjd_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed with SJD code {rc}: {msg}")
