"""ctypes binding of include/boxdreamer_b200.h.  No torch types cross this boundary: only raw
pointers (`tensor.data_ptr()`), sizes and the raw cudaStream_t.

The product path fails loudly when the CUDA library is missing; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BD_LIB_PATH") or os.path.join(HERE, "libboxdreamer_b200.so")  # BD_LIB_PATH: A/B builds (scripts/build_variant.sh)

BD_F32, BD_BF16 = 0, 1
PRECISION_EXACT, PRECISION_BF16 = 0, 1
EPI_F32, EPI_GELU, EPI_RESID, EPI_QKV, EPI_ACT = 0, 1, 2, 3, 4
PROF_CATS = ["gemm_qkv", "attention", "gemm_proj", "gemm_fc1", "gemm_fc2", "gemm_other", "layernorm", "glue", "topk", "pnp", "attention_dino", "attention_window"]

# every symbol include/boxdreamer_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "bd_last_error", "bd_version", "bd_create", "bd_destroy", "bd_load_weight", "bd_finalize_weights",
    "bd_dino_forward", "bd_decoder_forward", "bd_corners_topk", "bd_pnp", "bd_forward", "bd_forward_packed", "bd_forward_host", "bd_forward_host_submit", "bd_forward_host_wait",
    "bd_make_bbox_features", "bd_forward_host_px", "bd_pose_metrics",
    "bd_gemm", "bd_qkv_project", "bd_attention", "bd_layernorm", "bd_launch_count", "bd_profile_enable", "bd_profile_read", "bd_debug_attention_trace",
]


class BdConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "img_size", "patch_size", "d_model", "dec_layers", "dec_heads", "dino_layers", "dino_heads",
        "dino_registers", "dino_pretrain_grid", "precision", "attn_variant", "max_batch", "max_views")]


class BdPnpOpts(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n_hyp", C.c_int32), ("thr_px", C.c_float), ("seed", C.c_uint32),
                ("max_iter", C.c_int32)]


class BoxDreamerLibError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = False) -> C.CDLL:
    """Loads the C-ABI library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build
            _build.build()
        else:
            raise BoxDreamerLibError(
                f"{LIB_PATH} is missing: build it with `python -m boxdreamer_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    lib.bd_last_error.restype = C.c_char_p
    lib.bd_last_error.argtypes = []
    lib.bd_version.restype = C.c_int
    lib.bd_create.argtypes = [C.POINTER(vp), C.POINTER(BdConfig)]
    lib.bd_destroy.argtypes = [vp]
    lib.bd_load_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_int64), i32]
    lib.bd_finalize_weights.argtypes = [vp]
    lib.bd_dino_forward.argtypes = [vp, vp, i32, vp, i32, vp]
    lib.bd_decoder_forward.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, i32, vp]
    lib.bd_corners_topk.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
    lib.bd_pnp.argtypes = [vp, vp, vp, vp, vp, C.POINTER(BdPnpOpts), i32, i32, vp]
    lib.bd_forward.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(BdPnpOpts), i32, i32, vp]
    lib.bd_forward_host_submit.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(BdPnpOpts), i32, i32]
    lib.bd_forward_host_wait.argtypes = [vp, i32]
    lib.bd_forward_packed.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, C.POINTER(BdPnpOpts), i32, i32, vp]
    lib.bd_forward_host.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(BdPnpOpts), i32, i32]
    lib.bd_forward_host_px.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(BdPnpOpts), i32, i32]
    lib.bd_make_bbox_features.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.bd_pose_metrics.argtypes = [vp, vp, vp, vp, C.c_int64, vp, i32, i32, vp]
    lib.bd_gemm.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.bd_qkv_project.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.bd_attention.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, i32, vp]
    lib.bd_layernorm.argtypes = [vp, vp, vp, f32, vp, vp, i32, i32, vp]
    lib.bd_debug_attention_trace.argtypes = [vp]
    lib.bd_launch_count.argtypes = [vp]
    lib.bd_profile_enable.argtypes = [vp, i32]
    lib.bd_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), i32]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("bd_last_error", "bd_launch_count"):
            fn.restype = C.c_int
    lib.bd_launch_count.restype = C.c_longlong
    _lib = lib
    return lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().bd_last_error()
        raise BoxDreamerLibError(f"{what} failed (bd_status {code}): {msg.decode() if msg else ''}")


def ptr(t) -> C.c_void_p:
    """Raw device/host pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_contiguous(), "boxdreamer_b200: tensors crossing the C ABI must be contiguous"
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None) -> C.c_void_p:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
