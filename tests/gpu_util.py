"""Helpers shared by the -m gpu tests: raw-pointer calls through the C ABI and error reporting."""
import ctypes as C

import torch

from boxdreamer_b200 import _lib


def sp():
    return _lib.stream_ptr()


def report(name, got, ref, tol_abs=None, tol_rel=None):
    """Returns (ok, message) with max abs / scaled error and the worst index."""
    got = got.float().cpu()
    ref = ref.float().cpu()
    diff = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    mx = diff.max().item()
    idx = int(diff.reshape(-1).argmax())
    nan = int(torch.isnan(got).sum())
    msg = (f"{name}: max|d|={mx:.3e} scaled={mx / scale:.3e} ref_absmax={scale:.3e} worst_flat_idx={idx} "
           f"got={got.reshape(-1)[idx].item():.6g} ref={ref.reshape(-1)[idx].item():.6g} nan={nan} shape={tuple(got.shape)}")
    ok = nan == 0
    if tol_abs is not None:
        ok = ok and mx <= tol_abs
    if tol_rel is not None:
        ok = ok and mx / scale <= tol_rel
    return ok, msg


def gemm(A, W, bias, M, N, K, epi, precision, gamma=None, out=None):
    lib = _lib.load()
    dev = A.device
    if out is None:
        if epi in (_lib.EPI_F32, _lib.EPI_RESID):
            out = torch.zeros(M, N, device=dev, dtype=torch.float32)
        else:
            out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if precision == _lib.PRECISION_BF16 else torch.float32)
    _lib.check(lib.bd_gemm(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(out), M, N, K, epi, precision,
                           sp()), "bd_gemm")
    torch.cuda.synchronize()
    return out
