"""CPU: the product's PnP arithmetic without a GPU.  boxdreamer_b200/csrc/post.cu writes its solver as __host__ __device__
functions; tests/native/pnp_host_harness.cu (test infrastructure, never linked into the product library) compiles them for
the host.  Checked against (a) the cv2 fixture the GPU test uses (same gates: R within 1e-3 deg, >= 95 % at sigma = 5 px),
(b) numpy's eigen-decomposition for the Rayleigh-shifted inverse iteration, including near-degenerate spectra, and
(c) the Jacobi fallback."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "pnp_host_harness.cu")
OUT = os.path.join(HERE, "_build", "libpnp_host_harness.so")


@pytest.fixture(scope="module")
def harness():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC, os.path.join(ROOT, "boxdreamer_b200", "csrc", "post.cu"), os.path.join(ROOT, "boxdreamer_b200", "csrc", "bd_internal.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = [nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-gencode",
               "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", OUT, SRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    lib = C.CDLL(OUT)
    lib.test_pnp_iterative_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.test_smallest_eigvec12_host.argtypes = [C.c_void_p, C.c_void_p]
    lib.test_jacobi12_host.argtypes = [C.c_void_p, C.c_void_p]
    return lib


def _rot_err_deg(Ra, Rb):
    s = np.minimum(np.linalg.norm(Ra - Rb, axis=(-2, -1)) / (2.0 * np.sqrt(2.0)), 1.0)
    return np.degrees(2.0 * np.arcsin(s))


@pytest.mark.parametrize("tag,min_rate", [("s0", 1.0), ("s2", 1.0), ("s5", 0.95)])
def test_host_pnp_matches_cv2_fixture(harness, tag, min_rate):
    fx = np.load(os.path.join(HERE, "golden", "pnp_cv2.npz"))
    c2 = np.ascontiguousarray(fx[f"corners_{tag}"], dtype=np.float32)
    X3 = np.ascontiguousarray(fx[f"bbox3d_{tag}"], dtype=np.float32)
    Ks = np.ascontiguousarray(fx[f"K_{tag}"], dtype=np.float32)
    n = c2.shape[0]
    poses = np.zeros((n, 4, 4), dtype=np.float32)
    assert harness.test_pnp_iterative_host(c2.ctypes.data, X3.ctypes.data, Ks.ctypes.data, poses.ctypes.data, n, 8, 30) == 0
    R_cv = fx[f"R_{tag}"].astype(np.float64)
    d = _rot_err_deg(poses[:, :3, :3].astype(np.float64), R_cv)
    assert np.mean(d <= 1e-3) >= min_rate, f"{tag}: {np.sort(d)[-4:]}"


def test_shifted_inverse_iteration_vs_numpy(harness):
    rng = np.random.default_rng(5)
    n_ok = 0
    for case in range(300):
        Q, _ = np.linalg.qr(rng.normal(size=(12, 12)))
        lam = np.sort(rng.uniform(0.0, 1.0, size=12)) ** 3
        kind = case % 3
        if kind == 0:
            lam[0] = 1e-14                      # noise-free DLT: the smallest eigenvalue is ~0
        elif kind == 1:
            lam[1] = lam[0] * (1 + 1e-3) + 1e-9  # noisy corners: the two smallest eigenvalues nearly coincide
        A = (Q * lam) @ Q.T
        A = np.ascontiguousarray((A + A.T) / 2)
        x = np.zeros(12)
        ok = harness.test_smallest_eigvec12_host(A.ctypes.data, x.ctypes.data)
        w, V = np.linalg.eigh(A)
        if not ok:
            continue                             # the product then takes the Jacobi path (checked below)
        n_ok += 1
        assert abs(np.linalg.norm(x) - 1) < 1e-12
        resid = np.linalg.norm(A @ x - (x @ A @ x) * x)
        assert resid <= 1e-12 * max(np.trace(A), 1e-300), (case, resid)
        if kind != 1:                            # well separated: must be THE smallest eigenvector
            assert abs(abs(x @ V[:, 0]) - 1) < 1e-8, case
        else:                                    # nearly degenerate pair: anywhere in its 2-D eigenspace is a valid minimiser
            assert np.linalg.norm(x - V[:, :2] @ (V[:, :2].T @ x)) < 1e-6, case
    assert n_ok >= 280


def test_jacobi_fallback_vs_numpy(harness):
    rng = np.random.default_rng(6)
    for _ in range(20):
        M = rng.normal(size=(12, 12))
        A0 = M @ M.T
        A = np.ascontiguousarray(A0.copy())
        V = np.zeros((12, 12))
        harness.test_jacobi12_host(A.ctypes.data, V.ctypes.data)
        w = np.sort(np.diag(A))
        assert np.allclose(w, np.linalg.eigvalsh(A0), rtol=1e-10, atol=1e-12)
        assert np.allclose(V @ np.diag(np.diag(A)) @ V.T, A0, rtol=1e-9, atol=1e-10)
