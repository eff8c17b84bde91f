"""Latency of the PnP kernel at the batch size of one forward pass (B = 64): clean / noisy / garbage corners, LM cap sweep."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import _lib, synth
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64

def t(c2, X3, Ks, opts, n=20):
    poses = torch.empty(B, 4, 4, device="cuda")
    o = C.byref(opts) if opts is not None else None
    for _ in range(3):
        _lib.check(lib.bd_pnp(None, _lib.ptr(c2), _lib.ptr(X3), _lib.ptr(Ks), _lib.ptr(poses), o, B, 8, None))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        _lib.check(lib.bd_pnp(None, _lib.ptr(c2), _lib.ptr(X3), _lib.ptr(Ks), _lib.ptr(poses), o, B, 8, None))
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for name, sigma in (("clean", 0.0), ("sigma2", 2.0), ("sigma5", 5.0), ("garbage", -1.0)):
    c2, X3, Ks, gt = synth.synth_pnp_cases(B, max(sigma, 0.0), seed=99)
    if sigma < 0:
        c2 = np.random.default_rng(0).uniform(0, 224, size=c2.shape)
    c2c, X3c, Ksc = (torch.from_numpy(np.ascontiguousarray(x.astype(np.float32))).cuda() for x in (c2, X3, Ks))
    row = [f"{name:8s}"]
    for it in (0, 1, 5, 30):
        row.append(f"lm{it}: {t(c2c, X3c, Ksc, _lib.BdPnpOpts(0, 0, 2.0, 0, it)) * 1e3:7.1f} us")
    print("  ".join(row))
