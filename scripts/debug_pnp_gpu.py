"""GPU debug aid: per-query PnP errors against the cv2 fixture, single-thread launches vs batched."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import _lib
lib = _lib.load()
fx = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "pnp_cv2.npz"))
def rot_err(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.degrees(np.arccos(np.clip(c, -1, 1))))
def run(c2, X3, Ks):
    n = c2.shape[0]
    poses = torch.full((n, 4, 4), -7.0, device="cuda")
    _lib.check(lib.bd_pnp(None, _lib.ptr(c2), _lib.ptr(X3), _lib.ptr(Ks), _lib.ptr(poses), None, n, 8, None))
    torch.cuda.synchronize()
    return poses.cpu().numpy().astype(np.float64)
for tag in ("s0", "s2"):
    c2 = torch.from_numpy(fx[f"corners_{tag}"]).cuda(); X3 = torch.from_numpy(fx[f"bbox3d_{tag}"]).cuda(); Ks = torch.from_numpy(fx[f"K_{tag}"]).cuda()
    P = run(c2, X3, Ks)
    errs = [rot_err(P[i, :3, :3], fx[f"R_{tag}"][i]) for i in range(64)]
    print(tag, "batched errs:", " ".join(f"{e:.1e}" for e in errs))
    print(tag, "pose0 batched\n", P[0], "\nref R\n", fx[f"R_{tag}"][0], fx[f"t_{tag}"][0])
    single = []
    for i in range(8):
        Pi = run(c2[i:i+1].contiguous(), X3[i:i+1].contiguous(), Ks[i:i+1].contiguous())
        single.append(rot_err(Pi[0, :3, :3], fx[f"R_{tag}"][i]))
    print(tag, "single-launch errs:", " ".join(f"{e:.1e}" for e in single))
