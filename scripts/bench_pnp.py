"""BASELINE config 5: PnP kernel in isolation -- N queries, corner noise sigma in {0, 2, 5} px, both modes;
accuracy vs ground truth and vs cv2.solvePnP(ITERATIVE) / cv2.solvePnPRansac on a subsample (cv2 timed on one core)."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import _lib, synth
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
NSUB = 2000

def rot_err(Ra, Rb):
    s = np.minimum(np.linalg.norm(Ra - Rb, axis=(-2, -1)) / (2 * np.sqrt(2)), 1.0)
    return np.degrees(2 * np.arcsin(s))

def gpu_solve(c2, X3, Ks, opts):
    n = c2.shape[0]
    poses = torch.empty(n, 4, 4, device="cuda")
    o = C.byref(opts) if opts is not None else None
    for _ in range(2):
        _lib.check(lib.bd_pnp(None, _lib.ptr(c2), _lib.ptr(X3), _lib.ptr(Ks), _lib.ptr(poses), o, n, 8, None))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        _lib.check(lib.bd_pnp(None, _lib.ptr(c2), _lib.ptr(X3), _lib.ptr(Ks), _lib.ptr(poses), o, n, 8, None))
    b.record(); torch.cuda.synchronize()
    return poses.cpu().numpy().astype(np.float64), a.elapsed_time(b) / 3

out = {"n_queries": N}
try:
    import cv2
    cv2.setNumThreads(1)
except Exception:
    cv2 = None
base = synth.synth_pnp_cases(4096, 0.0, seed=4321)   # geometry pool (python loop is slow): tile it, re-noise per sigma
for sigma in (0.0, 2.0, 5.0):
    c2s, X3s, Kss, gts = synth.synth_pnp_cases(4096, sigma, seed=4321 + int(sigma))
    rep = (N + 4095) // 4096
    c2 = np.tile(c2s, (rep, 1, 1))[:N].copy(); X3 = np.tile(X3s, (rep, 1, 1))[:N]; Ks = np.tile(Kss, (rep, 1, 1))[:N]; gt = np.tile(gts, (rep, 1, 1))[:N]
    if sigma > 0:  # fresh noise for the tiled copies, quantised to the 0.05 px grid of top-20 means
        rng = np.random.Generator(np.random.PCG64(int(sigma * 10)))
        c2[4096:] = np.round((c2[4096:] + rng.normal(0, sigma * 0.3, size=c2[4096:].shape)) * 20) / 20
    c2c, X3c, Ksc = (torch.from_numpy(np.ascontiguousarray(x.astype(np.float32))).cuda() for x in (c2, X3, Ks))
    res = {}
    for name, opts in (("mode0_iterative", None), ("mode1_hyp154", _lib.BdPnpOpts(1, 154, 2.0, 0, 30)), ("mode1_hyp512", _lib.BdPnpOpts(1, 512, 2.0, 0, 30))):
        P, ms = gpu_solve(c2c, X3c, Ksc, opts)
        e = rot_err(P[:, :3, :3], gt[:, :, :3]); te = np.linalg.norm(P[:, :3, 3] - gt[:, :, 3], axis=1)
        nh = 1 if opts is None else opts.n_hyp
        res[name] = {"ms": round(ms, 3), "queries_per_s": round(N / ms * 1e3), "hypotheses_per_s": round(N * nh / ms * 1e3),
                     "rot_err_deg_median": float(np.median(e)), "rot_err_deg_p95": float(np.percentile(e, 95)), "t_err_median_m": float(np.median(te))}
        if name == "mode0_iterative":
            P0 = P
    if cv2 is not None:
        t0 = time.perf_counter(); Rs = []
        for i in range(NSUB):
            ok, rvec, tvec = cv2.solvePnP(X3[i].astype(np.float32), c2[i].astype(np.float32), Ks[i].astype(np.float32), None, flags=cv2.SOLVEPNP_ITERATIVE)
            Rs.append(cv2.Rodrigues(rvec)[0])
        dt = time.perf_counter() - t0
        d = rot_err(P0[:NSUB, :3, :3], np.stack(Rs))
        res["cv2_solvePnP_iterative_1core"] = {"queries_per_s": round(NSUB / dt), "gpu_mode0_vs_cv2_within_1e-3deg": float(np.mean(d <= 1e-3)),
                                               "rot_err_deg_median_vs_gt": float(np.median(rot_err(np.stack(Rs), gt[:NSUB, :, :3])))}
        t0 = time.perf_counter(); nok = 0
        for i in range(200):
            ok, rvec, tvec, inl = cv2.solvePnPRansac(X3[i].astype(np.float32), c2[i].astype(np.float32), Ks[i].astype(np.float32), None,
                                                     iterationsCount=512, reprojectionError=1.0, confidence=0.99, flags=cv2.SOLVEPNP_ITERATIVE)
            nok += int(ok)
        res["cv2_solvePnPRansac_512it_1core"] = {"queries_per_s": round(200 / (time.perf_counter() - t0)), "success": nok / 200}
    out[f"sigma_{sigma:g}px"] = res
print(json.dumps(out, indent=1))
