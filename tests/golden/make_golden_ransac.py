"""Golden vectors for the pooled robust PnP of the dense multi-round path (recover_pose_from_dense_bb8,
src/models/utils/box_utils.py:202-304): the reference's own call
    cv2.solvePnPRansac(pts3d, pts2d, K, None, reprojectionError=2.0, confidence=0.99, flags=SOLVEPNP_ITERATIVE,
                       iterationsCount=1000)
on pooled proposals (n_sub x 8 corners of one box, independent 0.5 px noise per proposal, quantised to the 0.05 px grid of
top-20 means) with 0 / 4 / 8 gross outliers.  OpenCV is third-party (pinned 4.11.0.86 in the reference, 4.13.0 here).

    python tests/golden/make_golden_ransac.py      # writes tests/golden/pnp_ransac_cv2.npz
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from boxdreamer_b200 import synth  # noqa: E402

cv2.setNumThreads(1)
cv2.setRNGSeed(0)
out = {}
for tag, n_sub, n_out in (("p2o0", 2, 0), ("p4o0", 4, 0), ("p4o4", 4, 4), ("p4o8", 4, 8), ("p8o8", 8, 8)):
    n_case = 48
    _, X3, Ks, gt = synth.synth_pnp_cases(n_case, 0.0, seed=700 + n_sub * 10 + n_out)
    rng = np.random.Generator(np.random.PCG64(900 + n_sub * 10 + n_out))
    P2, P3, Rs, ts, inl = [], [], [], [], []
    for i in range(n_case):
        R, t = gt[i][:, :3], gt[i][:, 3]
        cam = X3[i] @ R.T + t
        uv = (cam / cam[:, 2:3]) @ Ks[i].T
        pts2 = np.concatenate([uv[:, :2] + rng.normal(0, 0.5, size=(8, 2)) for _ in range(n_sub)])
        pts2 = np.round(pts2 * 20) / 20
        pts3 = np.tile(X3[i], (n_sub, 1))
        bad = rng.choice(n_sub * 8, size=n_out, replace=False)
        pts2[bad] = rng.uniform(0, 224, size=(n_out, 2))
        ok, rvec, tvec, inliers = cv2.solvePnPRansac(pts3.astype(np.float32), pts2.astype(np.float32), Ks[i].astype(np.float32), None,
                                                     reprojectionError=2.0, confidence=0.99, flags=cv2.SOLVEPNP_ITERATIVE,
                                                     iterationsCount=1000)
        assert ok
        m = np.zeros(n_sub * 8, dtype=bool)
        m[inliers.reshape(-1)] = True
        P2.append(pts2); P3.append(pts3); Rs.append(cv2.Rodrigues(rvec)[0]); ts.append(tvec.reshape(3)); inl.append(m)
    out[f"pts2d_{tag}"] = np.stack(P2).astype(np.float32)
    out[f"pts3d_{tag}"] = np.stack(P3).astype(np.float32)
    out[f"K_{tag}"] = Ks.astype(np.float32)
    out[f"R_{tag}"] = np.stack(Rs)
    out[f"t_{tag}"] = np.stack(ts)
    out[f"inliers_{tag}"] = np.stack(inl)
    out[f"gt_{tag}"] = gt
    print(tag, "mean inliers", np.stack(inl).sum(1).mean(), "of", n_sub * 8)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pnp_ransac_cv2.npz"), **out)
