"""Golden vectors for the device pose metrics (bd_pose_metrics) from the reference's own arithmetic
(src/lightning/utils/metrics/metric_utils.py): Metrics.query_pose_error and Metrics.project are CALLED on the unmodified
class; ADD / ADD-S / diameter follow process_single_bs_add (:370-392) line by line (that method itself reads CAD files).

    python tests/golden/make_golden_metrics.py      # writes tests/golden/pose_metrics.npz  (build container only)
"""
import os
import sys

import numpy as np
from scipy import spatial

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

ref_import.install()
import importlib  # noqa: E402

mu = importlib.import_module("src.lightning.utils.metrics.metric_utils")


class Cfg:
    metrics_list = []
    t_scale = "m"


M = mu.Metrics(Cfg())
rng = np.random.Generator(np.random.PCG64(2024))
B, N = 24, 1500
pts = (rng.normal(size=(N, 3)) * np.array([0.06, 0.04, 0.09])).astype(np.float32)


def rand_rot(scale):
    w = rng.normal(size=3) * scale
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx


gt, pred, Ks, out = [], [], [], []
for b in range(B):
    Rg = rand_rot(1.5)
    tg = np.array([rng.normal() * 0.05, rng.normal() * 0.05, 0.5 + rng.uniform() * 0.4])
    Rp = rand_rot([0.002, 0.05, 0.6][b % 3]) @ Rg
    tp = tg + rng.normal(size=3) * [0.001, 0.01, 0.05][b % 3]
    Pg = np.concatenate([Rg, tg[:, None]], 1).astype(np.float32)
    Pp = np.concatenate([Rp, tp[:, None]], 1).astype(np.float32)
    f = 600 + rng.uniform() * 200
    K = np.array([[f, 0, 320], [0, f, 240], [0, 0, 1]], dtype=np.float32)
    ang, te, inpl = M.query_pose_error(Pp, Pg)
    p2 = M.project(pts, K, Pp) - M.project(pts, K, Pg)
    proj = np.mean(np.linalg.norm(p2, axis=1))
    model_pred = (pts @ Pp[:, :3].T) + Pp[:, 3]
    model_gt = (pts @ Pg[:, :3].T) + Pg[:, 3]
    adds, _ = spatial.cKDTree(model_pred).query(model_gt, k=1)
    add = np.mean(np.linalg.norm(model_pred - model_gt, axis=-1))
    diam = np.linalg.norm(np.max(pts, axis=0) - np.min(pts, axis=0))
    gt.append(Pg); pred.append(Pp); Ks.append(K)
    out.append([ang, te / 100.0, inpl, proj, add, np.mean(adds), diam, 0.0])   # te: query_pose_error returns cm for t_scale "m"
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pose_metrics.npz"), pts=pts, pose_gt=np.stack(gt), pose_pred=np.stack(pred),
                    K=np.stack(Ks), out=np.array(out, dtype=np.float64))
print(np.array(out)[:4])
