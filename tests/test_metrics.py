"""Device pose metrics (boxdreamer_b200/metrics.py).

CPU: the AUC helpers against the reference's (sklearn-based) functions -- build container only.
GPU: bd_pose_metrics against tests/golden/pose_metrics.npz, generated from the unmodified reference class and the
line-by-line ADD / ADD-S arithmetic (tests/golden/make_golden_metrics.py).  Tolerances: distances 1e-4 relative (fp32
sums over 1500 points); angles 0.02 deg absolute + 1e-3 relative -- both sides take acos of a float32 trace, which
resolves small angles only to ~0.01 deg."""
import os

import numpy as np
import pytest
import torch

from boxdreamer_b200 import metrics as M
from oracle import ref_import

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present")
def test_auc_helpers_equal_reference():
    ref_import.install()
    import importlib
    mu = importlib.import_module("src.lightning.utils.metrics.metric_utils")
    rng = np.random.default_rng(3)
    add = np.abs(rng.normal(size=200)) * 0.06
    proj = np.abs(rng.normal(size=200)) * 25
    assert abs(M.auc_add(add) - mu.auc_add(add)) <= 1e-12
    assert abs(M.auc_proj2d(proj) - mu.auc_proj2d(proj)) <= 1e-12
    assert abs(M.compute_auc_sklearn(add) - mu.compute_auc_sklearn(add)) <= 1e-12


def test_auc_helpers_known_values():
    assert abs(M.auc_add(np.zeros(10)) - 1.0) <= 1e-12           # every error below every threshold
    assert M.auc_add(np.full(10, 1.0)) == 0.0
    assert abs(M.auc_proj2d(np.full(4, 20.0)) - 0.5) <= 2e-3     # step at the middle of [0, 40]


@pytest.mark.gpu
def test_pose_metrics_kernel_matches_reference_fixture():
    fx = np.load(os.path.join(GOLD, "pose_metrics.npz"))
    pts = torch.from_numpy(fx["pts"]).cuda()
    out = M.pose_metrics(torch.from_numpy(fx["pose_pred"]).cuda(), torch.from_numpy(fx["pose_gt"]).cuda(),
                         torch.from_numpy(fx["K"]).cuda(), pts).cpu().numpy().astype(np.float64)
    ref = fx["out"]
    for col, name in ((0, "rotation"), (2, "in-plane")):
        assert np.all(np.abs(out[:, col] - ref[:, col]) <= 0.02 + 1e-3 * np.abs(ref[:, col])), name
    for col, name in ((1, "translation"), (3, "proj2d"), (4, "ADD"), (5, "ADD-S"), (6, "diameter")):
        assert np.all(np.abs(out[:, col] - ref[:, col]) <= 1e-4 * np.abs(ref[:, col]) + 1e-7), (name, out[:3, col], ref[:3, col])
    # one cloud per query (strided) == shared cloud
    B = ref.shape[0]
    out2 = M.pose_metrics(torch.from_numpy(fx["pose_pred"]).cuda(), torch.from_numpy(fx["pose_gt"]).cuda(),
                          torch.from_numpy(fx["K"]).cuda(), pts.unsqueeze(0).expand(B, -1, -1).contiguous()).cpu().numpy()
    assert np.array_equal(out2, out.astype(np.float32))


@pytest.mark.gpu
def test_metrics_class_on_forward_output():
    """compute_metrics on a dict shaped like BoxDreamer.forward's output: identity transform / unit scale, prediction == ground
    truth for sample 0 -> all errors 0 and ADD-0.1d = 1; a shifted prediction for sample 1 -> ADD = |shift|."""
    B, T = 2, 3
    eye = torch.eye(4).repeat(B, T, 1, 1)
    eye[:, :, 2, 3] = 0.6
    pred = eye.clone()
    pred[1, 2, 0, 3] += 0.05
    data = {"query_idx": torch.tensor([2, 2]), "original_poses": eye.cuda(), "pred_poses": pred.cuda(),
            "scale": torch.ones(B, T, 3).cuda(), "coordinate_transform": torch.eye(4).repeat(B, 1, 1).cuda(),
            "original_intrinsics": torch.tensor([[500.0, 0, 320], [0, 500.0, 240], [0, 0, 1]]).repeat(B, T, 1, 1).cuda()}
    pts = (torch.rand(800, 3, generator=torch.Generator().manual_seed(1)) - 0.5) * 0.2
    m = M.Metrics({"t_scale": "m", "metrics_list": []})
    r = m.compute_metrics(data, pts.cuda()).cpu()
    assert float(r[0, :6].abs().max()) == 0.0
    assert abs(float(r[1, 4]) - 0.05) <= 1e-6 and abs(float(r[1, 1]) - 0.05) <= 1e-6
    assert m.metrics_result["ADD_0.1d_0"] == [1.0, 0.0] and abs(m.metrics_result["t_errs_0"][1] - 5.0) <= 1e-4
    assert float(r[1, 5]) <= float(r[1, 4])                       # nearest neighbour can only be closer
