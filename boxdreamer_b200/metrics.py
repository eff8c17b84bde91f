"""Evaluation metrics on the device: the step right after the hot path in `test_step`
(src/lightning/BoxDreamer_lightning_model.py:230-243).  The reference copies the whole batch to the CPU, deep-copies it and
walks it query by query with a thread pool and a cKDTree (src/lightning/utils/metrics/metric_utils.py); here the per-query
numbers come from one kernel launch (`bd_pose_metrics`) on device tensors, only the [B,8] result is read back.

Mirrors (reference file:line), same names and argument meaning
  Metrics.query_pose_error          metric_utils.py:162-210   (batched)
  Metrics.projection_2d_error_mp    metric_utils.py:255-329   (-> proj2D_metric)
  Metrics.add_metric_mp             metric_utils.py:331-448   (-> ADD_raw, ADDs_raw, ADD_0.1d, ADDs_0.1d)
  auc_add / auc_proj2d / compute_auc_sklearn   metric_utils.py:770-803
CAD-model point sampling (`get_cached_points`, file I/O) stays with the caller: pass the model points as a tensor.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["Metrics", "pose_metrics", "auc_add", "auc_proj2d", "compute_auc_sklearn"]


def pose_metrics(pose_pred: torch.Tensor, pose_gt: torch.Tensor, K: torch.Tensor, model_pts: torch.Tensor) -> torch.Tensor:
    """pose_pred / pose_gt [B,3|4,4], K [B,3,3], model_pts [N,3] (shared) or [B,N,3] -> [B,8] fp32 on the device:
    rot_deg, trans_norm, inplane_deg, proj2d_mean_px, add_mean, adds_mean, diameter, 0."""
    if not pose_pred.is_cuda:
        raise _lib.BoxDreamerLibError("pose_metrics: inputs must be CUDA tensors (no CPU fallback)")
    B = pose_pred.shape[0]
    pp = pose_pred[:, :3, :].float().contiguous()
    pg = pose_gt[:, :3, :].float().contiguous()
    Kc = K.float().contiguous()
    pts = model_pts.float().contiguous()
    N = pts.shape[-2]
    stride = 0 if pts.dim() == 2 else 3 * N
    out = torch.empty(B, 8, device=pose_pred.device, dtype=torch.float32)
    with torch.cuda.device(pose_pred.device):
        _lib.check(_lib.load().bd_pose_metrics(_lib.ptr(pp), _lib.ptr(pg), _lib.ptr(Kc), _lib.ptr(pts), stride, _lib.ptr(out), B, N,
                                               _lib.stream_ptr()), "bd_pose_metrics")
    return out


def _trapz_auc(x: np.ndarray, y: np.ndarray) -> float:
    """sklearn.metrics.auc for increasing x: the trapezoidal rule."""
    return float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) * 0.5))


def auc_add(metrics) -> float:
    thresholds = np.linspace(0.0, 0.10, 1000)
    results = np.asarray(metrics)
    acc = np.array([(results <= t).sum() / len(results) for t in thresholds])
    return _trapz_auc(thresholds, acc) / (thresholds.max() - thresholds.min())


def auc_proj2d(metrics) -> float:
    thresholds = np.linspace(0, 40.0, 1000)
    results = np.asarray(metrics)
    acc = np.array([(results <= t).sum() / len(results) for t in thresholds])
    return _trapz_auc(thresholds, acc) / (thresholds.max() - thresholds.min())


def compute_auc_sklearn(errs, max_val=0.1, step=0.001) -> float:
    errs = np.sort(np.array(errs))
    X = np.arange(0, max_val + step, step)
    Y = np.ones(len(X))
    for i, x in enumerate(X):
        y = (errs <= x).sum() / len(errs)
        Y[i] = y
        if y >= 1:
            break
    return _trapz_auc(X, Y) / (max_val * 1)


class Metrics:
    """Per-batch pose metrics with the reference's result keys; `metrics_config.t_scale` in {"m", "mm", other} scales the
    translation error to centimetres exactly as metric_utils.py:179-183."""

    def __init__(self, metrics_config=None):
        assert metrics_config is not None, "Metrics config is None!"
        self.metrics_config = metrics_config
        self.metrics_result = {}

    def reset(self):
        self.metrics_result = {}

    def _t_scale(self) -> float:
        ts = self.metrics_config["t_scale"] if isinstance(self.metrics_config, dict) else self.metrics_config.t_scale
        return 100.0 if ts == "m" else (0.1 if ts == "mm" else 1.0)

    def query_pose_error(self, pose_pred: torch.Tensor, pose_gt: torch.Tensor):
        """Batched: -> (angular distance deg [B], translation error cm [B], in-plane rotation error deg [B])."""
        B = pose_pred.shape[0]
        eye = torch.eye(3, device=pose_pred.device).expand(B, 3, 3)
        one = torch.zeros(1, 3, device=pose_pred.device)
        r = pose_metrics(pose_pred, pose_gt, eye, one)
        return r[:, 0], r[:, 1] * self._t_scale(), r[:, 2]

    @staticmethod
    def _query_rows(data, key):
        idx = data["query_idx"].to(data[key].device)
        return data[key][torch.arange(idx.shape[0], device=idx.device), idx]

    def compute_metrics(self, data: dict, model_pts: torch.Tensor, dataloader_id: int = 0) -> dict:
        """data: the dict BoxDreamer.forward returned (device tensors); model_pts [N,3] or [B,N,3] CAD-model points.
        Appends to metrics_result under the reference's keys and returns the per-query tensor [B,8]."""
        pose_gt = self._query_rows(data, "original_poses").float()
        pose_pred = self._query_rows(data, "pred_poses").float().clone()
        scale = self._query_rows(data, "scale").float()
        K = self._query_rows(data, "original_intrinsics").float()
        pose_pred[:, :3, 3] *= scale                                   # metric_utils.py:280, 366
        pose_pred = pose_pred @ data["coordinate_transform"].float()   # :281, 367
        r = pose_metrics(pose_pred, pose_gt, K, model_pts)
        h = r.cpu().numpy()
        thr = 0.1 * h[:, 6]
        res = self.metrics_result
        res.setdefault(f"R_errs_{dataloader_id}", []).extend(h[:, 0].tolist())
        res.setdefault(f"t_errs_{dataloader_id}", []).extend((h[:, 1] * self._t_scale()).tolist())
        res.setdefault(f"proj2D_metric_{dataloader_id}", []).extend(h[:, 3].tolist())
        res.setdefault(f"ADD_raw_{dataloader_id}", []).extend(h[:, 4].tolist())
        res.setdefault(f"ADDs_raw_{dataloader_id}", []).extend(h[:, 5].tolist())
        res.setdefault(f"ADD_0.1d_{dataloader_id}", []).extend((h[:, 4] < thr).astype(np.float64).tolist())
        res.setdefault(f"ADDs_0.1d_{dataloader_id}", []).extend((h[:, 5] < thr).astype(np.float64).tolist())
        return r
