"""Device-side input synthesis: the reference builds the 8-corner heat maps of every view on
the CPU in the dataset (`make_bbox_features`, src/datasets/utils/base/bbox_utils.py:263-303, called from
src/datasets/base.py:689-693) and uploads 8*S*S values per view; here the caller ships the 16 projected corner
coordinates and the maps are rasterised on the GPU (`bd_make_bbox_features`).

Mirrors (same names and argument meaning)
  make_bbox_features(bbox, type, shape)   datasets/utils/base/bbox_utils.py:218-303   (type="heatmap" only)
  make_proj_bbox(pose, intrinsic, bbox)   datasets/utils/base/camera_utils.py:62-84
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["make_bbox_features", "make_proj_bbox"]


def make_proj_bbox(pose: torch.Tensor, intrinsic: torch.Tensor, bbox: torch.Tensor) -> torch.Tensor:
    """pose [L,4,4] (world->camera), intrinsic [L,3,3], bbox [8,3] or [L,8,3] -> projected corners [L,8,2] in pixels:
    x = K (R X + t), divided by depth (reproj_pytorch, camera_utils.py:9-59).  fp32, batched, on the inputs' device."""
    pose, intrinsic, bbox = pose.float(), intrinsic.float(), bbox.float()
    L = pose.shape[0]
    if bbox.dim() == 2:
        bbox = bbox.unsqueeze(0).expand(L, 8, 3)
    cam = torch.matmul(bbox, pose[:, :3, :3].transpose(1, 2)) + pose[:, None, :3, 3]
    uvw = torch.matmul(cam, intrinsic.transpose(1, 2))
    return uvw[..., :2] / uvw[..., 2:3]


def make_bbox_features(bbox: torch.Tensor, type: str = "heatmap", shape=None, dtype: torch.dtype = torch.float32,
                       group: int | None = None) -> torch.Tensor:
    """bbox [L,8,2] projected corners in crop pixels (CUDA tensor) -> heat maps [L,8,H,W] in [-1,1], H == W.
    Like the reference, corner i's maps are normalised by their maximum over the whole call; `group` splits the call into
    blocks of that many consecutive views (one dataset sample = T views) normalised separately."""
    if type != "heatmap":
        raise NotImplementedError("make_bbox_features: only the 'heatmap' representation is built (bbox_representation: heatmap)")
    if not bbox.is_cuda:
        raise _lib.BoxDreamerLibError("make_bbox_features: the input must be a CUDA tensor (no CPU fallback)")
    H, W = shape
    if H != W:
        raise ValueError("make_bbox_features: square crops only")
    L = bbox.shape[0]
    px = bbox.reshape(L, 8, 2).float().contiguous()
    out = torch.empty(L, 8, H, W, device=bbox.device, dtype=dtype)
    code = _lib.BD_BF16 if dtype == torch.bfloat16 else _lib.BD_F32
    if dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"make_bbox_features: unsupported dtype {dtype}")
    with torch.cuda.device(bbox.device):
        _lib.check(_lib.load().bd_make_bbox_features(_lib.ptr(px), _lib.ptr(out), code, L, H, group or L, _lib.stream_ptr()),
                   "bd_make_bbox_features")
    return out
