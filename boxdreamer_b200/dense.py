"""Dense-reference processing around the hot path: reference selection by DINO
similarity, multi-round decoding over sub-batches of the reference set, pooled robust PnP and the fine pass on the
pose-nearest references.

Mirrors (reference file:line)
  sub_batchify                          src/models/utils/data_utils.py:5-95
  fetch_neighbors_by_pose_similarity    src/models/utils/data_utils.py:98-135
  filter_by_neighbor_mask               src/models/utils/data_processing.py:9-99   (+ update_filtered_data :102-171)
  dino_matching                         src/models/utils/matching.py:64-174        (similarity_type="dot_product")
  process_dense_input / normalize       src/models/utils/data_processing.py:174-231
  process_multi_round                   src/models/utils/dense_processing.py:8-158
  recover_pose_from_dense_bb8           src/models/utils/box_utils.py:202-304

Everything here is index bookkeeping on device tensors (plain torch, device-agnostic so the CPU tests can compare it
with the reference functions); the arithmetic -- encoder, decoder, top-20 corners, pooled PnP -- runs in the CUDA engine
behind the C ABI.  Reference behaviour that is kept on purpose:
  * sub-batches are zero-padded when (T-1) is not a multiple of sub_batch_size (data_utils.py:38-61);
  * dino_matching fills invalid pairs with -1e4 but averages over everything that is not -1e9, i.e. over ALL pairs
    (matching.py:131 vs :160-162);
  * with multi_round=True only fine_level=True yields a tensor for BoxDreamer.forward; the coarse-only branch returns the
    data dict, which the caller then scatters as if it were a heat map (dense_processing.py:145-158,
    BoxDreamerModel.py:343-344) -- that raises in the reference and raises here.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["sub_batchify", "fetch_neighbors_by_pose_similarity", "filter_by_neighbor_mask", "dino_matching",
           "process_dense_input", "process_multi_round", "normalize"]

# tensors of the data dict that carry one entry per view and have to follow a reference selection
_PER_VIEW_KEYS = (("original_poses", (4, 4)), ("intrinsics", (3, 3)), ("non_ndc_intrinsics", (3, 3)),
                  ("original_intrinsics", (3, 3)), ("scale", (3,)), ("bbox_3d", (8, 3)), ("bbox_proj_crop", (8, 2)))


def normalize(x: torch.Tensor) -> torch.Tensor:
    """data_processing.py:229-231 (the decoder ignores image_masks; kept for the call signature)."""
    return x / (x.sum() + 1e-6)


def _split_query(t: torch.Tensor, camera_mask: torch.Tensor):
    """[B,T,...] -> (references [B,T-1,...] in view order, query [B,...])."""
    B, T = camera_mask.shape
    return t[~camera_mask].reshape(B, T - 1, *t.shape[2:]), t[camera_mask]


def sub_batchify(pose_feat, frames, camera_mask, rgb_feature, image_masks, sub_batch_size):
    """Groups of `sub_batch_size` references, each followed by the query view: [B,T,...] -> [B,n_sub,sub+1,...].
    The last group is zero-padded; the query sits at index `sub_batch_size` of every group."""
    B, T = camera_mask.shape
    n_ref = T - 1
    n_sub = (n_ref + sub_batch_size - 1) // sub_batch_size
    outs = []
    for t in (pose_feat, frames, rgb_feature, image_masks):
        refs, query = _split_query(t, camera_mask)
        padded = refs.new_zeros(B, n_sub * sub_batch_size, *t.shape[2:])
        padded[:, :n_ref] = refs
        grouped = padded.view(B, n_sub, sub_batch_size, *t.shape[2:])
        q = query.unsqueeze(1).unsqueeze(2).expand(B, n_sub, 1, *t.shape[2:])
        outs.append(torch.cat([grouped, q], dim=2))
    new_mask = torch.zeros(B, n_sub, sub_batch_size + 1, dtype=torch.bool, device=camera_mask.device)
    new_mask[:, :, sub_batch_size] = True
    new_pose_feat, new_frames, new_rgb, new_image_masks = outs
    return new_pose_feat, new_frames, new_mask, new_rgb, new_image_masks


def fetch_neighbors_by_pose_similarity(gt_poses, pred_pose, topk=5):
    """Indices [B,topk] of the references whose pose is nearest to the predicted one:
    geodesic rotation angle + translation distance, smallest first."""
    B, N = gt_poses.shape[:2]
    pred = pred_pose.reshape(B, 1, 4, 4)
    rel = torch.matmul(pred[..., :3, :3], gt_poses[..., :3, :3].transpose(-1, -2))      # R_pred R_gt^T  [B,N,3,3]
    cos = (torch.diagonal(rel, dim1=-2, dim2=-1).sum(-1) - 1) / 2
    rot = torch.acos(torch.clamp(cos, -1, 1))
    trans = torch.norm(pred[..., :3, 3] - gt_poses[..., :3, 3], dim=-1)
    return torch.topk(rot + trans, k=topk, dim=1, largest=False).indices


def _mask_from_indices(indices: torch.Tensor, n: int) -> torch.Tensor:
    mask = torch.zeros(indices.shape[0], n, dtype=torch.bool, device=indices.device)
    mask.scatter_(1, indices, True)
    return mask


def _keep_refs(t: torch.Tensor, camera_mask: torch.Tensor, neighbor_mask: torch.Tensor) -> torch.Tensor:
    """[B,T,...] -> [B,k+1,...]: the selected references in view order, then the query."""
    B = camera_mask.shape[0]
    refs, query = _split_query(t, camera_mask)
    kept = refs[neighbor_mask].reshape(B, -1, *t.shape[2:])
    return torch.cat([kept, query.unsqueeze(1)], dim=1)


def filter_by_neighbor_mask(data, neighbor_mask, pose_feat, frames, camera_mask, rgb_feature, image_masks):
    """Keeps the references selected by neighbor_mask [B,T-1] (the same count in every row), moves the query to the
    last position and rewrites the per-view entries of `data` accordingly (update_filtered_data)."""
    B = frames.shape[0]
    new_pose_feat = _keep_refs(pose_feat, camera_mask, neighbor_mask)
    new_frames = _keep_refs(frames, camera_mask, neighbor_mask)
    new_rgb = _keep_refs(rgb_feature, camera_mask, neighbor_mask) if rgb_feature is not None else None
    new_image_masks = _keep_refs(image_masks, camera_mask, neighbor_mask)
    T = new_frames.shape[1]
    new_mask = torch.zeros(B, T, dtype=torch.bool, device=camera_mask.device)
    new_mask[:, -1] = True

    poses = data["poses"]
    data["bbox_feat"] = new_pose_feat.clone()
    data["images"] = new_frames.clone()
    data["query_idx"] = torch.full((B,), T - 1, dtype=torch.int64, device=poses.device)
    data["camera_mask"] = new_mask.clone()
    data["poses"] = _keep_refs(poses, camera_mask, neighbor_mask)
    for key, dims in _PER_VIEW_KEYS:
        if key in data:
            data[key] = _keep_refs(data[key].reshape(B, camera_mask.shape[1], *dims), camera_mask, neighbor_mask)
    if "original_images" in data:   # list over views of lists over samples (data_processing.py:139-171)
        org = data["original_images"]
        keep = neighbor_mask.cpu().numpy()
        n_keep = int(keep[0].sum())
        new_org = [[] for _ in range(n_keep + 1)]
        for b in range(len(org[0])):
            slot = 0
            for t in range(len(org) - 1):
                if keep[b, t]:
                    new_org[slot].append(org[t][b])
                    slot += 1
            new_org[-1].append(org[-1][b])
        data["original_images"] = new_org
    return data, new_pose_feat, new_frames, new_mask, new_rgb, new_image_masks


def _foreground_tokens(images: torch.Tensor, grid: int, threshold: float = 0.05) -> torch.Tensor:
    """[L,3,H,W] -> [L,grid*grid] {0,1}: luminance above the threshold, nearest-resized to the token grid."""
    lum = 0.299 * images[:, 0] + 0.587 * images[:, 1] + 0.114 * images[:, 2]
    fg = (lum > threshold).float()
    return F.interpolate(fg.unsqueeze(1), size=(grid, grid), mode="nearest").reshape(images.shape[0], -1)


def dino_matching(ref_features, query_features, ref_images, query_images, similarity_type="dot_product", topk=10,
                  similarity_params=None):
    """Boolean mask [B,N] of the `topk` references most similar to the query: mean over all token pairs of the cosine
    similarity between foreground-masked DINO tokens (pairs with a background token count as -1e4, see module docstring)."""
    if similarity_type != "dot_product":
        raise NotImplementedError("dino_matching: only the reference's default 'dot_product' similarity is built")
    B, N, L, D = ref_features.shape
    grid = int(round(L ** 0.5))
    q_mask = _foreground_tokens(query_images, grid)                                        # [B,L]
    r_mask = _foreground_tokens(ref_images.reshape(B * N, *ref_images.shape[2:]), grid)    # [B*N,L]
    q_mask = q_mask.unsqueeze(1).expand(B, N, L).reshape(B * N, L, 1)
    r_mask = r_mask.unsqueeze(-1)
    q_feat = query_features.unsqueeze(1).expand(B, N, L, D).reshape(B * N, L, D)
    r_feat = ref_features.reshape(B * N, L, D)
    q_n = F.normalize(q_feat * q_mask, dim=-1)
    r_n = F.normalize(r_feat * r_mask, dim=-1)
    sim = torch.bmm(q_n, r_n.transpose(-2, -1))
    valid = torch.bmm(q_mask, r_mask.transpose(-2, -1))
    sim = sim.masked_fill(valid == 0, -1e4)
    kept = sim.masked_fill(sim == -1e9, 0)                       # (no-op: the fill value above is -1e4)
    count = (sim != -1e9).float().sum(dim=[1, 2])
    mean_sim = (kept.sum(dim=[1, 2]) / count).reshape(B, N)
    mean_sim = torch.nan_to_num(mean_sim, 0.0, 0.0, 0.0)
    return _mask_from_indices(torch.topk(mean_sim, k=topk, dim=-1).indices, N)


def _cfg(dense_cfg, key):
    return dense_cfg[key] if isinstance(dense_cfg, dict) else getattr(dense_cfg, key)


def process_dense_input(data, pose_feat, frames, camera_mask, rgb_feature, image_masks, dense_cfg):
    """Optional DINO pre-selection of `filter_topk` references."""
    if _cfg(dense_cfg, "filter") == "dino" and _cfg(dense_cfg, "filter_enable"):
        ref_feat, q_feat = _split_query(rgb_feature, camera_mask)
        ref_img, q_img = _split_query(frames, camera_mask)
        neighbor_mask = dino_matching(ref_feat, q_feat, ref_img, q_img, topk=_cfg(dense_cfg, "filter_topk"))
        return filter_by_neighbor_mask(data, neighbor_mask, pose_feat, frames, camera_mask, rgb_feature, image_masks)
    return data, pose_feat, frames, camera_mask, rgb_feature, image_masks


MAX_POOLED_POINTS = 256   # PNP_MAXPTS_POOLED of csrc/post.cu (pooled robust PnP: n_sub sub-batches x 8 corners)


def process_multi_round(data, pose_feat, frames, camera_mask, rgb_feature, image_masks, decoder, dense_cfg, bbox_representation,
                        pooled_pose_fn):
    """Coarse round over sub-batches of the references -> pooled robust PnP -> fine round on the `fine_topk` references
    nearest to the coarse pose.  `decoder(pose_feat, frames, masks, rgb_feature, image_masks) -> [B',8,S,S]`;
    `pooled_pose_fn(heats [B,n_sub,8,S,S], bbox3d_q [B,8,3], K_q [B,3,3]) -> poses [B,4,4]` (recover_pose_from_dense_bb8).
    Also stores the coarse result in data["dense_coarse_poses"] (not a reference key)."""
    if bbox_representation != "heatmap":
        raise NotImplementedError("multi-round: heatmap representation only")
    if not _cfg(dense_cfg, "fine_level"):
        raise NotImplementedError(
            "dense_cfg.multi_round without fine_level returns the data dict where BoxDreamer.forward expects a heat map "
            "(dense_processing.py:145-158 vs BoxDreamerModel.py:343-344): that configuration fails in the reference too")
    B = frames.shape[0]
    poses = data["poses"]
    K_q = data["non_ndc_intrinsics"][camera_mask]
    bbox3d_q = data["bbox_3d"][camera_mask]
    sub = _cfg(dense_cfg, "sub_batch_size")
    n_ref = frames.shape[1] - 1
    n_sub_planned = (n_ref + sub - 1) // sub
    if n_sub_planned * 8 > MAX_POOLED_POINTS:   # checked before any GPU work: the coarse decoder round is the expensive part
        raise ValueError(
            f"dense multi-round: {n_ref} references in sub-batches of {sub} pool {n_sub_planned * 8} 2D-3D pairs, but bd_pnp takes at most "
            f"{MAX_POOLED_POINTS}; enable dense_cfg.filter_enable (filter_topk <= {MAX_POOLED_POINTS // 8 * sub}) or raise sub_batch_size")
    g_pose, g_frames, g_mask, g_rgb, g_img_masks = sub_batchify(pose_feat, frames, camera_mask, rgb_feature, image_masks, sub)
    n_sub = g_pose.shape[1]
    if _cfg(dense_cfg, "dense_mem_friendly"):
        rounds = [decoder(g_pose[:, i], g_frames[:, i], g_mask[:, i], g_rgb[:, i], normalize(g_img_masks[:, i])) for i in range(n_sub)]
        heats = torch.stack(rounds, dim=1)
    else:
        flat = lambda t: t.reshape(B * n_sub, *t.shape[2:])
        heats = decoder(flat(g_pose), flat(g_frames), flat(g_mask), flat(g_rgb), normalize(flat(g_img_masks)))
        heats = heats.reshape(B, n_sub, *heats.shape[1:])
    query_poses = pooled_pose_fn(heats, bbox3d_q, K_q)
    data["dense_coarse_poses"] = query_poses
    ref_poses, _ = _split_query(poses, camera_mask)
    idx = fetch_neighbors_by_pose_similarity(ref_poses.float(), query_poses.float(), topk=_cfg(dense_cfg, "fine_topk"))
    neighbor_mask = _mask_from_indices(idx, ref_poses.shape[1])
    data, pose_feat, frames, camera_mask, rgb_feature, image_masks = filter_by_neighbor_mask(
        data, neighbor_mask, pose_feat, frames, camera_mask, rgb_feature, image_masks)
    return decoder(pose_feat, frames, camera_mask, rgb_feature, normalize(image_masks))
