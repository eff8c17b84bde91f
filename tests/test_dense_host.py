"""Host-side bookkeeping of the dense-reference path (boxdreamer_b200/dense.py).

CPU, two layers:
  * properties that hold by construction (shapes, padding, query position, selection order) -- run everywhere;
  * equality with the UNMODIFIED reference functions (src/models/utils/{data_utils,data_processing,matching}.py) on
    seeded random tensors -- build container only (skipped where /root/reference is absent)."""
import copy

import pytest
import torch

from boxdreamer_b200 import dense
from oracle import ref_import


def _case(B=2, T=8, S=28, L=16, D=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.tensor([3, 0][:B] if B <= 2 else list(range(B)), dtype=torch.int64) % T
    mask = torch.zeros(B, T, dtype=torch.bool)
    mask[torch.arange(B), q] = True
    pose_feat = torch.randn(B, T, 8, S, S, generator=g)
    frames = torch.rand(B, T, 3, S, S, generator=g)
    frames[:, :, :, : S // 3] *= 0.02          # a dark band: exercises the foreground mask of dino_matching
    rgb = torch.randn(B, T, L, D, generator=g)
    img_masks = torch.ones(B, T, 1, S, S)
    poses = torch.eye(4).repeat(B, T, 1, 1)
    for b in range(B):
        for t in range(T):
            Qm, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
            if torch.det(Qm) < 0:
                Qm[:, 0] = -Qm[:, 0]
            poses[b, t, :3, :3] = Qm
            poses[b, t, :3, 3] = torch.randn(3, generator=g) * 0.3
    data = {
        "poses": poses, "images": frames.clone(), "bbox_feat": pose_feat.clone(), "query_idx": q.clone(),
        "intrinsics": torch.randn(B, T, 3, 3, generator=g), "non_ndc_intrinsics": torch.randn(B, T, 3, 3, generator=g),
        "bbox_3d": torch.randn(B, T, 8, 3, generator=g), "bbox_proj_crop": torch.randn(B, T, 8, 2, generator=g),
        "image_masks": img_masks.clone(), "camera_mask": mask.clone(),
    }
    return data, pose_feat, frames, mask, rgb, img_masks


def test_sub_batchify_layout_and_padding():
    data, pose_feat, frames, mask, rgb, img_masks = _case(T=8)          # 7 references, groups of 3 -> 3 groups, 2 padded slots
    gp, gf, gm, gr, gi = dense.sub_batchify(pose_feat, frames, mask, rgb, img_masks, 3)
    assert gp.shape == (2, 3, 4, 8, 28, 28) and gf.shape == (2, 3, 4, 3, 28, 28) and gr.shape == (2, 3, 4, 16, 12)
    assert gm.shape == (2, 3, 4) and bool(gm[:, :, 3].all()) and not bool(gm[:, :, :3].any())
    refs = frames[~mask].reshape(2, 7, 3, 28, 28)
    assert torch.equal(gf[:, 0, :3], refs[:, 0:3]) and torch.equal(gf[:, 1, :3], refs[:, 3:6]) and torch.equal(gf[:, 2, 0], refs[:, 6])
    assert float(gf[:, 2, 1:3].abs().max()) == 0.0 and float(gp[:, 2, 1:3].abs().max()) == 0.0   # zero padding
    for i in range(3):
        assert torch.equal(gf[:, i, 3], frames[mask]) and torch.equal(gr[:, i, 3], rgb[mask])     # the query closes every group


def test_filter_by_neighbor_mask_rewrites_data():
    data, pose_feat, frames, mask, rgb, img_masks = _case()
    keep = torch.zeros(2, 7, dtype=torch.bool)
    keep[0, [1, 4, 6]] = True
    keep[1, [0, 2, 3]] = True
    before = copy.deepcopy(data)
    data, pf, fr, cm, rf, im = dense.filter_by_neighbor_mask(data, keep, pose_feat, frames, mask, rgb, img_masks)
    assert fr.shape[1] == 4 and bool(cm[:, -1].all()) and int(cm.sum()) == 2
    assert torch.equal(data["query_idx"], torch.tensor([3, 3]))
    refs0 = before["images"][0][~mask[0]]
    assert torch.equal(fr[0, :3], refs0[[1, 4, 6]]) and torch.equal(fr[0, 3], before["images"][0, 3])
    for key in ("poses", "intrinsics", "non_ndc_intrinsics", "bbox_3d", "bbox_proj_crop"):
        assert data[key].shape[1] == 4
        assert torch.equal(data[key][1, -1], before[key][1, 0])                                  # sample 1: query was view 0
        assert torch.equal(data[key][1, :3], before[key][1][~mask[1]][[0, 2, 3]])


def test_fetch_neighbors_prefers_nearby_poses():
    data, *_ = _case()
    ref = data["poses"][:, 1:]
    pred = ref[:, 2].clone()                      # identical to reference 2 -> distance 0 -> must come first
    idx = dense.fetch_neighbors_by_pose_similarity(ref, pred, topk=3)
    assert idx.shape == (2, 3) and bool((idx[:, 0] == 2).all())


needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present")


def _ref_modules():
    ref_import.install()
    import importlib
    du = importlib.import_module("src.models.utils.data_utils")
    dp = importlib.import_module("src.models.utils.data_processing")
    mt = importlib.import_module("src.models.utils.matching")
    return du, dp, mt


@needs_ref
def test_matches_reference_functions():
    du, dp, mt = _ref_modules()
    for seed, T, sub in ((1, 8, 3), (2, 6, 5), (3, 12, 5)):
        data, pose_feat, frames, mask, rgb, img_masks = _case(T=T, seed=seed)
        ours = dense.sub_batchify(pose_feat, frames, mask, rgb, img_masks, sub)
        theirs = du.sub_batchify(pose_feat.clone(), frames.clone(), mask.clone(), rgb.clone(), img_masks.clone(), sub)
        for a, b in zip(ours, theirs):
            assert a.shape == b.shape and torch.equal(a.to(b.dtype), b)
        ref_p = data["poses"][~mask].reshape(2, T - 1, 4, 4)
        pred = data["poses"][mask] + 0.01
        assert torch.equal(dense.fetch_neighbors_by_pose_similarity(ref_p, pred, topk=3),
                           du.fetch_neighbors_by_pose_similarity(ref_p, pred, topk=3))
        ref_f = rgb[~mask].reshape(2, T - 1, *rgb.shape[2:])
        ref_i = frames[~mask].reshape(2, T - 1, *frames.shape[2:])
        m_ours = dense.dino_matching(ref_f, rgb[mask], ref_i, frames[mask], topk=3)
        m_ref = mt.dino_matching(ref_f, rgb[mask], ref_i, frames[mask], topk=3)
        assert torch.equal(m_ours, m_ref)
        d1, d2 = copy.deepcopy(data), copy.deepcopy(data)
        o = dense.filter_by_neighbor_mask(d1, m_ours, pose_feat, frames, mask, rgb, img_masks)
        r = dp.filter_by_neighbor_mask(d2, m_ref, pose_feat.clone(), frames.clone(), mask.clone(), rgb.clone(), img_masks.clone())
        for a, b in zip(o[1:], r[1:]):
            assert torch.equal(a, b)
        for key in d2:
            if torch.is_tensor(d2[key]):
                assert torch.equal(d1[key], d2[key]), key


def test_multi_round_rejects_more_pooled_points_than_bd_pnp_takes_before_any_work():
    """the pooled robust PnP takes n_sub * 8 <= 256 2D-3D pairs (32 sub-batches); a larger reference set must fail with a
    clear message BEFORE the coarse decoder round runs, not with BD_ERR_INVALID after it.  Sub-batches up to the limit pass the
    check (here the decoder stub then stops the run)."""
    class Ran(Exception):
        pass

    def decoder(*a, **k):
        raise Ran()

    cfg = {"sub_batch_size": 1, "fine_level": True, "fine_topk": 2, "dense_mem_friendly": False}
    data, pose_feat, frames, mask, rgb, img_masks = _case(B=1, T=34, S=14, L=4, D=4)        # 33 references -> 33 sub-batches -> 264 pairs
    with pytest.raises(ValueError, match="264 2D-3D pairs"):
        dense.process_multi_round(data, pose_feat, frames, mask, rgb, img_masks, decoder, cfg, "heatmap", lambda *a: None)
    data, pose_feat, frames, mask, rgb, img_masks = _case(B=1, T=33, S=14, L=4, D=4)        # 32 sub-batches -> 256 pairs: accepted
    with pytest.raises(Ran):
        dense.process_multi_round(data, pose_feat, frames, mask, rgb, img_masks, decoder, cfg, "heatmap", lambda *a: None)
