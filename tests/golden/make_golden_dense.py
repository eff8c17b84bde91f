"""Golden fixture for the dense multi-round path from the UNMODIFIED reference (build container only):
BoxDreamer.forward with dense_cfg {enable, filter dino top-5, multi_round, sub_batch_size 3, fine_level, fine_topk 2} on
B=1, T=8 (7 references), 224 px, fp32 CPU, synth weights (seed 0) / inputs (seed 4321).  Captured at the seams of
dense_processing.py / data_processing.py: the DINO pre-selection mask, the sub-batch heat maps, the pooled key points, the
cv2.solvePnPRansac pose (randomised, third-party: recorded so the GPU test can inject it), the fine-pass neighbours and the
final outputs.

    python tests/golden/make_golden_dense.py      # writes tests/golden/dense_b1t8.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from boxdreamer_b200 import synth  # noqa: E402
from boxdreamer_b200.config import make_config  # noqa: E402
from oracle import ref_import  # noqa: E402

ref_import.install()
import cv2  # noqa: E402

cv2.setRNGSeed(0)
from src.models.modules.encoder import dinov2 as ref_dino  # noqa: E402

orig = ref_dino.DinoV2Wrapper.load_model
ref_dino.DinoV2Wrapper.load_model = lambda self, device="cpu": orig(self, device="cpu")
from src.models.BoxDreamerModel import BoxDreamer  # noqa: E402
import src.models.utils.data_processing as dp  # noqa: E402
import src.models.utils.dense_processing as dn  # noqa: E402

cfg = make_config(224)
cfg["modules"]["dense_cfg"].update(dict(enable=True, filter_enable=True, filter="dino", filter_topk=5, multi_round=True,
                                        sub_batch_size=3, fine_level=True, fine_topk=2, dense_mem_friendly=False))
model = BoxDreamer(cfg).eval()
model.load_state_dict(synth.synth_decoder_state_dict(0), strict=True)
model.rgb_encoder.model.load_state_dict(synth.synth_dino_state_dict(0), strict=True)

B, T, SEED = 1, 8, 4321
data = synth.synth_inputs(B, T, 224, seed=SEED)
data["query_idx"] = torch.tensor([3], dtype=torch.int64)
rec = {}

_dm = dp.dino_matching
def dino_matching(*a, **k):
    m = _dm(*a, **k)
    rec["filter_mask"] = m.clone()
    return m
dp.dino_matching = dino_matching

_rp = dn.recover_pose_from_dense_bb8
def recover(bbox_feat, bbox_3d, K, rep):
    rec["coarse_heat"] = bbox_feat.clone()                 # [B, n_sub, 1, H, W, 8]
    poses, kp = _rp(bbox_feat, bbox_3d, K, rep)
    rec["coarse_pose"] = poses.clone()
    rec["pooled_kp_norm"] = kp.clone()
    return poses, kp
dn.recover_pose_from_dense_bb8 = recover

_fn = dn.fetch_neighbors_by_pose_similarity
def fetch(gt, pred, topk=5):
    idx = _fn(gt, pred, topk=topk)
    rec["fine_idx"] = idx.clone()
    return idx
dn.fetch_neighbors_by_pose_similarity = fetch

with torch.no_grad():
    out = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()})

heat = rec["coarse_heat"][:, :, 0].permute(0, 1, 4, 2, 3).contiguous()      # [B, n_sub, 8, H, W]
mask = out["camera_mask"]
final = out["pred_bbox"][mask]
cs = lambda t: np.array([t.double().sum().item(), t.double().abs().sum().item()])
np.savez_compressed(
    os.path.join(HERE, "dense_b1t8.npz"),
    input_seed=np.array(SEED), query_idx=data["query_idx"].numpy(),
    filter_mask=rec["filter_mask"].numpy(), coarse_heat_sub=heat[:, :, :, ::4, ::4].numpy(), coarse_heat_cs=cs(heat),
    pooled_kp_norm=rec["pooled_kp_norm"].numpy(), coarse_pose=rec["coarse_pose"].numpy(), fine_idx=rec["fine_idx"].numpy(),
    final_heat_sub=final[:, :, ::4, ::4].numpy(), final_heat_cs=cs(final), final_images_cs=cs(out["images"]),
    regression_boxes=out["regression_boxes"].numpy(), out_query_idx=out["query_idx"].numpy())
print("filter mask", rec["filter_mask"].int().tolist(), "fine idx", rec["fine_idx"].tolist(), "views out", out["images"].shape[1])
print("coarse pose ok", float(rec["coarse_pose"][0, 0, 3, 3]), "heat", tuple(heat.shape))
