"""Generates the committed golden fixtures from the UNMODIFIED reference, run in the build
container (needs /root/reference; see oracle/ref_import.py for the stub recipe).

    python tests/golden/make_golden.py

Writes
  tests/golden/forward_b1t2.npz   BASELINE config 1 (B=1, T=2, 224 px, fp32 CPU)
  tests/golden/forward_b2t3.npz   a ragged-ish second case (B=2, T=3; query_idx = [2, 0])
  tests/golden/pnp_cv2.npz        cv2.solvePnP(SOLVEPNP_ITERATIVE) (box_utils.py:171-183) on
                                  synthetic box corners, sigma in {0, 2, 5} px, 64 cases each

Weights/inputs are the deterministic streams of boxdreamer_b200/synth.py (seed recorded in the
file), loaded into the reference modules with load_state_dict(strict=True).  Seams are captured
with forward hooks on the reference modules and stored sub-sampled
(strides recorded) together with float64 checksums of the full tensors.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from boxdreamer_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

WEIGHT_SEED = 0


def checksum(t: torch.Tensor):
    t = t.double()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def run_case(model, B, T, seed, query_idx=None):
    data = synth.synth_inputs(B, T, 224, seed=seed)
    if query_idx is not None:
        data["query_idx"] = torch.tensor(query_idx, dtype=torch.int64)
    seams = {}
    hooks = []

    def grab(name, take_input=False):
        def fn(mod, inp, out):
            seams[name] = (inp[0] if take_input else out).detach().clone()
        return fn

    dm = model.rgb_encoder.model
    dec = model.decoder
    for i in (0, 5, 11):
        hooks.append(dm.blocks[i].register_forward_hook(grab(f"dino_block{i}")))
        hooks.append(dec.attn[i].register_forward_hook(grab(f"dec_block{i}")))
    hooks.append(dec.attn[0].register_forward_hook(grab("fused", take_input=True)))
    hooks.append(dec.bbox_proj.register_forward_hook(grab("logits")))
    hooks.append(dm.norm.register_forward_hook(grab("dino_norm")))
    with torch.no_grad():
        ret = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()})
    for h in hooks:
        h.remove()
    mask = ret["camera_mask"]
    query_ret = ret["pred_bbox"][mask]  # [B,8,S,S]
    hm = ((query_ret.float() + 1) / 2).reshape(B, 8, -1)
    vals, idx = torch.topk(hm, k=21, dim=2)
    assert (vals[:, :, 19] > vals[:, :, 20]).all(), "tie at the top-20 boundary: pick another seed"
    kp_norm = ret["regression_boxes"][mask]
    out = {
        "weight_seed": np.array(WEIGHT_SEED), "input_seed": np.array(seed), "B": np.array(B), "T": np.array(T),
        "query_idx": data["query_idx"].numpy(),
        "stride_tok": np.array(16), "stride_ch": np.array(4),
        "dino_feats_sub": seams["dino_norm"][:, 5:][:, ::16, ::4].numpy(),
        "dino_feats_cs": checksum(seams["dino_norm"][:, 5:]),
        "fused_sub": seams["fused"][:, ::16, ::4].numpy(), "fused_cs": checksum(seams["fused"]),
        "logits_sub": seams["logits"][:, ::4, ::7].numpy(), "logits_cs": checksum(seams["logits"]),
        "logits_absmax": np.array(seams["logits"].abs().max().item()),
        "query_ret_sub": query_ret[:, :, ::4, ::4].numpy(), "query_ret_cs": checksum(query_ret),
        "topk_vals": vals.numpy(), "topk_idx": idx.numpy(),
        "keypoints_norm": kp_norm.numpy(),
        "regression_boxes": ret["regression_boxes"].numpy(),
        "pred_poses": ret["pred_poses"].numpy(),
        "camera_mask": mask.numpy(),
    }
    for i in (0, 5, 11):
        out[f"dino_block{i}_sub"] = seams[f"dino_block{i}"][:, ::16, ::4].numpy()
        out[f"dino_block{i}_cs"] = checksum(seams[f"dino_block{i}"])
        out[f"dec_block{i}_sub"] = seams[f"dec_block{i}"][:, ::16, ::4].numpy()
        out[f"dec_block{i}_cs"] = checksum(seams[f"dec_block{i}"])
    return out


def pnp_fixture():
    import cv2

    out = {"cv2_version": np.array(cv2.__version__)}
    for sigma in (0.0, 2.0, 5.0):
        c2, X3, Ks, gt = synth.synth_pnp_cases(64, sigma)
        Rs = np.zeros((64, 3, 3))
        ts = np.zeros((64, 3))
        for i in range(64):
            ok, rvec, tvec = cv2.solvePnP(X3[i], c2[i], Ks[i], None, flags=cv2.SOLVEPNP_ITERATIVE)
            assert ok
            Rs[i], _ = cv2.Rodrigues(rvec)
            ts[i] = tvec.ravel()
        tag = f"s{int(sigma)}"
        out[f"corners_{tag}"], out[f"bbox3d_{tag}"], out[f"K_{tag}"] = c2, X3, Ks
        out[f"R_{tag}"], out[f"t_{tag}"], out[f"gt_{tag}"] = Rs, ts, gt
    return out


def main():
    torch.manual_seed(0)
    model = ref_import.build_reference()
    model.load_state_dict(synth.synth_decoder_state_dict(WEIGHT_SEED), strict=True)
    model.rgb_encoder.model.load_state_dict(synth.synth_dino_state_dict(WEIGHT_SEED), strict=True)
    np.savez_compressed(os.path.join(HERE, "forward_b1t2.npz"), **run_case(model, 1, 2, 1235))
    np.savez_compressed(os.path.join(HERE, "forward_b2t3.npz"), **run_case(model, 2, 3, 1236, query_idx=[2, 0]))
    np.savez_compressed(os.path.join(HERE, "pnp_cv2.npz"), **pnp_fixture())
    for f in ("forward_b1t2.npz", "forward_b2t3.npz", "pnp_cv2.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
