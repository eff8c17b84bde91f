"""CPU: the oracle (oracle/boxdreamer_oracle.py) against the committed golden fixtures that were generated from the
unmodified reference (tests/golden/make_golden.py), and the PnP restatement against cv2 4.13 outputs."""
import os

import numpy as np
import pytest
import torch

from boxdreamer_b200 import synth
from oracle import boxdreamer_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def weights():
    return synth.synth_decoder_state_dict(0), synth.synth_dino_state_dict(0)


def _scaled(got, ref):
    got = torch.as_tensor(got).float()
    ref = torch.as_tensor(ref).float()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-30))


def test_oracle_forward_matches_reference_golden_b1t2(weights):
    gold = np.load(os.path.join(GOLD, "forward_b1t2.npz"))
    dec, dino = weights
    data = synth.synth_inputs(1, 2, 224, seed=int(gold["input_seed"]))
    seams = {}
    with torch.no_grad():
        out = O.forward(data, dec, dino, seams=seams)
    st, sc = int(gold["stride_tok"]), int(gold["stride_ch"])
    assert _scaled(seams["dino_feats"].reshape(2, 256, 768)[:, ::st, ::sc], gold["dino_feats_sub"]) <= 1e-5
    assert _scaled(seams["fused"][:, ::st, ::sc], gold["fused_sub"]) <= 1e-5
    for i in (0, 5, 11):
        assert _scaled(seams[f"dino_block{i}"][:, ::st, ::sc], gold[f"dino_block{i}_sub"]) <= 1e-5
        assert _scaled(seams[f"dec_block{i}"][:, ::st, ::sc], gold[f"dec_block{i}_sub"]) <= 1e-5
    assert _scaled(out["logits"][:, ::4, ::7], gold["logits_sub"]) <= 1e-5
    assert _scaled(out["query_ret"][:, :, ::4, ::4], gold["query_ret_sub"]) <= 1e-5
    cs = gold["logits_cs"]
    assert abs(out["logits"].double().sum().item() - cs[0]) <= 1e-5 * cs[1]
    # integer part: bit-exact
    ref_idx = torch.from_numpy(gold["topk_idx"][:, :, :20]).long()
    assert torch.equal(torch.sort(out["topk_idx"], dim=2).values, torch.sort(ref_idx, dim=2).values)
    assert (gold["topk_vals"][:, :, 19] > gold["topk_vals"][:, :, 20]).all(), "fixture must have no tie at the top-20 boundary"
    assert torch.equal(out["keypoints_norm"], torch.from_numpy(gold["keypoints_norm"]))
    assert torch.equal(out["regression_boxes"], torch.from_numpy(gold["regression_boxes"]))
    assert torch.equal(out["camera_mask"], torch.from_numpy(gold["camera_mask"]))


def _rot_err_deg(Ra, Rb):
    """Geodesic angle between two rotations, computed from the chordal distance so that it stays accurate for the
    float32-rounded matrices the C ABI returns (acos((tr-1)/2) loses half the digits near 0)."""
    s = min(np.linalg.norm(np.asarray(Ra, dtype=np.float64) - np.asarray(Rb, dtype=np.float64)) / (2.0 * np.sqrt(2.0)), 1.0)
    return float(np.degrees(2.0 * np.arcsin(s)))


@pytest.mark.parametrize("tag,min_rate", [("s0", 1.0), ("s2", 1.0), ("s5", 0.95)])
def test_pnp_oracle_matches_cv2_fixture(tag, min_rate):
    fx = np.load(os.path.join(GOLD, "pnp_cv2.npz"))
    n = 32
    good = 0
    for i in range(n):
        R, t = O.solve_pnp_iterative(fx[f"bbox3d_{tag}"][i], fx[f"corners_{tag}"][i], fx[f"K_{tag}"][i])
        re = _rot_err_deg(R, fx[f"R_{tag}"][i])
        te = np.linalg.norm(t - fx[f"t_{tag}"][i]) / np.linalg.norm(fx[f"t_{tag}"][i])
        good += int(re <= 1e-3 and te <= 1e-4)
    assert good / n >= min_rate


def test_projection_known_answer():
    """The one known-answer vector the reference's tests hold for this geometry (tests/dataset/test_base.py:143-151)."""
    K = np.array([[1000.0, 0, 320], [0, 1000, 240], [0, 0, 1]])
    pts = np.array([[0.0, 0, 5], [1, 1, 5], [-1, -1, 5]])
    uv = synth.project(K, np.eye(3), np.zeros(3), pts)
    assert np.allclose(uv, [[320, 240], [520, 440], [120, 40]], atol=1e-5, rtol=1e-5)


def test_topk_tie_rule_and_mean():
    heat = torch.full((1, 8, 16, 16), -1.0)
    heat[0, :, 3, 5] = 0.5
    heat[0, :, 3, 6] = 0.5
    idx, kp, nm = O.corners_topk(heat)
    assert idx[0, 0, 0].item() == 3 * 16 + 5 and idx[0, 0, 1].item() == 3 * 16 + 6  # tie -> lower index first
    assert idx[0, 0, 2].item() == 0  # then the lowest indices of the flat background
