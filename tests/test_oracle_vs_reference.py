"""CPU, build container only: pins the oracle against the UNMODIFIED reference's own BoxDreamer.forward
(imported from /root/reference with the stub recipe of oracle/ref_import.py).  Skipped where the tree is absent."""
import pytest
import torch

from boxdreamer_b200 import synth
from oracle import boxdreamer_oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present")


def test_oracle_equals_reference_forward():
    model = ref_import.build_reference()
    dec, dino = synth.synth_decoder_state_dict(0), synth.synth_dino_state_dict(0)
    model.load_state_dict(dec, strict=True)
    model.rgb_encoder.model.load_state_dict(dino, strict=True)
    data = synth.synth_inputs(2, 2, 224, seed=4242)
    data["query_idx"] = torch.tensor([0, 1], dtype=torch.int64)
    with torch.no_grad():
        ref = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()})
        out = O.forward(data, dec, dino, with_pnp=False)
    assert torch.equal(ref["camera_mask"], out["camera_mask"])
    scale = ref["pred_bbox"].abs().max()
    assert (ref["pred_bbox"] - out["pred_bbox"]).abs().max() <= 1e-6 * scale
    assert torch.equal(ref["regression_boxes"], out["regression_boxes"])


def test_reference_state_dict_layout():
    model = ref_import.build_reference()
    assert list(model.state_dict().keys()) == list(synth.decoder_param_shapes().keys())
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == {k: tuple(v) for k, v in synth.decoder_param_shapes().items()}
    dsd = model.rgb_encoder.model.state_dict()
    assert {k: tuple(v.shape) for k, v in dsd.items()} == {k: tuple(v) for k, v in synth.dino_param_shapes().items()}


def test_input_synthesis_restatement_equals_reference():
    """synth.make_heatmaps / inputs.make_proj_bbox (the CPU side of the device rasteriser's parity test) against the
    dataset's own make_bbox_features(type='heatmap') and make_proj_bbox (bbox_utils.py:263-303, camera_utils.py:62-84)."""
    ref_import.install()
    from src.datasets.utils.base.bbox_utils import make_bbox_features as ref_feat
    from src.datasets.utils.base.camera_utils import make_proj_bbox as ref_proj
    from boxdreamer_b200.inputs import make_proj_bbox
    data = synth.synth_inputs(2, 3, 224, seed=99)
    poses = data["poses"].view(6, 4, 4)
    K = data["non_ndc_intrinsics"].view(6, 3, 3)
    X = data["bbox_3d"].view(6, 8, 3)
    proj_ref = ref_proj(poses, K, X)
    proj = make_proj_bbox(poses, K, X)
    assert torch.allclose(proj, proj_ref, atol=1e-3, rtol=1e-6)      # pixels; fp32 matmul association differs
    heat_ref = ref_feat(proj_ref, "heatmap", (224, 224))
    heat = synth.make_heatmaps(proj_ref, 224, group=6)      # one reference call = one normalisation group
    assert heat.shape == heat_ref.shape == (6, 8, 224, 224)
    assert torch.equal(heat, heat_ref)
