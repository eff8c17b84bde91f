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
