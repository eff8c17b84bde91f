"""CPU: the C-ABI library loads and exports every symbol include/boxdreamer_b200.h declares (no compute calls
without a GPU), the product fails loudly without CUDA, and the host-side mirror keeps the reference's interface."""
import ctypes as C
import os
import re

import pytest
import torch

from boxdreamer_b200 import BoxDreamer, _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "boxdreamer_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bd_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/boxdreamer_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.bd_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    h = C.c_void_p()
    cfg = _lib.BdConfig(224, 14, 768, 12, 8, 12, 12, 4, 37, 0, 1, 1, 2)
    rc = lib.bd_create(C.byref(h), C.byref(cfg))
    assert rc == -2 and b"no CUDA device" in lib.bd_last_error()
    from boxdreamer_b200.config import make_config
    m = BoxDreamer(make_config())
    with pytest.raises(_lib.BoxDreamerLibError):
        m(synth.synth_inputs(1, 2))


def test_state_dict_layout_matches_reference_contract():
    from boxdreamer_b200.config import make_config
    m = BoxDreamer(make_config())
    sd = m.state_dict()
    exp = synth.decoder_param_shapes()
    assert list(sd.keys()) == list(exp.keys()) and len(sd) == 177
    assert all(tuple(sd[k].shape) == tuple(exp[k]) for k in exp)
    assert sum(v.numel() for v in sd.values()) == 88_649_504  # README.md:354-355 "88.6M"
    # Lightning checkpoints carry the "BoxDreamer." prefix (BoxDreamer_lightning_model.py:34); demo.py:564-573 strips it
    ck = {"BoxDreamer." + k: v for k, v in synth.synth_decoder_state_dict(3).items()}
    m.load_state_dict({k[len("BoxDreamer."):]: v for k, v in ck.items()}, strict=True)
    dsd = m.rgb_encoder.model.state_dict()
    assert {k: tuple(v.shape) for k, v in dsd.items()} == {k: tuple(v) for k, v in synth.dino_param_shapes().items()}
    assert hasattr(m.rgb_encoder, "get_device") and hasattr(m.rgb_encoder, "to_device") and hasattr(m.rgb_encoder, "predict")


def test_config_validation_mirrors_reference():
    from boxdreamer_b200.config import make_config
    cfg = make_config()
    cfg["modules"]["decoder"]["patch_size"] = 16
    with pytest.raises(AssertionError):
        BoxDreamer(cfg)
    cfg = make_config()
    cfg["modules"]["use_tracking"] = True
    with pytest.raises(NotImplementedError):
        BoxDreamer(cfg)


def test_synth_inputs_contract():
    d = synth.synth_inputs(2, 3, 224, seed=5)
    assert d["images"].shape == (2, 3, 3, 224, 224) and d["bbox_feat"].shape == (2, 3, 8, 224, 224)
    assert float(d["bbox_feat"].max()) == 1.0 and float(d["bbox_feat"].min()) >= -1.0
    assert d["query_idx"].tolist() == [2, 2]
    d2 = synth.synth_inputs(2, 3, 224, seed=5)
    assert all(torch.equal(d[k], d2[k]) for k in d if torch.is_tensor(d[k]))
    w1, w2 = synth.synth_decoder_state_dict(0), synth.synth_decoder_state_dict(0)
    assert all(torch.equal(w1[k], w2[k]) for k in w1)


def test_interpolated_pos_embed_matches_oracle():
    from oracle import boxdreamer_oracle as O
    from boxdreamer_b200.config import make_config
    m = BoxDreamer(make_config())
    dino = synth.synth_dino_state_dict(0)
    m.rgb_encoder.model.load_state_dict(dino)
    for S in (224, 336, 518):
        assert torch.equal(m.rgb_encoder.model.interpolated_pos_embed(S), O.dino_pos_embed(dino["pos_embed"], S))
