"""Checkpoint I/O + predict step (boxdreamer_b200/checkpoint.py).  CPU: file formats, Lightning
prefix stripping, rank-0-only loading with a world_size-2 gloo broadcast.  GPU: predict_step equals the module call."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from boxdreamer_b200 import BoxDreamer, checkpoint as ck, synth
from boxdreamer_b200.config import make_config


def test_load_checkpoint_formats(tmp_path):
    sd = synth.synth_decoder_state_dict(0, num_layers=1)
    from safetensors.torch import save_file
    p1 = str(tmp_path / "w.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, p1)
    p2 = str(tmp_path / "lightning.ckpt")    # Lightning: {"state_dict": {"BoxDreamer.<key>": ...}, plus things that are not weights}
    torch.save({"state_dict": {**{"BoxDreamer." + k: v for k, v in sd.items()}, "loss.weight": torch.ones(1)}, "epoch": 3}, p2)
    p3 = str(tmp_path / "plain.pth")
    torch.save(sd, p3)
    p4 = str(tmp_path / "BoxDreamer-vitb.safetensor")      # the upstream file name (run.py:172-183: singular suffix)
    save_file({k: v.contiguous() for k, v in sd.items()}, p4)
    p5 = str(tmp_path / "weights.bin")                    # safetensors content under a foreign name: sniffed from the header
    save_file({k: v.contiguous() for k, v in sd.items()}, p5)
    for p in (p1, p2, p3, p4, p5):
        got = ck.load_checkpoint(p)
        assert set(got.keys()) == set(sd.keys())          # safetensors returns its keys sorted; load_model orders by the model
        assert all(torch.equal(got[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        ck.load_checkpoint(str(tmp_path / "missing.ckpt"))


def test_untrusted_pickle_is_refused(tmp_path):
    """torch checkpoints are read with weights_only=True: a pickle that needs arbitrary code is rejected, not executed."""
    class Boom:
        def __reduce__(self):
            return (os.system, ("true",))
    p = str(tmp_path / "evil.ckpt")
    torch.save({"state_dict": {"decoder.x": torch.ones(1)}, "hook": Boom()}, p)
    with pytest.raises(Exception):
        ck.load_checkpoint(p)


def _worker_bad(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = BoxDreamer(make_config(224, num_layers=1))
        try:
            ck.load_model(model, "/nonexistent/ckpt.safetensor" if rank == 0 else None)
            ret[rank] = "no error"
        except FileNotFoundError:
            ret[rank] = "FileNotFoundError"
        except RuntimeError:
            ret[rank] = "RuntimeError"
    finally:
        dist.destroy_process_group()


def test_bad_path_fails_on_every_rank_instead_of_hanging():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_bad, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret) == {0: "FileNotFoundError", 1: "RuntimeError"}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = BoxDreamer(make_config(224, num_layers=1))
        ck.load_model(model, path if rank == 0 else None)          # only rank 0 may touch the file
        ref = synth.synth_decoder_state_dict(0, num_layers=1)
        ret[rank] = all(torch.equal(v, ref[k]) for k, v in model.state_dict().items())
    finally:
        dist.destroy_process_group()


def test_rank0_load_and_broadcast(tmp_path):
    from safetensors.torch import save_file
    path = str(tmp_path / "w.safetensors")
    save_file({k: v.contiguous() for k, v in synth.synth_decoder_state_dict(0, num_layers=1).items()}, path)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), path, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


@pytest.mark.gpu
def test_predict_step_equals_module_call():
    cfg = make_config(224)
    m = BoxDreamer(cfg, precision="exact")
    m.load_state_dict(synth.synth_decoder_state_dict(0), strict=True)
    m.rgb_encoder.model.load_state_dict(synth.synth_dino_state_dict(0), strict=True)
    m = m.cuda().eval()
    data = synth.synth_inputs(2, 3, 224, seed=90)
    res = ck.predict_step(m, data, batch_idx=7)                      # host batch in, as a dataloader would hand it over
    full = m({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()})
    mask = full["camera_mask"]
    assert torch.equal(res["pred_poses"], full["pred_poses"][mask]) and torch.equal(res["regression_boxes"], full["regression_boxes"][mask])
    assert res["batch_idx"] == 7 and res["pred_poses"].shape == (2, 4, 4)

    class Wrapper(ck.PredictMixin):
        BoxDreamer = m
    assert torch.equal(Wrapper().predict_step(data, 0)["pred_poses"], res["pred_poses"])
