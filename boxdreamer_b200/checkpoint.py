"""Checkpoint I/O and the predict step.

Mirrors (reference file:line)
  warp_model                      src/demo/demo.py:564-573       (Lightning checkpoints prefix every key with "BoxDreamer.")
  checkpoint download / load      run.py:172-184                 (every rank loads the file itself)
  PL_BoxDreamer.test_step         src/lightning/BoxDreamer_lightning_model.py:228-243   (there is no predict_step upstream)

`load_checkpoint` reads a .safetensors file (scripts/tools/make_safetensor.py's output) or a torch / Lightning checkpoint and
returns the 177-tensor decoder state dict; `load_model` does that on rank 0 only and hands the weights to the other ranks
with ONE broadcast of a flat blob (boxdreamer_b200.dist.broadcast_state) when torch.distributed is initialised;
`predict_step` is the missing Lightning hook: batch in, the query-view predictions out, nothing leaves the device except
what the caller asks for.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import dist as bdist

__all__ = ["warp_model", "load_checkpoint", "load_model", "predict_step", "PredictMixin"]

PREFIX = "BoxDreamer."


def warp_model(state_dict: dict) -> dict:
    """Unwraps {"state_dict": ...} and strips the Lightning attribute prefix (BoxDreamer_lightning_model.py:34)."""
    sd = state_dict.get("state_dict", state_dict) if isinstance(state_dict, dict) else state_dict
    return {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in sd.items()}


def _is_safetensors(path: str) -> bool:
    """Upstream names its files `.safetensor` (run.py:172-183 downloads `BoxDreamer-vitb.safetensor`, scripts/tools/
    make_safetensor.py writes `<ckpt>.safetensor`); the library's own suffix is `.safetensors`.  Accept both, and sniff the
    header (8-byte little-endian length followed by a JSON object) for files with any other name."""
    if path.endswith((".safetensors", ".safetensor")):
        return True
    try:
        with open(path, "rb") as f:
            head = f.read(9)
        n = int.from_bytes(head[:8], "little")
        return len(head) == 9 and head[8:9] == b"{" and 0 < n < os.path.getsize(path)
    except OSError:
        return False


def _load_file(path: str) -> dict:
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    if _is_safetensors(path):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    # weights_only: tensors and plain containers only -- a checkpoint is untrusted input, never unpickle arbitrary objects
    return torch.load(path, map_location="cpu", weights_only=True)


def load_checkpoint(path: str) -> dict:
    """-> {"decoder.*": fp32 CPU tensors}; keys outside the model (optimizer state, loss buffers, ...) are dropped."""
    raw = _load_file(path)
    sd = warp_model(raw)
    return {k: v.detach().to(torch.float32) for k, v in sd.items() if torch.is_tensor(v) and k.startswith("decoder.")}


def load_model(model, path: str | None, dino_path: str | None = None, device=None, group=None):
    """Loads decoder (and optionally DINOv2) weights into a boxdreamer_b200.BoxDreamer.  With an initialised process
    group only rank 0 touches the file system; the other ranks pass path=None or simply ignore it."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0       # rank inside `group`; group rank 0 is the loader
    dec_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    dino_shapes = {k: tuple(v.shape) for k, v in model.rgb_encoder.model.state_dict().items()}
    dec = dino = None
    error = None
    if rank == 0:
        try:
            dec = load_checkpoint(path)
            missing = [k for k in dec_shapes if k not in dec]
            if missing:
                raise KeyError(f"checkpoint {path} lacks {len(missing)} decoder tensors, e.g. {missing[:3]}")
            if dino_path is not None:
                dino = _load_file(dino_path)
                missing = [k for k in dino_shapes if k not in dino]
                if missing:
                    raise KeyError(f"DINOv2 checkpoint {dino_path} lacks {len(missing)} tensors, e.g. {missing[:3]}")
        except Exception as exc:   # do not leave the other ranks hanging in the collective: tell them first
            error = exc
    if world > 1:
        if device is None:         # NCCL cannot broadcast CPU tensors: default to the model's device
            backend = dist.get_backend(group)
            mdev = next(model.parameters()).device
            device = mdev if (backend == "nccl" and mdev.type == "cuda") else (torch.device("cuda", torch.cuda.current_device())
                                                                                if backend == "nccl" else None)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        # status word from the loader: 0 = failed, 1 = decoder only, 2 = decoder + DINOv2
        flag = torch.tensor([0 if error is not None else (2 if dino is not None else 1)] if rank == 0 else [0], device=device)
        dist.broadcast(flag, src=src, group=group)
        status = int(flag.item())
        if status == 0:
            if error is not None:
                raise error
            raise RuntimeError("load_model: the loading rank failed to read the checkpoint (see its traceback)")
        dec = bdist.broadcast_state(dec, dec_shapes, src=src, device=device, group=group)
        if status == 2:
            dino = bdist.broadcast_state(dino, dino_shapes, src=src, device=device, group=group)
    elif error is not None:
        raise error
    model.load_state_dict({k: dec[k] for k in dec_shapes}, strict=True)
    if dino is not None:
        model.rgb_encoder.model.load_state_dict({k: dino[k] for k in dino_shapes}, strict=True)
    return model


@torch.no_grad()
def predict_step(model, batch: dict, batch_idx: int = 0, dataloader_idx: int = 0, keep_heatmaps: bool = False) -> dict:
    """One inference step on a batch dict of the reference's dataset (src/datasets/base.py:725-765): tensors are moved to the
    model's device, the hot path runs, and the query-view predictions come back as a small dict:
    pred_poses [B,4,4], regression_boxes [B,8,2], query_idx [B] (+ pred_bbox [B,8,S,S] when keep_heatmaps)."""
    dev = next(model.parameters()).device
    data = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
    was_training = model.training
    model.eval()
    try:
        out = model(data)
    finally:
        model.train(was_training)
    mask = out["camera_mask"]
    res = {"pred_poses": out["pred_poses"][mask], "regression_boxes": out["regression_boxes"][mask], "query_idx": out["query_idx"],
           "batch_idx": batch_idx, "dataloader_idx": dataloader_idx}
    if keep_heatmaps:
        res["pred_bbox"] = out["pred_bbox"][mask]
    return res


class PredictMixin:
    """Mix into the Lightning module next to `self.BoxDreamer` (BoxDreamer_lightning_model.py:34):
        class PL_BoxDreamer(PredictMixin, pl.LightningModule): ...
    `trainer.predict(model, dataloader)` then gathers the dicts of `predict_step`."""

    def predict_step(self, batch, batch_idx, dataloader_idx=0):
        return predict_step(self.BoxDreamer, batch, batch_idx, dataloader_idx)
