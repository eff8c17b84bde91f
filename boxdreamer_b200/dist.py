"""Query-batch sharding across GPUs: one process per GPU, `torch.distributed` plumbing.

Every query is independent (no cross-sample op in BoxDreamer.forward), so the data path has no collective:
  * weights: rank 0 holds the checkpoint, everyone else receives it with ONE broadcast of a flat fp32 blob
    (replaces "every rank loads the checkpoint itself", run.py:172-184);
  * inputs: contiguous split of the batch dimension;
  * results: one all-gather per batch of a packed [B_local, 28] fp32 record (R|t 12 floats + 8 corners x 2)
    = 112 B/query (replaces the gloo pickle gather of src/utils/comm.py:179-219).
Backend-agnostic: NCCL over NVLink on the GPU box, gloo in the CPU tests (world_size 2).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RECORD = 28  # 12 (R|t, row-major 3x4) + 16 (8 corners x (x, y) in normalised [-1, 1] crop coordinates)


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous split of n items over `world` ranks; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(data: dict, world: int, rank: int) -> dict:
    """Slices every tensor whose first dim is the batch (all keys of the input dict are, base.py:725-765)."""
    B = data["query_idx"].shape[0]
    lo, hi = shard_bounds(B, world, rank)
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B else v) for k, v in data.items()}


def flatten_state(sd: dict, keys=None):
    keys = list(sd.keys()) if keys is None else keys
    flat = torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in keys])
    return flat, keys


def unflatten_state(flat: torch.Tensor, shapes: dict) -> dict:
    out, off = {}, 0
    for k, shp in shapes.items():
        n = 1
        for s in shp:
            n *= int(s)
        out[k] = flat[off:off + n].view(*shp)
        off += n
    assert off == flat.numel(), "blob size does not match the state-dict layout"
    return out


def broadcast_state(sd: dict | None, shapes: dict, src: int = 0, device=None, group=None) -> dict:
    """Rank `src` (a GLOBAL rank, as torch.distributed.broadcast takes it) passes its state dict; every rank of `group`
    returns an identical one (a single collective)."""
    total = 0
    for shp in shapes.values():
        n = 1
        for s in shp:
            n *= int(s)
        total += n
    rank = dist.get_rank()   # global rank: `src` is global too (a group rank would be wrong for non-default groups)
    if rank == src:
        flat, _ = flatten_state(sd, list(shapes.keys()))
        flat = flat.to(device) if device is not None else flat
        assert flat.numel() == total
    else:
        flat = torch.empty(total, dtype=torch.float32, device=device)
    dist.broadcast(flat, src=src, group=group)
    return unflatten_state(flat, shapes)


def pack_results(poses: torch.Tensor, corners_norm: torch.Tensor) -> torch.Tensor:
    """poses [B,4,4], corners_norm [B,8,2] -> [B,28] fp32."""
    B = poses.shape[0]
    return torch.cat([poses[:, :3, :].reshape(B, 12).float(), corners_norm.reshape(B, 16).float()], dim=1).contiguous()


def unpack_results(rec: torch.Tensor):
    B = rec.shape[0]
    poses = torch.zeros(B, 4, 4, dtype=rec.dtype, device=rec.device)
    poses[:, :3, :] = rec[:, :12].view(B, 3, 4)
    # a failed solve is the all-zero matrix (box_utils.py:136): keep [3,3] = 0 there
    ok = rec[:, :12].abs().sum(dim=1) > 0
    poses[:, 3, 3] = ok.to(rec.dtype)
    return poses, rec[:, 12:].view(B, 8, 2)


def gather_records(rec_local: torch.Tensor, out: torch.Tensor | None = None, group=None, async_op: bool = False):
    """Equal shards (every rank holds [B, 28], as Engine.forward_packed writes it): ONE collective into `out`
    [world * B, 28] in rank order, nothing else on the stream -- no packing, padding or re-assembly kernels.  With
    async_op=True the returned work handle is waited on when the records are consumed, so the gather of step k overlaps
    the encoder of step k + 1 (use alternating `rec_local` / `out` buffers).  Returns (out, work | None)."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty(world * rec_local.shape[0], rec_local.shape[1], dtype=rec_local.dtype, device=rec_local.device)
    work = dist.all_gather_into_tensor(out, rec_local, group=group, async_op=async_op)
    return out, (work if async_op else None)


def all_gather_results(rec_local: torch.Tensor, counts, group=None) -> torch.Tensor:
    """All-gather of ragged shards: every rank returns [sum(counts), 28] in rank order."""
    world = dist.get_world_size(group)
    mx = max(counts)
    pad = torch.zeros(mx, RECORD, dtype=rec_local.dtype, device=rec_local.device)
    pad[: rec_local.shape[0]] = rec_local
    out = torch.empty(world * mx, RECORD, dtype=rec_local.dtype, device=rec_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)
