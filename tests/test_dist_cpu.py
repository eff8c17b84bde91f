"""CPU, gloo, world_size 2: the N>1 host logic of boxdreamer_b200/dist.py (shard bounds, one-collective weight
broadcast, ragged all-gather of packed pose/corner records).  The same code runs over NCCL on the GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from boxdreamer_b200 import dist as bdist
from boxdreamer_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) weights: rank 0 owns them, one broadcast
        shapes = synth.decoder_param_shapes(num_layers=1)
        sd = synth.synth_decoder_state_dict(0, num_layers=1) if rank == 0 else None
        got = bdist.broadcast_state(sd, shapes, src=0)
        ref = synth.synth_decoder_state_dict(0, num_layers=1)
        ok_w = all(torch.equal(got[k], ref[k]) for k in ref) and list(got.keys()) == list(ref.keys())
        # (2) inputs: contiguous shard of a 5-query batch (ragged: 3 + 2)
        data = synth.synth_inputs(5, 2, 224, seed=9, with_images=False)
        local = bdist.shard_batch(data, world, rank)
        lo, hi = bdist.shard_bounds(5, world, rank)
        ok_s = local["poses"].shape[0] == hi - lo and torch.equal(local["poses"], data["poses"][lo:hi])
        # (3) results: pack -> ragged all-gather -> unpack, every rank sees the full batch in order
        poses = data["poses"][:, -1].clone()
        poses[0] = 0  # a failed solve stays the zero matrix
        corners = data["bbox_proj_crop"][:, -1]
        rec = bdist.pack_results(poses[lo:hi], corners[lo:hi])
        counts = [bdist.shard_bounds(5, world, r)[1] - bdist.shard_bounds(5, world, r)[0] for r in range(world)]
        allrec = bdist.all_gather_results(rec, counts)
        P, Cn = bdist.unpack_results(allrec)
        ok_g = allrec.shape == (5, bdist.RECORD) and torch.allclose(P, poses) and torch.allclose(Cn, corners)
        # (4) equal shards: one collective, double-buffered, waited on one step later (what bench.py --gpus N does)
        recs = [torch.full((3, bdist.RECORD), float(10 * rank + k)) for k in range(2)]
        outs, works = [None, None], [None, None]
        for k in range(2):
            outs[k], works[k] = bdist.gather_records(recs[k], async_op=True)
        ok_e = True
        for k in range(2):
            works[k].wait()
            want = torch.cat([torch.full((3, bdist.RECORD), float(10 * r + k)) for r in range(world)])
            ok_e = ok_e and torch.equal(outs[k], want)
        sync_out, none = bdist.gather_records(recs[0])
        ok_e = ok_e and none is None and torch.equal(sync_out, outs[0])
        ret[rank] = (ok_w, ok_s, ok_g and ok_e)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_roundtrip():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == (True, True, True), f"rank {r}: {ret[r]}"


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 64, 511, 512):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
