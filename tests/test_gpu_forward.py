"""-m gpu: the whole hot path through the drop-in module / C ABI against (a) the committed golden fixtures generated
from the unmodified reference and (b) the CPU oracle run on the same seeded inputs.

Tolerances (BASELINE.json north_star): corner indices bit-exact; heat-map logits within 1e-4 relative
(max|d| / max|ref|) on the fp32 'exact' path; R|t within 1e-3 deg / 1e-4 relative against the same-corner PnP.
The bf16 tensor path is compared statistically (SURVEY.md section 7 'Tolerance vs precision').
"""
import os

import numpy as np
import pytest
import torch

from boxdreamer_b200 import BoxDreamer, _lib, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _config(img_size=224):
    from boxdreamer_b200.config import make_config
    return make_config(img_size)


@pytest.fixture(scope="module")
def weights():
    return synth.synth_decoder_state_dict(0), synth.synth_dino_state_dict(0)


def _model(weights, precision):
    dec, dino = weights
    m = BoxDreamer(_config(), precision=precision)
    m.load_state_dict(dec, strict=True)
    m.rgb_encoder.model.load_state_dict(dino, strict=True)
    return m.cuda().eval()


def _to_cuda(data):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}


def _scaled(got, ref):
    return float((got.float().cpu() - ref.float().cpu()).abs().max() / (ref.float().abs().max() + 1e-30))


def _rot_err_deg(Ra, Rb):
    """Geodesic angle between two rotations, computed from the chordal distance so that it stays accurate for the
    float32-rounded matrices the C ABI returns (acos((tr-1)/2) loses half the digits near 0)."""
    s = min(np.linalg.norm(np.asarray(Ra, dtype=np.float64) - np.asarray(Rb, dtype=np.float64)) / (2.0 * np.sqrt(2.0)), 1.0)
    return float(np.degrees(2.0 * np.arcsin(s)))


@pytest.mark.parametrize("name,B,T,seed,qidx", [("forward_b1t2.npz", 1, 2, 1235, None), ("forward_b2t3.npz", 2, 3, 1236, [2, 0])])
def test_exact_path_matches_reference_golden(weights, name, B, T, seed, qidx):
    gold = np.load(os.path.join(GOLD, name))
    m = _model(weights, "exact")
    data = synth.synth_inputs(B, T, 224, seed=seed)
    if qidx is not None:
        data["query_idx"] = torch.tensor(qidx, dtype=torch.int64)
    d = _to_cuda(data)
    eng = m._engine_for(d["images"], B, T)
    feats = eng.dino_forward(d["images"].view(B * T, 3, 224, 224).contiguous())
    st, sc = int(gold["stride_tok"]), int(gold["stride_ch"])
    e = _scaled(feats[:, ::st, ::sc], torch.from_numpy(gold["dino_feats_sub"]))
    assert e <= 1e-4, f"DINOv2 patch tokens vs reference: scaled err {e:.3e}"
    cs = gold["dino_feats_cs"]
    assert abs(feats.double().sum().item() - cs[0]) <= 1e-4 * cs[1]
    heat, logits = eng.decoder_forward(d["bbox_feat"].contiguous(), feats, d["query_idx"], want_logits=True)
    lg = logits.view(B, 256, 1568)
    e = _scaled(lg[:, ::4, ::7], torch.from_numpy(gold["logits_sub"]))
    assert e <= 1e-4, f"heat-map logits vs reference: scaled err {e:.3e} (tolerance 1e-4)"
    cs = gold["logits_cs"]
    assert abs(lg.double().sum().item() - cs[0]) <= 1e-4 * cs[1]
    e = _scaled(heat[:, :, ::4, ::4], torch.from_numpy(gold["query_ret_sub"]))
    assert e <= 1e-4, f"query_ret vs reference: scaled err {e:.3e}"
    px, nm, idx = eng.corners_topk(heat, want_idx=True)
    ref_idx = torch.from_numpy(gold["topk_idx"][:, :, :20]).long()
    got_sorted = torch.sort(idx.cpu().long(), dim=2).values
    ref_sorted = torch.sort(ref_idx, dim=2).values
    assert torch.equal(got_sorted, ref_sorted), "top-20 index sets must equal the reference's (bit-exact)"
    assert torch.allclose(nm.cpu(), torch.from_numpy(gold["keypoints_norm"]), atol=1e-6, rtol=0)
    # the full module call (drop-in API): dict in, same dict out
    out = m(d)
    assert out is d
    assert torch.equal(out["camera_mask"].cpu(), torch.from_numpy(gold["camera_mask"]))
    assert torch.allclose(out["regression_boxes"].cpu(), torch.from_numpy(gold["regression_boxes"]), atol=1e-6, rtol=0)
    for k in ("pred_bbox", "pred_poses", "pred_intrinsics", "regression_boxes", "camera_mask"):
        assert k in out
    assert out["pred_bbox"].shape == (B, T, 8, 224, 224) and out["pred_poses"].shape == (B, T, 4, 4)


def test_exact_path_matches_oracle_full_tensors(weights):
    from oracle import boxdreamer_oracle as O
    dec, dino = weights
    B, T = 2, 3
    data = synth.synth_inputs(B, T, 224, seed=77)
    data["query_idx"] = torch.tensor([1, 2], dtype=torch.int64)
    with torch.no_grad():
        ref = O.forward(data, dec, dino)
    m = _model(weights, "exact")
    out = m(_to_cuda(data))
    e = _scaled(out["pred_bbox"], ref["pred_bbox"])
    assert e <= 1e-4, f"pred_bbox scaled err {e:.3e}"
    assert torch.allclose(out["regression_boxes"].cpu(), ref["regression_boxes"], atol=1e-6, rtol=0)
    # PnP on identical corners: GPU (fp64 DLT+LM) vs the oracle's numpy restatement
    mask = ref["camera_mask"]
    got = out["pred_poses"].cpu()[mask].double().numpy()
    exp = ref["pred_poses"][mask].double().numpy()
    for b in range(B):
        assert _rot_err_deg(got[b, :3, :3], exp[b, :3, :3]) <= 1e-3
        assert np.linalg.norm(got[b, :3, 3] - exp[b, :3, 3]) <= 1e-4 * max(np.linalg.norm(exp[b, :3, 3]), 1e-6)
    # non-query rows keep the input poses
    assert torch.equal(out["pred_poses"].cpu()[~mask], data["poses"][~mask])


def test_bf16_tensor_path_statistical_agreement(weights):
    """bf16 tcgen05 path vs the fp32 oracle at a small ragged shape.  Stated tolerance: logits mean|d| <= 2.5e-3 * max|ref| and
    max|d| <= 1.5e-2 * max|ref| after 24 transformer layers in bf16 -- about twice the values measured on B200 (1.0e-3 /
    5.9e-3).  With these random-init weights the maps are noise, so corner agreement is only printed here; corners and poses of
    the bf16 path are gated on well-conditioned maps in tests/test_gpu_bf16_parity.py (config-2 shape)."""
    from oracle import boxdreamer_oracle as O
    dec, dino = weights
    B, T = 2, 3
    data = synth.synth_inputs(B, T, 224, seed=78)
    with torch.no_grad():
        ref = O.forward(data, dec, dino, with_pnp=False)
    m = _model(weights, "bf16")
    d = _to_cuda({k: (v.to(torch.bfloat16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()})
    eng = m._engine_for(d["images"], B, T)
    feats = eng.dino_forward(d["images"].view(B * T, 3, 224, 224).contiguous())
    heat, logits = eng.decoder_forward(d["bbox_feat"].contiguous(), feats, d["query_idx"], want_logits=True)
    ref_l = ref["logits"].reshape(-1, 1568)
    diff = (logits.cpu() - ref_l).abs()
    scale = ref_l.abs().max().item()
    print(f"bf16 logits: max|d|/max|ref| = {diff.max().item() / scale:.3e}, mean|d|/max|ref| = {diff.mean().item() / scale:.3e}")
    assert not torch.isnan(logits).any()
    assert diff.mean().item() <= 2.5e-3 * scale
    assert diff.max().item() <= 1.5e-2 * scale
    px, nm = eng.corners_topk(heat)
    dist = (px.cpu() - ref["keypoints_px"]).norm(dim=-1)
    print(f"bf16 corners: median dist {dist.median().item():.3f} px, frac<2px {float((dist < 2).float().mean()):.3f}")
    out = m(d)
    assert out["pred_poses"].dtype == torch.bfloat16 and out["pred_bbox"].dtype == torch.bfloat16


def test_host_buffer_entry_matches_device_entry(weights):
    m = _model(weights, "exact")
    B, T = 1, 2
    data = synth.synth_inputs(B, T, 224, seed=1235)
    d = _to_cuda(data)
    eng = m._engine_for(d["images"], B, T)
    mask = torch.zeros(B, T, dtype=torch.bool)
    mask[torch.arange(B), data["query_idx"]] = True
    K_q = data["non_ndc_intrinsics"][mask].float().contiguous()
    X_q = data["bbox_3d"][mask].float().contiguous()
    heat, px, nm, poses = eng.forward(d["images"].contiguous(), d["bbox_feat"].contiguous(), d["query_idx"], X_q.cuda(), K_q.cuda())
    torch.cuda.synchronize()
    hh, hpx, hnm, hposes = eng.forward_host(data["images"].contiguous(), data["bbox_feat"].contiguous(), data["query_idx"], X_q, K_q,
                                           want_heat=True)
    assert torch.equal(hpx, px.cpu()) and torch.equal(hposes, poses.cpu()) and torch.equal(hh, heat.cpu())


def test_errors_are_loud(weights):
    m = _model(weights, "exact")
    data = synth.synth_inputs(1, 2, 224, seed=1)
    with pytest.raises(_lib.BoxDreamerLibError):
        m(data)  # CPU tensors: no fallback
    bad = _to_cuda(synth.synth_inputs(1, 2, 210, seed=1))
    with pytest.raises(AssertionError):
        m(bad)  # betr.py:269-271


def test_336px_long_sequence_config(weights):
    """BASELINE config 4 geometry at a test-sized batch: 336 px crops (P = 576, DINOv2 pos-embed resampled to 24x24,
    581 DINOv2 tokens) and N = T*P = 1728 decoder tokens (13.5 key tiles: exercises the trimmed tail tile)."""
    from oracle import boxdreamer_oracle as O
    from boxdreamer_b200.config import make_config
    dec, dino = weights
    B, T, S = 1, 3, 336
    data = synth.synth_inputs(B, T, S, seed=91)
    with torch.no_grad():
        ref = O.forward(data, dec, dino, with_pnp=False)
    for precision, dtype, tol_max in (("exact", torch.float32, 1e-4), ("bf16", torch.bfloat16, 2e-2)):
        m = BoxDreamer(make_config(S), precision=precision)
        m.load_state_dict(dec, strict=True)
        m.rgb_encoder.model.load_state_dict(dino, strict=True)
        m = m.cuda().eval()
        d = _to_cuda({k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()})
        eng = m._engine_for(d["images"], B, T)
        feats = eng.dino_forward(d["images"].view(B * T, 3, S, S).contiguous())
        heat, logits = eng.decoder_forward(d["bbox_feat"].contiguous(), feats, d["query_idx"], want_logits=True)
        e = _scaled(logits.view(B, 576, 1568), ref["logits"])
        print(f"336px {precision}: logits scaled max err {e:.3e}")
        assert e <= tol_max, f"{precision}: {e:.3e}"
        if precision == "exact":
            px, nm, idx = eng.corners_topk(heat, want_idx=True)
            assert torch.equal(torch.sort(idx.cpu().long(), dim=2).values, torch.sort(ref["topk_idx"], dim=2).values)


def test_forward_from_projected_corners_equals_forward_from_heatmaps(weights):
    """Device-side input synthesis: a batch that carries 'bbox_proj_px' instead of 'bbox_feat' must give the same result
    (exact precision; the rasterised maps differ from the CPU-made ones by <= 4e-6, tests/test_gpu_simt.py)."""
    B, T = 2, 3
    data = synth.synth_inputs(B, T, 224, seed=79)
    px = (data["bbox_proj_crop"].float() + 1) / 2 * 224
    data["bbox_feat"] = synth.make_heatmaps(px.view(B * T, 8, 2), 224, group=T).view(B, T, 8, 224, 224)   # dataset semantics
    m = _model(weights, "exact")
    ref = m(_to_cuda(data))
    alt = {k: v for k, v in _to_cuda(data).items() if k != "bbox_feat"}
    alt["bbox_proj_px"] = px
    out = m(alt)
    assert _scaled(out["pred_bbox"], ref["pred_bbox"]) <= 1e-4
    assert torch.allclose(out["regression_boxes"], ref["regression_boxes"], atol=0.2 / 224 * 2)   # top-20 means: <= 0.2 px
    # host-buffer entry with corners (64 B per view) == host-buffer entry with maps
    eng = m._engine_for(ref["pred_bbox"], B, T)
    mask = ref["camera_mask"].cpu()
    args = (data["query_idx"], data["bbox_3d"][mask].float().contiguous(), data["non_ndc_intrinsics"][mask].float().contiguous())
    h1, px1, _, p1 = eng.forward_host(data["images"].contiguous(), data["bbox_feat"].contiguous(), *args, want_heat=True)
    h2, px2, _, p2 = eng.forward_host_px(data["images"].contiguous(), alt["bbox_proj_px"].contiguous(), *args, want_heat=True)
    assert float((h1 - h2).abs().max()) <= 1e-4 * float(h1.abs().max())
    assert float((px1 - px2).abs().max()) <= 0.2


def test_reference_feature_cache_equals_full_forward(weights):
    """Rank-2 widening: queries that share a reference set re-use its encoder tokens.  Same kernels, exact precision ->
    the cached path must reproduce the full forward bit for bit."""
    B, R = 3, 2
    base = synth.synth_inputs(1, R + 1, 224, seed=80)                     # one sample: R references + a query slot
    other = synth.synth_inputs(B, R + 1, 224, seed=81)                    # B different query crops / intrinsics
    full = {}
    for k, v in other.items():
        if torch.is_tensor(v) and v.dim() >= 2 and v.shape[1] == R + 1:
            w = v.clone()
            w[:, :R] = base[k][0, :R]                                      # every sample lists the same references
            full[k] = w
        else:
            full[k] = v
    full["query_idx"] = torch.full((B,), R, dtype=torch.int64)
    m = _model(weights, "exact")
    ref = m(_to_cuda(full))
    cache = m.encode_references(base["images"][0, :R].cuda(), base["bbox_feat"][0, :R].cuda())
    out = m.forward_with_references(full["images"][:, R].cuda(), cache, full["bbox_3d"][:, R].cuda(), full["non_ndc_intrinsics"][:, R].cuda())
    mask = ref["camera_mask"]
    assert torch.equal(out["pred_bbox"], ref["pred_bbox"][mask])
    assert torch.equal(out["regression_boxes"], ref["regression_boxes"][mask])
    assert torch.equal(out["pred_poses"], ref["pred_poses"][mask])


@pytest.mark.parametrize("precision,dtype", [("exact", torch.float32), ("bf16", torch.bfloat16)])
def test_graph_replay_equals_eager_launches(weights, monkeypatch, precision, dtype):
    """bd_forward at small shapes stages its inputs and replays a captured CUDA graph of the launch chain (call 1 eager,
    call 2 captured, call 3+ replayed): every call must return bit-identical results, equal to an engine created with
    BOXDREAMER_B200_GRAPHS=0, also for fresh inputs (the replay must read the staged inputs, not captured pointers), on a
    non-default stream, and through the host-buffer entry (per-stage graphs)."""
    B, T = 2, 3
    def inputs(seed):
        data = synth.synth_inputs(B, T, 224, seed=seed)
        d = _to_cuda({k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()})
        mask = torch.zeros(B, T, dtype=torch.bool)
        mask[torch.arange(B), data["query_idx"]] = True
        return (d["images"].contiguous(), d["bbox_feat"].contiguous(), d["query_idx"],
                data["bbox_3d"][mask].float().cuda().contiguous(), data["non_ndc_intrinsics"][mask].float().cuda().contiguous())
    monkeypatch.setenv("BOXDREAMER_B200_GRAPHS", "0")
    m0 = _model(weights, precision)
    a, b = inputs(91), inputs(92)
    eng0 = m0._engine_for(a[0], B, T)
    ref_a = [t.clone() for t in eng0.forward(*a)]
    ref_b = [t.clone() for t in eng0.forward(*b)]
    monkeypatch.setenv("BOXDREAMER_B200_GRAPHS", "1")
    m1 = _model(weights, precision)
    eng1 = m1._engine_for(a[0], B, T)
    assert eng1.handle.value != eng0.handle.value
    l0 = eng1.lib.bd_launch_count(eng1.handle)
    for i, (x, ref) in enumerate([(a, ref_a), (a, ref_a), (a, ref_a), (b, ref_b), (a, ref_a)]):
        got = eng1.forward(*x)
        torch.cuda.synchronize()
        for g, r, nm in zip(got, ref, ("heat", "corners_px", "corners_norm", "poses")):
            assert torch.equal(g, r), f"call {i}: {nm} differs between graph replay and eager launches"
    per_call = (eng1.lib.bd_launch_count(eng1.handle) - l0) / 5
    assert per_call > 100, "replayed launches are not counted"
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        got = eng1.forward(*b)
    side.synchronize()
    assert all(torch.equal(g, r) for g, r in zip(got, ref_b)), "graph replay on a side stream differs"
    # host-buffer entry: encoder and decoder stages are replayed separately
    host = [t.cpu().contiguous() for t in b]
    for i in range(3):
        got = eng1.forward_host(*host, want_heat=True)
        for g, r, nm in zip(got, ref_b, ("heat", "corners_px", "corners_norm", "poses")):
            assert torch.equal(g.cpu(), r.cpu()), f"host call {i}: {nm} differs"


@pytest.mark.parametrize("B,T", [(2, 3), (5, 6)])
def test_forward_packed_record_equals_packed_forward(weights, B, T):
    """bd_forward_packed: the [B, 28] record written by the PnP kernel's epilogue equals dist.pack_results of bd_forward's
    poses and normalised corners (staged / graph-replayed shape and eagerly launched shape)."""
    from boxdreamer_b200 import dist as bdist
    m = _model(weights, "bf16")
    data = synth.synth_inputs(B, T, 224, seed=77)
    d = _to_cuda({k: (v.to(torch.bfloat16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()})
    mask = torch.zeros(B, T, dtype=torch.bool)
    mask[torch.arange(B), data["query_idx"]] = True
    X = data["bbox_3d"][mask].float().cuda().contiguous()
    K = data["non_ndc_intrinsics"][mask].float().cuda().contiguous()
    eng = m._engine_for(d["images"], B, T)
    args = (d["images"].contiguous(), d["bbox_feat"].contiguous(), d["query_idx"], X, K)
    for _ in range(3):   # eager, captured, replayed (small shape)
        _, _, nm, poses = eng.forward(*args, want_heat=False)
        rec = eng.forward_packed(*args)
        torch.cuda.synchronize()
        assert torch.equal(rec, bdist.pack_results(poses, nm))
        P, Cn = bdist.unpack_results(rec)
        assert torch.equal(P[:, :3], poses[:, :3]) and torch.equal(Cn, nm)


def test_pipelined_host_entry_matches_blocking_call(weights):
    """bd_forward_host_submit / _wait on alternating staging slots (batch k+1 submitted before batch k is waited for) returns,
    for every batch, exactly what the blocking bd_forward_host returns -- for the map inputs and for the projected-corner
    (_px) inputs, with different batches in flight at the same time."""
    m = _model(weights, "bf16")
    B, T = 2, 3
    batches = []
    for seed in (201, 202, 203, 204):
        data = synth.synth_inputs(B, T, 224, seed=seed, dtype=torch.bfloat16)
        mask = torch.zeros(B, T, dtype=torch.bool)
        mask[torch.arange(B), data["query_idx"]] = True
        batches.append((data["images"].contiguous().pin_memory(), data["bbox_feat"].contiguous().pin_memory(),
                        data["query_idx"].contiguous(), data["bbox_3d"][mask].float().contiguous(),
                        data["non_ndc_intrinsics"][mask].float().contiguous(),
                        ((data["bbox_proj_crop"].float() + 1) / 2 * 224).contiguous()))
    eng = m._engine_for(batches[0][0].cuda(), B, T)
    for use_px in (False, True):
        ref = []
        for im, bb, qi, X, K, px in batches:
            out = eng.forward_host_px(im, px, qi, X, K, want_heat=True) if use_px else eng.forward_host(im, bb, qi, X, K, want_heat=True)
            ref.append([t.clone() for t in out])
        for _ in range(2):   # second round: the per-slot stage graphs are replayed
            got = [None] * len(batches)
            for i, (im, bb, qi, X, K, px) in enumerate(batches):
                got[i] = eng.forward_host_submit(i & 1, im, None if use_px else bb, qi, X, K, bbox_px=px if use_px else None, want_heat=True)
                if i > 0:
                    eng.forward_host_wait((i - 1) & 1)
                    assert all(torch.equal(g, r) for g, r in zip(got[i - 1], ref[i - 1])), f"batch {i - 1} (px={use_px})"
            eng.forward_host_wait((len(batches) - 1) & 1)
            assert all(torch.equal(g, r) for g, r in zip(got[-1], ref[-1]))


@pytest.mark.parametrize("B,T,qidx", [(3, 3, [2, 0, 1]), (2, 6, [5, 3])])
def test_last_decoder_block_on_query_rows_is_bit_identical(weights, monkeypatch, B, T, qidx):
    """The tensor path runs the decoder's last block only for the query view's tokens (attention from a query window over all
    keys, proj / LayerNorm / MLP on the gathered rows) -- the other rows are never read (betr.py:419-430).  Per-row arithmetic
    is unchanged, so the logits must not change by a single bit against BD_LAST_LAYER_PRUNE=0."""
    data = synth.synth_inputs(B, T, 224, seed=311)
    data["query_idx"] = torch.tensor(qidx, dtype=torch.int64)
    d = _to_cuda({k: (v.to(torch.bfloat16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()})
    m = _model(weights, "bf16")
    eng = m._engine_for(d["images"], B, T)
    feats = eng.dino_forward(d["images"].view(B * T, 3, 224, 224).contiguous())
    outs = []
    for flag in ("1", "0", "1"):
        monkeypatch.setenv("BD_LAST_LAYER_PRUNE", flag)
        heat, logits = eng.decoder_forward(d["bbox_feat"].contiguous(), feats, d["query_idx"], want_logits=True)
        torch.cuda.synchronize()
        outs.append((heat.clone(), logits.clone()))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0]), "pruned last block changed the logits"
    assert torch.equal(outs[0][1], outs[2][1])


def test_last_decoder_block_on_query_rows_336px(monkeypatch):
    """Same bit-identity at 336 px: P = 576 tokens per view is not a multiple of the 128-row query tile, so the query window ends
    inside a tile (rows beyond it belong to the next view or to the padding and are clipped by the compact O tensor map)."""
    dec, dino = synth.synth_decoder_state_dict(0), synth.synth_dino_state_dict(0)
    m = BoxDreamer(_config(336), precision="bf16")
    m.load_state_dict(dec, strict=True)
    m.rgb_encoder.model.load_state_dict(dino, strict=True)
    m = m.cuda().eval()
    B, T = 2, 3
    data = synth.synth_inputs(B, T, 336, seed=312)
    data["query_idx"] = torch.tensor([2, 0], dtype=torch.int64)   # last view: the window ends at the sequence end
    d = _to_cuda({k: (v.to(torch.bfloat16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()})
    eng = m._engine_for(d["images"], B, T)
    feats = eng.dino_forward(d["images"].view(B * T, 3, 336, 336).contiguous())
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("BD_LAST_LAYER_PRUNE", flag)
        heat, logits = eng.decoder_forward(d["bbox_feat"].contiguous(), feats, d["query_idx"], want_logits=True)
        torch.cuda.synchronize()
        outs.append(logits.clone())
    assert torch.isfinite(outs[0]).all() and torch.equal(outs[0], outs[1])


def test_mixed_host_and_device_entries_share_the_workspace_safely(weights):
    """A host batch in flight (bd_forward_host_submit, internal streams) and device-pointer calls on the caller's stream use the
    same workspace: each side must wait for the other on the device.  Interleave them without any host synchronisation in
    between and compare every result with the same call made alone."""
    m = _model(weights, "bf16")
    B, T = 2, 3
    def make(seed):
        data = synth.synth_inputs(B, T, 224, seed=seed, dtype=torch.bfloat16)
        mask = torch.zeros(B, T, dtype=torch.bool)
        mask[torch.arange(B), data["query_idx"]] = True
        return (data["images"].contiguous(), data["bbox_feat"].contiguous(), data["query_idx"].contiguous(),
                data["bbox_3d"][mask].float().contiguous(), data["non_ndc_intrinsics"][mask].float().contiguous())
    host = [tuple(t.pin_memory() for t in make(401 + i)) for i in range(3)]
    dev = [tuple(t.cuda() for t in make(501 + i)) for i in range(3)]
    eng = m._engine_for(dev[0][0], B, T)
    ref_h = [[t.clone() for t in eng.forward_host(*h, want_heat=True)] for h in host]
    ref_d = [[t.clone() for t in eng.forward(*d)] for d in dev]
    torch.cuda.synchronize()
    for rnd in range(2):
        got_h, got_d = [], []
        for i in range(3):
            got_h.append(eng.forward_host_submit(i & 1, *host[i], want_heat=True))
            got_d.append(eng.forward(*dev[i]))          # no synchronisation: ordered on the device
            if i > 0:
                eng.forward_host_wait((i - 1) & 1)
                got_h[i - 1] = [t.clone() for t in got_h[i - 1]]   # a slot's pinned result buffers are reused by its next submit
        eng.forward_host_wait(0)
        eng.forward_host_wait(1)
        got_h[2] = [t.clone() for t in got_h[2]]
        torch.cuda.synchronize()
        for i in range(3):
            assert all(torch.equal(g, r) for g, r in zip(got_h[i], ref_h[i])), f"round {rnd}: host batch {i} corrupted by a device call"
            assert all(torch.equal(g, r) for g, r in zip(got_d[i], ref_d[i])), f"round {rnd}: device call {i} corrupted by a host batch"


def test_full_config2_size_properties(weights, monkeypatch):
    """BASELINE config 2 at its full size (64 queries x 6 views, 224 px, bf16) -- too large for the CPU oracle, so checked through
    size-independent properties of the path: (a) determinism (two runs bit-identical), (b) every query is independent: permuting
    the batch permutes the results bit for bit (no cross-sample operation), and a query evaluated alone (B = 1, graph
    replay path) equals its row of the batch, (c) the query-window last block equals the full one at this size, (d) poses are
    rigid (R orthonormal, det +1) or the zero matrix of a failed solve."""
    B, T = 64, 6
    data = synth.synth_inputs(B, T, 224, seed=4242, dtype=torch.bfloat16)
    mask = torch.zeros(B, T, dtype=torch.bool)
    mask[torch.arange(B), data["query_idx"]] = True
    img, bb, qi = data["images"].cuda().contiguous(), data["bbox_feat"].cuda().contiguous(), data["query_idx"].cuda()
    X = data["bbox_3d"][mask].float().cuda().contiguous()
    K = data["non_ndc_intrinsics"][mask].float().cuda().contiguous()
    m = _model(weights, "bf16")
    eng = m._engine_for(img, B, T)
    a = [t.clone() for t in eng.forward(img, bb, qi, X, K)]
    b = eng.forward(img, bb, qi, X, K)
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(a, b)), "not deterministic"
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(5)).cuda()
    p = eng.forward(img[perm].contiguous(), bb[perm].contiguous(), qi[perm].contiguous(), X[perm].contiguous(), K[perm].contiguous())
    torch.cuda.synchronize()
    for x, y, nm in zip(a, p, ("heat", "corners_px", "corners_norm", "poses")):
        assert torch.equal(x[perm], y), f"{nm}: batch permutation is not a permutation of the results"
    for i in (0, 37):
        one = eng.forward(img[i:i + 1].contiguous(), bb[i:i + 1].contiguous(), qi[i:i + 1].contiguous(), X[i:i + 1].contiguous(),
                          K[i:i + 1].contiguous())
        one = eng.forward(img[i:i + 1].contiguous(), bb[i:i + 1].contiguous(), qi[i:i + 1].contiguous(), X[i:i + 1].contiguous(),
                          K[i:i + 1].contiguous())   # second call: captured graph
        torch.cuda.synchronize()
        assert torch.equal(one[1][0], a[1][i]) and torch.equal(one[3][0], a[3][i]), f"query {i} alone differs from its batch row"
    feats = eng.dino_forward(img.view(B * T, 3, 224, 224))
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("BD_LAST_LAYER_PRUNE", flag)
        outs.append(eng.decoder_forward(bb, feats, qi, want_logits=True)[1].clone())
    assert torch.equal(outs[0], outs[1]), "query-window last block differs from the full block at B = 64"
    R = a[3][:, :3, :3].double()
    ok = a[3][:, 3, 3] == 1
    assert ok.any()
    eye = torch.eye(3, dtype=torch.float64, device=R.device)
    assert float((R[ok] @ R[ok].transpose(1, 2) - eye).abs().max()) < 1e-5 and float((torch.linalg.det(R[ok]) - 1).abs().max()) < 1e-5
    assert float(a[3][~ok].abs().max()) == 0.0 if (~ok).any() else True


def test_config4_shape_properties(monkeypatch):
    """BASELINE config 4's shape (16 references, 336 px: N = 17 * 576 = 9792 decoder tokens per query) on a small batch: the
    long-sequence path of the attention kernel (102 key tiles, trimmed tail), checked through the same size-independent properties
    -- determinism, a query alone equals its batch row, query-window last block == full block."""
    dec, dino = synth.synth_decoder_state_dict(0), synth.synth_dino_state_dict(0)
    m = BoxDreamer(_config(336), precision="bf16")
    m.load_state_dict(dec, strict=True)
    m.rgb_encoder.model.load_state_dict(dino, strict=True)
    m = m.cuda().eval()
    B, T = 3, 17
    data = synth.synth_inputs(B, T, 336, seed=4343, dtype=torch.bfloat16)
    data["query_idx"] = torch.tensor([16, 0, 9], dtype=torch.int64)
    mask = torch.zeros(B, T, dtype=torch.bool)
    mask[torch.arange(B), data["query_idx"]] = True
    img, bb, qi = data["images"].cuda().contiguous(), data["bbox_feat"].cuda().contiguous(), data["query_idx"].cuda()
    X = data["bbox_3d"][mask].float().cuda().contiguous()
    K = data["non_ndc_intrinsics"][mask].float().cuda().contiguous()
    eng = m._engine_for(img, B, T)
    a = [t.clone() for t in eng.forward(img, bb, qi, X, K)]
    b = eng.forward(img, bb, qi, X, K)
    torch.cuda.synchronize()
    assert all(torch.isfinite(x).all() for x in a) and all(torch.equal(x, y) for x, y in zip(a, b))
    for i in range(B):
        one = eng.forward(img[i:i + 1].contiguous(), bb[i:i + 1].contiguous(), qi[i:i + 1].contiguous(), X[i:i + 1].contiguous(),
                          K[i:i + 1].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(one[0][0], a[0][i]) and torch.equal(one[3][0], a[3][i]), f"query {i} alone differs from its batch row"
    feats = eng.dino_forward(img.view(B * T, 3, 336, 336))
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("BD_LAST_LAYER_PRUNE", flag)
        outs.append(eng.decoder_forward(bb, feats, qi, want_logits=True)[1].clone())
    assert torch.equal(outs[0], outs[1])


def test_query_without_references_matches_oracle(weights):
    """Edge case T = 1 (a query view and no reference view): the decoder sequence is the query's P tokens alone.  Exact path
    against the oracle at the 1e-4 gate, bf16 path at its statistical gate; corners bit-exact on the exact path."""
    from oracle import boxdreamer_oracle as O
    dec, dino = weights
    B, T = 2, 1
    data = synth.synth_inputs(B, T, 224, seed=93)
    with torch.no_grad():
        ref = O.forward(data, dec, dino, with_pnp=False)
    for precision, dtype, tol in (("exact", torch.float32, 1e-4), ("bf16", torch.bfloat16, 1.5e-2)):
        m = _model(weights, precision)
        d = _to_cuda({k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()})
        eng = m._engine_for(d["images"], B, T)
        feats = eng.dino_forward(d["images"].view(B * T, 3, 224, 224).contiguous())
        heat, logits = eng.decoder_forward(d["bbox_feat"].contiguous(), feats, d["query_idx"], want_logits=True)
        assert _scaled(logits.view(B, 256, 1568), ref["logits"]) <= tol, precision
        if precision == "exact":
            px, nm = eng.corners_topk(heat)
            assert torch.allclose(px.cpu(), ref["keypoints_px"], atol=1e-6, rtol=0)


def test_new_entries_validate_their_arguments(weights):
    """bd_forward_packed / bd_forward_host_submit / _wait fail loudly (status + bd_last_error) instead of guessing: robust PnP mode
    with a packed record, staging slots other than 0 / 1, null result pointers; waiting on an idle slot is a no-op."""
    import ctypes as C
    m = _model(weights, "bf16")
    B, T = 1, 2
    data = synth.synth_inputs(B, T, 224, seed=95, dtype=torch.bfloat16)
    mask = torch.zeros(B, T, dtype=torch.bool)
    mask[torch.arange(B), data["query_idx"]] = True
    img, bb, qi = data["images"].cuda().contiguous(), data["bbox_feat"].cuda().contiguous(), data["query_idx"].cuda()
    X = data["bbox_3d"][mask].float().cuda().contiguous()
    K = data["non_ndc_intrinsics"][mask].float().cuda().contiguous()
    eng = m._engine_for(img, B, T)
    with pytest.raises(_lib.BoxDreamerLibError, match="mode 0"):
        eng.forward_packed(img, bb, qi, X, K, opts=_lib.BdPnpOpts(1, 64, 2.0, 0, 30))
    host = (img.cpu().pin_memory(), bb.cpu().pin_memory(), qi.cpu(), X.cpu(), K.cpu())
    with pytest.raises(_lib.BoxDreamerLibError, match="slot"):
        eng.forward_host_submit(2, *host)
    assert eng.lib.bd_forward_host_wait(eng.handle, 1) == 0          # nothing submitted on slot 1: returns at once
    assert eng.lib.bd_forward_host_wait(eng.handle, 5) != 0
    rc = eng.lib.bd_forward_packed(eng.handle, _lib.ptr(img), _lib.ptr(bb), 1, _lib.ptr(qi), _lib.ptr(X), _lib.ptr(K), None, None, B, T, None)
    assert rc != 0 and b"null" in eng.lib.bd_last_error()
    rec = eng.forward_packed(img, bb, qi, X, K)                       # and the valid call still works afterwards
    assert rec.shape == (B, 28) and torch.isfinite(rec).all()
