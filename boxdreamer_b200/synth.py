"""Deterministic synthetic weights and inputs for BoxDreamer's hot path.

There is no network on the build or GPU boxes, so neither the HF checkpoint
(`yyh929/BoxDreamer`, run.py:172-183) nor the torch.hub DINOv2 weights (dinov2.py:35) exist.
Everything here is derived from integer PCG64 streams (bit-identical on every host), keyed by
the state-dict key name, so that the reference module (build container), the oracle and the
CUDA engine (GPU box) all see the same parameters without shipping 700 MB of tensors.

State-dict layouts follow the reference exactly:
  decoder: 177 tensors, `BoxDreamer.state_dict()` (BoxDreamerModel.py:110; betr.py:139-176; blocks.py:808-868)
  DINOv2 ViT-B/14+4reg: 176 tensors (src/models/sources/DINOv2/vision_transformer.py:44-170)
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch

D = 768
PATCH = 14


def _uniform(key: str, shape, std: float, seed: int, mean: float = 0.0) -> torch.Tensor:
    """Uniform with the requested std, from 24-bit integers (exact in fp32, libm-free)."""
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))
    n = int(np.prod(shape)) if len(shape) else 1
    ints = rng.integers(0, 1 << 24, size=n, dtype=np.int64)
    u = (ints.astype(np.float64) / float(1 << 23)) - 1.0  # [-1, 1)
    vals = (u * (std * math.sqrt(3.0)) + mean).astype(np.float32)
    return torch.from_numpy(vals.reshape(shape))


def decoder_param_shapes(num_layers: int = 12, d: int = D, heads: int = 8, patch: int = PATCH, box_dim: int = 8):
    """Key -> shape, in the reference's registration order (SURVEY.md section 8a 'State-dict layout')."""
    hd = d // heads
    pp = patch * patch * box_dim
    shapes = {"decoder.bbox_learnable_query": (1, d)}
    for i in range(num_layers):
        p = f"decoder.attn.{i}."
        shapes[p + "norm1.weight"] = (d,)
        shapes[p + "norm1.bias"] = (d,)
        shapes[p + "attn.qkv.weight"] = (3 * d, d)
        shapes[p + "attn.qkv.bias"] = (3 * d,)
        shapes[p + "attn.q_norm.weight"] = (hd,)
        shapes[p + "attn.k_norm.weight"] = (hd,)
        shapes[p + "attn.proj.weight"] = (d, d)
        shapes[p + "attn.proj.bias"] = (d,)
        shapes[p + "norm2.weight"] = (d,)
        shapes[p + "norm2.bias"] = (d,)
        shapes[p + "mlp.fc1.weight"] = (4 * d, d)
        shapes[p + "mlp.fc1.bias"] = (4 * d,)
        shapes[p + "mlp.fc2.weight"] = (d, 4 * d)
        shapes[p + "mlp.fc2.bias"] = (d,)
    shapes["decoder.bbox_proj.weight"] = (pp, d)
    shapes["decoder.bbox_proj.bias"] = (pp,)
    shapes["decoder.input_transform.fc1.weight"] = (d, d)
    shapes["decoder.input_transform.fc1.bias"] = (d,)
    shapes["decoder.input_transform.fc2.weight"] = (d, d)
    shapes["decoder.input_transform.fc2.bias"] = (d,)
    shapes["decoder.bbox_emb.weight"] = (d, pp)
    shapes["decoder.bbox_emb.bias"] = (d,)
    return shapes


def dino_param_shapes(depth: int = 12, d: int = D, patch: int = PATCH, pretrain_grid: int = 37, n_reg: int = 4):
    shapes = {
        "cls_token": (1, 1, d),
        "pos_embed": (1, pretrain_grid * pretrain_grid + 1, d),
        "register_tokens": (1, n_reg, d),
        "mask_token": (1, d),
        "patch_embed.proj.weight": (d, 3, patch, patch),
        "patch_embed.proj.bias": (d,),
    }
    for i in range(depth):
        p = f"blocks.{i}."
        shapes[p + "norm1.weight"] = (d,)
        shapes[p + "norm1.bias"] = (d,)
        shapes[p + "attn.qkv.weight"] = (3 * d, d)
        shapes[p + "attn.qkv.bias"] = (3 * d,)
        shapes[p + "attn.proj.weight"] = (d, d)
        shapes[p + "attn.proj.bias"] = (d,)
        shapes[p + "ls1.gamma"] = (d,)
        shapes[p + "norm2.weight"] = (d,)
        shapes[p + "norm2.bias"] = (d,)
        shapes[p + "mlp.fc1.weight"] = (4 * d, d)
        shapes[p + "mlp.fc1.bias"] = (4 * d,)
        shapes[p + "mlp.fc2.weight"] = (d, 4 * d)
        shapes[p + "mlp.fc2.bias"] = (d,)
        shapes[p + "ls2.gamma"] = (d,)
    shapes["norm.weight"] = (d,)
    shapes["norm.bias"] = (d,)
    return shapes


def _init_one(key: str, shape, seed: int) -> torch.Tensor:
    leaf = key.split(".")[-1]
    if key.endswith("norm.weight") or ".norm1.weight" in key or ".norm2.weight" in key \
            or "q_norm.weight" in key or "k_norm.weight" in key or leaf == "gamma":
        return _uniform(key, shape, 0.1, seed, mean=1.0)
    if leaf == "bias":
        return _uniform(key, shape, 0.05, seed)
    if key.endswith("qkv.weight"):
        return _uniform(key, shape, 0.04, seed)
    if leaf == "weight":
        return _uniform(key, shape, 0.02, seed)
    # tokens / tables: bbox_learnable_query, cls_token, pos_embed, register_tokens, mask_token
    return _uniform(key, shape, 0.2, seed)


def synth_decoder_state_dict(seed: int = 0, num_layers: int = 12) -> dict:
    return {k: _init_one(k, s, seed) for k, s in decoder_param_shapes(num_layers).items()}


def synth_dino_state_dict(seed: int = 0, depth: int = 12) -> dict:
    return {k: _init_one("dino." + k, s, seed) for k, s in dino_param_shapes(depth).items()}


# ----------------------------------------------------------------------------------------------
# inputs


def box_corners(ext: np.ndarray) -> np.ndarray:
    """8 corners of an axis-aligned box centred at 0, order of vis_utils.py:1156-1165."""
    hx, hy, hz = (ext / 2.0).tolist()
    mn = (-hx, -hy, -hz)
    mx = (hx, hy, hz)
    return np.array([
        [mn[0], mn[1], mn[2]], [mn[0], mx[1], mn[2]], [mx[0], mx[1], mn[2]], [mx[0], mn[1], mn[2]],
        [mn[0], mn[1], mx[2]], [mn[0], mx[1], mx[2]], [mx[0], mx[1], mx[2]], [mx[0], mn[1], mx[2]],
    ], dtype=np.float64)


def random_rotation(rng: np.random.Generator) -> np.ndarray:
    """Uniform-ish rotation from a unit quaternion of integer-derived components."""
    while True:
        q = rng.integers(-(1 << 20), 1 << 20, size=4).astype(np.float64)
        n = math.sqrt(float((q * q).sum()))
        if n > 1e3:
            break
    w, x, y, z = (q / n).tolist()
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def _urand(rng, lo, hi, size=None):
    ints = rng.integers(0, 1 << 24, size=size)
    return lo + (hi - lo) * (np.asarray(ints, dtype=np.float64) / float(1 << 24))


def project(K: np.ndarray, R: np.ndarray, t: np.ndarray, X: np.ndarray) -> np.ndarray:
    """Pinhole projection, same convention as camera_utils.py:9-59 (x = K [R|t] X)."""
    Xc = X @ R.T + t[None, :]
    uv = Xc @ K.T
    return uv[:, :2] / uv[:, 2:3]


def make_heatmaps(corners_px: torch.Tensor, S: int, group: int = 1) -> torch.Tensor:
    """GT 8-corner heatmaps in [-1,1]; follows datasets/utils/base/bbox_utils.py:263-303.

    corners_px [L,8,2] (pixel coords of the crop) -> [L,8,S,S] fp32.
    `group`: number of consecutive views normalised together.  The reference divides corner i's maps by their maximum
    over the WHOLE call (`bbox_map[..., i].max()`, bbox_utils.py:296), and the dataset calls it once per sample, i.e.
    over that sample's T views: group=T is the dataset's semantics (what `bd_make_bbox_features` implements and what
    tests/test_oracle_vs_reference.py pins).  group=1 (per-view maximum, the default) is what the synthetic inputs of
    the committed golden fixtures were generated with; the two differ by ~1e-3 in the map values.
    """
    L = corners_px.shape[0]
    c = corners_px.to(torch.float32)
    ix = torch.arange(S, dtype=torch.float32).view(1, 1, 1, S)
    iy = torch.arange(S, dtype=torch.float32).view(1, 1, S, 1)
    dx = c[:, :, 0].view(L, 8, 1, 1) - ix
    dy = c[:, :, 1].view(L, 8, 1, 1) - iy
    dist = torch.sqrt(dx ** 2 + dy ** 2)
    center = c.mean(dim=1)
    dis = torch.sqrt((center[:, None, 0] - c[:, :, 0]) ** 2 + (center[:, None, 1] - c[:, :, 1]) ** 2)
    scale = (dis / 10) ** 2
    m = torch.exp(-dist / scale.view(L, 8, 1, 1))
    if group == 1:
        m = m / m.amax(dim=(2, 3), keepdim=True)
    else:
        assert L % group == 0
        mg = m.view(L // group, group, 8, S, S)
        m = (mg / mg.amax(dim=(1, 3, 4), keepdim=True)).view(L, 8, S, S)
    return m * 2 - 1


def synth_inputs(B: int, T: int, S: int = 224, seed: int = 1234, dtype=torch.float32,
                 with_images: bool = True) -> dict:
    """The input dict `BoxDreamer.forward` reads (SURVEY.md section 8a row 0, 8d 'Synthetic inputs')."""
    rng = np.random.Generator(np.random.PCG64([seed, B, T, S]))
    poses = np.zeros((B, T, 4, 4), dtype=np.float64)
    Ks = np.zeros((B, T, 3, 3), dtype=np.float64)
    bbox3d = np.zeros((B, T, 8, 3), dtype=np.float64)
    proj = np.zeros((B, T, 8, 2), dtype=np.float64)
    for b in range(B):
        ext = _urand(rng, 0.05, 0.25, size=3)
        X = box_corners(ext)
        f = float(_urand(rng, 250.0, 400.0)) * S / 224.0
        K = np.array([[f, 0, S / 2.0], [0, f, S / 2.0], [0, 0, 1.0]])
        for t in range(T):
            R = random_rotation(rng)
            tv = np.array([0.0, 0.0, float(_urand(rng, 0.4, 0.8))]) + _urand(rng, -0.017, 0.017, size=3)
            poses[b, t, :3, :3] = R
            poses[b, t, :3, 3] = tv
            poses[b, t, 3, 3] = 1.0
            Ks[b, t] = K
            bbox3d[b, t] = X
            proj[b, t] = project(K, R, tv, X)
    proj_t = torch.from_numpy(proj).to(torch.float32)
    bbox_feat = make_heatmaps(proj_t.view(B * T, 8, 2), S).view(B, T, 8, S, S)
    norm_proj = torch.clamp(proj_t / float(S) * 2 - 1, min=-5, max=5)
    data = {
        "bbox_feat": bbox_feat.to(dtype),
        "query_idx": torch.full((B,), T - 1, dtype=torch.int64),
        "poses": torch.from_numpy(poses).to(dtype),
        "non_ndc_intrinsics": torch.from_numpy(Ks).to(dtype),
        "intrinsics": torch.from_numpy(Ks).to(dtype),
        "bbox_3d": torch.from_numpy(bbox3d).to(dtype),
        "bbox_proj_crop": norm_proj.to(dtype),
        "crop_parameters": torch.zeros(B, T, 4, dtype=dtype),
        "image_masks": torch.ones(B, T, 1, S, S, dtype=dtype),
    }
    if with_images:
        n = B * T * 3 * S * S
        ints = rng.integers(0, 256, size=n, dtype=np.int64).astype(np.float32) / 256.0
        data["images"] = torch.from_numpy(ints.reshape(B, T, 3, S, S)).to(dtype)
    return data


def synth_pnp_cases(n: int, sigma: float, S: int = 224, seed: int = 1239, quantise: bool = True):
    """BASELINE config 5: box corners projected by a random pose + Gaussian pixel noise.

    Returns float32 arrays (corners_px [n,8,2], bbox3d [n,8,3], K [n,3,3]) and the fp64 GT pose [n,3,4].
    Gaussian noise is Box-Muller on integer-derived uniforms (host libm; used for accuracy
    statistics, not for bit-exact comparison).
    """
    rng = np.random.Generator(np.random.PCG64([seed, n, int(sigma * 1000)]))
    c2 = np.zeros((n, 8, 2), dtype=np.float64)
    X3 = np.zeros((n, 8, 3), dtype=np.float64)
    Ks = np.zeros((n, 3, 3), dtype=np.float64)
    gt = np.zeros((n, 3, 4), dtype=np.float64)
    for i in range(n):
        ext = _urand(rng, 0.05, 0.25, size=3)
        X = box_corners(ext)
        f = float(_urand(rng, 250.0, 400.0)) * S / 224.0
        K = np.array([[f, 0, S / 2.0], [0, f, S / 2.0], [0, 0, 1.0]])
        R = random_rotation(rng)
        tv = np.array([0.0, 0.0, float(_urand(rng, 0.4, 0.8))]) + _urand(rng, -0.017, 0.017, size=3)
        uv = project(K, R, tv, X)
        if sigma > 0:
            u1 = (rng.integers(1, 1 << 24, size=16).astype(np.float64)) / float(1 << 24)
            u2 = (rng.integers(0, 1 << 24, size=16).astype(np.float64)) / float(1 << 24)
            g = np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * np.pi * u2)
            uv = uv + sigma * g.reshape(8, 2)
        if quantise:
            uv = np.round(uv * 20.0) / 20.0  # top-20 means are multiples of 0.05 px
        c2[i], X3[i], Ks[i] = uv, X, K
        gt[i, :, :3], gt[i, :, 3] = R, tv
    return c2.astype(np.float32), X3.astype(np.float32), Ks.astype(np.float32), gt
