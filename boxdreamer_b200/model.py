"""Drop-in for the reference's `src/models` module API on the inference hot path.

    from boxdreamer_b200 import BoxDreamer          # instead of src.models.BoxDreamerModel.BoxDreamer
    model = BoxDreamer(config).cuda().eval()         # same config["modules"] tree (transformer.yaml:10-71)
    model.load_state_dict(ckpt)                      # same 177 decoder keys ("decoder.*")
    data = model(data)                               # same in-place-mutating forward(dict) -> dict

Mirrors (reference file:line)
  BoxDreamer          src/models/BoxDreamerModel.py:21-384
  BETR                src/models/modules/backbone/betr.py:11-437   (forward seam :249-308)
  DinoV2Wrapper       src/models/modules/encoder/dinov2.py:6-60    (plain object, weights outside the state_dict)
  process_prediction  src/models/utils/prediction_utils.py:63-103
All arithmetic runs in the CUDA engine behind the C ABI (include/boxdreamer_b200.h); this file only
holds parameters, checks shapes and scatters results into the caller's dict.  Inference only: no autograd.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

__all__ = ["BoxDreamer", "BETR", "DinoV2Wrapper", "Engine", "validate_model_config", "setup_camera_params"]


# ----------------------------------------------------------------------------------------------
# config helpers (src/models/utils/config_utils.py:10-96)


def validate_model_config(config):
    config = config.copy()
    assert config["pose_representation"] in ["plucker", "vector", "bb8"]
    assert config["bbox_representation"] in ["heatmap", "voting", "cornernet"]
    if config["bbox_representation"] in ["cornernet"]:
        config["bbox_representation"] = "heatmap"
    assert config["coordinate"] in ["first_camera", "object"]
    if config["use_rgb"] and config["encoder"]["name"] == "dino":
        assert config["decoder"]["patch_size"] == 14, "Dinov2 only supports patch size 14"
    assert (config["patchify_rays"] and config["use_rgb"]) or (
        not config["patchify_rays"] and not config["use_rgb"]
    ), "patchify_rays should be True when use_rgb is True"
    return config


def setup_camera_params(config):
    assert config["rotation_type"] is None, "boxdreamer_b200 builds the bb8 path only (rotation_type: null)"
    assert config["pose_representation"] == "bb8"
    dec = config["decoder"]
    dec["rotation_type"] = None
    dec["camera_dim"] = 0
    dec["rotation_length"] = 0
    dec["use_pretrained"] = config["use_rgb"]
    dec["patchify_rays"] = config["patchify_rays"]
    dec["pose_representation"] = config["pose_representation"]
    dec["bbox_representation"] = config["bbox_representation"]
    if config["use_rgb"] and dec["diff_emb"]:
        dec["diff_emb"] = False
    return config, 0, 0


# ----------------------------------------------------------------------------------------------
# engine wrapper


def _on_device(fn):
    """Runs an Engine method with the engine's device current (kernels, `current_stream()` and allocations then all refer to
    the device the workspace lives on, whatever device the caller had selected)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        if torch.cuda.current_device() == self.device:
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapper


class Engine:
    """One bd_handle: workspace for (max_batch, max_views) at one precision on the current device."""

    def __init__(self, img_size, patch, d_model, dec_layers, dec_heads, precision, max_batch, max_views,
                 attn_variant=None, dino_layers=12, dino_heads=12, dino_registers=4):
        if not torch.cuda.is_available():
            raise _lib.BoxDreamerLibError("boxdreamer_b200 needs a CUDA device: the hot path has no CPU fallback")
        self.lib = _lib.load()
        if attn_variant is None:
            attn_variant = int(os.environ.get("BOXDREAMER_B200_ATTN_VARIANT", "2"))
        self.cfg = _lib.BdConfig(img_size, patch, d_model, dec_layers, dec_heads, dino_layers, dino_heads,
                                 dino_registers, 37, precision, attn_variant, max_batch, max_views)
        self.handle = C.c_void_p()
        self.device = torch.cuda.current_device()
        _lib.check(self.lib.bd_create(C.byref(self.handle), C.byref(self.cfg)), "bd_create")
        self.precision = precision
        self.max_batch, self.max_views = max_batch, max_views
        self.S, self.patch, self.d = img_size, patch, d_model
        self.P = (img_size // patch) ** 2
        self.weights_version = -1

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.bd_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @_on_device
    def load_weights(self, named_tensors: dict):
        for name, t in named_tensors.items():
            t = t.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.bd_load_weight(self.handle, name.encode(), _lib.ptr(t), shape, t.dim()),
                       f"bd_load_weight({name})")
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        _lib.check(self.lib.bd_finalize_weights(self.handle), "bd_finalize_weights")

    # thin typed wrappers -----------------------------------------------------------------
    @staticmethod
    def _dt(t):
        if t.dtype == torch.float32:
            return _lib.BD_F32
        if t.dtype == torch.bfloat16:
            return _lib.BD_BF16
        raise TypeError(f"unsupported dtype {t.dtype}")

    @_on_device
    def dino_forward(self, images):
        L = images.shape[0]
        feats = torch.empty(L, self.P, self.d, device=images.device, dtype=torch.float32)
        _lib.check(self.lib.bd_dino_forward(self.handle, _lib.ptr(images), self._dt(images), _lib.ptr(feats), L,
                                            _lib.stream_ptr()), "bd_dino_forward")
        return feats

    @_on_device
    def decoder_forward(self, bbox_feat, feats, query_idx, want_logits=False):
        B, T = bbox_feat.shape[:2]
        heat = torch.empty(B, 8, self.S, self.S, device=bbox_feat.device, dtype=torch.float32)
        logits = torch.empty(B * self.P, self.patch * self.patch * 8, device=bbox_feat.device,
                             dtype=torch.float32) if want_logits else None
        _lib.check(self.lib.bd_decoder_forward(self.handle, _lib.ptr(bbox_feat), self._dt(bbox_feat), _lib.ptr(feats),
                                               _lib.ptr(query_idx), _lib.ptr(heat), _lib.ptr(logits), B, T,
                                               _lib.stream_ptr()), "bd_decoder_forward")
        return (heat, logits) if want_logits else heat

    @_on_device
    def corners_topk(self, heat, want_idx=False):
        B, _, S, _ = heat.shape
        px = torch.empty(B, 8, 2, device=heat.device, dtype=torch.float32)
        nm = torch.empty(B, 8, 2, device=heat.device, dtype=torch.float32)
        idx = torch.empty(B, 8, 20, device=heat.device, dtype=torch.int32) if want_idx else None
        _lib.check(self.lib.bd_corners_topk(self.handle, _lib.ptr(heat), _lib.ptr(px), _lib.ptr(nm), _lib.ptr(idx), B, S,
                                            _lib.stream_ptr()), "bd_corners_topk")
        return (px, nm, idx) if want_idx else (px, nm)

    @_on_device
    def pnp(self, corners_px, bbox3d, K, opts=None):
        B, n = corners_px.shape[:2]
        poses = torch.empty(B, 4, 4, device=corners_px.device, dtype=torch.float32)
        o = C.byref(opts) if opts is not None else None
        _lib.check(self.lib.bd_pnp(self.handle, _lib.ptr(corners_px), _lib.ptr(bbox3d), _lib.ptr(K), _lib.ptr(poses), o, B, n,
                                   _lib.stream_ptr()), "bd_pnp")
        return poses

    @_on_device
    def forward(self, images, bbox_feat, query_idx, bbox3d_q, K_q, want_heat=True, opts=None):
        B, T = images.shape[:2]
        dev = images.device
        heat = torch.empty(B, 8, self.S, self.S, device=dev, dtype=torch.float32) if want_heat else None
        px = torch.empty(B, 8, 2, device=dev, dtype=torch.float32)
        nm = torch.empty(B, 8, 2, device=dev, dtype=torch.float32)
        poses = torch.empty(B, 4, 4, device=dev, dtype=torch.float32)
        o = C.byref(opts) if opts is not None else None
        _lib.check(self.lib.bd_forward(self.handle, _lib.ptr(images), _lib.ptr(bbox_feat), self._dt(images),
                                       _lib.ptr(query_idx), _lib.ptr(bbox3d_q), _lib.ptr(K_q), _lib.ptr(heat), _lib.ptr(px),
                                       _lib.ptr(nm), _lib.ptr(poses), o, B, T, _lib.stream_ptr()), "bd_forward")
        return heat, px, nm, poses

    @_on_device
    def forward_packed(self, images, bbox_feat, query_idx, bbox3d_q, K_q, out=None, opts=None):
        """The forward, returning the packed [B, 28] result record (dist.RECORD: R|t, normalised corners) written by the PnP
        kernel itself -- the payload of the multi-GPU result gather (dist.gather_records)."""
        B, T = images.shape[:2]
        rec = out if out is not None else torch.empty(B, 28, device=images.device, dtype=torch.float32)
        o = C.byref(opts) if opts is not None else None
        _lib.check(self.lib.bd_forward_packed(self.handle, _lib.ptr(images), _lib.ptr(bbox_feat), self._dt(images),
                                              _lib.ptr(query_idx), _lib.ptr(bbox3d_q), _lib.ptr(K_q), _lib.ptr(rec), o, B, T,
                                              _lib.stream_ptr()), "bd_forward_packed")
        return rec

    def _pinned_results(self, tag, B, want_heat):
        """Pinned result buffers (heat | None, corners_px, corners_norm, poses), allocated once per (tag, B) and reused: pinning /
        unpinning host memory inside the loop synchronises the device and would undo the pipelining of the host entries."""
        cache = self.__dict__.setdefault("_pinned_cache", {})
        key = (tag, B)
        if key not in cache:
            cache[key] = [None, torch.empty(B, 8, 2, dtype=torch.float32).pin_memory(),
                          torch.empty(B, 8, 2, dtype=torch.float32).pin_memory(), torch.empty(B, 4, 4, dtype=torch.float32).pin_memory()]
        bufs = cache[key]
        if want_heat and bufs[0] is None:
            bufs[0] = torch.empty(B, 8, self.S, self.S, dtype=torch.float32).pin_memory()
        return (bufs[0] if want_heat else None), bufs[1], bufs[2], bufs[3]

    @_on_device
    def forward_host(self, images, bbox_feat, query_idx, bbox3d_q, K_q, want_heat=False, opts=None):
        """Host tensors in, host tensors out (H2D/D2H inside the call)."""
        B, T = images.shape[:2]
        heat, px, nm, poses = self._pinned_results("sync", B, want_heat)
        o = C.byref(opts) if opts is not None else None
        _lib.check(self.lib.bd_forward_host(self.handle, _lib.ptr(images), _lib.ptr(bbox_feat), self._dt(images),
                                            _lib.ptr(query_idx), _lib.ptr(bbox3d_q), _lib.ptr(K_q), _lib.ptr(heat),
                                            _lib.ptr(px), _lib.ptr(nm), _lib.ptr(poses), o, B, T), "bd_forward_host")
        return (heat.clone() if heat is not None else None), px.clone(), nm.clone(), poses.clone()

    @_on_device
    def forward_host_submit(self, slot, images, bbox_feat, query_idx, bbox3d_q, K_q, bbox_px=None, want_heat=False, opts=None):
        """Pipelined host entry: enqueue one batch on staging slot 0 / 1 and return the (pinned) result tensors, which are
        filled once `forward_host_wait(slot)` returns and stay valid until the same slot is submitted again (they are reused).
        Pass `bbox_px` [B,T,8,2] instead of `bbox_feat` (None) to have the reference heat maps rasterised on the device.  The
        input tensors must stay alive until the wait."""
        B, T = images.shape[:2]
        heat, px, nm, poses = self._pinned_results(("slot", slot), B, want_heat)
        o = C.byref(opts) if opts is not None else None
        _lib.check(self.lib.bd_forward_host_submit(self.handle, slot, _lib.ptr(images), _lib.ptr(bbox_feat), _lib.ptr(bbox_px),
                                                   self._dt(images), _lib.ptr(query_idx), _lib.ptr(bbox3d_q), _lib.ptr(K_q),
                                                   _lib.ptr(heat), _lib.ptr(px), _lib.ptr(nm), _lib.ptr(poses), o, B, T),
                   "bd_forward_host_submit")
        self._host_keepalive = getattr(self, "_host_keepalive", {})
        self._host_keepalive[slot] = (images, bbox_feat, bbox_px, query_idx, bbox3d_q, K_q)
        return heat, px, nm, poses

    @_on_device
    def forward_host_wait(self, slot):
        _lib.check(self.lib.bd_forward_host_wait(self.handle, slot), "bd_forward_host_wait")
        getattr(self, "_host_keepalive", {}).pop(slot, None)

    @_on_device
    def forward_host_px(self, images, bbox_px, query_idx, bbox3d_q, K_q, want_heat=False, opts=None):
        """forward_host with the reference heat maps rasterised on the device from bbox_px [B,T,8,2] (host, fp32)."""
        B, T = images.shape[:2]
        heat, px, nm, poses = self._pinned_results("sync", B, want_heat)
        o = C.byref(opts) if opts is not None else None
        _lib.check(self.lib.bd_forward_host_px(self.handle, _lib.ptr(images), _lib.ptr(bbox_px), self._dt(images),
                                               _lib.ptr(query_idx), _lib.ptr(bbox3d_q), _lib.ptr(K_q), _lib.ptr(heat),
                                               _lib.ptr(px), _lib.ptr(nm), _lib.ptr(poses), o, B, T), "bd_forward_host_px")
        return (heat.clone() if heat is not None else None), px.clone(), nm.clone(), poses.clone()


# ----------------------------------------------------------------------------------------------
# parameter containers with the reference's state_dict layout


class _Attn(nn.Module):
    def __init__(self, d, hd):
        super().__init__()
        self.qkv = nn.Linear(d, 3 * d)
        self.q_norm = _Scale(hd)
        self.k_norm = _Scale(hd)
        self.proj = nn.Linear(d, d)


class _Scale(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))


class _Mlp(nn.Module):
    def __init__(self, d_in, d_hidden, d_out):
        super().__init__()
        self.fc1 = nn.Linear(d_in, d_hidden)
        self.fc2 = nn.Linear(d_hidden, d_out)


class _Block(nn.Module):
    """Holds the tensors of SelfAttentionBlock (blocks.py:808-868)."""

    def __init__(self, d, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(d, 1e-5)
        self.attn = _Attn(d, d // heads)
        self.norm2 = nn.LayerNorm(d, 1e-5)
        self.mlp = _Mlp(d, 4 * d, d)


class BETR(nn.Module):
    """Parameter layout + forward seam of the reference BETR (betr.py:11-437), bb8/heatmap/use_pretrained."""

    def __init__(self, d_model=512, nhead=8, num_decoder_layers=6, **kwargs):
        super().__init__()
        self.d_model, self.nhead, self.att_depth = d_model, nhead, num_decoder_layers
        self.patch_size = kwargs["patch_size"]
        self.img_size = kwargs["img_size"]
        self.pose_representation = kwargs.get("pose_representation", "bb8")
        self.bbox_representation = kwargs.get("bbox_representation", "voting")
        self.use_pretrained = kwargs["use_pretrained"]
        if not (self.pose_representation == "bb8" and self.bbox_representation == "heatmap" and self.use_pretrained):
            raise NotImplementedError("boxdreamer_b200 builds pose_representation=bb8, bbox_representation=heatmap, use_rgb=True")
        assert kwargs.get("nvs_supervision", False) or kwargs.get("ray_supervision", False), \
            "At least one supervision should be True"
        self.box_dim = 8
        pp = self.patch_size ** 2 * self.box_dim
        self.attn = nn.Sequential(*[_Block(d_model, nhead) for _ in range(num_decoder_layers)])
        self.bbox_proj = nn.Linear(d_model, pp)
        self.input_transform = _Mlp(d_model, d_model, d_model)
        self.bbox_learnable_query = nn.Parameter(torch.zeros(1, d_model))
        self.bbox_emb = nn.Linear(pp, d_model)
        self._owner = None  # set by BoxDreamer: provides the engine

    def forward(self, pose_feat, rgbs=None, masks=None, pretrain_rgb_feat=None, image_masks=None):
        assert rgbs is not None, "rgbs input should not be None"
        B, N, C_, H, W = rgbs.shape
        assert H == W == self.img_size, f"H and W should be equal to img_size {self.img_size}, got {H}x{W}"
        if self._owner is None:
            raise RuntimeError("BETR must be owned by a boxdreamer_b200.BoxDreamer (it provides the CUDA engine)")
        if pretrain_rgb_feat is None:
            raise NotImplementedError("use_rgb=False path is not built")
        if not bool((masks.sum(dim=1) == 1).all()):
            raise ValueError("exactly one query view per sample is required (betr.py:288-290)")
        query_idx = masks.to(torch.int64).argmax(dim=1).contiguous()
        eng = self._owner._engine_for(pose_feat, B, N)
        bbox = self._owner._as_engine_input(pose_feat)
        feats = pretrain_rgb_feat.to(torch.float32).contiguous()
        return eng.decoder_forward(bbox, feats, query_idx)


class _DinoBlock(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.norm1 = nn.LayerNorm(d, 1e-6)
        self.attn = nn.Module()
        self.attn.qkv = nn.Linear(d, 3 * d)
        self.attn.proj = nn.Linear(d, d)
        self.ls1 = nn.Module()
        self.ls1.gamma = nn.Parameter(torch.ones(d))
        self.norm2 = nn.LayerNorm(d, 1e-6)
        self.mlp = _Mlp(d, 4 * d, d)
        self.ls2 = nn.Module()
        self.ls2.gamma = nn.Parameter(torch.ones(d))


class DinoParams(nn.Module):
    """state_dict layout of dinov2_vitb14_reg (DINOv2 vision_transformer.py:44-170)."""

    def __init__(self, d=768, depth=12, patch=14, pretrain_grid=37, n_reg=4):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, d))
        self.pos_embed = nn.Parameter(torch.zeros(1, pretrain_grid * pretrain_grid + 1, d))
        self.register_tokens = nn.Parameter(torch.zeros(1, n_reg, d))
        self.mask_token = nn.Parameter(torch.zeros(1, d))
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(3, d, kernel_size=patch, stride=patch)
        self.blocks = nn.ModuleList([_DinoBlock(d) for _ in range(depth)])
        self.norm = nn.LayerNorm(d, 1e-6)
        self.patch_size = patch

    def interpolated_pos_embed(self, S: int) -> torch.Tensor:
        """vision_transformer.py:179-211 (interpolate_offset=0.0, antialias=True)."""
        pe = self.pos_embed.detach().float()
        N = pe.shape[1] - 1
        g = S // self.patch_size
        if g * g == N:
            return pe
        M = int(N ** 0.5)
        d = pe.shape[-1]
        patch_pe = F.interpolate(pe[:, 1:].reshape(1, M, M, d).permute(0, 3, 1, 2), mode="bicubic", antialias=True,
                                 size=(g, g))
        patch_pe = patch_pe.permute(0, 2, 3, 1).reshape(1, -1, d)
        return torch.cat((pe[:, :1], patch_pe), dim=1)


class DinoV2Wrapper:
    """encoder/dinov2.py:6-60: plain object (its weights are NOT in BoxDreamer.state_dict()).

    There is no network here, so weights are not fetched from torch.hub: load them with
    `wrapper.model.load_state_dict(torch.load(<dinov2_vitb14_reg4_pretrain.pth>))`.
    """

    def __init__(self, ckpt_path=None, cfg=None):
        cfg = cfg or {}
        self.model_type = cfg.get("model_type", "dinov2_vits14_reg")
        assert self.model_type == "dinov2_vitb14_reg", "boxdreamer_b200 builds dinov2_vitb14_reg only"
        self.freeze = cfg.get("freeze", True)
        self.device = None
        self.model = DinoParams()
        self.model.eval()
        for p in self.model.parameters():
            p.requires_grad = False
        if ckpt_path is not None and os.path.isfile(str(ckpt_path)):
            self.model.load_state_dict(torch.load(ckpt_path, map_location="cpu", weights_only=True))
        self._owner = None

    def get_device(self):
        return self.device

    def to_device(self, device):
        self.model = self.model.to(device)
        self.device = device

    def predict(self, input_tensor):
        flag = input_tensor.dim() == 5
        if flag:
            B, T = input_tensor.shape[:2]
            input_tensor = input_tensor.flatten(0, 1)
        if self._owner is None:
            raise RuntimeError("DinoV2Wrapper must be owned by a boxdreamer_b200.BoxDreamer")
        L = input_tensor.shape[0]
        eng = self._owner._engine_for(input_tensor, 1, 1, images=L)
        ret = eng.dino_forward(self._owner._as_engine_input(input_tensor))
        if flag:
            ret = ret.view(B, T, *ret.shape[1:])
        return ret


# ----------------------------------------------------------------------------------------------


class BoxDreamer(nn.Module):
    """B200-native BoxDreamer (BoxDreamerModel.py:21): same config, state_dict and forward contract."""

    def __init__(self, config, precision: str | None = None):
        super().__init__()
        self.config = config
        mc = config["modules"]
        self.use_matching = mc["use_matching"]
        self.use_tracking = mc["use_tracking"]
        self.use_rgb = mc["use_rgb"]
        self.roatation_type = mc["rotation_type"]
        self.coordinate = mc["coordinate"]
        self.pose_representation = mc["pose_representation"]
        self.image_size = mc["decoder"]["img_size"]
        self.patch_size = mc["decoder"]["patch_size"]
        mc = validate_model_config(mc)
        self.bbox_representation = mc["bbox_representation"]
        self.dense_cfg = mc.get("dense_cfg", None)
        mc, self.camera_dim, self.rotation_length = setup_camera_params(mc)
        self.module_configs = mc
        if self.use_tracking:
            raise NotImplementedError("Tracking is not supported yet")
        if self.use_matching:
            raise NotImplementedError("use_matching (LoFTR) is outside the hot path built here")
        if not (self.use_rgb and mc["encoder"]["name"] == "dino"):
            raise NotImplementedError("boxdreamer_b200 builds the DINOv2 encoder path only")
        self.rgb_encoder = DinoV2Wrapper(**mc["encoder"]["dino"])
        self.rgb_encoder._owner = self
        self.decoder = BETR(**mc["decoder"])
        object.__setattr__(self.decoder, "_owner", self)
        # "exact" (fp32 SIMT kernels), "bf16" (tcgen05 tensor path) or None = follow the inputs/autocast like the reference
        self.precision = precision or os.environ.get("BOXDREAMER_B200_PRECISION") or None
        self.write_pred_bbox = True
        self._engines = {}
        self._weights_version = 0
        self.register_load_state_dict_post_hook(lambda m, k: m._bump())
        self.rgb_encoder.model.register_load_state_dict_post_hook(lambda m, k: self._bump())

    # -- engine management ------------------------------------------------------------------
    def _bump(self):
        self._weights_version += 1

    def sync_weights(self):
        """Call after mutating parameters in place (load_state_dict is tracked automatically)."""
        self._bump()

    def _pick_precision(self, t: torch.Tensor) -> int:
        p = self.precision or getattr(self, "_forward_precision", None)   # fixed once per forward (see forward())
        if p is None:
            low = t.dtype in (torch.bfloat16, torch.float16) or torch.is_autocast_enabled()
            p = "bf16" if low else "exact"
        if p not in ("bf16", "exact"):
            raise ValueError(f"precision must be 'bf16' or 'exact', got {p}")
        return _lib.PRECISION_BF16 if p == "bf16" else _lib.PRECISION_EXACT

    def _as_engine_input(self, t: torch.Tensor) -> torch.Tensor:
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.float()
        return t.contiguous()

    def _engine_for(self, like: torch.Tensor, B: int, T: int, images: int = 0) -> Engine:
        """The engine (workspace) for B sequences of T views; `images` additionally asks for room for that many images in
        the encoder (its calls are not tied to the B x T shape).  The workspace scales with max_batch * max_views images,
        so when it has to grow it keeps that product at what is needed instead of multiplying the old maxima: the dense path
        mixes (B*T images, 1 view), (B*n_sub, sub+1) and (B, fine_topk+1) calls and would otherwise allocate ~6x too much."""
        if not like.is_cuda:
            raise _lib.BoxDreamerLibError("boxdreamer_b200: inputs must be CUDA tensors (no CPU fallback); "
                                          "call model.cuda() and move the batch to the GPU")
        prec = self._pick_precision(like)
        key = (prec, like.device.index)
        eng = self._engines.get(key)
        need = max(B * T, images)
        with torch.cuda.device(like.device):
            if eng is None or eng.max_batch * eng.max_views < need or eng.max_batch < B or eng.max_views < T:
                old_images = eng.max_batch * eng.max_views if eng is not None else 0
                mv = max(T, eng.max_views if eng else 0)
                mb = max(B, -(-max(need, old_images) // mv))
                if eng is not None:
                    eng.close()
                eng = Engine(self.image_size, self.patch_size, self.decoder.d_model, self.decoder.att_depth,
                             self.decoder.nhead, prec, mb, mv)
                self._engines[key] = eng
            if eng.weights_version != self._weights_version:
                eng.load_weights(self._named_weights())
                eng.weights_version = self._weights_version
        return eng

    def _named_weights(self) -> dict:
        out = {k: v for k, v in self.state_dict().items()}
        dm = self.rgb_encoder.model
        for k, v in dm.state_dict().items():
            if k in ("mask_token",):
                continue
            out["dino." + k] = v
        out["dino.pos_embed"] = dm.interpolated_pos_embed(self.image_size)
        return out

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        # DinoV2Wrapper is a plain object in the reference too; keep it on the module's device for convenience
        try:
            dev = next(self.parameters()).device
            self.rgb_encoder.to_device(dev)
        except StopIteration:
            pass
        return r

    # -- forward ----------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, data):
        poses = data["poses"]
        images = data["images"]
        B, T = poses.shape[:2]
        query_idx = data["query_idx"]
        if not torch.is_tensor(query_idx):
            raise NotImplementedError("Query index must be specified")
        dev = images.device
        if tuple(images.shape[-2:]) != (self.image_size, self.image_size):
            raise AssertionError(f"H and W should be equal to img_size {self.image_size}, got {tuple(images.shape[-2:])}")
        query_idx = query_idx.to(device=dev, dtype=torch.int64).contiguous()
        # the precision is decided once, from the images (or self.precision): later engine look-ups keyed by fp32 intermediates
        # (heat maps, pooled proposals of the dense path) must not build a second engine at another precision
        if self.precision is None:
            low = images.dtype in (torch.bfloat16, torch.float16) or torch.is_autocast_enabled()
            self._forward_precision = "bf16" if low else "exact"
        try:
            return self._forward_impl(data, poses, images, B, T, query_idx, dev)
        finally:
            self._forward_precision = None

    def _forward_impl(self, data, poses, images, B, T, query_idx, dev):
        camera_mask = torch.zeros(poses.shape[:2], dtype=torch.bool, device=dev)
        camera_mask[torch.arange(B, device=dev), query_idx] = True
        data["camera_mask"] = camera_mask.clone()

        if "bbox_feat" not in data:
            # device-side input synthesis: the dataset ships the projected corners in crop
            # pixels ("bbox_proj_px" [B,T,8,2], fp32) instead of the 8*S*S heat maps (src/datasets/base.py:689-693)
            if "bbox_proj_px" not in data:
                raise KeyError("data needs 'bbox_feat' (reference contract) or 'bbox_proj_px' (device-side synthesis)")
            from .inputs import make_bbox_features
            feat_dtype = images.dtype if images.dtype in (torch.float32, torch.bfloat16) else torch.float32
            data["bbox_feat"] = make_bbox_features(data["bbox_proj_px"].to(dev).reshape(B * T, 8, 2), "heatmap",
                                                   (self.image_size, self.image_size), dtype=feat_dtype, group=T).view(
                B, T, 8, self.image_size, self.image_size)
        if self.dense_cfg is not None and bool(self.dense_cfg["enable"]):
            return self._forward_dense(data, camera_mask)

        eng = self._engine_for(images, B, T)
        imgs = self._as_engine_input(images)
        bbox_feat = data["bbox_feat"]
        bbox = self._as_engine_input(bbox_feat)
        if bbox.dtype != imgs.dtype:
            bbox = bbox.to(imgs.dtype)
        K_q = data["non_ndc_intrinsics"][camera_mask].float().contiguous()
        bbox3d_q = data["bbox_3d"][camera_mask].float().contiguous()

        if self.training:
            feats = eng.dino_forward(imgs.view(B * T, 3, self.image_size, self.image_size))
            heat = eng.decoder_forward(bbox, feats, query_idx)
            _, nm = eng.corners_topk(heat)
            qposes = None
        else:
            heat, _, nm, qposes = eng.forward(imgs, bbox, query_idx, bbox3d_q, K_q, want_heat=True)

        # _update_predictions (BoxDreamerModel.py:341-344)
        if self.write_pred_bbox:
            data["pred_bbox"] = bbox_feat.clone()
            data["pred_bbox"][camera_mask] = heat.to(bbox_feat.dtype)
        else:
            data["pred_bbox_query"] = heat
        # process_prediction / calculate_bb8_projections (prediction_utils.py:88-103, 106-136)
        data["regression_boxes"] = data["bbox_proj_crop"].clone()
        data["regression_boxes"][camera_mask] = nm.to(data["regression_boxes"].dtype)
        pred_poses = poses.clone()
        if qposes is not None:
            pred_poses[camera_mask] = qposes.to(pred_poses.dtype)
            pred_poses = torch.nan_to_num(pred_poses, nan=0.0, posinf=0.0, neginf=0.0)
        data["pred_poses"] = pred_poses
        data["pred_intrinsics"] = data["intrinsics"]
        return data

    # -- reference-feature cache ------------------------------------------------
    @torch.no_grad()
    def encode_references(self, images, bbox_feat):
        """Encodes a reference set once: images [R,3,S,S], bbox_feat [R,8,S,S] (the views every query of a video / demo
        session shares; the reference re-encodes them for every query, BoxDreamerModel.py:274-285).  Returns the cache to
        pass to `forward_with_references`."""
        R = images.shape[0]
        eng = self._engine_for(images, 1, 1, images=R)
        feats = eng.dino_forward(self._as_engine_input(images))
        return {"feats": feats, "bbox_feat": bbox_feat, "n_ref": R}

    @torch.no_grad()
    def forward_with_references(self, query_images, cache, bbox3d_q, K_q):
        """query_images [B,3,S,S] against the cached reference set: only the B query crops go through the encoder
        (1/(R+1) of the encoder work), then the usual decoder over [R references + query] per sample, top-20 corners and
        PnP.  Returns the query-view entries of the reference's output dict:
        pred_bbox [B,8,S,S], regression_boxes [B,8,2] (normalised), keypoints [B,8,2] (pixels), pred_poses [B,4,4].
        Identical to `forward` on a batch whose every sample lists the cached references followed by the query."""
        B = query_images.shape[0]
        R = cache["n_ref"]
        T = R + 1
        dev = query_images.device
        eng = self._engine_for(query_images, B, T)
        q_feats = eng.dino_forward(self._as_engine_input(query_images))                     # [B,P,d]
        feats = torch.cat([cache["feats"].unsqueeze(0).expand(B, R, *q_feats.shape[1:]), q_feats.unsqueeze(1)], dim=1).contiguous()
        ref_maps = cache["bbox_feat"]
        bbox = torch.cat([ref_maps.unsqueeze(0).expand(B, R, *ref_maps.shape[1:]),
                          torch.zeros(B, 1, *ref_maps.shape[1:], device=dev, dtype=ref_maps.dtype)], dim=1).contiguous()
        bbox = self._as_engine_input(bbox)
        if bbox.dtype != self._as_engine_input(query_images).dtype:
            bbox = bbox.to(self._as_engine_input(query_images).dtype)
        query_idx = torch.full((B,), R, dtype=torch.int64, device=dev)
        heat = eng.decoder_forward(bbox, feats, query_idx)       # the query's own map is replaced by the learnable token
        px, nm = eng.corners_topk(heat)
        out = {"pred_bbox": heat, "regression_boxes": nm, "keypoints": px}
        if not self.training:
            out["pred_poses"] = torch.nan_to_num(eng.pnp(px, bbox3d_q.float().contiguous(), K_q.float().contiguous()), nan=0.0, posinf=0.0, neginf=0.0)
        return out

    # -- dense reference sets (BoxDreamerModel.py:289-330) -------------------
    def _pooled_pose(self, heats, bbox3d_q, K_q, n_hyp=512, thr_px=2.0):
        """recover_pose_from_dense_bb8 (box_utils.py:202-304): top-20 corners of every proposal, all n_sub*8 2D-3D
        pairs of a query pooled into one robust PnP (inlier threshold 2 px, as the reference's solvePnPRansac call)."""
        B, n_sub = heats.shape[:2]
        eng = self._engine_for(heats, 1, 1)    # top-20 and PnP need no workspace
        px, _ = eng.corners_topk(heats.reshape(B * n_sub, *heats.shape[2:]).float().contiguous())
        pts2d = px.reshape(B, n_sub * 8, 2).contiguous()
        pts3d = bbox3d_q.float().unsqueeze(1).expand(B, n_sub, 8, 3).reshape(B, n_sub * 8, 3).contiguous()
        opts = _lib.BdPnpOpts(1, n_hyp, thr_px, 0, 30)
        return eng.pnp(pts2d, pts3d, K_q.float().contiguous(), opts)

    def _forward_dense(self, data, camera_mask):
        from . import dense
        cfg = self.dense_cfg
        frames = data["images"]
        pose_feat = data["bbox_feat"]
        image_masks = data["image_masks"] if "image_masks" in data else torch.ones_like(frames[:, :, :1])
        rgb_feature = self.rgb_encoder.predict(frames)     # every view is encoded once; selection works on the tokens
        data, pose_feat, frames, camera_mask, rgb_feature, image_masks = dense.process_dense_input(
            data, pose_feat, frames, camera_mask, rgb_feature, image_masks, cfg)
        if bool(dense._cfg(cfg, "multi_round")):
            heat = dense.process_multi_round(data, pose_feat, frames, camera_mask, rgb_feature, image_masks, self.decoder, cfg,
                                             self.bbox_representation, self._pooled_pose)
        else:
            heat = self.decoder(pose_feat, frames, camera_mask, rgb_feature, dense.normalize(image_masks))
        # the selection rewrote the per-view entries of `data`: everything below reads the rewritten ones
        camera_mask = data["camera_mask"]
        poses = data["poses"]
        bbox_feat = data["bbox_feat"]
        B, T = poses.shape[:2]
        eng = self._engine_for(heat, B, T)
        px, nm = eng.corners_topk(heat)
        data["pred_bbox"] = bbox_feat.clone()
        data["pred_bbox"][camera_mask] = heat.to(bbox_feat.dtype)
        data["regression_boxes"] = data["bbox_proj_crop"].clone()
        data["regression_boxes"][camera_mask] = nm.to(data["regression_boxes"].dtype)
        pred_poses = poses.clone()
        if not self.training:
            K_q = data["non_ndc_intrinsics"][camera_mask].float().contiguous()
            bbox3d_q = data["bbox_3d"][camera_mask].float().contiguous()
            qposes = eng.pnp(px, bbox3d_q, K_q)
            pred_poses[camera_mask] = qposes.to(pred_poses.dtype)
            pred_poses = torch.nan_to_num(pred_poses, nan=0.0, posinf=0.0, neginf=0.0)
        data["pred_poses"] = pred_poses
        data["pred_intrinsics"] = data["intrinsics"]
        return data
