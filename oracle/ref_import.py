"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference (`/root/reference/src/models/BoxDreamerModel.py`)
in this CPU-only container so that

  * `tests/golden/make_golden.py` can dump golden vectors from the reference's own
    `BoxDreamer.forward`, and
  * `tests/test_oracle_vs_reference.py` can pin `oracle/boxdreamer_oracle.py` against it.

`/root/reference` does not exist on the GPU box; nothing run there may import this module
(the importing tests skip themselves when the directory is absent).

Recipe:
  1. a `sys.meta_path` finder returns inert stub packages for the third-party packages the
     reference imports at module scope but never touches on the bb8/heatmap eval path;
  2. `timm.models.vision_transformer.Mlp` and `timm.layers.DropPath` are real stand-ins
     (blocks.py:28-29 needs them; semantics of vggsfm/models/modules.py:127-162);
  3. `flash_attn` is hidden so betr.py:67-75 picks the SDPA branch (blocks.py:273-285);
  4. `torch.hub.load` returns the vendored DINOv2 ViT-B/14 + 4 registers
     (src/models/sources/DINOv2/vision_transformer.py) with the upstream hub kwargs.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("BOXDREAMER_REFERENCE", "/root/reference")

_STUBBED = (
    "hydra omegaconf pytorch3d minipytorch3d kornia timm imageio pycolmap lightglue torchmetrics "
    "pytorch_lightning lightning roma trimesh open3d visdom poselib pyceres hloc gradio matplotlib "
    "plotly accelerate xformers apex lmdb albumentations natsort seaborn plyfile pyquaternion "
    "transforms3d h5py decord rerun cupy sam2 groundingdino dust3r mast3r wis3d yacs addict "
    "simple_parsing ray loguru"
).split()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


class _Dummy:
    """Callable / instantiable / sub-classable inert object."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        return type(name, (_Dummy,), {})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root not in self.roots:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _TimmMlp(nn.Module):
    """Stand-in for timm.models.vision_transformer.Mlp (timm==1.0.15): fc1 -> act -> fc2."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0.0, use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return x


from boxdreamer_b200.config import AttrDict, make_config  # noqa: E402,F401  (config tree shared with the product)


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("XFORMERS_DISABLED", "1")
    roots = []
    for name in _STUBBED:
        try:
            if importlib.util.find_spec(name) is None:
                roots.append(name)
        except (ImportError, ValueError):
            roots.append(name)
    for must in ("timm", "xformers"):
        if must not in roots:
            roots.append(must)
    sys.meta_path.insert(0, _StubFinder(roots))
    # real stand-ins
    import timm.models.vision_transformer as tv  # stub
    import timm.layers as tl  # stub
    tv.Mlp = _TimmMlp
    tl.DropPath = _DropPath
    sys.modules["flash_attn"] = None  # hide: CPU must take the SDPA branch
    # light-weight registration of src.datasets (its __init__ drags matplotlib)
    for pkg, sub in (("src.datasets", "src/datasets"), ("src.datasets.utils", "src/datasets/utils"),
                     ("src.datasets.utils.base", "src/datasets/utils/base")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REFERENCE_ROOT, sub)]
        sys.modules[pkg] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    from src.models.sources.DINOv2 import vision_transformer as vit

    def _hub_load(repo, model_type, *a, **k):
        assert model_type == "dinov2_vitb14_reg", model_type
        return vit.vit_base(patch_size=14, num_register_tokens=4, img_size=518, init_values=1.0,
                            block_chunks=0, interpolate_antialias=True, interpolate_offset=0.0)

    torch.hub.load = _hub_load
    _installed = True


def build_reference(img_size=224, num_layers=12):
    """Returns the reference's own `BoxDreamer` (eval mode, CPU, fp32)."""
    install()
    from src.models.modules.encoder import dinov2 as ref_dino
    # DinoV2Wrapper.load_model defaults to device='cuda' (dinov2.py:26); CPU here.
    orig = ref_dino.DinoV2Wrapper.load_model
    ref_dino.DinoV2Wrapper.load_model = lambda self, device="cpu": orig(self, device="cpu")
    from src.models.BoxDreamerModel import BoxDreamer
    model = BoxDreamer(make_config(img_size, num_layers)).eval()
    return model
