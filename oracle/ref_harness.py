"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (yinnhao/TAIR) in-place.

Only usable where the reference tree exists (``$TAIR_REF`` or ``/root/reference``,
i.e. the build container; never on the GPU box).  It is used by
``tests/golden/make_golden.py`` to generate the committed fixtures and by the
``requires_reference`` tests that pin ``oracle/`` against the real modules.
Nothing under ``tair_b200/`` may import this file.

The reference needs a handful of packages that are not installed here; each gets
the smallest stub that lets the *inference* path import:
ftfy, timm, torchsde, omegaconf, accelerate, pyiqa, wandb, detectron2 and the
compiled ``testr.adet._C`` (replaced by the reference's own pure-PyTorch
``ms_deform_attn_core_pytorch``, testr/adet/layers/ms_deform_attn.py:39-59).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import typing
from types import SimpleNamespace

import torch

REF_ROOT = os.environ.get("TAIR_REF", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "terediff"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave as a package
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(_stub(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


class _Instances:
    """Attribute bag with the semantics of detectron2.structures.Instances
    (detectron2/detectron2/structures/instances.py:8-144) that the hot path uses."""

    def __init__(self, image_size, **kw):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in kw.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            object.__setattr__(self, name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError(name)
        return self._fields[name]

    def set(self, name, value):
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get_fields(self):
        return self._fields

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")


_installed = False


def install() -> None:
    """Put the reference on sys.path behind the stubs.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    def _try(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not _try("ftfy"):
        _stub("ftfy", fix_text=lambda s: s)
    if not _try("timm"):
        def _to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        class DropPath(torch.nn.Identity):
            def __init__(self, *a, **k):
                super().__init__()
        _stub("timm")
        _stub("timm.models")
        _stub("timm.models.layers", DropPath=DropPath, to_2tuple=_to_2tuple,
              trunc_normal_=torch.nn.init.trunc_normal_)
    if not _try("torchsde"):
        _stub("torchsde")
    if not _try("omegaconf"):
        class ListConfig(list):
            pass

        class OmegaConf:  # only .load of a yaml into attribute dicts is used by val_patches.py
            @staticmethod
            def load(path):
                import yaml

                def wrap(o):
                    if isinstance(o, dict):
                        return _AttrDict({k: wrap(v) for k, v in o.items()})
                    if isinstance(o, list):
                        return [wrap(v) for v in o]
                    return o
                with open(path) as f:
                    return wrap(yaml.safe_load(f))
        _stub("omegaconf", OmegaConf=OmegaConf)
        _stub("omegaconf.listconfig", ListConfig=ListConfig)
    for name in ("accelerate", "accelerate.utils", "pyiqa", "wandb"):
        if not _try(name):
            _stub(name, Accelerator=object, DistributedDataParallelKwargs=object, set_seed=lambda *_a, **_k: None,
                  create_metric=lambda *_a, **_k: None)
    if not _try("detectron2"):
        _stub("detectron2")
        _stub("detectron2.structures", Instances=_Instances, ImageList=object, Boxes=object)
        _stub("detectron2.utils")
        _stub("detectron2.utils.comm", get_world_size=lambda: 1)
        _stub("detectron2.config", CfgNode=dict)
    if not hasattr(torch, "Tuple"):
        torch.Tuple = typing.Tuple  # terediff/sampler/edm_sampler.py:145

    for p in (REF_ROOT, os.path.join(REF_ROOT, "testr")):
        if p not in sys.path:
            sys.path.insert(0, p)

    # compiled MSDeformAttn op has no CPU path: register an empty _C and route the autograd
    # Function through the reference's own pure-PyTorch restatement.
    adet = importlib.import_module("testr.adet")  # namespace package of the reference
    fake_c = types.ModuleType("testr.adet._C")
    sys.modules["testr.adet._C"] = fake_c
    adet._C = fake_c
    msda = importlib.import_module("testr.adet.layers.ms_deform_attn")

    def _apply(value, shapes, start, loc, w, _step):
        return msda.ms_deform_attn_core_pytorch(value, shapes.tolist(), loc, w)
    msda._MSDeformAttnFunction.apply = staticmethod(_apply)
    _installed = True


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


# --- reference constructors --------------------------------------------------------------------

def unet_cfg(model_channels=320, context_dim=1024, channel_mult=(1, 2, 4, 4), num_head_channels=64):
    """configs/val/val_terediff.yaml:6-20 (unet_cfg) with overridable width for small test cases."""
    return dict(use_checkpoint=False, image_size=32, in_channels=4, out_channels=4, model_channels=model_channels,
                attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=list(channel_mult),
                num_head_channels=num_head_channels, use_spatial_transformer=True, use_linear_in_transformer=True,
                transformer_depth=1, context_dim=context_dim, legacy=False)


def controlnet_cfg(**kw):
    """configs/val/val_terediff.yaml:53-67 (controlnet_cfg)."""
    c = unet_cfg(**kw)
    c.pop("out_channels")
    c["hint_channels"] = 4
    return c


def build_unet(**kw):
    install()
    from terediff.model.controlnet import ControlledUnetModel
    return ControlledUnetModel(**unet_cfg(**kw)).eval()


def build_controlnet(**kw):
    install()
    from terediff.model.controlnet import ControlNet
    return ControlNet(**controlnet_cfg(**kw)).eval()


def testr_cfg(device="cpu"):
    """yacs cfg subset read by TESTR / TransformerDetector: testr/adet/config/defaults.py:340-369 overlaid
    with testr/configs/TESTR/{Base-TESTR,TESTR_R_50_Polygon}.yaml."""
    loss = SimpleNamespace(AUX_LOSS=True, POINT_CLASS_WEIGHT=2.0, POINT_COORD_WEIGHT=5.0, POINT_TEXT_WEIGHT=4.0,
                           BOX_CLASS_WEIGHT=2.0, BOX_COORD_WEIGHT=5.0, BOX_GIOU_WEIGHT=2.0, FOCAL_ALPHA=0.25,
                           FOCAL_GAMMA=2.0)
    tr = SimpleNamespace(ENABLED=True, INFERENCE_TH_TEST=0.5, VOC_SIZE=96, NUM_CHARS=25, AUX_LOSS=True,
                         ENC_LAYERS=6, DEC_LAYERS=6, DIM_FEEDFORWARD=1024, HIDDEN_DIM=256, DROPOUT=0.1, NHEADS=8,
                         NUM_QUERIES=100, ENC_N_POINTS=4, DEC_N_POINTS=4, POSITION_EMBEDDING_SCALE=6.283185307179586,
                         NUM_FEATURE_LEVELS=4, USE_POLYGON=True, NUM_CTRL_POINTS=16, LOSS=loss)
    return SimpleNamespace(MODEL=SimpleNamespace(DEVICE=device, TRANSFORMER=tr))


def build_testr(device="cpu"):
    install()
    from testr.adet.modeling.transformer_detector import TransformerDetector
    return TransformerDetector(testr_cfg(device)).eval()
