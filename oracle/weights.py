"""TEST INFRASTRUCTURE — seeded, non-degenerate weights for parity tests and benchmarks.

The reference constructors zero the second conv of every ResBlock, every transformer
``proj_out``, the UNet output conv, all ControlNet zero-convs and the MSDeformAttn
offset/weight projections (unet.py:177-179, attention.py:331, controlnet.py:318-321,
ms_deform_attn.py:101-110), so a default-initialised network is numerically vacuous
(SURVEY.md §8a hazard 1).  Every tensor is therefore overwritten from a generator keyed
by (seed, parameter name): independent of iteration order, identical on every host.

    >=2-D ".weight"/in_proj_weight : N(0, gain^2 / fan_in); gain 0.5 for the residual-closing
                                    layers (ResBlock out_layers.3, transformer proj_out)
    1-D  "*.weight" (norm scales)  : 1 + N(0, 0.02^2)
    1-D  other (biases)            : N(0, 0.02^2); TESTR class-head biases 0 so detections exist
    embeddings / level_embed       : N(0, 1)
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Sequence

import torch

_EMBED_SUFFIXES = ("ctrl_point_embed.weight", "text_embed.weight", "level_embed")
_HALF_GAIN_SUFFIXES = ("out_layers.3.weight", "proj_out.weight")
_ZERO_BIAS_PARTS = ("ctrl_point_class", "bbox_class")


def init_tensor(name: str, shape: Sequence[int], seed: int = 1234, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFFFFFFFFFF)
    shape = tuple(int(s) for s in shape)
    z = torch.randn(shape, generator=g, dtype=torch.float32)
    if name.endswith(_EMBED_SUFFIXES):
        out = z
    elif len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        gain = 0.5 if name.endswith(_HALF_GAIN_SUFFIXES) else 1.0
        out = z * (gain / fan_in ** 0.5)
    elif name.endswith(".weight"):
        out = 1.0 + 0.02 * z
    elif name.endswith(".bias") and any(p in name for p in _ZERO_BIAS_PARTS):
        out = torch.zeros(shape)
    else:
        out = 0.02 * z
    return out.to(dtype)


def seeded_state_dict(manifest: Mapping[str, Sequence[int]], seed: int = 1234) -> Dict[str, torch.Tensor]:
    """manifest: parameter/buffer name -> shape (e.g. ``{k: v.shape for k, v in module.state_dict().items()}``)."""
    return {k: init_tensor(k, shp, seed) for k, shp in manifest.items()}


def manifest_of(module: torch.nn.Module) -> Dict[str, list]:
    return {k: list(v.shape) for k, v in module.state_dict().items() if v.dtype.is_floating_point}


def seeded_randn(shape, seed: int, scale: float = 1.0) -> torch.Tensor:
    """Deterministic CPU N(0, scale^2) tensor; the input generator shared by make_golden.py and the tests."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(tuple(shape), generator=g) * scale
