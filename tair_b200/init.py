"""Seeded non-degenerate initialisation for synthetic-weight runs (bench / smoke).

The reference constructors zero half of the network (``zero_module`` on the closing conv of every ResBlock, every
transformer ``proj_out``, the UNet output conv and all ControlNet zero-convs), so a default-initialised model
computes v == 0.  For throughput runs with random weights every tensor is overwritten on the device:
matrices ~ N(0, gain^2/fan_in) (gain 0.5 on the residual-closing layers), norm scales ~ 1 + N(0, 0.02^2),
biases ~ N(0, 0.02^2).  Real checkpoints simply ``load_state_dict`` instead.
"""
from __future__ import annotations

import torch
from torch import nn

_HALF_GAIN = ("out_layers.3.weight", "proj_out.weight")


@torch.no_grad()
def nondegenerate_init_(module: nn.Module, seed: int = 1234) -> nn.Module:
    dev = next(module.parameters()).device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for name, p in module.named_parameters():
        z = torch.randn(p.shape, device=dev, generator=g, dtype=torch.float32)
        if p.dim() >= 2:
            fan_in = p[0].numel()
            gain = 0.5 if name.endswith(_HALF_GAIN) else 1.0
            p.copy_(z * (gain / fan_in ** 0.5))
        elif name.endswith(".weight"):
            p.copy_(1.0 + 0.02 * z)
        else:
            p.copy_(0.02 * z)
    return module
