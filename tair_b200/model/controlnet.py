"""ControlledUnetModel and IRControlNet (terediff/model/controlnet.py:16-56, :59-337) on the sm_100a kernels.

Public ``forward`` signatures, argument meaning and return values follow the reference:
    ControlNet.forward(x, hint, timesteps, context) -> list of 13 control tensors
    ControlledUnetModel.forward(x, timesteps, context, control, only_mid_control) -> (out, [4 decoder features])
with reference-facing tensors in (B,C,H,W) fp32.  Internally both networks run channels-last bf16; the
``*_nhwc`` methods expose that fast path to ``ControlLDM.forward`` so no layout round-trip happens between them.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from .. import ops
from .unet import TimestepEmbedSequential, UNetModel
from .util import BF16, Conv1x1, zero_module


def _is_nhwc_bf16(t: torch.Tensor) -> bool:
    return t.dtype == BF16


class ControlledUnetModel(UNetModel):
    extract_idx = (2, 5, 8, 11)  # controlnet.py:44

    def forward_nhwc(self, x_nhwc: torch.Tensor, timesteps: torch.Tensor, context: torch.Tensor,
                     control: Optional[List[torch.Tensor]] = None, only_mid_control: bool = False):
        """x_nhwc: padded channels-last latent; control: 13 channels-last bf16 tensors (consumed from the end)."""
        enc = self.encode_nhwc(x_nhwc, timesteps, context)
        return self.decode_nhwc(enc, control, only_mid_control)

    def encode_nhwc(self, x_nhwc: torch.Tensor, timesteps: torch.Tensor, context: torch.Tensor):
        """Input blocks + middle block: the part of the UNet that does not depend on the ControlNet residuals
        (controlnet.py:18-56 adds them only after the middle block), so ControlLDM can run it beside the ControlNet."""
        step = self._begin_step(timesteps, context)
        hs = []
        h = x_nhwc
        for module in self.input_blocks:
            h = module(h, step)
            hs.append(h)
        h = self.middle_block(h, step)
        return step, hs, h

    def decode_nhwc(self, enc, control: Optional[List[torch.Tensor]] = None, only_mid_control: bool = False):
        step, hs, h = enc
        control = None if control is None else list(control)
        if control is not None:
            h = ops.add(h, control.pop())
        feats = []
        for i, module in enumerate(self.output_blocks):
            skip = hs.pop()
            if only_mid_control or control is None:
                h = ops.concat_add(h, skip)
            else:
                h = ops.concat_add(h, skip, control.pop())
            h = module(h, step)
            if i in self.extract_idx:
                feats.append(h)
        h = self.out[0](h, act=ops.ACT_SILU)
        out = self.out[2](h)
        return out, feats

    def forward(self, x, timesteps=None, context=None, control=None, only_mid_control=False, **kwargs):
        if control is not None:
            control = [c if _is_nhwc_bf16(c) else ops.nchw_to_nhwc(c.float()) for c in control]
        out, feats = self.forward_nhwc(self._embed_input(x), timesteps, context, control, only_mid_control)
        return ops.nhwc_to_nchw(out, self.out_channels), [ops.nhwc_to_nchw(f) for f in feats]


class ControlNet(UNetModel):
    """Encoder copy of the UNet taking latent || hint (8 channels), every block followed by a 1x1 'zero' conv."""

    def __init__(self, image_size=32, in_channels=4, model_channels=320, hint_channels=4, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), **kw):
        kw.pop("out_channels", None)
        super().__init__(image_size=image_size, in_channels=in_channels, model_channels=model_channels,
                         out_channels=in_channels, num_res_blocks=num_res_blocks,
                         attention_resolutions=attention_resolutions, _build_decoder=False,
                         _hint_channels=hint_channels, **kw)
        self.hint_channels = hint_channels
        self.zero_convs = nn.ModuleList([self.make_zero_conv(c) for c in self._enc_channels])
        self.middle_block_out = self.make_zero_conv(self._mid_channels)

    @staticmethod
    def make_zero_conv(channels: int) -> TimestepEmbedSequential:
        """controlnet.py:318-321."""
        return TimestepEmbedSequential(zero_module(Conv1x1(channels, channels)))

    def forward_nhwc(self, xh_nhwc: torch.Tensor, timesteps: torch.Tensor, context: torch.Tensor) -> List[torch.Tensor]:
        step = self._begin_step(timesteps, context)
        outs = []
        h = xh_nhwc
        for module, zero_conv in zip(self.input_blocks, self.zero_convs):
            h = module(h, step)
            outs.append(zero_conv[0](h))
        h = self.middle_block(h, step)
        outs.append(self.middle_block_out[0](h))
        return outs

    def forward(self, x, hint, timesteps, context, **kwargs):
        xh = torch.cat((x, hint), dim=1)  # controlnet.py:326
        outs = self.forward_nhwc(self._embed_input(xh), timesteps, context)
        return [ops.nhwc_to_nchw(o) for o in outs]
