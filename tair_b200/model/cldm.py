"""ControlLDM — the restoration model wrapper (terediff/model/cldm.py:20-217) on the sm_100a kernels.

``forward(x_noisy, t, cond) -> (eps, extracted_feats[4])`` keeps the reference contract (cldm.py:160-179):
IRControlNet on (x_noisy || c_img) -> per-block ``control_scales`` -> controlled UNet.  Both networks exchange
channels-last bf16 tensors directly; only the public inputs/outputs are (B,C,H,W) fp32.

The VAE (``tair_b200.model.vae.AutoencoderKL``, once per tile) and the OpenCLIP text encoder
(``tair_b200.model.clip.FrozenOpenCLIPEmbedder``, once per step) are built from ``vae_cfg`` / ``clip_cfg`` on the same kernels.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
from torch import nn

from .. import ops
from .controlnet import ControlledUnetModel, ControlNet
from .util import BF16


class ControlLDM(nn.Module):
    def __init__(self, unet_cfg: dict, vae_cfg: Optional[dict] = None, clip_cfg: Optional[dict] = None,
                 controlnet_cfg: Optional[dict] = None, latent_scale_factor: float = 0.18215):
        """Argument order of the reference constructor (cldm.py:22-31).  ``vae_cfg`` builds the AutoencoderKL on our
        kernels and ``clip_cfg`` the OpenCLIP text encoder (``tair_b200.model.clip.FrozenOpenCLIPEmbedder``); either may be
        None and attached later.  ``ControlLDM(unet_cfg, controlnet_cfg)`` is accepted as a shorthand."""
        super().__init__()
        if controlnet_cfg is None and isinstance(vae_cfg, dict) and "hint_channels" in vae_cfg:
            vae_cfg, controlnet_cfg = None, vae_cfg
        if controlnet_cfg is None:
            raise ValueError("ControlLDM needs a controlnet_cfg")
        self.unet = ControlledUnetModel(**unet_cfg)
        self.vae: Optional[nn.Module] = None
        if vae_cfg is not None:
            from .vae import AutoencoderKL
            self.vae = AutoencoderKL(**vae_cfg)
        self.clip: Optional[nn.Module] = None
        if clip_cfg is not None:
            from .clip import FrozenOpenCLIPEmbedder
            self.clip = FrozenOpenCLIPEmbedder(**clip_cfg)
        self.controlnet = ControlNet(**controlnet_cfg)
        self.scale_factor = latent_scale_factor
        self.control_scales = [1.0] * 13  # cldm.py:30
        self.return_nhwc_feats = False  # True: hand channels-last bf16 features to the TESTR head (no round trip)
        self.overlap_controlnet = os.environ.get("TAIR_OVERLAP_CONTROLNET", "1") != "0"
        self._side: Dict[torch.device, "torch.cuda.Stream"] = {}

    def _side_stream(self, device) -> "torch.cuda.Stream":
        s = self._side.get(device)
        if s is None:
            s = self._side[device] = torch.cuda.Stream(device=device)
        return s

    # -- checkpoint loading (host-side key mapping; cldm.py:33-90) -----------------------------------
    @torch.no_grad()
    def load_pretrained_sd(self, sd: Dict[str, torch.Tensor]):
        """cldm.py:33-61: pick ``model.diffusion_model.*`` / ``first_stage_model.*`` / ``cond_stage_model.*`` out of a
        Stable Diffusion 2.1 checkpoint -> (unused, missing) key sets; the three modules are frozen in eval mode."""
        module_map = {"unet": "model.diffusion_model", "vae": "first_stage_model", "clip": "cond_stage_model"}
        used, missing = set(), set()
        mods = [(n, getattr(self, n)) for n in ("unet", "vae", "clip") if isinstance(getattr(self, n), nn.Module)]
        for name, module in mods:
            init_sd = {}
            for key in module.state_dict():
                target = ".".join([module_map[name], key])
                if target not in sd:
                    missing.add(target)
                    continue
                init_sd[key] = sd[target].clone()
                used.add(target)
            module.load_state_dict(init_sd, strict=False)
        for _, module in mods:
            module.eval()
            module.train = lambda mode=True, _m=module: _m      # disabled_train (cldm.py:14-17)
            for p in module.parameters():
                p.requires_grad = False
        return set(sd.keys()) - used, missing

    @torch.no_grad()
    def load_controlnet_from_ckpt(self, sd: Dict[str, torch.Tensor]) -> None:
        """cldm.py:63-65."""
        self.controlnet.load_state_dict(sd, strict=True)

    @torch.no_grad()
    def load_controlnet_from_unet(self):
        """cldm.py:67-90: initialise the ControlNet from the UNet encoder; the 8-channel input conv gets the UNet's
        4 input channels plus zeros -> (keys widened with zeros, keys kept from scratch)."""
        unet_sd, scratch = self.unet.state_dict(), self.controlnet.state_dict()
        init_sd, widened, kept = {}, set(), set()
        for key, this in scratch.items():
            if key in unet_sd:
                target = unet_sd[key]
                if this.size() == target.size():
                    init_sd[key] = target.clone()
                else:
                    oc, _, h, w = this.size()
                    zeros = torch.zeros((oc, this.size(1) - target.size(1), h, w), dtype=target.dtype, device=target.device)
                    init_sd[key] = torch.cat((target, zeros), dim=1)
                    widened.add(key)
            else:
                init_sd[key] = this.clone()
                kept.add(key)
        self.controlnet.load_state_dict(init_sd, strict=True)
        return widened, kept

    def cast_dtype(self, dtype: torch.dtype) -> "ControlLDM":
        """cldm.py:181-217 switches the reference's blocks to fp16.  The kernels here always compute in bf16 with fp32
        accumulation and keep fp32 master parameters, so there is nothing to convert; kept for API compatibility."""
        self.unet.dtype = dtype
        self.controlnet.dtype = dtype
        return self

    # -- optional non-hot-path submodules ---------------------------------------------------------
    def attach_vae(self, vae: nn.Module) -> None:
        self.vae = vae

    def attach_clip(self, clip) -> None:
        """Plug in a text encoder: a module (registered as a child) or any object with ``encode(list[str])``."""
        self._modules.pop("clip", None)
        self.__dict__.pop("clip", None)
        if isinstance(clip, nn.Module) or clip is None:
            self.clip = clip
        else:
            object.__setattr__(self, "clip", clip)

    @torch.no_grad()
    def vae_decode(self, z: torch.Tensor) -> torch.Tensor:
        """cldm.py:121-141 (untiled)."""
        if self.vae is None:
            raise RuntimeError("ControlLDM.vae_decode: no VAE attached")
        return self.vae.decode(z / self.scale_factor)

    @torch.no_grad()
    def vae_encode(self, image: torch.Tensor, sample: bool = True) -> torch.Tensor:
        """cldm.py:92-119 (untiled)."""
        if self.vae is None:
            raise RuntimeError("ControlLDM.vae_encode: no VAE attached")
        post = self.vae.encode(image)
        return (post.sample() if sample else post.mode()) * self.scale_factor

    @torch.no_grad()
    def prepare_condition(self, cond_img: torch.Tensor, txt: List[str]) -> Dict[str, torch.Tensor]:
        """cldm.py:143-158."""
        if self.clip is None:
            raise RuntimeError("ControlLDM.prepare_condition: no text encoder attached")
        return dict(c_txt=self.clip.encode(txt), c_img=self.vae_encode(cond_img * 2 - 1, sample=False))

    # -- the hot path -------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x_noisy: torch.Tensor, t: torch.Tensor, cond: Dict[str, torch.Tensor]):
        c_txt = cond["c_txt"]
        unet, cn = self.unet, self.controlnet
        control = None
        if "c_img" in cond and self.overlap_controlnet:
            # The UNet encoder does not depend on the ControlNet (its residuals enter after the middle block), so the
            # two run on two streams: every kernel is a persistent grid, and the second stream's CTAs fill the SMs that
            # the first one's tail wave and its small 8x8 / 16x16 layers leave idle.  Fork/join is by events, which a
            # CUDA-graph capture records as parallel branches; results are unchanged (no atomics anywhere).
            main = torch.cuda.current_stream()
            side = self._side_stream(x_noisy.device)
            xh = torch.cat((x_noisy, cond["c_img"]), dim=1)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                control = cn.forward_nhwc(cn._embed_input(xh), t, c_txt)
                if any(s != 1.0 for s in self.control_scales):
                    control = [(c.float() * s).to(BF16) for c, s in zip(control, self.control_scales)]
            enc = unet.encode_nhwc(unet._embed_input(x_noisy), t, c_txt)
            main.wait_stream(side)
            out, feats = unet.decode_nhwc(enc, control, False)
        else:
            if "c_img" in cond:
                xh = torch.cat((x_noisy, cond["c_img"]), dim=1)
                control = cn.forward_nhwc(cn._embed_input(xh), t, c_txt)
                if any(s != 1.0 for s in self.control_scales):
                    control = [(c.float() * s).to(BF16) for c, s in zip(control, self.control_scales)]
            out, feats = unet.forward_nhwc(unet._embed_input(x_noisy), t, c_txt, control, False)
        eps = ops.nhwc_to_nchw(out, unet.out_channels)
        if not self.return_nhwc_feats:
            feats = [ops.nhwc_to_nchw(f) for f in feats]
        return eps, feats
