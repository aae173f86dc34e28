"""Noise schedule of the diffusion process (host side, float64).

Mirrors the inference-relevant part of terediff/model/gaussian_diffusion.py: ``make_beta_schedule`` (:9-36),
``enforce_zero_terminal_snr`` (:49-72) and the ``Diffusion`` container (:75-110) whose ``betas`` feed the sampler
(val_patches.py:239-241).  The training loss (``p_losses``) is out of scope.
"""
from __future__ import annotations

import numpy as np


def make_beta_schedule(schedule: str, n_timestep: int, linear_start: float = 1e-4, linear_end: float = 2e-2,
                       cosine_s: float = 8e-3) -> np.ndarray:
    if schedule == "linear":
        return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    if schedule == "sqrt_linear":
        return np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
    if schedule == "sqrt":
        return np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64) ** 0.5
    if schedule == "cosine":
        t = np.arange(n_timestep + 1, dtype=np.float64) / n_timestep + cosine_s
        a = np.cos(t / (1 + cosine_s) * np.pi / 2) ** 2
        a = a / a[0]
        return np.clip(1 - a[1:] / a[:-1], 0, 0.999)
    raise ValueError(f"schedule '{schedule}' unknown.")


def enforce_zero_terminal_snr(betas: np.ndarray) -> np.ndarray:
    """Shift/scale sqrt(alpha_bar) so the last step has zero SNR while the first keeps its value."""
    s = np.sqrt(np.cumprod(1.0 - betas))
    first, last = s[0].copy(), s[-1].copy()
    s = (s - last) * (first / (first - last))
    abar = s ** 2
    return 1.0 - np.concatenate([abar[:1], abar[1:] / abar[:-1]])


class Diffusion:
    def __init__(self, timesteps=1000, beta_schedule="linear", loss_type="l2", linear_start=1e-4, linear_end=2e-2,
                 cosine_s=8e-3, parameterization="eps", zero_snr=False):
        self.num_timesteps = timesteps
        self.parameterization = parameterization
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end,
                                   cosine_s=cosine_s)
        if zero_snr:
            betas = enforce_zero_terminal_snr(betas)
        self.betas = betas
        abar = np.cumprod(1.0 - betas, axis=0)
        self.sqrt_alphas_cumprod = np.sqrt(abar).astype(np.float32)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - abar).astype(np.float32)

    def q_sample(self, z_0, t, noise):
        """gaussian_diffusion.py:124-129 (forward diffusion; used by the legacy pipeline's start point / noise_aug)."""
        import torch
        a = torch.as_tensor(self.sqrt_alphas_cumprod, device=z_0.device)[t].view(-1, *([1] * (z_0.dim() - 1)))
        b = torch.as_tensor(self.sqrt_one_minus_alphas_cumprod, device=z_0.device)[t].view(-1, *([1] * (z_0.dim() - 1)))
        return a * z_0 + b * noise

    # the reference Diffusion is an nn.Module whose buffers are only used by the training loss; inference code calls
    # ``diffusion.to(device)`` (val_patches.py:240) and reads ``.betas`` (numpy) — keep both working
    def to(self, *args, **kwargs) -> "Diffusion":
        return self

    def eval(self) -> "Diffusion":
        return self


def val_diffusion() -> Diffusion:
    """configs/val/val_terediff.yaml:87-94."""
    return Diffusion(linear_start=0.00085, linear_end=0.0120, timesteps=1000, zero_snr=True, parameterization="v")
