"""TEST INFRASTRUCTURE — restatement of the spaced DDPM sampler arithmetic (numpy float64 schedule, torch fp32 step).

Follows terediff/model/gaussian_diffusion.py:9-73 (beta schedule, zero-terminal-SNR rescale) and
terediff/sampler/spaced_sampler.py:14-65 (space_timesteps), :77-121 (make_schedule), :141-147 (x0 from v),
:123-131 (posterior), :167-189 (p_sample).  Pinned by tests/golden/schedule_50.json (generated from the imported
reference) and tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch


def linear_betas(n: int = 1000, start: float = 0.00085, end: float = 0.0120) -> np.ndarray:
    """make_beta_schedule('linear') — gaussian_diffusion.py:12-18: linspace over sqrt(beta), squared."""
    return np.linspace(start ** 0.5, end ** 0.5, n, dtype=np.float64) ** 2


def zero_terminal_snr(betas: np.ndarray) -> np.ndarray:
    """enforce_zero_terminal_snr — gaussian_diffusion.py:49-72 (done in float64 like torch.from_numpy(betas))."""
    abar_sqrt = np.sqrt(np.cumprod(1.0 - betas))
    first, last = abar_sqrt[0].copy(), abar_sqrt[-1].copy()
    abar_sqrt = (abar_sqrt - last) * (first / (first - last))
    abar = abar_sqrt ** 2
    alphas = np.concatenate([abar[0:1], abar[1:] / abar[:-1]])
    return 1.0 - alphas


def diffusion_betas(timesteps=1000, linear_start=0.00085, linear_end=0.0120, zero_snr=True) -> np.ndarray:
    """Diffusion.__init__ — gaussian_diffusion.py:75-110 with the val config (configs/val/val_terediff.yaml:87-94)."""
    b = linear_betas(timesteps, linear_start, linear_end)
    return zero_terminal_snr(b) if zero_snr else b


def space_timesteps(num_timesteps: int, count: int) -> List[int]:
    """spaced_sampler.py:14-65 for a single section: `count` steps with fractional stride, python round()."""
    if count <= 1:
        stride = 1.0
    else:
        stride = (num_timesteps - 1) / (count - 1)
    cur, out = 0.0, []
    for _ in range(count):
        out.append(round(cur))
        cur += stride
    return sorted(set(out))


def make_schedule(training_betas: np.ndarray, num_steps: int) -> Dict[str, np.ndarray]:
    """spaced_sampler.py:77-121 — float64 arithmetic, tables cast to fp32 (Sampler.register, sampler.py:26-29)."""
    abar_train = np.cumprod(1.0 - training_betas, axis=0)
    used = space_timesteps(len(training_betas), num_steps)
    betas, last = [], 1.0
    for i in used:
        betas.append(1 - abar_train[i] / last)
        last = abar_train[i]
    betas = np.array(betas, dtype=np.float64)
    alphas = 1.0 - betas
    abar = np.cumprod(alphas, axis=0)
    abar_prev = np.append(1.0, abar[:-1])
    with np.errstate(divide="ignore"):
        post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
        tables = dict(
            sqrt_alphas_cumprod=np.sqrt(abar),
            sqrt_one_minus_alphas_cumprod=np.sqrt(1 - abar),
            sqrt_recip_alphas_cumprod=np.sqrt(1.0 / abar),            # inf at the last index (abar = 0): unused for 'v'
            sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / abar - 1),
            posterior_variance=post_var,
            posterior_log_variance_clipped=np.log(np.append(post_var[1], post_var[1:])) if len(post_var) > 1
            else np.log(np.append(post_var[0], post_var[0])),
            posterior_mean_coef1=betas * np.sqrt(abar_prev) / (1.0 - abar),
            posterior_mean_coef2=(1.0 - abar_prev) * np.sqrt(alphas) / (1.0 - abar),
        )
    out = {k: v.astype(np.float32) for k, v in tables.items()}
    out["timesteps"] = np.array(used, dtype=np.int32)
    return out


def p_sample_update(tables: Dict[str, torch.Tensor], x: torch.Tensor, v: torch.Tensor, t: torch.Tensor,
                    noise: torch.Tensor, v_uncond: Optional[torch.Tensor] = None, cfg_scale: float = 1.0):
    """One reverse step given the network output(s): spaced_sampler.py:141-147,123-131,180-188 (+ the CFG combine of
    :161-163 applied to the v tensors).  tables: fp32 torch vectors; t int64 [B].  Returns (x_prev, pred_x0)."""
    def ex(name):
        return tables[name].gather(-1, t).reshape(-1, *([1] * (x.dim() - 1)))
    if v_uncond is not None:
        v = v_uncond + cfg_scale * (v - v_uncond)
    x0 = ex("sqrt_alphas_cumprod") * x - ex("sqrt_one_minus_alphas_cumprod") * v
    mean = ex("posterior_mean_coef1") * x0 + ex("posterior_mean_coef2") * x
    var = ex("posterior_variance")
    nonzero = (t != 0).float().reshape(-1, *([1] * (x.dim() - 1)))
    return mean + nonzero * torch.sqrt(var) * noise, x0


def tables_to_torch(sched: Dict[str, np.ndarray], device="cpu") -> Dict[str, torch.Tensor]:
    return {k: torch.tensor(v, dtype=torch.float32, device=device) for k, v in sched.items() if k != "timesteps"}


def sample_loop(model_fn, sched: Dict[str, np.ndarray], x_T: torch.Tensor, noises, cond_fn=None, trace=None):
    """val_sample / sample loop skeleton — spaced_sampler.py:270-296: model_t = original timestep, t = 49..0 index;
    ``noises[i]`` replaces torch.randn_like at loop iteration i (noise injection for parity).
    ``model_fn(x, model_t) -> (v, feats)``;  ``cond_fn(i, feats)`` is called after each step (TESTR/prompt feedback);
    ``trace`` (a list) receives the latent after every step."""
    tabs = tables_to_torch(sched, x_T.device)
    ts = np.flip(sched["timesteps"])
    total = len(ts)
    x = x_T
    for i, cur in enumerate(ts):
        B = x.shape[0]
        model_t = torch.full((B,), int(cur), device=x.device, dtype=torch.long)
        t = torch.full((B,), total - i - 1, device=x.device, dtype=torch.long)
        v, feats = model_fn(x, model_t)
        x, _ = p_sample_update(tabs, x, v, t, noises[i])
        if trace is not None:
            trace.append(x)
        if cond_fn is not None:
            cond_fn(i, feats)
    return x
