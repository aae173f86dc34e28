"""Spaced DDPM ancestral sampler driving the sm_100a denoising step.

Drop-in surface of terediff/sampler/sampler.py:10-38 (``Sampler``) and spaced_sampler.py:67-328
(``SpacedSampler``): ``make_schedule``, ``apply_model``, ``p_sample``, ``sample``, ``val_sample`` keep their
argument lists and return values.  Differences, all behaviour-preserving on the reference's working paths:

* the per-step arithmetic (x0 from v, posterior mean/variance, noise injection, optional CFG combine) is ONE
  kernel (``tair_sampler_update``) instead of ~9 elementwise launches, bit-identical in fp32;
* classifier-free guidance actually works: the reference's CFG branch does arithmetic on ``(eps, feats)`` tuples
  and raises (spaced_sampler.py:161-163, SURVEY.md §8a hazard 6).  Here cond/uncond are evaluated as one
  stacked batch, v = v_u + s (v_c - v_u), features are taken from the cond half;
* ``val_sample`` accepts B > 1 tiles: the reference builds the prompt from ``results[0]`` only (:298) and therefore
  is batch-1 in practice; here every tile gets its own prompt (for B == 1 this is the reference behaviour);
* the whole step (ControlNet + UNet + update [+ TESTR head]) can be captured once in a CUDA graph and replayed
  (``use_cuda_graph``), removing the several thousand eager launches per step of the reference;
* step noise comes from ``noise_fn(i, x)`` when given (parity tests inject the oracle's noise), else randn_like.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from .. import ops

_TABLES = ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
           "posterior_variance")
_TABLES_EPS = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod") + _TABLES[2:]


def space_timesteps(num_timesteps: int, section_counts) -> set:
    """Evenly spaced subset of the training timesteps (spaced_sampler.py:14-65; 'ddimN' striding included)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                picked = range(0, num_timesteps, stride)
                if len(picked) == want:
                    return set(picked)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(v) for v in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    steps, offset = [], 0
    for k, count in enumerate(section_counts):
        size = base + (1 if k < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            steps.append(offset + round(pos))
            pos += stride
        offset += size
    return set(steps)


class Sampler(nn.Module):
    """sampler.py:10-38."""

    def __init__(self, betas: np.ndarray, parameterization: str, rescale_cfg: bool):
        super().__init__()
        self.num_timesteps = len(betas)
        self.training_betas = betas
        self.training_alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.context = {}
        self.parameterization = parameterization
        self.rescale_cfg = rescale_cfg

    def register(self, name: str, value: np.ndarray, dtype: torch.dtype = torch.float32) -> None:
        new = torch.tensor(value, dtype=dtype)
        old = getattr(self, name, None)
        if isinstance(old, torch.Tensor) and old.shape == new.shape and old.dtype == new.dtype:
            old.copy_(new)  # keep the device address stable: captured CUDA graphs read these tables
        else:
            # a new allocation: graphs captured against the old tables would read freed memory
            getattr(self, "_graphs", {}).clear()
            self.register_buffer(name, new)

    def get_cfg_scale(self, default_cfg_scale: float, model_t: int) -> float:
        if self.rescale_cfg and default_cfg_scale > 1:
            return 1 + default_cfg_scale * ((1 - math.cos(math.pi * ((1000 - model_t) / 1000) ** 5.0)) / 2)
        return default_cfg_scale


class SpacedSampler(Sampler):
    def __init__(self, betas: np.ndarray, parameterization: str = "v", rescale_cfg: bool = False):
        super().__init__(betas, parameterization, rescale_cfg)
        if parameterization not in ("v", "eps"):
            raise ValueError(f"unknown parameterization {parameterization!r} (spaced_sampler.py:176-179 knows 'eps' and 'v')")
        self.noise_fn: Optional[Callable[[int, torch.Tensor], torch.Tensor]] = None
        self.trace: Optional[list] = None   # parity tests set a list: one record per step (latent, head outputs)
        self._graphs: Dict[tuple, "_StepGraph"] = {}
        self._warm: Dict[torch.device, "torch.cuda.Stream"] = {}

    def _warm_stream(self, device) -> "torch.cuda.Stream":
        """One warm-up stream per device for every graph capture: the per-stream GroupNorm / split-K workspaces of
        ``ops`` are keyed by stream handle, so a fresh stream per capture would leak one workspace each."""
        dev = torch.device(device)
        s = self._warm.get(dev)
        if s is None:
            s = self._warm[dev] = torch.cuda.Stream(device=dev)
        return s

    # ---- schedule (host, float64; spaced_sampler.py:77-121) -------------------------------------------------
    def make_schedule(self, num_steps: int) -> None:
        keep = sorted(space_timesteps(self.num_timesteps, str(num_steps)))
        abar_kept = self.training_alphas_cumprod[keep]
        prev_kept = np.concatenate([[1.0], abar_kept[:-1]])
        betas = 1 - abar_kept / prev_kept
        self.timesteps = np.array(keep, dtype=np.int32)
        alphas = 1.0 - betas
        abar = np.cumprod(alphas, axis=0)
        abar_prev = np.append(1.0, abar[:-1])
        with np.errstate(divide="ignore"):
            var = betas * (1.0 - abar_prev) / (1.0 - abar)
            self.register("sqrt_alphas_cumprod", np.sqrt(abar))
            self.register("sqrt_one_minus_alphas_cumprod", np.sqrt(1 - abar))
            self.register("sqrt_recip_alphas_cumprod", np.sqrt(1.0 / abar))        # inf at the last step: never used for 'v'
            self.register("sqrt_recipm1_alphas_cumprod", np.sqrt(1.0 / abar - 1))
            self.register("posterior_variance", var)
            tail = var[1] if len(var) > 1 else var[0]
            self.register("posterior_log_variance_clipped", np.log(np.append(tail, var[1:] if len(var) > 1 else var[:1])))
            self.register("posterior_mean_coef1", betas * np.sqrt(abar_prev) / (1.0 - abar))
            self.register("posterior_mean_coef2", (1.0 - abar_prev) * np.sqrt(alphas) / (1.0 - abar))

    def _tables(self) -> List[torch.Tensor]:
        """x0 = A[t]*x - B[t]*model_output for both parameterisations (spaced_sampler.py:133-147): 'v' uses
        (sqrt_ac, sqrt_1mac), 'eps' uses (sqrt_recip_ac, sqrt_recipm1_ac) — the latter are inf at the last index under
        zero terminal SNR exactly as in the reference (SURVEY.md §8a hazard 5)."""
        names = _TABLES if self.parameterization == "v" else _TABLES_EPS
        return [getattr(self, n) for n in names]

    def _graph_key(self, *parts) -> tuple:
        return parts + (len(self.timesteps), self.parameterization) + tuple(t.data_ptr() for t in self._tables())

    # ---- one step ------------------------------------------------------------------------------------------
    def apply_model(self, model, x, model_t, cond, uncond, cfg_scale):
        """-> (v_cond, v_uncond | None, feats).  cfg == 1 or no uncond: single forward (spaced_sampler.py:158-159)."""
        if uncond is None or cfg_scale == 1.0:
            v, feats = model(x, model_t, cond)
            return v, None, feats
        B = x.shape[0]
        both = {k: torch.cat([cond[k], uncond[k]], 0) for k in cond}
        v, feats = model(torch.cat([x, x], 0), torch.cat([model_t, model_t], 0), both)
        return v[:B].contiguous(), v[B:].contiguous(), None if feats is None else [f[:B] for f in feats]

    @torch.no_grad()
    def p_sample(self, model, x, model_t, t, cond, uncond, cfg_scale, noise: Optional[torch.Tensor] = None,
                 cfg_scale_dev: Optional[torch.Tensor] = None):
        """spaced_sampler.py:167-189.  ``cfg_scale_dev`` (1-element fp32 CUDA tensor) overrides the host scale inside the
        fused update, so one captured graph serves every guidance scale of a ``rescale_cfg`` run."""
        v, v_u, feats = self.apply_model(model, x, model_t, cond, uncond, cfg_scale)
        if noise is None:
            noise = torch.randn_like(x)
        x_prev = ops.sampler_update(x.contiguous(), v.contiguous(), noise.contiguous(), t, self._tables(),
                                    v_uncond=v_u, cfg_scale=float(cfg_scale), cfg_scale_dev=cfg_scale_dev)
        return x_prev, feats

    def _noise(self, i: int, x: torch.Tensor) -> torch.Tensor:
        return self.noise_fn(i, x) if self.noise_fn is not None else torch.randn_like(x)

    def _stepper(self, kind: str, model, x, cond, uncond, cfg_scale, head=None, extra=()):
        """Cached ``_StepGraph`` for this (model, shapes, schedule tables).  The key holds the table addresses and the
        step count, and the entry keeps a weak reference to the model: a graph is never replayed against re-allocated
        tables or against a different model that happens to reuse a freed ``id()``."""
        guided = uncond is not None and cfg_scale != 1.0
        key = self._graph_key(kind, id(model), id(head), tuple(x.shape), guided, extra,
                              tuple(sorted((k, tuple(v.shape)) for k, v in cond.items())))
        st = self._graphs.get(key)
        if st is not None and (st.model_ref() is not model or (head is not None and st.head_ref() is not head)):
            st = None
        if st is None:
            st = self._graphs[key] = _StepGraph(self, model, x, cond, uncond if guided else None, head=head)
        return st

    # ---- loops ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, model, device, steps, x_size, cond, uncond, cfg_scale, tiled=False, tile_size=-1, tile_stride=-1,
               x_T=None, progress=True, cfg=None, use_cuda_graph: bool = False):
        """spaced_sampler.py:192-243 -> (x, sampled_unet_feats).  ``tiled=True`` wraps the model in the latent-tiling
        function of the legacy pipeline (terediff/utils/common.py:125-234, used by pipeline.py:150-165)."""
        self.make_schedule(steps)
        self.to(device)
        x = torch.randn(x_size, device=device, dtype=torch.float32) if x_T is None else x_T.to(device).float()
        if tiled:
            from ..legacy import tiled_model
            model = tiled_model(model, tile_size, tile_stride)
            use_cuda_graph = False
        order = np.flip(self.timesteps)
        total = len(order)
        bs = x_size[0]
        want = [] if cfg is None else list(cfg.exp_args["unet_feat_sampling_timestep"])
        kept = []
        # The decoder features of a step are only returned for the steps listed in cfg (none by default): let the model
        # hand them over channels-last bf16 and convert the few that are kept, instead of four NHWC->NCHW fp32
        # transposes on every step (1 % of the step).
        lazy = getattr(model, "return_nhwc_feats", None) is False
        if lazy:
            model.return_nhwc_feats = True
        to_nchw = (lambda fs: [ops.nhwc_to_nchw(f) for f in fs]) if lazy else (lambda fs: fs)
        try:
            return self._sample_loop(model, device, x, order, total, bs, cond, uncond, cfg_scale, want, kept, to_nchw,
                                     use_cuda_graph, lazy)
        finally:
            if lazy:
                model.return_nhwc_feats = False

    def _sample_loop(self, model, device, x, order, total, bs, cond, uncond, cfg_scale, want, kept, to_nchw,
                     use_cuda_graph, lazy):
        stepper = None
        if use_cuda_graph:
            stepper = self._stepper("sample", model, x, cond, uncond, cfg_scale, extra=(lazy,))
            stepper.load_cond(cond, uncond)
        for i, cur in enumerate(order):
            cur = int(cur)
            scale = self.get_cfg_scale(cfg_scale, cur)
            if stepper is not None:
                x, feats = stepper.run(x, cur, total - i - 1, self._noise(i, x), scale)
            else:
                model_t = torch.full((bs,), cur, device=device, dtype=torch.long)
                t = torch.full((bs,), total - i - 1, device=device, dtype=torch.long)
                x, feats = self.p_sample(model, x, model_t, t, cond, uncond, scale, noise=self._noise(i, x))
            if self.trace is not None:
                self.trace.append(dict(i=i, timestep=cur, x=x.clone()))
            if i + 1 in want:
                kept.append((i + 1, cur, to_nchw([f.clone() for f in feats] if stepper is not None else feats)))
        return x, kept

    @torch.no_grad()
    def val_sample(self, model, device, steps, x_size, cond, uncond, cfg_scale, tiled=False, tile_size=-1,
                   tile_stride=-1, x_T=None, progress=True, cfg=None, pure_cldm=None, ts_model=None, val_prompt=None,
                   use_cuda_graph: bool = False):
        """spaced_sampler.py:246-328 -> (x, ts_results): every step runs the text-spotting head on the step's decoder
        features, decodes the recognised strings and re-encodes the prompt that conditions the NEXT step."""
        from ..prompt import build_prompt, decode_texts
        assert ts_model is not None, "Text-spotting model must be provided for validation sampling."
        self.make_schedule(steps)
        self.to(device)
        x = torch.randn(x_size, device=device, dtype=torch.float32) if x_T is None else x_T.to(device).float()
        order = np.flip(self.timesteps)
        total = len(order)
        bs = x_size[0]
        mode = cfg.exp_args.mode if cfg is not None else "VAL"
        style = cfg.exp_args.prompt_style if cfg is not None else "CAPTION"
        ts_results = []
        ours = hasattr(ts_model, "testr") and hasattr(ts_model, "inference")   # tair_b200.testr.TransformerDetector
        # our head reads channels-last bf16 features directly: skip the NHWC->NCHW fp32->NHWC bf16 round trip per step
        lazy = ours and getattr(model, "return_nhwc_feats", None) is False
        if lazy:
            model.return_nhwc_feats = True
        try:
            stepper = None
            if use_cuda_graph and ours:
                # one graph per step: ControlNet + UNet + sampler update + the dense part of the text-spotting head;
                # only the data-dependent post-processing (thresholding, string decode, prompt, CLIP) stays outside
                stepper = self._stepper("val", model, x, cond, uncond, cfg_scale, head=ts_model.testr, extra=(lazy,))
            for i, cur in enumerate(order):
                cur = int(cur)
                scale = self.get_cfg_scale(cfg_scale, cur)
                dense = None
                if stepper is not None:
                    stepper.load_cond(cond, uncond)
                    x, feats, dense = stepper.run(x, cur, total - i - 1, self._noise(i, x), scale)
                else:
                    model_t = torch.full((bs,), cur, device=device, dtype=torch.long)
                    t = torch.full((bs,), total - i - 1, device=device, dtype=torch.long)
                    x, feats = self.p_sample(model, x, model_t, t, cond, uncond, scale, noise=self._noise(i, x))
                    if ours:
                        dense = ts_model.testr(feats)
                results = None
                if dense is not None and self.trace is None and hasattr(ts_model, "detect_host"):
                    texts, polys = ts_model.detect_host(dense, (512, 512))     # one kernel + one D2H for the whole batch
                else:
                    if dense is not None:
                        results = ts_model.inference(dense["pred_logits"], dense["pred_ctrl_points"], dense["pred_texts"],
                                                     [(512, 512)] * bs)
                    else:
                        _, results = ts_model(feats, None, mode)
                    texts, polys = decode_texts(results)
                prompts = [build_prompt(tx, style) for tx in texts]
                cond["c_txt"] = pure_cldm.clip.encode(prompts if bs > 1 else prompts[0])   # mutated in place like :317
                ts_results.append(dict(timestep=cur, pred_texts=texts[0], pred_prompt=prompts[0], pred_polys=polys[0],
                                       batch_texts=texts, batch_prompts=prompts))
                if self.trace is not None:
                    rec = dict(i=i, timestep=cur, x=x.clone(), results=results, c_txt=cond["c_txt"].clone())
                    if dense is not None:
                        rec.update(pred_logits=dense["pred_logits"].clone(), pred_ctrl_points=dense["pred_ctrl_points"].clone(),
                                   pred_texts=dense["pred_texts"].clone(),
                                   topk=dense["enc_outputs"]["topk_indices"].clone(),
                                   enc_logits=dense["enc_outputs"]["pred_logits"][..., 0].clone())
                    self.trace.append(rec)
            return x, ts_results
        finally:
            if lazy:
                model.return_nhwc_feats = False


class _StepGraph:
    """One denoising step (model forward(s) + sampler update [+ text-spotting head]) captured in a CUDA graph and replayed
    per step.  Static inputs: x, model_t, t, noise, cond tensors and the guidance scale (a device scalar read by the
    fused update), so ONE graph serves every step and every scale of a run."""

    def __init__(self, sampler: SpacedSampler, model, x, cond, uncond, head=None):
        import weakref
        self.s, self.model, self.head = sampler, model, head
        self.model_ref = weakref.ref(model)
        self.head_ref = weakref.ref(head) if head is not None else (lambda: None)
        self.x = x.clone()
        B = x.shape[0]
        self.model_t = torch.zeros((B,), device=x.device, dtype=torch.long)
        self.t = torch.zeros((B,), device=x.device, dtype=torch.long)
        self.noise = torch.zeros_like(x)
        self.scale = torch.ones((1,), device=x.device, dtype=torch.float32)
        self.cond = {k: v.clone() for k, v in cond.items()}
        self.uncond = None if uncond is None else {k: v.clone() for k, v in uncond.items()}
        self.graph = None

    def _step(self):
        # a host scale != 1 selects the stacked cond/uncond forward; its value is irrelevant (the kernel reads self.scale)
        host_scale = 1.0 if self.uncond is None else 2.0
        out, feats = self.s.p_sample(self.model, self.x, self.model_t, self.t, self.cond, self.uncond, host_scale,
                                     noise=self.noise, cfg_scale_dev=self.scale)
        dense = self.head(feats) if self.head is not None else None
        return out, feats, dense

    def _capture(self):
        side = self.s._warm_stream(self.x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture: weight packing, smem attributes, autotuner, allocator
            for _ in range(2):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        # capture on the SAME stream the warm-up ran on: the per-stream GroupNorm / split-K workspaces of ``ops`` were
        # allocated there outside the capture, so no workspace ends up inside one graph's private memory pool
        with torch.cuda.graph(g, stream=side):
            out, feats, dense = self._step()
        self.graph = (g, out, feats, dense)

    def load_cond(self, cond, uncond):
        for k, v in cond.items():
            self.cond[k].copy_(v)
        if uncond is not None and self.uncond is not None:
            for k, v in uncond.items():
                self.uncond[k].copy_(v)

    def run(self, x, model_t: int, t: int, noise, scale: float):
        if self.graph is None:
            self._capture()
        g, out, feats, dense = self.graph
        self.x.copy_(x)
        self.model_t.fill_(model_t)
        self.t.fill_(t)
        self.noise.copy_(noise)
        self.scale.fill_(float(scale))
        g.replay()
        if self.head is not None:
            return out.clone(), feats, dense
        return out.clone(), feats
