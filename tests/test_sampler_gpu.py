"""Parity of the denoising loop (tair_b200.sampler.SpacedSampler over tair_b200.model.ControlLDM) with the oracle
loop on identical seeded weights, start noise and injected per-step noise (SURVEY.md §8d parity gates ii / iii)."""
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STEP_TOL = 2e-2   # teacher-forced: max-abs error of x_{t-1} relative to max-abs of x_{t-1} (bf16 network, fp32 update)
PSNR_MIN = 40.0   # free-running 50 steps: PSNR of the decoded image against the oracle's, dB (north star)


def full_cfgs():
    u = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
             channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
             use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, legacy=False)
    c = dict(u)
    c.pop("out_channels")
    c["hint_channels"] = 4
    return u, c


@pytest.fixture(scope="module")
def setup(cuda_lib, manifests):
    from oracle import sampler as OS, weights
    from tair_b200.model import ControlLDM
    from tair_b200.model.gaussian_diffusion import val_diffusion
    from tair_b200.sampler import SpacedSampler
    usd = weights.seeded_state_dict(manifests["unet_full"])
    csd = weights.seeded_state_dict(manifests["controlnet_full"])
    m = ControlLDM(*full_cfgs())
    m.unet.load_state_dict(usd)
    m.controlnet.load_state_dict(csd)
    m = m.cuda().eval()
    usd = {k: v.cuda() for k, v in usd.items()}
    csd = {k: v.cuda() for k, v in csd.items()}
    sampler = SpacedSampler(val_diffusion().betas, "v", False)
    sched = OS.make_schedule(OS.diffusion_betas(), 50)
    return m, usd, csd, sampler, sched


def inputs(B, seed=0):
    from oracle import weights
    return (weights.seeded_randn((B, 4, 64, 64), 100 + seed).cuda(), weights.seeded_randn((B, 4, 64, 64), 200 + seed).cuda(),
            weights.seeded_randn((B, 77, 1024), 300 + seed).cuda())


def test_schedule_matches_oracle(setup):
    _, _, _, sampler, sched = setup
    sampler.make_schedule(50)
    assert np.array_equal(sampler.timesteps, sched["timesteps"])
    for k in ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_variance", "posterior_mean_coef1",
              "posterior_mean_coef2", "posterior_log_variance_clipped", "sqrt_recip_alphas_cumprod"):
        assert np.array_equal(getattr(sampler, k).cpu().numpy(), sched[k], equal_nan=True), k


@pytest.mark.parametrize("i", [0, 1, 25, 49])
def test_teacher_forced_step(setup, i):
    from oracle import sampler as OS, unet as OU, weights
    m, usd, csd, sampler, sched = setup
    sampler.make_schedule(50)
    sampler.to("cuda")
    tabs = OS.tables_to_torch(sched, "cuda")
    x, c_img, c_txt = inputs(2, seed=i)
    noise = weights.seeded_randn((2, 4, 64, 64), 400 + i).cuda()
    cur = int(sched["timesteps"][::-1][i])
    model_t = torch.full((2,), cur, device="cuda", dtype=torch.long)
    t = torch.full((2,), 49 - i, device="cuda", dtype=torch.long)
    with torch.no_grad():
        v, _ = OU.cldm_forward(usd, csd, x, model_t, c_txt, c_img)
        ref, _ = OS.p_sample_update(tabs, x, v, t, noise)
    out, feats = sampler.p_sample(m, x, model_t, t, dict(c_txt=c_txt, c_img=c_img), None, 1.0, noise=noise)
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    assert len(feats) == 4 and err < STEP_TOL, err


def test_cfg_step_matches_oracle(setup):
    """Classifier-free guidance (stacked cond/uncond batch, fused combine) vs two oracle forwards."""
    from oracle import sampler as OS, unet as OU, weights
    m, usd, csd, sampler, sched = setup
    sampler.make_schedule(50)
    sampler.to("cuda")
    tabs = OS.tables_to_torch(sched, "cuda")
    x, c_img, c_txt = inputs(1, seed=7)
    un_txt = weights.seeded_randn((1, 77, 1024), 777).cuda()
    noise = weights.seeded_randn((1, 4, 64, 64), 778).cuda()
    model_t = torch.full((1,), 500, device="cuda", dtype=torch.long)
    t = torch.full((1,), 25, device="cuda", dtype=torch.long)
    with torch.no_grad():
        vc, _ = OU.cldm_forward(usd, csd, x, model_t, c_txt, c_img)
        vu, _ = OU.cldm_forward(usd, csd, x, model_t, un_txt, c_img)
        ref, _ = OS.p_sample_update(tabs, x, vc, t, noise, v_uncond=vu, cfg_scale=4.0)
    out, _ = sampler.p_sample(m, x, model_t, t, dict(c_txt=c_txt, c_img=c_img), dict(c_txt=un_txt, c_img=c_img), 4.0,
                              noise=noise)
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 2 * STEP_TOL


def test_free_running_50_steps_latent_and_psnr(setup, manifests):
    """50 free-running steps with injected noise; CUDA-graph replay path; decoded-image PSNR against the oracle."""
    from oracle import sampler as OS, unet as OU, vae as OV, weights
    m, usd, csd, sampler, sched = setup
    B = 1
    x_T, c_img, c_txt = inputs(B, seed=11)
    noises = [weights.seeded_randn((B, 4, 64, 64), 1000 + i).cuda() for i in range(50)]
    with torch.no_grad():
        ref = OS.sample_loop(lambda x, mt: OU.cldm_forward(usd, csd, x, mt, c_txt, c_img), sched, x_T, noises)
    sampler.noise_fn = lambda i, x: noises[i]
    try:
        out, _ = sampler.sample(m, "cuda", 50, (B, 4, 64, 64), dict(c_txt=c_txt, c_img=c_img), None, 1.0, x_T=x_T,
                                progress=False, use_cuda_graph=True)
        eager, _ = sampler.sample(m, "cuda", 50, (B, 4, 64, 64), dict(c_txt=c_txt, c_img=c_img), None, 1.0, x_T=x_T,
                                  progress=False, use_cuda_graph=False)
    finally:
        sampler.noise_fn = None
    assert torch.equal(out, eager), "CUDA-graph replay must reproduce the eager loop bit for bit"
    lat_err = ((out - ref).abs().max() / ref.abs().max()).item()
    vsd = {k: v.cuda() for k, v in weights.seeded_state_dict(manifests["vae_decoder"]).items()}
    with torch.no_grad():
        img_ref, img = OV.latent_to_image(vsd, ref), OV.latent_to_image(vsd, out)
    p = OV.psnr(img, img_ref).min().item()
    print(f"free-running 50 steps: latent rel max-abs err {lat_err:.3e}, decoded PSNR {p:.1f} dB")
    assert p >= PSNR_MIN, f"PSNR {p:.1f} dB (latent err {lat_err:.3e})"


class HashClip:
    """Deterministic stand-in for FrozenOpenCLIPEmbedder.encode: prompt string -> (n,77,1024)."""

    def encode(self, prompts):
        if isinstance(prompts, str):
            prompts = [prompts]
        out = []
        for p in prompts:
            g = torch.Generator().manual_seed(zlib.crc32(p.encode()))
            out.append(torch.randn((77, 1024), generator=g))
        return torch.stack(out).cuda()


def test_val_sample_with_text_spotting_feedback(setup, manifests):
    """val_sample contract (spaced_sampler.py:246-328): TESTR runs on every step's features, prompts are rebuilt and
    re-encoded, cond['c_txt'] is mutated in place; 3 steps, B=2 (per-tile prompts)."""
    from types import SimpleNamespace
    from oracle import weights
    from tair_b200.testr import TransformerDetector, default_cfg
    m, _, _, sampler, sched = setup
    det = TransformerDetector(default_cfg("cuda"))
    det.load_state_dict(weights.seeded_state_dict(manifests["testr"]))
    det = det.cuda().eval()
    m.attach_clip(HashClip())
    B = 2
    x_T, c_img, _ = inputs(B, seed=21)
    cond = dict(c_txt=m.clip.encode([""] * B), c_img=c_img)
    first_ctx = cond["c_txt"].clone()
    cfg = SimpleNamespace(exp_args=SimpleNamespace(mode="VAL", prompt_style="CAPTION"))
    x, res = sampler.val_sample(m, "cuda", 3, (B, 4, 64, 64), cond, None, 1.0, x_T=x_T, progress=False, cfg=cfg,
                                pure_cldm=m, ts_model=det)
    assert x.shape == (B, 4, 64, 64) and torch.isfinite(x).all()
    assert len(res) == 3 and [r["timestep"] for r in res] == [999, 500, 0]
    for r in res:
        assert set(r) >= {"timestep", "pred_texts", "pred_prompt", "pred_polys"}
        assert r["pred_prompt"].startswith("A realistic scene where the texts ")
        assert len(r["pred_texts"]) == len(r["pred_polys"]) and len(r["batch_prompts"]) == B
        assert all(isinstance(s, str) and len(s) <= 25 for s in r["pred_texts"])
        assert all(p.shape == (16, 2) and p.dtype == np.int32 for p in r["pred_polys"])
    assert cond["c_txt"].shape == (B, 77, 1024) and not torch.equal(cond["c_txt"], first_ctx)


def test_val_sample_with_kernel_clip_encoder(setup, manifests):
    """Same loop with the real text encoder on the kernels (tair_b200.model.clip) in place of the stand-in: the prompt of
    every tile is tokenised (hash tokenizer here — the CLIP merge table is not shipped), run through the 23 causal
    blocks, memoised by prompt string, and conditions the next step."""
    from types import SimpleNamespace
    from oracle import clip as OC, weights
    from tair_b200 import ops
    from tair_b200.model.clip import FrozenOpenCLIPEmbedder
    from tair_b200.testr import TransformerDetector, default_cfg
    m, _, _, sampler, sched = setup
    det = TransformerDetector(default_cfg("cuda"))
    det.load_state_dict(weights.seeded_state_dict(manifests["testr"]))
    det = det.cuda().eval()
    csd = weights.seeded_state_dict(manifests["clip_text"])
    clip = FrozenOpenCLIPEmbedder(1024, None, dict(context_length=77, vocab_size=49408, width=1024, heads=16, layers=24),
                                  layer="penultimate")
    clip.load_state_dict(csd)
    clip = clip.cuda().eval()

    def hash_tokenizer(texts):
        out = torch.zeros((len(texts), 77), dtype=torch.long)
        for i, s in enumerate(texts):
            ids = [49406] + [zlib.crc32(w.encode()) % 49000 for w in s.split()][:75] + [49407]
            out[i, :len(ids)] = torch.tensor(ids)
        return out
    clip.attach_tokenizer(hash_tokenizer)
    old = m.clip
    m.attach_clip(clip)
    try:
        B = 2
        x_T, c_img, _ = inputs(B, seed=22)
        cond = dict(c_txt=clip.encode([""] * B), c_img=c_img)
        assert torch.equal(cond["c_txt"][0], cond["c_txt"][1])
        cfg = SimpleNamespace(exp_args=SimpleNamespace(mode="VAL", prompt_style="CAPTION"))
        x, res = sampler.val_sample(m, "cuda", 3, (B, 4, 64, 64), cond, None, 1.0, x_T=x_T, progress=False, cfg=cfg,
                                    pure_cldm=m, ts_model=det, use_cuda_graph=True)
        assert torch.isfinite(x).all() and len(res) == 3
        prompts = res[-1]["batch_prompts"]
        with torch.no_grad():
            ref = OC.encode_tokens({k: v.cuda() for k, v in csd.items()}, hash_tokenizer(prompts).cuda())
        err = ((cond["c_txt"] - ref).abs().max() / ref.abs().max()).item()
        assert err < 4e-2, err
        # memoised: encoding prompts that were already seen launches no kernel
        ops.reset_launch_count()
        again = clip.encode(prompts)
        assert ops.launch_count() == 0 and torch.equal(again, cond["c_txt"])
    finally:
        m.attach_clip(old)
