"""Pin the oracle (oracle/*.py) against fixtures produced by the imported reference (tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import msda, sampler, tiles, unet, weights


def test_schedule_tables_bit_equal(golden):
    g = golden("schedule_50.npz")
    s = sampler.make_schedule(sampler.diffusion_betas(), 50)
    assert np.array_equal(s["timesteps"], g["timesteps"])
    assert list(s["timesteps"][:5]) == [0, 20, 41, 61, 82] and s["timesteps"][-1] == 999
    for k in s:
        assert np.array_equal(s[k], g[k], equal_nan=True), k
    # fixed points noted in SURVEY.md §8c
    assert s["posterior_mean_coef1"][0] == 1.0 and s["posterior_mean_coef2"][0] == 0.0
    assert s["posterior_variance"][0] == 0.0 and s["sqrt_alphas_cumprod"][49] == 0.0
    assert np.isinf(s["sqrt_recip_alphas_cumprod"][49])


def test_p_sample_update_bit_equal(golden):
    g = golden("schedule_50.npz")
    tabs = sampler.tables_to_torch(sampler.make_schedule(sampler.diffusion_betas(), 50))
    x, v, noise = (weights.seeded_randn((2, 4, 8, 8), s) for s in (11, 12, 13))
    for idx in (49, 7, 0):
        t = torch.full((2,), idx, dtype=torch.long)
        xp, _ = sampler.p_sample_update(tabs, x, v, t, noise)
        assert np.array_equal(xp.numpy(), g[f"x_prev_t{idx}"]), idx


def test_unet_controlnet_narrow_matches_reference_fixture(golden, manifests):
    g = golden("unet_narrow.npz")
    usd = weights.seeded_state_dict(manifests["unet_narrow"])
    csd = weights.seeded_state_dict(manifests["controlnet_narrow"])
    x, hint, ctx = weights.seeded_randn((2, 4, 32, 32), 1), weights.seeded_randn((2, 4, 32, 32), 2), \
        weights.seeded_randn((2, 77, 128), 3)
    t = torch.from_numpy(g["t"])
    with torch.no_grad():
        ctrl = unet.controlnet_forward(csd, x, hint, t, ctx)
        out, feats = unet.unet_forward(usd, x, t, ctx, ctrl)
    assert np.abs(out.numpy() - g["out"]).max() < 2e-4
    assert np.abs(ctrl[0][:, ::4, ::2, ::2].numpy() - g["ctrl0"]).max() < 2e-4
    assert np.abs(ctrl[12].numpy()[:, ::4] - g["ctrl12"]).max() < 2e-4
    for i, f in enumerate(feats):
        assert np.abs(f[:, ::4, ::2, ::2].numpy() - g[f"feat{i}"]).max() < 5e-4, i


def test_msda_core_matches_reference_fixture(golden):
    g = golden("msda_case.npz")
    shapes = [tuple(int(v) for v in r) for r in g["shapes"]]
    S = sum(h * w for h, w in shapes)
    B, M, D, Lq, L, P = 2, 8, 32, 24, 3, 4
    value = weights.seeded_randn((B, S, M, D), 21)
    loc = torch.rand((B, Lq, M, L, P, 2), generator=torch.Generator().manual_seed(22)) * 1.4 - 0.2
    w = torch.softmax(weights.seeded_randn((B, Lq, M, L * P), 23), -1).view(B, Lq, M, L, P)
    for fn in (msda.msda_core, msda.msda_core_grid_sample):
        assert np.abs(fn(value, shapes, loc, w).numpy() - g["out"]).max() < 1e-5


def test_msda_edge_cases_agree():
    # samples exactly on / outside the borders, single-pixel level
    shapes = [(1, 1), (2, 5)]
    S = 11
    value = weights.seeded_randn((1, S, 2, 8), 5)
    loc = torch.tensor([0.0, 1.0, -0.4, 1.4, 0.5, 0.999, 0.25, 0.75]).view(1, 1, 1, 1, 4, 2).expand(1, 3, 2, 2, 4, 2).contiguous()
    w = torch.full((1, 3, 2, 2, 4), 0.125)
    a, b = msda.msda_core(value, shapes, loc, w), msda.msda_core_grid_sample(value, shapes, loc, w)
    assert torch.allclose(a, b, atol=1e-6)


def test_split_and_merge_match_reference_fixture(golden):
    g = golden("merge_case.npz")
    for name in ("a", "b", "c"):
        oh, ow, n = (int(v) for v in g[f"{name}_size"])
        img = np.zeros((oh, ow, 3), np.uint8)
        assert len(tiles.split_image(img)) == n
        gen = torch.Generator().manual_seed(31)
        tl = [torch.rand((1, 3, 512, 512), generator=gen) for _ in range(n)]
        m = tiles.merge_tiles(tl, (oh, ow))
        assert m.shape == (1, 3, 4 * oh, 4 * ow)
        assert np.array_equal(m[0, :, ::37, ::41].numpy(), g[f"{name}_sample"])
        assert m.double().sum().item() == float(g[f"{name}_sum"])


def test_tile_grid_counts():
    # shipped example 500x881 -> 5 x 8 = 40 tiles; 512x512 -> 5 x 5; 2160x3840 -> 20 x 35
    assert tiles.tile_grid(500, 881)[:2] == (5, 8)
    assert tiles.tile_grid(512, 512)[:2] == (5, 5)
    assert tiles.tile_grid(2160, 3840)[:2] == (20, 35)
    assert tiles.tile_grid(128, 128) == (1, 1, 128, 128)


def test_split_pads_with_zeros_and_orders_row_major():
    img = (np.arange(130 * 250 * 3) % 251).astype(np.uint8).reshape(130, 250, 3)
    ps = tiles.split_image(img)
    assert len(ps) == 2 * 3 and all(p.shape == (128, 128, 3) for p in ps)
    assert np.array_equal(ps[0], img[:128, :128])
    assert np.array_equal(ps[1][:, :, 0], img[:128, 112:240, 0])
    assert ps[5][18:, :, :].sum() == 0  # rows below the image are zero padding


def test_testr_head_matches_reference_fixture(golden, manifests):
    from oracle import testr as OT
    g = golden("testr_full.npz")
    sd = weights.seeded_state_dict(manifests["testr"])
    feats = [weights.seeded_randn(s, i) for s, i in (((1, 1280, 16, 16), 41), ((1, 1280, 32, 32), 42),
                                                      ((1, 640, 64, 64), 43), ((1, 320, 64, 64), 44))]
    with torch.no_grad():
        out = OT.testr_forward(sd, feats)
    assert np.abs(out["pred_logits"].numpy() - g["pred_logits"]).max() < 1e-4
    assert np.abs(out["pred_ctrl_points"].numpy() - g["pred_ctrl_points"]).max() < 1e-4
    assert np.abs(out["pred_texts"].numpy() - g["pred_texts"]).max() < 1e-4
    assert np.abs(out["enc_logits"].numpy()[:, ::8] - g["enc_logits"]).max() < 1e-4
    r = OT.inference(out)[0]
    assert len(r["scores"]) == int(g["n_inst"]) > 0
    assert np.array_equal(r["recs"].numpy(), g["recs"])
    assert np.abs(r["polygons"].numpy() - g["polygons"]).max() < 1e-2


def test_vae_decoder_matches_reference_fixture(golden, manifests):
    from oracle import vae
    g = golden("vae_decode.npz")
    sd = weights.seeded_state_dict(manifests["vae_decoder"])
    with torch.no_grad():
        img = vae.vae_decode(sd, weights.seeded_randn((1, 4, 16, 16), 51))
    assert img.shape == (1, 3, 128, 128)
    assert np.abs(img.numpy()[:, :, ::2, ::2] - g["img"]).max() < 1e-3 * max(1.0, np.abs(g["img"]).max())
    a = torch.rand(2, 3, 8, 8)
    assert torch.allclose(vae.psnr(a, a), torch.full((2,), 80.0, dtype=torch.float64))
    full = weights.seeded_state_dict(manifests["vae"])
    with torch.no_grad():
        mom = vae.vae_encode_moments(full, weights.seeded_randn((1, 3, 64, 64), 52).clamp(-1, 1))
    assert mom.shape == (1, 8, 8, 8)
    assert np.abs(mom.numpy() - g["moments"]).max() < 1e-3 * max(1.0, np.abs(g["moments"]).max())


def test_clip_text_encoder_matches_reference_fixture(golden, manifests):
    from oracle import clip as OC
    g = golden("clip_text.npz")
    sd = weights.seeded_state_dict(manifests["clip_text"])
    with torch.no_grad():
        z = OC.encode_tokens(sd, torch.from_numpy(g["tokens"]))
    assert z.shape == (2, 77, 1024)
    assert np.abs(z.numpy()[:, ::4, ::2] - g["z"]).max() < 2e-3


def test_swinir_matches_reference_fixture(golden, manifests):
    from oracle import swinir as OS
    g = golden("swinir.npz")
    sd = _swinir_state(manifests)
    with torch.no_grad():
        y = OS.swinir_forward(sd, torch.from_numpy(g["x"]))
    assert y.shape == (1, 3, 128, 128)
    assert np.abs(y.numpy() - g["y"]).max() < 2e-4


def _swinir_state(manifests):
    return weights.seeded_state_dict(manifests["swinir"])
