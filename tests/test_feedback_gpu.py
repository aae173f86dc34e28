"""Numeric parity of the text-spotting feedback loop of ``SpacedSampler.val_sample`` (spaced_sampler.py:246-328):
TESTR on every step's decoder features -> detections -> strings -> prompt -> text embedding -> next step's context.

Three layers of evidence:
  1. the fp32 oracle loop reproduces the fixture of the UNMODIFIED reference ``val_sample`` (strings, prompts, polygons,
     latents) — on the CPU in tests/test_feedback_oracle.py and here again on the GPU;
  2. the product loop (bf16 kernels, CUDA-graph step) against the oracle loop, DECISION-ALIGNED: the loop is discontinuous
     in the network outputs (top-100 proposals, 0.5 score threshold, arg-max characters), so every discrete decision of the
     product is first checked against the oracle's margins — it must be the oracle's own choice or lie within the stated
     bf16 distance of the oracle's decision boundary — and the oracle then continues with the product's choice, so that
     both follow one trajectory on which every continuous quantity (latent after each step, dense head outputs,
     polygons) and every host-side product (strings, prompt, embedding) is compared;
  3. batch independence: tile 0 of a 2-tile run equals the 1-tile run bit for bit (per-tile prompts included).
"""
import json
import os

import numpy as np
import pytest
import torch

from test_feedback_oracle import check_against_fixture, load_fixture, run_oracle

pytestmark = pytest.mark.gpu

LATENT_TOL = 3e-2      # max-abs error of the latent after each step relative to its max-abs (bf16 network, 3 steps)
ENC_TOL = 5e-2         # encoder proposal logits, relative to their max-abs (same bound as test_testr_gpu.py)
SCORE_TOL = 0.04       # |oracle detection probability - 0.5| below which a differing keep decision is admissible
TEXT_TOL = 8e-2        # character logits, relative to their max-abs (same bound as test_testr_gpu.py)
POLY_TOL_PX = 0.02 * 512 + 1.0


def full_cfgs():
    u = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
             channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
             use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, legacy=False)
    c = dict(u)
    c.pop("out_channels")
    c["hint_channels"] = 4
    return u, c


@pytest.fixture(scope="module")
def product(cuda_lib, manifests):
    from oracle import val_loop as VL, weights
    from tair_b200.model import ControlLDM
    from tair_b200.model.gaussian_diffusion import val_diffusion
    from tair_b200.sampler import SpacedSampler
    from tair_b200.testr import TransformerDetector, default_cfg
    m = ControlLDM(*full_cfgs())
    m.unet.load_state_dict(weights.seeded_state_dict(manifests["unet_full"]))
    m.controlnet.load_state_dict(weights.seeded_state_dict(manifests["controlnet_full"]))
    m = m.cuda().eval()
    m.attach_clip(VL.HashClip("cuda"))
    det = TransformerDetector(default_cfg("cuda"))
    det.load_state_dict(weights.seeded_state_dict(manifests["testr"]))
    det = det.cuda().eval()
    return m, det, SpacedSampler(val_diffusion().betas, "v", False)


def run_product(product, B, graph):
    from types import SimpleNamespace
    from oracle import weights
    m, det, sampler = product
    steps = load_fixture()[0]["steps"]
    x_T = weights.seeded_randn((1, 4, 64, 64), 9800).cuda()
    c_img = weights.seeded_randn((1, 4, 64, 64), 9900).cuda()
    noises = [weights.seeded_randn((1, 4, 64, 64), 10000 + i).cuda() for i in range(steps)]
    if B > 1:   # tile 0 is the fixture's tile, the others are different tiles
        x_T = torch.cat([x_T] + [weights.seeded_randn((1, 4, 64, 64), 9810 + b).cuda() for b in range(1, B)])
        c_img = torch.cat([c_img] + [weights.seeded_randn((1, 4, 64, 64), 9910 + b).cuda() for b in range(1, B)])
        noises = [torch.cat([n] + [weights.seeded_randn((1, 4, 64, 64), 10100 + 10 * b + i).cuda() for b in range(1, B)])
                  for i, n in enumerate(noises)]
    cond = dict(c_txt=m.clip.encode([""] * B), c_img=c_img)
    cfg = SimpleNamespace(exp_args=SimpleNamespace(mode="VAL", prompt_style="CAPTION"))
    sampler.noise_fn = lambda i, x: noises[i]
    sampler.trace = []
    try:
        x, res = sampler.val_sample(m, "cuda", steps, (B, 4, 64, 64), cond, None, 1.0, x_T=x_T, progress=False, cfg=cfg,
                                    pure_cldm=m, ts_model=det, use_cuda_graph=graph)
        return x, res, sampler.trace, cond
    finally:
        sampler.noise_fn, sampler.trace = None, None


def test_oracle_feedback_loop_on_gpu_matches_reference_fixture(cuda_lib, manifests):
    x, res, trace = run_oracle(manifests, "cuda")
    check_against_fixture(x, res, trace, 2e-3)


def test_product_feedback_loop_decision_aligned_with_oracle(product, manifests):
    from oracle import val_loop as VL
    x_p, res_p, tr_p, cond_p = run_product(product, 1, graph=True)
    # the product's discrete decisions of every step, in the form the oracle loop accepts
    decisions = []
    for tp in tr_p:
        keep = tp["pred_logits"][0].mean(-2).sigmoid().max(-1)[0] >= 0.5
        assert int(keep.sum()) == len(tp["results"][0])
        decisions.append(dict(topk=tp["topk"][:1], keep=keep, recs=tp["results"][0].recs))
    x_o, res_o, tr_o = run_oracle(manifests, "cuda", decisions=decisions)
    stats = []
    for i, (tp, to, rp, ro, d) in enumerate(zip(tr_p, tr_o, res_p, res_o, decisions)):
        # (a) latent after the step
        err = ((tp["x"] - to["x"]).abs().max() / to["x"].abs().max()).item()
        assert err < LATENT_TOL, f"step {i}: latent rel err {err:.3e}"
        # (b) proposal selection: picked proposals are the oracle's top-100 up to the logit tolerance, in an order the
        #     oracle's logits accept up to the same tolerance
        enc = to["enc_logits"]
        tol = ENC_TOL * enc.abs().max().item()
        picked = enc[d["topk"][0]]
        kth = enc.topk(100)[0][-1].item()
        assert (picked >= kth - 2 * tol).all(), f"step {i}: a picked proposal is below the oracle's 100th logit by > tol"
        own = set(to["free_topk"][0].tolist())
        n_same = len(own & set(d["topk"][0].tolist()))
        assert ((picked[:-1] - picked[1:]) >= -2 * tol).all(), f"step {i}: proposal order contradicts the oracle's logits"
        # (c) dense head outputs on the common proposals
        dn = to["dense"]
        rel = lambda a, b: ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()  # noqa: E731
        assert rel(tp["pred_logits"], dn["pred_logits"]) < 1.5e-1
        assert (tp["pred_ctrl_points"] - dn["pred_ctrl_points"]).abs().max().item() < 2e-2
        assert rel(tp["pred_texts"], dn["pred_texts"]) < TEXT_TOL
        # (d) keep decisions: equal to the oracle's, or the oracle's score is within SCORE_TOL of the threshold
        own_keep = to["score"] >= 0.5
        flips = own_keep != d["keep"]
        assert ((to["score"] - 0.5).abs()[flips] < SCORE_TOL).all(), f"step {i}: a detection decision differs beyond the margin"
        # (e) characters of the kept instances: equal to the oracle's arg-max, or within the logit tolerance of it
        lg = dn["pred_texts"][0][d["keep"]]
        ttol = 2 * TEXT_TOL * lg.abs().max().item() if lg.numel() else 0.0
        chosen = lg.gather(-1, d["recs"][..., None]).squeeze(-1)
        cflips = d["recs"] != to["own_recs"]
        assert ((lg.max(-1)[0] - chosen)[cflips] <= ttol).all(), f"step {i}: a recognised character differs beyond the margin"
        # (f) polygons of the kept instances (pixel units, before the int32 cast) and their int32 form
        pp = tp["results"][0].polygons
        assert pp.shape == to["polygons"].shape
        if pp.numel():
            assert (pp - to["polygons"]).abs().max().item() < POLY_TOL_PX
        for a, b in zip(rp["pred_polys"], ro["pred_polys"]):
            assert a.dtype == np.int32 and a.shape == (16, 2) and np.abs(a - b).max() <= int(POLY_TOL_PX) + 1
        # (g) host side: strings and prompt from the same characters are identical; so is the re-encoded context
        assert rp["pred_texts"] == ro["pred_texts"] and rp["pred_prompt"] == ro["pred_prompt"], f"step {i}: strings / prompt differ"
        assert rp["timestep"] == ro["timestep"]
        assert torch.equal(tp["c_txt"], product[0].clip.encode(ro["pred_prompt"]))
        stats.append(dict(step=i, latent_rel_err=err, proposals_shared=n_same, kept=int(d["keep"].sum()),
                          keep_flips=int(flips.sum()), chars=int(d["recs"].numel()), char_flips=int(cflips.sum())))
    # admissible flips must stay the exception: a systematic disagreement would hide behind the margins otherwise
    assert sum(s["keep_flips"] for s in stats) <= 0.1 * 100 * len(stats)
    assert sum(s["char_flips"] for s in stats) <= 0.25 * max(1, sum(s["chars"] for s in stats))
    assert all(s["proposals_shared"] >= 80 for s in stats)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        json.dump(stats, open(os.path.join(out, "feedback_parity.json"), "w"), indent=1)
    print("feedback loop parity:", stats)


def test_feedback_loop_graph_equals_eager_and_is_batch_independent(product):
    x1, res1, tr1, _ = run_product(product, 1, graph=True)
    xe, rese, _, _ = run_product(product, 1, graph=False)
    assert torch.equal(x1, xe), "CUDA-graph replay of the val_sample step must equal the eager loop bit for bit"
    assert [r["pred_prompt"] for r in res1] == [r["pred_prompt"] for r in rese]
    x2, res2, tr2, cond2 = run_product(product, 2, graph=True)
    assert torch.equal(x2[:1], x1), "tile 0 of a 2-tile batch must equal the 1-tile run bit for bit"
    for a, b in zip(res1, res2):
        assert a["pred_texts"] == b["pred_texts"] and a["pred_prompt"] == b["batch_prompts"][0]
        assert len(b["batch_prompts"]) == 2
    assert cond2["c_txt"].shape == (2, 77, 1024)
