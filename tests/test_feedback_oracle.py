"""The oracle's text-spotting feedback loop (oracle/val_loop.py) against the fixture produced by the UNMODIFIED
reference ``SpacedSampler.val_sample`` (tests/golden/make_golden.py feedback): same strings, same prompts, same
int32 polygons, latents to fp32 round-off after every step.  Runs on the CPU (and again on the GPU in fp32 from
tests/test_feedback_gpu.py)."""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN


def load_fixture():
    meta = json.load(open(os.path.join(GOLDEN, "val_feedback.json")))
    arr = np.load(os.path.join(GOLDEN, "val_feedback.npz"))
    return meta, arr


def run_oracle(manifests, device, decisions=None):
    from oracle import sampler as OS, unet as OU, val_loop as VL, weights
    usd = {k: v.to(device) for k, v in weights.seeded_state_dict(manifests["unet_full"]).items()}
    csd = {k: v.to(device) for k, v in weights.seeded_state_dict(manifests["controlnet_full"]).items()}
    tsd = {k: v.to(device) for k, v in weights.seeded_state_dict(manifests["testr"]).items()}
    meta, _ = load_fixture()
    steps = meta["steps"]
    sched = OS.make_schedule(OS.diffusion_betas(), steps)
    x_T, c_img = weights.seeded_randn((1, 4, 64, 64), 9800).to(device), weights.seeded_randn((1, 4, 64, 64), 9900).to(device)
    noises = [weights.seeded_randn((1, 4, 64, 64), 10000 + i).to(device) for i in range(steps)]
    clip = VL.HashClip(device)
    cond = dict(c_txt=clip.encode(""), c_img=c_img)
    with torch.no_grad():
        return VL.val_sample(lambda x, mt, c: OU.cldm_forward(usd, csd, x, mt, c["c_txt"], c["c_img"]), tsd, sched, x_T,
                             noises, cond, clip, decisions=decisions)


def check_against_fixture(x, res, trace, tol):
    meta, arr = load_fixture()
    assert [r["timestep"] for r in res] == meta["timesteps"]
    for i, r in enumerate(res):
        assert r["pred_texts"] == meta["pred_texts"][i], f"step {i}: recognised strings differ from the reference"
        assert r["pred_prompt"] == meta["pred_prompt"][i]
        want = arr[f"polys{i}"]
        got = np.stack(r["pred_polys"]) if r["pred_polys"] else np.zeros((0, 16, 2), np.int32)
        assert got.shape == want.shape
        # (int32) truncation of a pixel coordinate: a value within fp32 round-off of an integer may land on either side
        assert np.abs(got - want).max(initial=0) <= 1 and (got != want).mean() < 1e-2
        err = (trace[i]["x"].cpu() - torch.from_numpy(arr["x"][i:i + 1])).abs().max().item()
        assert err < tol, f"step {i}: latent differs from the reference by {err}"


def test_oracle_feedback_loop_matches_reference_fixture(manifests):
    x, res, trace = run_oracle(manifests, "cpu")
    check_against_fixture(x, res, trace, 2e-4)
    # the decisions of this fixture are not borderline at fp32 precision (they must survive a change of host / device)
    for tr in trace:
        assert (tr["score"] - 0.5).abs().min().item() > 1e-4
