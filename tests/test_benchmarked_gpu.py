"""Parity on the configurations bench.py measures ("the benchmarked workload has no parity check").

  * BASELINE configs[1]'s sampler: 25 UniPC(bh2) + CFG steps at 28x50 against tests/golden/rollout25_full.pt, which
    oracle/make_golden.py produced with the reference's own network classes inside the loop (B = 1);
  * BASELINE configs[1]'s batch: one step of B = 8 scenes -- scene 0 carries the golden's inputs and must reproduce the B = 1
    result, other scenes are checked against the oracle port on the host;
  * BASELINE configs[3]'s geometry (448x800, latent 56x100): one step against tests/golden/step_hd.pt (reference classes);
  * a branch without box tokens (`bboxes_3d_data=None`, dataset/utils.py:235-237) against the oracle.
Tolerances (SURVEY.md section 8c, bf16 kernels vs the fp32 reference): one step -- noise prediction cosine >= 0.999 and
rel-L2 <= 2e-2; final latents after the 25-step rollout -- cosine >= 0.99."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402

pytestmark = pytest.mark.gpu
COS_MIN, REL_L2_MAX = 0.999, 2e-2


@pytest.fixture(scope="module")
def world():
    unet, nets, sds = common.build_models()
    dev = torch.device("cuda:0")
    for m in [unet] + nets:
        m.pack(dev)
    return dict(unet=unet, nets=nets, sds=sds, dev=dev)


def _denoiser(w, inp_cpu, steps, graph, scheduler=None):
    from dualdiff_b200.pipeline import DualDiffDenoiser
    inp = common.to_dev(inp_cpu, w["dev"])
    den = DualDiffDenoiser(w["unet"], w["nets"], scheduler=scheduler, guidance_scale=2.0, use_cuda_graph=graph)
    den.prepare(inp["latents"], inp["prompt_embeds"], inp["camera_param"], [inp["boxes_bg"], inp["boxes_fg"]],
                [inp["cond_bg"], inp["cond_fg"]], num_inference_steps=steps)
    return den


def _one_step_eps(w, inp_cpu, t, h, wd):
    """raw (uncond, cond) noise prediction of one loop body at timestep t: NCHW fp32 on the host"""
    from dualdiff_b200 import ops
    den = _denoiser(w, inp_cpu, 4, graph=False)
    den.t_cur.fill_(float(t))
    den.coef_cur.copy_(den.coef_table[0])
    rows = den._step_kernels()
    torch.cuda.synchronize()
    B = inp_cpu["latents"].shape[0]
    return ops.rows_to_nchw(rows, 2 * B * 6, (h, wd)).cpu()


def test_25_step_unipc_rollout_matches_the_reference_rollout(world):
    from dualdiff_b200 import synthetic as S
    fix = torch.load(os.path.join(common.GOLDEN, "rollout25_full.pt"))
    c = fix["config"]
    for name, m in (("unet", world["unet"]), ("bg", world["nets"][0]), ("fg", world["nets"][1])):
        assert S.manifest_digest(S.manifest_of(m)) == fix["manifest_digest"][name]
    inp = S.make_inputs(c["B"], c["h"], c["w"], seed=1, L_bg=c["L_bg"], L_fg=c["L_fg"])
    den = _denoiser(world, inp, c["n_steps"], graph=True)
    assert torch.equal(den.scheduler.timesteps.cpu().long(), fix["timesteps"].long()), "timestep schedule differs"
    report = {}
    for i in range(c["n_steps"]):
        den.step(i)
        if i + 1 in fix["latents_at"]:
            torch.cuda.synchronize()
            cur = den.latents.reshape(c["B"], 6, 4, c["h"], c["w"]).float().cpu()
            report[i + 1] = common.metrics(cur, fix["latents_at"][i + 1])
    for k, m in report.items():
        print(f"after step {k:2d}: cos {m['cos']:.6f} rel-L2 {m['rel_l2']:.4f}")
    final = den.latents.reshape(c["B"], 6, 4, c["h"], c["w"]).float().cpu()
    m = common.metrics(final, fix["final_latents"])
    print("final latents after 25 UniPC+CFG steps vs the reference rollout:", m)
    assert torch.isfinite(final).all()
    assert m["cos"] >= 0.99, m


def test_batch_of_8_scenes_one_step(world):
    """the bench batch: scene 0 = the golden's scene (must equal the B = 1 result on the GPU and the reference golden),
    two more scenes of the batch against the oracle port"""
    from oracle import dualdiff_oracle as O
    fix = torch.load(os.path.join(common.GOLDEN, "step_full.pt"))
    c = fix["config"]
    h, wd, t = c["h"], c["w"], c["t"]
    seeds = [1, 11, 12, 13, 14, 15, 16, 17]
    batch = common.make_scene_batch(seeds, h, wd, c["L_bg"], c["L_fg"])
    B = len(seeds)
    eps8 = _one_step_eps(world, batch, t, h, wd).reshape(2, B, 6, 4, h, wd)
    assert torch.isfinite(eps8).all()
    # scene 0 against the reference golden (B = 1 run of the reference's own classes)
    e0 = eps8[:, 0].reshape(12, 4, h, wd)
    m_gold = common.metrics(e0, fix["eps_raw"])
    print("scene 0 of the B=8 step vs reference golden:", m_gold)
    assert m_gold["cos"] >= COS_MIN and m_gold["rel_l2"] <= REL_L2_MAX, m_gold
    # ... and against the same scene computed alone on the GPU.  Not bit-equal: the tile schedule of the GEMMs (stream-K
    # splits, CTA pairing) depends on the number of row tiles, so fp32 sums are formed in another order, bf16 roundings flip
    # and the network amplifies them to the same level as the distance to the fp32 reference (measured rel-L2 1.1e-2 against
    # 1.2e-2 to the reference).  The bound is therefore the bf16 tolerance, like for the reference itself.
    e1 = _one_step_eps(world, common.scene_of(batch, 0), t, h, wd)
    m_b1 = common.metrics(e0, e1)
    print("scene 0 of the B=8 step vs the B=1 step on the GPU:", m_b1, "bit-equal:", torch.equal(e0, e1))
    assert m_b1["cos"] >= COS_MIN and m_b1["rel_l2"] <= REL_L2_MAX, m_b1
    # two other scenes against the oracle port (fp32, host CPU)
    for i in (3, 7):
        sc = common.scene_of(batch, i)
        with torch.no_grad():
            ref = O.noise_prediction(world["sds"]["unet"], world["sds"]["bg"], world["sds"]["fg"], sc["latents"], t, sc, 2.0, True)
        m = common.metrics(eps8[:, i].reshape(12, 4, h, wd), ref["eps_raw"])
        print(f"scene {i} of the B=8 step vs oracle:", m)
        assert m["cos"] >= COS_MIN and m["rel_l2"] <= REL_L2_MAX, (i, m)


def test_hd_448x800_step_matches_reference(world):
    """latent 56x100 (BASELINE config 4 geometry): 5600 tokens -> 117 key tiles per level-0 attention, 5600-pixel GroupNorm
    slabs; golden from the reference's own classes"""
    from dualdiff_b200 import synthetic as S
    fix = torch.load(os.path.join(common.GOLDEN, "step_hd.pt"))
    c = fix["config"]
    inp = S.make_inputs(c["B"], c["h"], c["w"], seed=1, L_bg=c["L_bg"], L_fg=c["L_fg"])
    eps = _one_step_eps(world, inp, c["t"], c["h"], c["w"])
    assert torch.isfinite(eps).all()
    e_u, e_c = eps.chunk(2)
    m = common.metrics(e_u + 2.0 * (e_c - e_u), fix["eps"])
    print("guided eps at 56x100 vs reference:", m)
    assert m["cos"] >= COS_MIN and m["rel_l2"] <= 2 * REL_L2_MAX, m
    m_raw = common.metrics(common.strided_sample(eps.contiguous(), 65536), fix["eps_raw_sample"])
    print("raw eps sample at 56x100 vs reference:", m_raw)
    assert m_raw["cos"] >= COS_MIN and m_raw["rel_l2"] <= REL_L2_MAX, m_raw


def test_branch_without_box_tokens(world):
    """`bboxes_3d_data=None` for the bg branch (no visible box in the batch, dataset/utils.py:235-237): the branch runs on
    [camera | text] tokens only (unet_addon_rawbox.py:892-895,1066-1069) and the UNet reads those 78 tokens"""
    from dualdiff_b200 import synthetic as S
    from oracle import dualdiff_oracle as O
    h, wd, t = 8, 12, 500
    inp = S.make_inputs(1, h, wd, seed=21, L_bg=5, L_fg=4)
    inp["boxes_bg"] = None
    with torch.no_grad():
        ref = O.noise_prediction(world["sds"]["unet"], world["sds"]["bg"], world["sds"]["fg"], inp["latents"], t, inp, 2.0, True)
    eps = _one_step_eps(world, inp, t, h, wd)
    m = common.metrics(eps, ref["eps_raw"])
    print("bg branch without box tokens vs oracle:", m)
    assert m["cos"] >= COS_MIN and m["rel_l2"] <= REL_L2_MAX, m
