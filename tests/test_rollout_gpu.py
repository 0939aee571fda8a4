"""Multi-step / edge-case parity of the CUDA path against the fp32 oracle run on the host CPU of the GPU box
(full SDv1.5-shaped weights, small latents so the oracle finishes in seconds).
Tolerances: single step eps cosine >= 0.999 / rel-L2 <= 2e-2; final latents after N sampler steps cosine >= 0.99
(SURVEY.md §8c)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    unet, nets, sds = common.build_models()
    dev = torch.device("cuda:0")
    for m in [unet] + nets:
        m.pack(dev)
    return dict(unet=unet, nets=nets, sds=sds, dev=dev)


def _oracle_rollout(sds, inp, sched, steps, cfg, scale):
    from oracle import dualdiff_oracle as O
    lat = inp["latents"].clone()
    with torch.no_grad():
        for t in sched.timesteps[:steps]:
            cur = dict(inp)
            cur["latents"] = lat
            lat, _ = O.denoise_step(sds["unet"], sds["bg"], sds["fg"], sched, lat, int(t), cur, scale, cfg)
    return lat


def _cuda_rollout(w, inp_cpu, scheduler, steps, scale, graph):
    from dualdiff_b200.pipeline import DualDiffDenoiser
    inp = common.to_dev(inp_cpu, w["dev"])
    den = DualDiffDenoiser(w["unet"], w["nets"], scheduler=scheduler, guidance_scale=scale, use_cuda_graph=graph)
    den.prepare(inp["latents"], inp["prompt_embeds"], inp["camera_param"], [inp["boxes_bg"], inp["boxes_fg"]],
                [inp["cond_bg"], inp["cond_fg"]], num_inference_steps=scheduler_steps(scheduler, steps))
    for i in range(steps):
        den.step(i)
    torch.cuda.synchronize()
    B = inp_cpu["latents"].shape[0]
    return den.latents.reshape(B, 6, 4, *inp_cpu["latents"].shape[-2:]).float().cpu()


def scheduler_steps(scheduler, steps):
    return getattr(scheduler, "_n_total", steps)


@pytest.mark.parametrize("kind,n_total,steps,scale", [("unipc", 6, 4, 2.0), ("ddim", 5, 3, 2.0), ("unipc", 4, 2, 1.0)])
def test_sampler_rollout_matches_oracle(world, kind, n_total, steps, scale):
    """several full sampler steps (both branches + UNet + CFG + scheduler) on 2 scenes at latent 8x12"""
    from dualdiff_b200 import scheduler as SCH, synthetic as S
    from oracle import dualdiff_oracle as O
    inp = S.make_inputs(2, 8, 12, seed=5, L_bg=7, L_fg=3)
    cfg = scale > 1.0
    osch = O.UniPC() if kind == "unipc" else O.DDIM()
    osch.set_timesteps(n_total)
    ref = _oracle_rollout(world["sds"], inp, osch, steps, cfg, scale)
    sch = SCH.UniPCMultistepScheduler() if kind == "unipc" else SCH.DDIMScheduler()
    sch._n_total = n_total
    out = _cuda_rollout(world, inp, sch, steps, scale, graph=True)
    m = common.metrics(out, ref)
    print(f"{kind} {steps}/{n_total} steps scale={scale}: {m}")
    assert torch.isfinite(out).all()
    assert m["cos"] >= 0.99 and m["rel_l2"] <= 5e-2, m


@pytest.mark.parametrize("h,w,L_bg,L_fg", [(8, 12, 1, 1), (6, 10, 40, 64), (5, 7, 3, 2)])
def test_single_step_ragged_shapes(world, h, w, L_bg, L_fg):
    """odd latent sizes (stride-2 chains 5x7 -> 3x4 -> 2x2 -> 1x1), minimal and large token counts"""
    from dualdiff_b200 import ops, synthetic as S
    from dualdiff_b200.pipeline import DualDiffDenoiser
    from oracle import dualdiff_oracle as O
    inp = S.make_inputs(1, h, w, seed=9, L_bg=L_bg, L_fg=L_fg)
    with torch.no_grad():
        ref = O.noise_prediction(world["sds"]["unet"], world["sds"]["bg"], world["sds"]["fg"], inp["latents"], 500, inp, 2.0, True)
    d = common.to_dev(inp, world["dev"])
    den = DualDiffDenoiser(world["unet"], world["nets"], guidance_scale=2.0, use_cuda_graph=False)
    den.prepare(d["latents"], d["prompt_embeds"], d["camera_param"], [d["boxes_bg"], d["boxes_fg"]],
                [d["cond_bg"], d["cond_fg"]], num_inference_steps=4)
    den.t_cur.fill_(500.0)
    den.coef_cur.copy_(den.coef_table[0])
    eps = ops.rows_to_nchw(den._step_kernels(), 12, (h, w)).cpu()
    m = common.metrics(eps, ref["eps_raw"])
    print(f"latent {h}x{w} L=({L_bg},{L_fg}): {m}")
    assert m["cos"] >= 0.999 and m["rel_l2"] <= 2e-2, m


def test_hd_resolution_runs_and_is_deterministic(world):
    """448x800 (latent 56x100, BASELINE config 4 geometry): T = 5600 tokens, 44 KV tiles per attention"""
    from dualdiff_b200 import synthetic as S
    inp = S.make_inputs(1, 56, 100, seed=2, L_bg=28, L_fg=32)
    from dualdiff_b200 import scheduler as SCH
    a = _cuda_rollout(world, inp, SCH.UniPCMultistepScheduler(), 2, 2.0, graph=True)
    b = _cuda_rollout(world, inp, SCH.UniPCMultistepScheduler(), 2, 2.0, graph=False)
    assert torch.isfinite(a).all() and torch.equal(a, b)
