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


def test_pipeline_call_matches_oracle(world):
    """StableDiffusionBEVControlNetPipeline.__call__ end to end (prompt strings -> CLIP -> 3 UniPC steps with CFG -> VAE
    decode) against the oracle chain clip_oracle.encode_prompt -> dualdiff_oracle.denoise_step x 3 -> vae_oracle.
    Tolerances: prompt embeddings / final latents cosine >= 0.99, decoded images max-abs <= 0.1 of the
    [0, 1] range and mean-abs <= 0.02."""
    from dualdiff_b200 import scheduler as SCH, synthetic as S
    from dualdiff_b200.networks import AutoencoderKLDecoder, CLIPTextModel
    from dualdiff_b200.pipeline_bev_controlnet import BEVStableDiffusionPipelineOutput, StableDiffusionBEVControlNetPipeline
    from oracle import clip_oracle as CO, dualdiff_oracle as O, vae_oracle as V
    clip_cfg = dict(vocab_size=1000, hidden_size=768, intermediate_size=3072, num_hidden_layers=2, num_attention_heads=12,
                    max_position_embeddings=77)
    tok = CO.HashTokenizer(vocab_size=1000)
    sd_clip = S.init_state_dict(CO.manifest(**clip_cfg), seed=3)
    enc = CLIPTextModel(**clip_cfg, eos_token_id=tok.eos_token_id)
    enc.load_state_dict(sd_clip, strict=True)
    sd_vae = S.init_state_dict(V.manifest(), seed=4)
    with torch.device("meta"):
        vae = AutoencoderKLDecoder()
    vae.load_state_dict(sd_vae, strict=True, assign=True)
    pipe = StableDiffusionBEVControlNetPipeline(vae, enc, world["unet"], world["nets"], SCH.UniPCMultistepScheduler(), tok)
    pipe.to("cuda:0")
    pipe.set_progress_bar_config(disable=True)
    pipe.enable_xformers_memory_efficient_attention()
    inp = S.make_inputs(2, 8, 12, seed=5, L_bg=7, L_fg=3)
    prompts = ["a driving scene image at singapore-onenorth. rain, many pedestrians", "a driving scene image at boston-seaport. night"]
    kwargs = dict(prompt=prompts, image=[inp["cond_bg"], inp["cond_fg"]], camera_param=inp["camera_param"], height=64,
                  width=96, num_inference_steps=3, guidance_scale=2.0,
                  bev_controlnet_kwargs={"bboxes_3d_data": [inp["boxes_bg"], inp["boxes_fg"]], "use_aug_text": False})
    seen = []
    out = pipe(**kwargs, generator=torch.Generator().manual_seed(7), output_type="latent",
               callback=lambda i, t, lat: seen.append((i, int(t), tuple(lat.shape))))
    assert isinstance(out, BEVStableDiffusionPipelineOutput) and out.nsfw_content_detected is None
    lat = out.images.float().cpu()
    assert lat.shape == (2, 6, 4, 8, 12) and [s[0] for s in seen] == [0, 1, 2] and seen[0][2] == (2, 6, 4, 8, 12)
    # ---- the oracle chain on the host ----
    with torch.no_grad():
        pe = CO.encode_prompt(sd_clip, tok, prompts)
        pe_cuda = pipe._encode_prompt(prompts, pipe.device, 1, True).float().cpu()
        m = common.metrics(pe_cuda, pe)
        assert pe.shape == (4, 77, 768) and m["cos"] > 0.999 and m["rel_l2"] < 2e-2, m
        lat0 = torch.randn(2, 4, 8, 12, generator=torch.Generator().manual_seed(7))
        ref_inp = dict(inp)
        ref_inp["latents"] = torch.stack([lat0] * 6, dim=1)
        ref_inp["prompt_embeds"] = pe
        osch = O.UniPC()
        osch.set_timesteps(3)
        ref = _oracle_rollout(world["sds"], ref_inp, osch, 3, True, 2.0)
        ref_img = V.decode_latents(sd_vae, ref.reshape(12, 4, 8, 12)).reshape(2, 6, 3, 64, 96).permute(0, 1, 3, 4, 2)
    m = common.metrics(lat, ref)
    print("pipeline latents vs oracle:", m)
    assert torch.isfinite(lat).all() and m["cos"] >= 0.99 and m["rel_l2"] <= 5e-2, m
    # ---- decoded output, both output types ----
    res = pipe(**kwargs, generator=torch.Generator().manual_seed(7), output_type="np", return_dict=False)
    img = torch.from_numpy(res[0])
    assert img.shape == (2, 6, 64, 96, 3) and img.min() >= 0 and img.max() <= 1 and res[1] is None
    d = (img - ref_img).abs()
    print("pipeline images vs oracle: max", d.max().item(), "mean", d.mean().item())
    assert d.max() <= 0.1 and d.mean() <= 0.02
    pil = pipe(**kwargs, generator=torch.Generator().manual_seed(7), num_images_per_prompt=1).images
    assert len(pil) == 2 and len(pil[0]) == 6 and pil[0][0].size == (96, 64)
    # ---- bbox_max_length (reference :358,366): the call pads the box lists with masked tokens; it must equal the same call on
    # hand-padded lists bit for bit, and differ from the unpadded one (the null tokens join the keys of the text cross-attention)
    from dualdiff_b200.pipeline_bev_controlnet import pad_boxes
    padded = [pad_boxes(b, 9, 2, 6, torch.device("cpu")) for b in (inp["boxes_bg"], inp["boxes_fg"])]
    assert padded[0]["bboxes"].shape[2] == 9 and padded[1]["bboxes"].shape[2] == 9
    kw_pad = dict(kwargs, bev_controlnet_kwargs={"bboxes_3d_data": padded, "use_aug_text": False})
    a = pipe(**kwargs, generator=torch.Generator().manual_seed(7), output_type="latent", bbox_max_length=9).images
    b = pipe(**kw_pad, generator=torch.Generator().manual_seed(7), output_type="latent").images
    assert torch.isfinite(a).all() and torch.equal(a, b) and not torch.equal(a.float().cpu(), lat)


def test_pipeline_rejects_unbuilt_options(world):
    from dualdiff_b200 import scheduler as SCH
    from dualdiff_b200.pipeline_bev_controlnet import StableDiffusionBEVControlNetPipeline
    pipe = StableDiffusionBEVControlNetPipeline(None, None, world["unet"], world["nets"], SCH.UniPCMultistepScheduler(), None)
    base = dict(prompt=None, image=[None, None], camera_param=None, height=64, width=96)
    with pytest.raises(NotImplementedError, match="guess_mode"):
        pipe(**base, guess_mode=True)
    with pytest.raises(NotImplementedError, match="cross_attention_kwargs"):
        pipe(**base, cross_attention_kwargs={"scale": 0.5})
    with pytest.raises(ValueError, match="bboxes_3d_data"):       # use_aug_text itself is accepted (per-view prompts)
        pipe(**base, bev_controlnet_kwargs={"use_aug_text": True})
    with pytest.raises(ValueError, match="bboxes_3d_data"):
        pipe(**base, bev_controlnet_kwargs={"use_aug_text": False})
    with pytest.raises(AssertionError):
        StableDiffusionBEVControlNetPipeline(None, None, None, None, None, None, safety_checker=object())
