"""Host-side logic of the pipeline mirror (no GPU): prompt layout, latent preparation, condition batching, option checks.
The text encoder is stubbed by the oracle here only to exercise the host code around it."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _OracleEncoder:
    def __init__(self, sd, heads):
        self.sd, self.heads = sd, heads

    def __call__(self, ids, attention_mask=None):
        from oracle import clip_oracle as CO
        assert attention_mask is None
        return (CO.text_model(self.sd, ids, self.heads),)


def _pipe():
    from dualdiff_b200 import scheduler as SCH, synthetic as S
    from dualdiff_b200.pipeline_bev_controlnet import StableDiffusionBEVControlNetPipeline
    from oracle import clip_oracle as CO
    g = torch.load(os.path.join(ROOT, "tests", "golden", "clip_small.pt"))
    sd = S.init_state_dict(CO.manifest(**g["config"]), seed=g["seed"])
    tok = CO.HashTokenizer(vocab_size=100, model_max_length=16)
    pipe = StableDiffusionBEVControlNetPipeline(None, _OracleEncoder(sd, 2), None, [torch.nn.Identity(), torch.nn.Identity()],
                                                SCH.UniPCMultistepScheduler(), tok)
    return pipe, sd, tok


def test_encode_prompt_matches_the_restated_diffusers_function():
    from oracle import clip_oracle as CO
    pipe, sd, tok = _pipe()
    with torch.no_grad():
        a = pipe._encode_prompt(["x y", "z"], torch.device("cpu"), 2, True, negative_prompt=["bad", "worse"])
        b = CO.encode_prompt(sd, tok, ["x y", "z"], 2, True, negative_prompt=["bad", "worse"], num_heads=2)
        c = pipe._encode_prompt("x y", torch.device("cpu"), 1, False)
    assert a.shape == (8, 16, 128) and torch.equal(a, b)
    assert c.shape == (1, 16, 128) and torch.equal(c[0], a[4])
    with pytest.raises(ValueError):
        pipe._encode_prompt(["x", "y"], torch.device("cpu"), 1, True, negative_prompt=["only one"])
    pre = torch.randn(2, 16, 128)
    d = pipe._encode_prompt(None, torch.device("cpu"), 1, True, prompt_embeds=pre, negative_prompt_embeds=-pre)
    assert torch.equal(d, torch.cat([-pre, pre]))


def test_prepare_latents_and_images():
    pipe, _, _ = _pipe()
    dev = torch.device("cpu")
    a = pipe.prepare_latents(2, 4, 224, 400, torch.float32, dev, torch.Generator().manual_seed(3))
    assert a.shape == (2, 4, 28, 50) and torch.equal(a, torch.randn(2, 4, 28, 50, generator=torch.Generator().manual_seed(3)))
    gens = [torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)]
    b = pipe.prepare_latents(2, 4, 64, 96, torch.float32, dev, gens)
    assert torch.equal(b[0], b[1])                      # fix_seed_within_batch (misc/test_utils.py:291-302)
    with pytest.raises(ValueError):
        pipe.prepare_latents(3, 4, 64, 96, torch.float32, dev, gens)
    given = torch.ones(2, 4, 8, 12)
    assert torch.equal(pipe.prepare_latents(2, 4, 64, 96, torch.float32, dev, None, given), given)
    img = pipe._prepare_image(torch.rand(1, 3, 8, 48), 4, 2, dev)
    assert img.shape == (4, 3, 8, 48)
    img = pipe._prepare_image(torch.rand(2, 3, 8, 48), 4, 2, dev)
    assert img.shape == (4, 3, 8, 48) and torch.equal(img[0], img[1])


def test_output_helpers():
    from dualdiff_b200.pipeline_bev_controlnet import BEVStableDiffusionPipelineOutput
    pipe, _, _ = _pipe()
    imgs = np.random.RandomState(0).rand(2, 6, 8, 12, 3).astype("float32")
    pil = pipe.numpy_to_pil_double(imgs)
    assert len(pil) == 2 and len(pil[0]) == 6 and pil[1][5].size == (12, 8)
    out = BEVStableDiffusionPipelineOutput(images=pil, nsfw_content_detected=None)
    assert out.images is pil and out[0] is pil and out.to_tuple()[1] is None


def test_call_checks_before_touching_the_gpu():
    pipe, _, _ = _pipe()
    base = dict(prompt=["a"], image=[torch.zeros(1), torch.zeros(1)], camera_param=None, height=64, width=96)
    with pytest.raises(NotImplementedError, match="guess_mode"):
        pipe(**base, guess_mode=True)
    with pytest.raises(ValueError, match="bboxes_3d_data"):      # per-view prompts are supported: the next check fires
        pipe(**base, bev_controlnet_kwargs={"use_aug_text": True})
    with pytest.raises(ValueError, match="bboxes_3d_data"):
        pipe(**base, bev_controlnet_kwargs={"use_aug_text": False})
    with pytest.raises(ValueError, match="image"):
        pipe(**dict(base, image=torch.zeros(1)), bev_controlnet_kwargs={"use_aug_text": False})
    with pytest.raises(AssertionError):
        type(pipe)(None, None, None, None, None, None, safety_checker=object())


def test_bbox_max_length_pads_like_the_reference():
    """`bbox_max_length` (reference pipeline :358,366 -> add_uncond_to_kwargs(max_len=...), unet_addon_rawbox.py:683-760):
    zero boxes / class 0 / mask False appended along the token axis, an absent dict becomes an all-masked one; restated
    inline from the reference for the conditional half (the unconditional half is built from the padded shapes)."""
    from dualdiff_b200.pipeline_bev_controlnet import pad_boxes
    g = torch.Generator().manual_seed(0)
    data = {"bboxes": torch.randn(2, 6, 5, 8, 3, generator=g), "classes": torch.randint(0, 10, (2, 6, 5), generator=g),
            "masks": torch.rand(2, 6, 5, generator=g) > 0.3}
    out = pad_boxes(data, 9, 2, 6, torch.device("cpu"))
    for key in ("bboxes", "classes", "masks"):
        v = data[key]
        to_pad = torch.zeros_like(v)[:, :, 1].unsqueeze(2).expand(-1, -1, 4, *v.shape[3:])        # reference :731-736
        ref = torch.cat([v, to_pad], dim=2)
        assert out[key].dtype == v.dtype and torch.equal(out[key], ref), key
    assert pad_boxes(data, 5, 2, 6, torch.device("cpu"))["bboxes"] is data["bboxes"]              # nothing to pad
    none = pad_boxes(None, 7, 2, 6, torch.device("cpu"))                                            # reference :708-716
    assert none["bboxes"].shape == (2, 6, 7, 8, 3) and none["classes"].dtype == torch.long and not none["masks"].any()
    with pytest.raises(AssertionError):
        pad_boxes(data, 4, 2, 6, torch.device("cpu"))
