"""StableDiffusionBEVControlNetPipeline — the caller of the denoising step, kept as a drop-in (SURVEY.md §8b "Callers",
§8f rank 3).

Reference: pipeline/pipeline_bev_controlnet.py:39-559.  Same constructor, same `__call__` signature and argument meaning,
same `BEVStableDiffusionPipelineOutput`; what happens inside is the B200 path end to end:

  prompt  --tokenizer (host)-->  ids  --CLIPTextModel on the GEMM / LayerNorm / dd_seq_attention kernels-->  prompt_embeds
  (`_encode_prompt`, reference :273-281: negative half first)
  latents = randn(generator) * init_noise_sigma, one latent stacked over the six views (:327-345)
  DualDiffDenoiser.prepare(...)  — uncond camera / zero boxes in front (:346-375), tokens, text K/V, condition embedding, SFA
  DualDiffDenoiser.step(i)       — one CUDA-graph replay per sampler step (:378-504), `callback(i, t, latents)` honoured
  AutoencoderKLDecoder.decode_latents  — `decode_latents` with 5-dim latents (:101-113), then PIL / numpy as asked

Options that the reference path never takes raise NotImplementedError naming the option (guess_mode, a single
non-dual ControlNet, cross_attention_kwargs, controlnet_conditioning_scale != 1, safety checker).
"""
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Union

import numpy as np
import torch

from .networks.output_cls import _Output
from .pipeline import DualDiffDenoiser


@dataclass
class BEVStableDiffusionPipelineOutput(_Output):
    """reference :21-36"""
    images: Union[List[List[Any]], np.ndarray]
    nsfw_content_detected: Optional[List[bool]]


class MultiControlNetModel(torch.nn.Module):
    """diffusers.pipelines.controlnet.MultiControlNetModel: the container the reference wraps the two branches in
    (`isinstance(self.controlnet, MultiControlNetModel)`, :354,405)"""

    def __init__(self, controlnets):
        super().__init__()
        self.nets = torch.nn.ModuleList(controlnets)


class _NullBar:
    def __init__(self, total=None, **kw):
        self.n, self.total = 0, total

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def update(self, n=1):
        self.n += n


def pad_boxes(data, max_len: int, batch: int, n_cam: int, device):
    """`bbox_max_length` (reference pipeline :358,366 -> `BEVControlNetModel.add_uncond_to_kwargs(max_len=...)`,
    unet_addon_rawbox.py:683-760): the box / map-vector lists are padded along the token axis with zero boxes, class 0 and mask
    False up to `max_len` tokens -- masked tokens become the learned null token, so the key count of the text cross-attention
    follows `max_len`; a missing dict (no visible boxes) becomes an all-masked one.  The unconditional half is built from the
    padded shapes by `DualDiffDenoiser.prepare` as for unpadded boxes."""
    if data is None:
        return {"bboxes": torch.zeros([batch, n_cam, max_len, 8, 3], device=device),
                "classes": torch.zeros([batch, n_cam, max_len], device=device, dtype=torch.long),
                "masks": torch.zeros([batch, n_cam, max_len], device=device, dtype=torch.bool)}
    out = {}
    for key in ("bboxes", "classes", "masks"):
        v = data[key]
        extra = max_len - v.shape[2]
        assert extra >= 0, f"bbox_max_length={max_len} < {v.shape[2]} {key} tokens"
        out[key] = torch.cat([v, torch.zeros_like(v[:, :, :1]).expand(-1, -1, extra, *v.shape[3:])], dim=2) if extra else v
    for key, v in data.items():
        out.setdefault(key, v)
    return out


class StableDiffusionBEVControlNetPipeline:
    def __init__(self, vae, text_encoder, unet, controlnet, scheduler, tokenizer, safety_checker=None,
                 feature_extractor=None, requires_safety_checker: bool = False):
        assert safety_checker is None, "Please do not use safety_checker."      # reference :62
        if isinstance(controlnet, (list, tuple)):                                 # diffusers wraps lists itself
            controlnet = MultiControlNetModel(controlnet)
        self.vae, self.text_encoder, self.unet, self.controlnet = vae, text_encoder, unet, controlnet
        self.scheduler, self.tokenizer = scheduler, tokenizer
        self.safety_checker, self.feature_extractor = None, feature_extractor
        self.vae_scale_factor = 8            # 2 ** (len(vae.config.block_out_channels) - 1) for the SD VAE
        self._progress_bar_config: Dict[str, Any] = {}
        self._denoiser = None
        self._device = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, unet=None, controlnet=None, safety_checker=None,
                        feature_extractor=None, torch_dtype=None, tokenizer=None, **kwargs):
        """`pipe_cls.from_pretrained(cfg.model.pretrained_model_name_or_path, controlnet=..., unet=..., safety_checker=None,
        feature_extractor=None, torch_dtype=...)` (misc/test_utils.py:150-156): the VAE decoder, the text encoder, the
        tokenizer and the scheduler config come from the Stable-Diffusion-v1.5 directory (`vae/`, `text_encoder/`,
        `tokenizer/`, `scheduler/scheduler_config.json`); `unet` and `controlnet` are the already-loaded DualDiff modules.
        The scheduler is built as UniPC straight away (the reference replaces the checkpoint's scheduler by
        `UniPCMultistepScheduler.from_config(pipe.scheduler.config)` on the next line, :162).  `torch_dtype` is accepted and
        ignored: the kernels compute in bf16 with fp32 accumulation whatever the storage type of the checkpoint."""
        import json
        import os
        from .networks import AutoencoderKLDecoder, CLIPTextModel
        from .scheduler import UniPCMultistepScheduler
        if unet is None or controlnet is None:
            raise ValueError("pass the loaded DualDiff modules: unet=UNet2DConditionModelMultiview, controlnet=[bg, fg]")
        root = pretrained_model_name_or_path
        vae = AutoencoderKLDecoder.from_pretrained(root, subfolder="vae")
        text_encoder = CLIPTextModel.from_pretrained(root, subfolder="text_encoder")
        if tokenizer is None:
            from transformers import CLIPTokenizer       # host-side BPE, used as is
            tokenizer = CLIPTokenizer.from_pretrained(os.path.join(root, "tokenizer"))
        sched_cfg = {}
        sched_path = os.path.join(root, "scheduler", "scheduler_config.json")
        if os.path.exists(sched_path):
            with open(sched_path) as fh:
                sched_cfg = json.load(fh)
        scheduler = UniPCMultistepScheduler.from_config({k: v for k, v in sched_cfg.items() if k != "_class_name"})
        return cls(vae, text_encoder, unet, controlnet, scheduler, tokenizer, safety_checker=safety_checker,
                   feature_extractor=feature_extractor)

    # ---- diffusers DiffusionPipeline plumbing the callers use (misc/test_utils.py:157-171) -------------------------
    def to(self, device):
        device = torch.device(device)
        for m in (self.vae, self.text_encoder, self.unet, self.controlnet):
            if m is not None:
                m.to(device)
        self._device = device
        return self

    @property
    def device(self):
        return self._device if self._device is not None else self.unet.device

    _execution_device = device

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        """accepted, no-op: attention always runs on the tcgen05 kernels (misc/test_utils.py:164-165)"""

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def progress_bar(self, iterable=None, total=None):
        if self._progress_bar_config.get("disable", False):
            return _NullBar(total=total)
        try:
            from tqdm.auto import tqdm
        except Exception:  # pragma: no cover
            return _NullBar(total=total)
        return tqdm(iterable, total=total, **self._progress_bar_config) if iterable is not None else \
            tqdm(total=total, **self._progress_bar_config)

    # ---- reference helpers ------------------------------------------------------------------------------------------
    @staticmethod
    def numpy_to_pil(images):
        """diffusers DiffusionPipeline.numpy_to_pil"""
        from PIL import Image
        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        return [Image.fromarray(im.squeeze(), mode="L") if im.shape[-1] == 1 else Image.fromarray(im) for im in images]

    def numpy_to_pil_double(self, images):
        """5-dim input -> 2-dim list (reference :72-80)"""
        return [self.numpy_to_pil(imgs) for imgs in images]

    def decode_latents(self, latents):
        """(b, n_cam, 4, h, w) -> numpy (b, n_cam, 8h, 8w, 3) in [0, 1] (reference :101-113)"""
        bs = len(latents)
        flat = latents.reshape(-1, *latents.shape[-3:])
        image = self.vae.decode_latents(flat)                         # scale, decode, (x / 2 + 0.5).clamp(0, 1)
        image = image.reshape(bs, -1, *image.shape[1:])
        return image.permute(0, 1, 3, 4, 2).float().cpu().numpy()

    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt=None,
                       prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None):
        """diffusers 0.17.1 `_encode_prompt` (called at reference :273-281); negative embeddings in front"""
        if prompt is not None and isinstance(prompt, str):
            prompt = [prompt]
        batch = len(prompt) if prompt is not None else prompt_embeds.shape[0]

        def enc(texts, max_length):
            ids = self.tokenizer(texts, padding="max_length", max_length=max_length, truncation=True,
                                 return_tensors="pt").input_ids
            return self.text_encoder(ids.to(device), attention_mask=None)[0]

        if prompt_embeds is None:
            prompt_embeds = enc(prompt, self.tokenizer.model_max_length)
        prompt_embeds = prompt_embeds.to(device=device)
        b, L, C = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, L, C)
        if do_classifier_free_guidance:
            if negative_prompt_embeds is None:
                if negative_prompt is None:
                    uncond = [""] * batch
                elif prompt is not None and type(prompt) is not type(negative_prompt) and not isinstance(negative_prompt, str):
                    raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got "
                                    f"{type(negative_prompt)} != {type(prompt)}.")
                elif isinstance(negative_prompt, str):
                    uncond = [negative_prompt]
                else:
                    uncond = list(negative_prompt)
                if len(uncond) != batch:
                    raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(uncond)}, but `prompt` has "
                                     f"batch size {batch}. Please make sure that passed `negative_prompt` matches the batch "
                                     "size of `prompt`.")
                negative_prompt_embeds = enc(uncond, L)
            negative_prompt_embeds = negative_prompt_embeds.to(device=device, dtype=prompt_embeds.dtype)
            negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_images_per_prompt, 1).view(
                batch * num_images_per_prompt, L, -1)
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds])
        return prompt_embeds

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        """diffusers prepare_latents: host generators as in run_one_batch_pipe (misc/test_utils.py:286-304)"""
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective "
                             f"batch size of {batch_size}. Make sure the batch size matches the length of the generators.")
        if latents is None:
            if isinstance(generator, list):
                latents = torch.cat([torch.randn((1,) + shape[1:], generator=g, device=g.device, dtype=torch.float32)
                                     for g in generator])
            else:
                gdev = generator.device if generator is not None else device
                latents = torch.randn(shape, generator=generator, device=gdev, dtype=torch.float32)
        latents = latents.to(device=device, dtype=torch.float32)
        return latents * self.scheduler.init_noise_sigma

    def _prepare_image(self, image, batch_size, num_images_per_prompt, device):
        """`prepare_image` for tensor input with do_resize / do_normalize off (reference :286-313): repeat to the batch;
        the CFG duplication is done inside the denoiser (both halves share the condition, :351-373)"""
        if not torch.is_tensor(image):
            raise TypeError("the condition inputs must be tensors (bg panorama, fg ORS tensor)")
        if image.dim() == 3:
            image = image[None]
        repeat_by = batch_size if image.shape[0] == 1 else num_images_per_prompt
        if repeat_by != 1:
            image = image.repeat_interleave(repeat_by, dim=0)
        return image.to(device=device, dtype=torch.float32)

    # ---- the call ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str]], image, camera_param: Optional[torch.Tensor], height: int, width: int,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5,
                 negative_prompt: Optional[Union[str, List[str]]] = None, num_images_per_prompt: Optional[int] = 1,
                 eta: float = 0.0, generator=None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "pil", return_dict: bool = True,
                 callback: Optional[Callable[[int, int, torch.Tensor], None]] = None, callback_steps: int = 1,
                 cross_attention_kwargs: Optional[Dict[str, Any]] = None, controlnet_conditioning_scale: float = 1,
                 guess_mode: bool = False, use_zero_map_as_unconditional: bool = False, bev_controlnet_kwargs={},
                 bbox_max_length=None):
        if guess_mode:
            raise NotImplementedError("guess_mode is unused on the reference's dual-branch path")
        if cross_attention_kwargs is not None:
            raise NotImplementedError("cross_attention_kwargs are unused on the reference path")
        if controlnet_conditioning_scale != 1:
            raise NotImplementedError("controlnet_conditioning_scale != 1 (configs/runner/default.yaml keeps the default)")
        if not isinstance(self.controlnet, MultiControlNetModel) or len(self.controlnet.nets) != 2:
            raise NotImplementedError("the B200 path is the dual-branch configuration: controlnet = [bg branch, fg branch]")
        if not isinstance(image, (list, tuple)) or len(image) != 2:
            raise ValueError("dual branch: `image` must be [bg occupancy panorama, fg ORS tensor]")
        bev_controlnet_kwargs = dict(bev_controlnet_kwargs)
        use_aug_text = bool(bev_controlnet_kwargs.get("use_aug_text", False))   # one prompt per VIEW (configs/exp/occ_bg_augtext.yaml)
        bboxes_3d_data = bev_controlnet_kwargs.get("bboxes_3d_data")
        if not isinstance(bboxes_3d_data, (list, tuple)) or len(bboxes_3d_data) != 2:
            raise ValueError("dual branch: bev_controlnet_kwargs['bboxes_3d_data'] must be [bg boxes, fg map vectors]")
        nets = list(self.controlnet.nets)
        # 2. call parameters (reference :249-269)
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt) if not use_aug_text else len(prompt) // 6      # reference :250
        else:
            batch_size = prompt_embeds.shape[0] if not use_aug_text else prompt_embeds.shape[0] // 6
        device = self.device
        if device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: call pipe.to('cuda') first")
        do_cfg = guidance_scale > 1.0
        if camera_param is None:     # learned unconditional camera, CFG off (reference :262-267)
            camera_param = nets[0].uncond_cam_param((batch_size, 6))
            do_cfg = False
        # 3. prompt
        prompt_embeds = self._encode_prompt(prompt, device, num_images_per_prompt, do_cfg, negative_prompt,
                                            prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds)
        # 4. condition inputs
        images = [self._prepare_image(im, batch_size * num_images_per_prompt, num_images_per_prompt, device) for im in image]
        if use_zero_map_as_unconditional and do_cfg:
            raise NotImplementedError                                   # as the reference (:315-316)
        # 5./6. timesteps and latents
        lat = self.prepare_latents(batch_size * num_images_per_prompt, self.unet.config.in_channels, height, width,
                                   prompt_embeds.dtype, device, generator, latents)
        assert camera_param.shape[0] == batch_size, \
            f"Except {batch_size} camera params, but you have bs={len(camera_param)}"
        n_cam = camera_param.shape[1]
        lat = torch.stack([lat] * n_cam, dim=1)                        # bs, 6, 4, h, w (:345)
        camera_param = camera_param.to(device)
        move = lambda d: None if d is None else {k: v.to(device) for k, v in d.items()}
        boxes = [move(b) for b in bboxes_3d_data]
        if bbox_max_length is not None:                                # pad the boxes to max_len (:358,366 -> add_uncond_to_kwargs)
            boxes = [pad_boxes(b, bbox_max_length, batch_size * num_images_per_prompt, n_cam, device) for b in boxes]
        # 8. denoising loop: one DualDiffDenoiser (CUDA graph per step)
        den = self._denoiser
        if den is None or den.unet is not self.unet or den.nets != nets or den.scheduler is not self.scheduler:
            den = self._denoiser = DualDiffDenoiser(self.unet, nets, scheduler=self.scheduler, guidance_scale=guidance_scale)
        den.guidance_scale = guidance_scale
        den.scheduler.guidance_scale = guidance_scale if do_cfg else 1.0
        den.cfg = do_cfg
        den.prepare(lat, prompt_embeds, camera_param, boxes, images, num_inference_steps)
        timesteps = self.scheduler.timesteps
        with self.progress_bar(total=num_inference_steps) as bar:
            for i in range(len(timesteps)):
                cur = den.step(i)
                bar.update()
                if callback is not None and i % callback_steps == 0:
                    callback(i, timesteps[i], cur.reshape(batch_size * num_images_per_prompt, n_cam, *cur.shape[1:]))
        latents = den.latents.reshape(batch_size * num_images_per_prompt, n_cam, *den.latents.shape[1:])
        if output_type == "latent":
            out_image, nsfw = latents, None
        else:
            out_image, nsfw = self.decode_latents(latents), None
            if output_type == "pil":
                out_image = self.numpy_to_pil_double(out_image)
        if not return_dict:
            return (out_image, nsfw)
        return BEVStableDiffusionPipelineOutput(images=out_image, nsfw_content_detected=nsfw)
