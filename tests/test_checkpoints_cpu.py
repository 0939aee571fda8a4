"""Checkpoint loading of the pipeline components the way misc/test_utils.py:97-171 (`build_pipe`) does it: the text encoder
from a directory written by transformers itself, the VAE decoder from an `AutoencoderKL` checkpoint in the old (diffusers
<= 0.17 AttentionBlock) key naming of the SD-v1.5 files, the scheduler through `from_config`, the pipeline through
`from_pretrained`.  Host logic only (no GPU)."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SMALL = dict(vocab_size=100, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
             max_position_embeddings=16)
PNDM = {"_class_name": "PNDMScheduler", "_diffusers_version": "0.6.0", "beta_end": 0.012, "beta_schedule": "scaled_linear",
        "beta_start": 0.00085, "clip_sample": False, "num_train_timesteps": 1000, "set_alpha_to_one": False,
        "skip_prk_steps": True, "steps_offset": 1, "trained_betas": None}


def test_text_encoder_loads_a_directory_written_by_transformers(tmp_path):
    tr = pytest.importorskip("transformers")
    from dualdiff_b200.networks import CLIPTextModel
    cfg = tr.CLIPTextConfig(**SMALL, hidden_act="quick_gelu", layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=98, eos_token_id=99)
    torch.manual_seed(0)
    hf = tr.CLIPTextModel(cfg).eval()
    hf.save_pretrained(tmp_path / "text_encoder")
    ours = CLIPTextModel.from_pretrained(str(tmp_path), subfolder="text_encoder")
    assert ours.config.num_hidden_layers == 2 and ours.config.eos_token_id == 99
    ref = {k: v for k, v in hf.state_dict().items() if "position_ids" not in k}
    got = ours.state_dict()
    assert set(got) == set(ref) and all(torch.equal(got[k], ref[k]) for k in ref)
    # and back: a directory written by the mirror loads into transformers
    ours.save_pretrained(tmp_path / "again")
    hf2 = tr.CLIPTextModel.from_pretrained(tmp_path / "again")
    assert all(torch.equal(hf2.state_dict()[k], ref[k]) for k in ref)
    # a truncated checkpoint is refused
    from safetensors.torch import load_file, save_file
    sd = load_file(str(tmp_path / "again" / "model.safetensors"))
    sd.pop("text_model.final_layer_norm.weight")
    save_file(sd, str(tmp_path / "again" / "model.safetensors"))
    with pytest.raises(KeyError, match="final_layer_norm"):
        CLIPTextModel.from_pretrained(str(tmp_path / "again"))


def _vae_sd():
    from dualdiff_b200 import synthetic as S
    from oracle import vae_oracle as V
    return S.init_state_dict(V.manifest(), seed=4)


def test_vae_decoder_loads_an_old_style_autoencoderkl_checkpoint(tmp_path):
    from dualdiff_b200.networks import AutoencoderKLDecoder
    sd = _vae_sd()
    old = {}
    for k, v in sd.items():     # SD-v1.5 vae/diffusion_pytorch_model.bin naming + tensors of the encode half
        for new, o in ((".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn.")):
            k = k.replace(new, o)
        old[k] = v.half()
    old["encoder.conv_in.weight"] = torch.zeros(128, 3, 3, 3)
    old["quant_conv.weight"] = torch.zeros(8, 8, 1, 1)
    d = tmp_path / "vae"
    os.makedirs(d)
    torch.save(old, d / "diffusion_pytorch_model.bin")
    cfg = {"_class_name": "AutoencoderKL", "act_fn": "silu", "block_out_channels": [128, 256, 512, 512], "in_channels": 3,
           "latent_channels": 4, "layers_per_block": 2, "norm_num_groups": 32, "out_channels": 3, "sample_size": 512,
           "down_block_types": ["DownEncoderBlock2D"] * 4, "up_block_types": ["UpDecoderBlock2D"] * 4}
    with open(d / "config.json", "w") as fh:
        json.dump(cfg, fh)
    vae = AutoencoderKLDecoder.from_pretrained(str(tmp_path), subfolder="vae", torch_dtype=torch.float16)
    got = vae.state_dict()
    assert vae.scaling_factor == 0.18215 and set(got) == set(sd)
    assert all(torch.equal(got[k], sd[k].half().float()) for k in sd)
    with open(d / "config.json", "w") as fh:
        json.dump(dict(cfg, block_out_channels=[128, 256, 512]), fh)
    with pytest.raises(NotImplementedError, match="block_out_channels"):
        AutoencoderKLDecoder.from_pretrained(str(d))


def test_vae_decoder_save_load_roundtrip(tmp_path):
    from dualdiff_b200.networks import AutoencoderKLDecoder
    sd = _vae_sd()
    with torch.device("meta"):
        vae = AutoencoderKLDecoder(scaling_factor=0.2)
    vae.load_state_dict(sd, strict=True, assign=True)
    vae.save_pretrained(tmp_path / "vae")
    again = AutoencoderKLDecoder.from_pretrained(str(tmp_path / "vae"))
    assert again.scaling_factor == 0.2 and all(torch.equal(again.state_dict()[k], sd[k]) for k in sd)


def test_scheduler_from_config():
    from dualdiff_b200 import scheduler as SCH
    u = SCH.UniPCMultistepScheduler.from_config(PNDM)            # misc/test_utils.py:162
    assert u.config.solver_type == "bh2" and u.config["beta_start"] == 0.00085 and u.order == 1
    assert isinstance(SCH.UniPCMultistepScheduler.from_config(u.config), SCH.UniPCMultistepScheduler)
    assert isinstance(SCH.DDIMScheduler.from_config(u.config), SCH.DDIMScheduler)
    for bad in (dict(PNDM, beta_schedule="linear"), dict(PNDM, prediction_type="v_prediction"), dict(u.config, solver_order=3)):
        with pytest.raises(NotImplementedError):
            SCH.UniPCMultistepScheduler.from_config(bad)


def test_pipeline_from_pretrained(tmp_path):
    from dualdiff_b200 import scheduler as SCH
    from dualdiff_b200.networks import AutoencoderKLDecoder, CLIPTextModel
    from dualdiff_b200.pipeline_bev_controlnet import MultiControlNetModel, StableDiffusionBEVControlNetPipeline
    from oracle import clip_oracle as CO
    enc = CLIPTextModel(**SMALL, eos_token_id=99)
    torch.manual_seed(1)
    for p in enc.parameters():
        torch.nn.init.normal_(p, std=0.02)
    enc.save_pretrained(tmp_path / "text_encoder")
    with torch.device("meta"):
        vae = AutoencoderKLDecoder()
    vae.load_state_dict(_vae_sd(), strict=True, assign=True)
    vae.save_pretrained(tmp_path / "vae")
    os.makedirs(tmp_path / "scheduler")
    with open(tmp_path / "scheduler" / "scheduler_config.json", "w") as fh:
        json.dump(PNDM, fh)
    unet, nets = torch.nn.Identity(), [torch.nn.Identity(), torch.nn.Identity()]
    pipe = StableDiffusionBEVControlNetPipeline.from_pretrained(str(tmp_path), unet=unet, controlnet=nets, safety_checker=None,
                                                                feature_extractor=None, torch_dtype=torch.float16,
                                                                tokenizer=CO.HashTokenizer(100, 16))
    assert isinstance(pipe.controlnet, MultiControlNetModel) and list(pipe.controlnet.nets) == nets and pipe.unet is unet
    assert isinstance(pipe.scheduler, SCH.UniPCMultistepScheduler) and pipe.vae.scaling_factor == 0.18215
    assert torch.equal(pipe.text_encoder.state_dict()["text_model.final_layer_norm.bias"],
                       enc.state_dict()["text_model.final_layer_norm.bias"])
    pipe.scheduler = SCH.UniPCMultistepScheduler.from_config(pipe.scheduler.config)       # what build_pipe does next
    with pytest.raises(ValueError, match="unet"):
        StableDiffusionBEVControlNetPipeline.from_pretrained(str(tmp_path))
