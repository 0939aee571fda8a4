"""Oracle (test infrastructure, CPU fp32) for the VAE decode that follows the sampler loop -- SURVEY.md §8f rank 1.

The reference calls diffusers' `AutoencoderKL.decode` (pipeline/pipeline_bev_controlnet.py:101-113, `decode_latents`:
`latents = 1 / 0.18215 * latents; image = vae.decode(latents).sample; image = (image / 2 + 0.5).clamp(0, 1)`).  diffusers
(0.17.1, MagicDrive fork) is not vendored under /root/reference, so this restates the public library code
(models/autoencoder_kl.py, models/vae.py:Decoder, unet_2d_blocks.py:UNetMidBlock2D / UpDecoderBlock2D, resnet.py,
attention_processor.py:Attention) for the SD-v1.5 VAE configuration.  Pinning: diffusers itself cannot be run here, but
AutoencoderKL's decoder is a port of the CompVis/LDM decoder, and an INDEPENDENT implementation of that architecture is
installed in this image (torchtitan's flux autoencoder `Decoder`): with the SD geometry (ch 128, ch_mult (1,2,4,4), 3 resnets
per level, z_channels 4) and the same weights through the diffusers<->LDM key map it reproduces this oracle to 2e-7
(oracle/make_golden_vae.py -> tests/golden/vae_small.pt; tests/test_vae.py, live and against the golden).  `post_quant_conv`
and the 0.18215 scaling are diffusers-specific and restated; the parameter count (49,490,179 + 20) is a second structure pin.
State-dict keys are the diffusers ones (post_quant_conv.*, decoder.conv_in.*, decoder.mid_block.*, decoder.up_blocks.*,
decoder.conv_norm_out.*, decoder.conv_out.*; attention as to_q / to_k / to_v / to_out.0 / group_norm)."""
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BLOCK_OUT = (128, 256, 512, 512)     # SD-v1.5 VAE block_out_channels; the decoder walks them reversed
LAYERS_PER_BLOCK = 2                 # decoder uses layers_per_block + 1 = 3 resnets per up block
GROUPS, EPS, LATENT_CH, OUT_CH, SCALING = 32, 1e-6, 4, 3, 0.18215


def manifest() -> Dict[str, tuple]:
    """key -> shape of the decoder half of diffusers' AutoencoderKL (SD-v1.5 configuration)"""
    m = {}

    def conv(p, ci, co, k):
        m[p + ".weight"] = (co, ci, k, k); m[p + ".bias"] = (co,)

    def norm(p, c):
        m[p + ".weight"] = (c,); m[p + ".bias"] = (c,)

    def lin(p, ci, co):
        m[p + ".weight"] = (co, ci); m[p + ".bias"] = (co,)

    def resnet(p, ci, co):
        norm(p + ".norm1", ci); conv(p + ".conv1", ci, co, 3); norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", ci, co, 1)

    conv("post_quant_conv", LATENT_CH, LATENT_CH, 1)
    top = BLOCK_OUT[-1]
    conv("decoder.conv_in", LATENT_CH, top, 3)
    resnet("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v"):
        lin(f"{a}.{n}", top, top)
    lin(a + ".to_out.0", top, top)
    resnet("decoder.mid_block.resnets.1", top, top)
    rev = BLOCK_OUT[::-1]
    prev = rev[0]
    for i, co in enumerate(rev):
        for j in range(LAYERS_PER_BLOCK + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else co, co)
        if i < len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        prev = co
    norm("decoder.conv_norm_out", BLOCK_OUT[0])
    conv("decoder.conv_out", BLOCK_OUT[0], OUT_CH, 3)
    return m


def _conv(sd, p, x, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def _gn(sd, p, x):
    return F.group_norm(x, GROUPS, sd[p + ".weight"], sd[p + ".bias"], EPS)


def resnet(sd: SD, p: str, x):
    """ResnetBlock2D(temb_channels=None, output_scale_factor=1)"""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h)))
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def mid_attention(sd: SD, p: str, x):
    """Attention(heads=1, dim_head=C, norm_num_groups=32, residual_connection=True, bias=True)"""
    n, c, h, w = x.shape
    t = _gn(sd, p + ".group_norm", x).reshape(n, c, h * w).transpose(1, 2)          # [n, HW, C]
    q = F.linear(t, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(t, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(t, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1) @ v
    o = F.linear(a, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(n, c, h, w)


def decode(sd: SD, z):
    """AutoencoderKL.decode(z).sample: z [n, 4, h, w] -> [n, 3, 8h, 8w]"""
    x = _conv(sd, "post_quant_conv", z, padding=0)
    x = _conv(sd, "decoder.conv_in", x)
    x = resnet(sd, "decoder.mid_block.resnets.0", x)
    x = mid_attention(sd, "decoder.mid_block.attentions.0", x)
    x = resnet(sd, "decoder.mid_block.resnets.1", x)
    for i in range(len(BLOCK_OUT)):
        for j in range(LAYERS_PER_BLOCK + 1):
            x = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i < len(BLOCK_OUT) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(_gn(sd, "decoder.conv_norm_out", x))
    return _conv(sd, "decoder.conv_out", x)


def decode_latents(sd: SD, latents):
    """pipeline_bev_controlnet.py decode_latents (inherited from StableDiffusionPipeline): scale, decode, to [0, 1]"""
    image = decode(sd, latents / SCALING)
    return (image / 2 + 0.5).clamp(0, 1)
