"""Writes tests/golden/vae_small.pt: a latent and the image an INDEPENDENT implementation of the SD VAE decoder produces for
it -- torchtitan's `experiments.flux.model.autoencoder.Decoder` (the original CompVis/LDM decoder architecture that
diffusers' AutoencoderKL ports; instantiated with the SD-v1.5 geometry ch=128, ch_mult=(1,2,4,4), num_res_blocks=2,
z_channels=4), loaded with the same seeded synthetic weights through the diffusers<->LDM key map below, with diffusers'
`post_quant_conv` applied in front.  diffusers itself is not installed here, so this is what pins oracle/vae_oracle.py.
Run:  python oracle/make_golden_vae.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ldm_key(k: str) -> str:
    """diffusers AutoencoderKL decoder key -> CompVis/LDM decoder key (the map of diffusers' convert_vae_pt_to_diffusers)"""
    assert k.startswith("decoder.")
    k = k[len("decoder."):]
    k = k.replace("mid_block.resnets.0.", "mid.block_1.").replace("mid_block.resnets.1.", "mid.block_2.")
    k = k.replace("mid_block.attentions.0.group_norm.", "mid.attn_1.norm.")
    for a, b in (("to_q", "q"), ("to_k", "k"), ("to_v", "v"), ("to_out.0", "proj_out")):
        k = k.replace(f"mid_block.attentions.0.{a}.", f"mid.attn_1.{b}.")
    if k.startswith("up_blocks."):
        p = k.split(".")
        lvl = 3 - int(p[1])                      # LDM stores the up path lowest resolution last
        k = f"up.{lvl}.block.{p[3]}." + ".".join(p[4:]) if p[2] == "resnets" else f"up.{lvl}.upsample.conv." + ".".join(p[5:])
    return k.replace("conv_shortcut.", "nin_shortcut.").replace("conv_norm_out.", "norm_out.")


def ldm_decode(sd, z):
    """decode with torchtitan's LDM decoder (raises ImportError where torchtitan is absent)"""
    from torchtitan.experiments.flux.model.autoencoder import Decoder
    dec = Decoder(ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, in_channels=3, resolution=256, z_channels=4).eval()
    own = dec.state_dict()
    dec.load_state_dict({ldm_key(k): v.float().reshape(own[ldm_key(k)].shape) for k, v in sd.items() if k.startswith("decoder.")},
                        strict=True)
    with torch.no_grad():
        pq = torch.nn.functional.conv2d(z, sd["post_quant_conv.weight"].float(), sd["post_quant_conv.bias"].float())
        return dec(pq)


def main():
    from dualdiff_b200 import synthetic as S
    from oracle import vae_oracle as V
    sd = S.init_state_dict(V.manifest(), seed=4)
    z = torch.randn(1, 4, 4, 6, generator=torch.Generator().manual_seed(2))
    out = ldm_decode(sd, z)
    path = os.path.join(ROOT, "tests", "golden", "vae_small.pt")
    torch.save({"seed": 4, "z": z, "image": out.clone(), "producer": "torchtitan flux Decoder (LDM architecture), z_channels=4"}, path)
    print("wrote", path, tuple(out.shape), float(out.abs().max()))


if __name__ == "__main__":
    main()
