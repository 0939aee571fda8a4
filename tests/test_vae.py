"""VAE decode: oracle structure check on CPU, CUDA path vs the oracle on the GPU.
diffusers is not vendored under /root/reference and not installed: the oracle restates AutoencoderKL's decoder and is pinned
to an INDEPENDENT implementation of the same (CompVis/LDM) decoder that is present in this image -- torchtitan's flux
autoencoder `Decoder` with the SD geometry -- live and through tests/golden/vae_small.pt, plus the parameter count.  Tolerance (bf16 kernels vs fp32 oracle, stated):
cosine >= 0.999, rel-L2 <= 2e-2 on the decoded image."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_oracle_manifest_is_the_sd_vae_decoder():
    from oracle import vae_oracle as V
    m = V.manifest()
    # AutoencoderKL(SD-v1.5): decoder 49,490,179 parameters + post_quant_conv 4*4+4
    assert sum(math.prod(s) for s in m.values()) == 49_490_179 + 20
    assert m["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"] == (256, 512, 1, 1)
    assert m["decoder.mid_block.attentions.0.to_q.weight"] == (512, 512)
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in m


def test_oracle_matches_an_independent_ldm_decoder_golden():
    """tests/golden/vae_small.pt was produced by torchtitan's LDM decoder (oracle/make_golden_vae.py), not by the oracle"""
    from dualdiff_b200 import synthetic as S
    from oracle import vae_oracle as V
    g = torch.load(os.path.join(ROOT, "tests", "golden", "vae_small.pt"))
    sd = S.init_state_dict(V.manifest(), seed=g["seed"])
    with torch.no_grad():
        out = V.decode(sd, g["z"])
    assert out.shape == g["image"].shape and (out - g["image"]).abs().max() < 1e-5 * max(1.0, float(g["image"].abs().max()))


def test_oracle_matches_an_independent_ldm_decoder_live():
    pytest.importorskip("torchtitan")
    from dualdiff_b200 import synthetic as S
    from oracle import make_golden_vae as MG, vae_oracle as V
    sd = S.init_state_dict(V.manifest(), seed=9)
    z = torch.randn(2, 4, 6, 8, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        out = V.decode(sd, z)
    ref = MG.ldm_decode(sd, z)
    assert (out - ref).abs().max() < 1e-5 * max(1.0, float(ref.abs().max()))


def test_product_module_has_the_oracle_keys():
    from dualdiff_b200.networks import AutoencoderKLDecoder
    from oracle import vae_oracle as V
    with torch.device("meta"):
        vae = AutoencoderKLDecoder()
    assert {k: tuple(v.shape) for k, v in vae.state_dict().items()} == V.manifest()


def test_oracle_decode_shapes_and_range():
    from dualdiff_b200 import synthetic as S
    from oracle import vae_oracle as V
    sd = S.init_state_dict(V.manifest(), seed=4)
    z = torch.randn(1, 4, 4, 6, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        img = V.decode_latents(sd, z * 0.18215)
    assert img.shape == (1, 3, 32, 48) and img.min() >= 0 and img.max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w", [(2, 8, 12), (7, 4, 6)])
def test_vae_decode_matches_oracle(n, h, w):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from dualdiff_b200 import synthetic as S
    from dualdiff_b200.networks import AutoencoderKLDecoder
    from oracle import vae_oracle as V
    sd = S.init_state_dict(V.manifest(), seed=4)
    with torch.device("meta"):
        vae = AutoencoderKLDecoder()
    vae.load_state_dict(sd, strict=True, assign=True)
    z = torch.randn(n, 4, h, w, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = V.decode(sd, z)
    out = vae.to("cuda:0").decode(z.cuda()).sample.cpu()
    assert out.shape == (n, 3, 8 * h, 8 * w)
    m = common.metrics(out, ref)
    assert m["cos"] > 0.999 and m["rel_l2"] < 2e-2, m
    img = vae.decode_latents(z.cuda() * 0.18215).cpu()
    with torch.no_grad():
        ref_img = V.decode_latents(sd, z * 0.18215)
    assert (img - ref_img).abs().max() < 5e-2 and img.min() >= 0 and img.max() <= 1
