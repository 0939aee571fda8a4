"""Deterministic synthetic weights and inputs for the DualDiff denoising step.

There is no network access for checkpoints or nuScenes, so parity tests and bench.py use random-init
SDv1.5-shaped weights and synthetic latents/conditions.  Weights are generated *per state-dict key* from
a CPU generator seeded with crc32(key), so the product modules, the oracle and the reference's own classes
get bit-identical tensors regardless of module construction order or of which box generates them.
Zero-initialised modules of the reference (13 zero convs per branch, cond `conv_out`, attn4 `connector`;
unet_addon_rawbox.py:230-281, map_embedder.py:110-112, blocks.py:83) are re-randomised N(0, 0.02) — otherwise
both branches, SFA and cross-view attention contribute exactly 0 and parity would be vacuous.

Conditioning note: the reference feeds *unscaled* camera intrinsics (~1.3e3) through an include-input Fourier
embedding into `cam2token` (unet_addon_rawbox.py:308-349).  With PyTorch-default init that makes the camera
token ~60x larger than the text tokens, every text cross-attention softmax saturates and the random-init
network becomes chaotic (measured: a 1e-6 relative input perturbation moves the fp32 output by 2e-3; fp32 vs
fp64 differ by 7e-4), so no reduced-precision implementation - the reference's own fp16 path included - could
be compared meaningfully.  `cam2token.weight` is therefore drawn 100x smaller, which puts the camera token at
O(1) as in a trained model (then: 1e-4 perturbation -> 2e-4 output change; bf16 weights -> cosine 0.99998).
"""
import math
import zlib
from typing import Dict, Tuple

import torch

ZERO_INIT_MARKERS = ("controlnet_down_blocks.", "controlnet_mid_block.", "controlnet_cond_embedding.conv_out.",
                     ".connector.")


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def init_tensor(key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = _gen(key, seed)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if any(m in key for m in ZERO_INIT_MARKERS):
        return torch.randn(shape, generator=g) * 0.02
    if key.endswith("_class_tokens") or key.startswith("uncond_cam"):
        return torch.randn(shape, generator=g)
    if "null_" in key:
        return torch.randn(shape, generator=g) * 0.02
    if key.endswith("_embedding.weight"):   # CLIP token / position tables (transformers init: N(0, 0.02))
        return torch.randn(shape, generator=g) * 0.02
    is_norm = ".norm" in key or key.startswith("conv_norm_out") or ".norm." in key or "layer_norm" in key
    if is_norm and len(shape) == 1:
        if leaf == "weight":
            return 1.0 + 0.05 * torch.randn(shape, generator=g)
        return 0.05 * torch.randn(shape, generator=g)
    if leaf == "weight":
        fan_in = int(math.prod(shape[1:])) if len(shape) > 1 else shape[0]
        bound = 1.0 / math.sqrt(fan_in)
        if key == "cam2token.weight":
            bound *= 0.01  # see the conditioning note in the module docstring
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if leaf == "bias":
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return torch.randn(shape, generator=g) * 0.02


def init_state_dict(manifest: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: init_tensor(k, tuple(s), seed) for k, s in manifest.items()}


def manifest_of(module: torch.nn.Module) -> Dict[str, Tuple[int, ...]]:
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def manifest_digest(manifest: Dict[str, Tuple[int, ...]]) -> str:
    import hashlib
    h = hashlib.sha256()
    for k in sorted(manifest):
        h.update(f"{k}:{tuple(manifest[k])};".encode())
    return h.hexdigest()[:16]


def make_inputs(B: int = 1, h: int = 28, w: int = 50, seed: int = 1, n_cam: int = 6, L_bg: int = 28, L_fg: int = 32,
                same_noise_across_views: bool = True):
    """Synthetic step inputs with the reference's layouts (dataset/utils.py:390-445,463-491)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    r = lambda *s: torch.randn(*s, generator=g)
    u = lambda *s: torch.rand(*s, generator=g)
    if same_noise_across_views:  # pipeline_bev_controlnet.py:345 stacks one latent over the views
        lat = r(B, 1, 4, h, w).expand(B, n_cam, 4, h, w).contiguous()
    else:
        lat = r(B, n_cam, 4, h, w)
    prompt = r(2 * B, 77, 768)  # uncond first
    # camera_param = cat[K (3x3), camera2lidar[:3, :4]]  (dataset/utils.py:434-437), unscaled intrinsics
    K = torch.tensor([[1266.0, 0, 816.0], [0, 1266.0, 491.0], [0, 0, 1]]).expand(B, n_cam, 3, 3).clone()
    K[..., :2, :] *= 1 + 0.05 * (u(B, n_cam, 2, 3) * 2 - 1)
    q, _ = torch.linalg.qr(r(B, n_cam, 3, 3))
    trans = (u(B, n_cam, 3, 1) * 4 - 2)
    cam = torch.cat([K, q, trans], dim=-1)  # (B, 6, 3, 7)
    boxes_bg = {
        "bboxes": u(B, n_cam, L_bg, 8, 3) * 100 - 50,
        "classes": torch.randint(0, 10, (B, n_cam, L_bg), generator=g),
        "masks": u(B, n_cam, L_bg) < 0.7,
    }
    vec = u(B, 1, L_fg, 8, 3) * 100 - 50
    vec[..., 2] = 0
    boxes_fg = {
        "bboxes": vec,
        "classes": torch.randint(0, 3, (B, 1, L_fg), generator=g),
        "masks": u(B, 1, L_fg) < 0.7,
    }
    cond_bg = u(B, 3, 8 * h, n_cam * 8 * w)  # occupancy-projection panorama in [0,1]
    ids = torch.randint(0, 17, (B * n_cam, 320, h, w), generator=g)
    ids[u(B * n_cam, 320, h, w) < 0.8] = 17
    cond_fg = ids.float() / 17.0  # ORS class ids / 17  (dataset/utils.py:412-420)
    return {"latents": lat, "prompt_embeds": prompt, "camera_param": cam, "boxes_bg": boxes_bg,
            "boxes_fg": boxes_fg, "cond_bg": cond_bg, "cond_fg": cond_fg}
