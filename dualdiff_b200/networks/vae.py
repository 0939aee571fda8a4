"""AutoencoderKLDecoder — the VAE decode that follows the sampler loop, on the B200 kernels.

Reference call site: pipeline/pipeline_bev_controlnet.py:101-113 (`decode_latents`: `latents / 0.18215`,
`self.vae.decode(latents).sample`, `(image / 2 + 0.5).clamp(0, 1)`); the network is diffusers' `AutoencoderKL`
(SD-v1.5 VAE), which is not vendored in the reference tree -- oracle/vae_oracle.py restates it (pinned to an
independent LDM decoder implementation, see its header).
This module keeps the decode half: same state-dict keys (`post_quant_conv.*`, `decoder.*`, attention as
`to_q / to_k / to_v / to_out.0 / group_norm`), `decode(z).sample`, `decode_latents(latents)`.

Everything runs on the hot-path kernels: 3x3 convs as implicit GEMM over the zero-haloed layout (GroupNorm+SiLU and the
nearest-2x upsample write that layout directly), 1x1 shortcuts / linears as GEMMs with the residual fused.  The single
512-wide attention head of the mid block does not fit the flash kernel (TMEM: S 128 + O 512 + P 64 columns > 512), so it
runs per image as S = Q K^T (fp32 out) -> dd_softmax_rows -> (P X) W_v^T, with V^T produced directly by a GEMM with the
operands swapped (no transpose) and the value bias folded into the output projection (rows of P sum to one).
Images are decoded in chunks (default 6 = one scene) to bound the 224x400x128-channel activations."""
from typing import Dict, Optional

import torch
import torch.nn as nn

from dataclasses import dataclass

from .output_cls import _Output

BF = torch.bfloat16
BLOCK_OUT = (128, 256, 512, 512)
SCALING_FACTOR = 0.18215


@dataclass
class DecoderOutput(_Output):
    sample: torch.Tensor


def _manifest() -> Dict[str, tuple]:
    m = {}

    def conv(p, ci, co, k):
        m[p + ".weight"] = (co, ci, k, k); m[p + ".bias"] = (co,)

    def vec(p, c):
        m[p + ".weight"] = (c,); m[p + ".bias"] = (c,)

    def lin(p, ci, co):
        m[p + ".weight"] = (co, ci); m[p + ".bias"] = (co,)

    def resnet(p, ci, co):
        vec(p + ".norm1", ci); conv(p + ".conv1", ci, co, 3); vec(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", ci, co, 1)

    conv("post_quant_conv", 4, 4, 1)
    top = BLOCK_OUT[-1]
    conv("decoder.conv_in", 4, top, 3)
    resnet("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    vec(a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{n}", top, top)
    resnet("decoder.mid_block.resnets.1", top, top)
    prev = top
    for i, co in enumerate(BLOCK_OUT[::-1]):
        for j in range(3):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else co, co)
        if i < 3:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        prev = co
    vec("decoder.conv_norm_out", BLOCK_OUT[0])
    conv("decoder.conv_out", BLOCK_OUT[0], 3, 3)
    return m


class _Node(nn.Module):
    """bare container: gives the flat diffusers key list a module tree (state_dict / load_state_dict / .to work)"""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container only")


class AutoencoderKLDecoder(nn.Module):
    def __init__(self, scaling_factor: float = SCALING_FACTOR, images_per_chunk: int = 6):
        super().__init__()
        self.scaling_factor, self.images_per_chunk = scaling_factor, images_per_chunk
        for key, shape in _manifest().items():
            node = self
            *path, leaf = key.split(".")
            for part in path:
                if not hasattr(node, part):
                    node.add_module(part, _Node())
                node = getattr(node, part)
            node.register_parameter(leaf, nn.Parameter(torch.empty(shape)))
        self._packed = None

    # ---- checkpoints (diffusers layout: config.json + diffusion_pytorch_model.safetensors | .bin) -----------------
    config_name = "config.json"
    weights_names = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin")
    # diffusers <= 0.17 AttentionBlock names (the SD-v1.5 VAE files on disk) -> Attention names kept by this mirror
    _OLD_ATTN = ((".query.", ".to_q."), (".key.", ".to_k."), (".value.", ".to_v."), (".proj_attn.", ".to_out.0."))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **kwargs):
        """load `<sd-v1-5>/vae` (an `AutoencoderKL` checkpoint) and keep its decode half: `decoder.*` and
        `post_quant_conv.*`; `encoder.*` / `quant_conv.*` are ignored.  Old-style attention keys are renamed."""
        import json
        import os
        root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        with open(os.path.join(root, cls.config_name)) as fh:
            cfg = json.load(fh)
        want = dict(block_out_channels=list(BLOCK_OUT), latent_channels=4, out_channels=3, layers_per_block=2,
                    norm_num_groups=32, act_fn="silu")
        for k, v in want.items():
            have = cfg.get(k, v)
            have = list(have) if isinstance(have, (list, tuple)) else have
            if have != v:
                raise NotImplementedError(f"AutoencoderKLDecoder: {k}={have!r}; only the SD-v1.5 VAE geometry ({v!r}) is built")
        model = cls(scaling_factor=cfg.get("scaling_factor", SCALING_FACTOR))
        sd = None
        for name in cls.weights_names:
            path = os.path.join(root, name)
            if os.path.exists(path):
                if name.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd = load_file(path)
                else:
                    sd = torch.load(path, map_location="cpu", weights_only=True)
                break
        if sd is None:
            raise FileNotFoundError(f"no {' / '.join(cls.weights_names)} under {root}")
        ren = {}
        for k, v in sd.items():
            if ".attentions." in k:
                for old, new in cls._OLD_ATTN:
                    k = k.replace(old, new)
            ren[k] = v
        own = model.state_dict()
        missing = [k for k in own if k not in ren or ren[k].numel() != own[k].numel()]
        if missing:
            raise KeyError(f"{root}: decoder tensors missing or mis-shaped: {missing[:4]}{' ...' if len(missing) > 4 else ''}")
        # (old AttentionBlock checkpoints store the projections as [C, C]; 1x1-conv variants as [C, C, 1, 1])
        model.load_state_dict({k: ren[k].float().reshape(own[k].shape) for k in own}, strict=True)
        return model.eval()

    def save_pretrained(self, save_directory, safe_serialization=True, **kwargs):
        import json
        import os
        os.makedirs(save_directory, exist_ok=True)
        cfg = dict(_class_name="AutoencoderKL", _diffusers_version="0.17.1", act_fn="silu", block_out_channels=list(BLOCK_OUT),
                   in_channels=3, out_channels=3, latent_channels=4, layers_per_block=2, norm_num_groups=32,
                   scaling_factor=self.scaling_factor)
        with open(os.path.join(save_directory, self.config_name), "w") as fh:
            json.dump(cfg, fh, indent=2)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, self.weights_names[0]))
        else:
            torch.save(sd, os.path.join(save_directory, self.weights_names[1]))

    # ---- packing ----------------------------------------------------------------------------------------
    def pack(self, device=None):
        from .. import engine
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: move the VAE decoder to a CUDA device before use")
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        pk = engine.Packer(sd, device)
        f = lambda k: sd[k].float()
        # post_quant_conv (1x1, 4 -> 4) as a centre-tap 3x3 over the 8-channel padded latent layout
        w = torch.zeros(8, 8, 3, 3)
        w[:4, :4, 1, 1] = f("post_quant_conv.weight")[:, :, 0, 0].cpu()
        from ..packing import pack_conv3x3
        pk.put("pq.w", pack_conv3x3(w))
        b = torch.zeros(8); b[:4] = f("post_quant_conv.bias").cpu()
        pk.put("pq.b", pk.f32(b))
        pk.conv3("decoder.conv_in", pad_cin_to=8)
        resnets = ["decoder.mid_block.resnets.0", "decoder.mid_block.resnets.1"] + \
                  [f"decoder.up_blocks.{i}.resnets.{j}" for i in range(4) for j in range(3)]
        for p in resnets:
            pk.resnet(p, [])
        for i in range(3):
            pk.conv3(f"decoder.up_blocks.{i}.upsamplers.0.conv")
        a = "decoder.mid_block.attentions.0"
        pk.norm(a + ".group_norm")
        pk.lin(a + ".to_q"); pk.lin(a + ".to_k")
        pk.put(a + ".to_v.w", f(a + ".to_v.weight").to(BF))
        wo, bo, bv = f(a + ".to_out.0.weight").double(), f(a + ".to_out.0.bias").double(), f(a + ".to_v.bias").double()
        pk.put(a + ".to_out.0.w", wo.float().to(BF))
        pk.put(a + ".to_out.0.b", pk.f32(wo @ bv + bo))            # softmax rows sum to 1: the value bias passes through
        pk.norm("decoder.conv_norm_out")
        pk.conv3("decoder.conv_out")
        self._packed = pk.out
        return self

    # ---- forward ----------------------------------------------------------------------------------------
    def _resnet(self, P, p, x, n, hw):
        from .. import ops
        g1 = ops.groupnorm(x, P[p + ".norm1.g"], P[p + ".norm1.b"], n_img=n, hw=hw, eps=1e-6, silu=True, padded_out=True)
        h = ops.gemm(g1, P[p + ".conv1.w"], bias=P[p + ".conv1.b"], taps=9, conv_hw=hw, n_img=n)
        g2 = ops.groupnorm(h, P[p + ".norm2.g"], P[p + ".norm2.b"], n_img=n, hw=hw, eps=1e-6, silu=True, padded_out=True)
        res = x
        if (p + ".conv_shortcut.w") in P:
            res = ops.gemm(x, P[p + ".conv_shortcut.w"], bias=P[p + ".conv_shortcut.b"])
        return ops.gemm(g2, P[p + ".conv2.w"], bias=P[p + ".conv2.b"], taps=9, conv_hw=hw, n_img=n, res1=res)

    def _mid_attention(self, P, p, x, n, hw):
        from .. import ops
        T, C = hw[0] * hw[1], x.shape[1]
        if T % 8 != 0:
            raise ValueError(f"latent size {hw} gives {T} tokens; the mid-block attention needs a multiple of 8")
        t = ops.groupnorm(x, P[p + ".group_norm.g"], P[p + ".group_norm.b"], n_img=n, hw=hw, eps=1e-6, silu=False)
        q = ops.gemm(t, P[p + ".to_q.w"], bias=P[p + ".to_q.b"])
        k = ops.gemm(t, P[p + ".to_k.w"], bias=P[p + ".to_k.b"])
        o = torch.empty((n * T, C), device=x.device, dtype=BF)
        for i in range(n):
            rows = slice(i * T, (i + 1) * T)
            vt = ops.gemm(P[p + ".to_v.w"], t[rows])                                   # V^T (without bias): [C, T]
            s = ops.gemm(q[rows], k[rows], out_f32=True)                               # [T, T] fp32 scores
            pr = ops.softmax_rows(s, float(C) ** -0.5)
            ops.gemm(pr, vt, out=o[rows])                                              # P V
        return ops.gemm(o, P[p + ".to_out.0.w"], bias=P[p + ".to_out.0.b"], res1=x)

    def _decode_chunk(self, z):
        from .. import ops
        P = self._packed
        n, c, h, w = z.shape
        zp = ops.nchw_to_padded(z.contiguous(), n_outer=1, n_view=n, c=4, h=h, w=w, cp=8, stride_outer=0,
                                stride_view=4 * h * w, stride_c=h * w, stride_h=w)
        x = ops.gemm(zp, P["pq.w"], bias=P["pq.b"], taps=9, conv_hw=(h, w), n_img=n)                    # post_quant_conv
        x = ops.gemm(ops.pad_rows(x, n_img=n, hw=(h, w)), P["decoder.conv_in.w"], bias=P["decoder.conv_in.b"], taps=9,
                     conv_hw=(h, w), n_img=n)
        x = self._resnet(P, "decoder.mid_block.resnets.0", x, n, (h, w))
        x = self._mid_attention(P, "decoder.mid_block.attentions.0", x, n, (h, w))
        x = self._resnet(P, "decoder.mid_block.resnets.1", x, n, (h, w))
        hw = (h, w)
        for i in range(4):
            for j in range(3):
                x = self._resnet(P, f"decoder.up_blocks.{i}.resnets.{j}", x, n, hw)
            if i < 3:
                hw2 = (2 * hw[0], 2 * hw[1])
                pad = ops.upsample_pad(x, n_img=n, hw=hw, hw2=hw2)
                p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                x = ops.gemm(pad, P[p + ".w"], bias=P[p + ".b"], taps=9, conv_hw=hw2, n_img=n)
                hw = hw2
        g = ops.groupnorm(x, P["decoder.conv_norm_out.g"], P["decoder.conv_norm_out.b"], n_img=n, hw=hw, eps=1e-6, silu=True,
                          padded_out=True)
        rows = ops.gemm(g, P["decoder.conv_out.w"], bias=P["decoder.conv_out.b"], taps=9, conv_hw=hw, n_img=n, out_f32=True)
        return ops.rows_to_nchw(rows, n, hw, out_dtype=torch.float32)

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """z: [n, 4, h, w] (already divided by the scaling factor) -> DecoderOutput(sample [n, 3, 8h, 8w] fp32)"""
        if not z.is_cuda:
            raise RuntimeError("dualdiff_b200 has no CPU path: `z` must be a CUDA tensor")
        if self._packed is None:
            self.pack(z.device)
        z = z.float()
        outs = [self._decode_chunk(z[i:i + self.images_per_chunk]) for i in range(0, z.shape[0], self.images_per_chunk)]
        out = torch.cat(outs) if len(outs) > 1 else outs[0]
        return DecoderOutput(sample=out) if return_dict else (out,)

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """pipeline `decode_latents`: scale, decode, map to [0, 1] (the affine map and clamp are torch elementwise ops)"""
        image = self.decode(latents / self.scaling_factor).sample
        return (image / 2 + 0.5).clamp_(0, 1)
