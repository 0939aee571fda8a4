"""BEVControlNetModel — one ControlNet-style condition branch of DualDiff (bg: occupancy-projection image,
fg: Occupancy-Ray-shape-Sampling tensor), each fused with the text tokens by Semantic Fusion Attention.

Reference: networks/unet_addon_rawbox.py:39-1082 (ctor :43-88, forward :794-811, uncond helpers :327-335,
:671-769).  Same constructor keywords, post-construction attributes set by the caller
(misc/test_utils.py:123-136: use_cam_in_temb / adm_proj / use_txt_con_fusion(p) / txt_con_fusion(p) /
use_occ_3d / controlnet_cond_embedding / use_box_adapter), forward signature and state-dict keys.  The forward
runs on the B200 engine; training-time condition dropping (drop_cond_ratio, :839-846) is out of scope.
"""
import importlib
import logging
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import _tree
from .embedder import get_embedder
from .output_cls import BEVControlNetOutput
from .txt_con_fusion import txt_con_XFormersAttn, txt_con_XFormersAttn_plus
from .unet_2d_condition_multiview import nchw_view

# the reference selects embedder classes by dotted path (misc/common.py:11-15); its own paths map to our mirrors
_ALIASES = {
    "magicdrive.networks.bbox_embedder.ContinuousBBoxWithTextEmbedding":
        "dualdiff_b200.networks.bbox_embedder.ContinuousBBoxWithTextEmbedding",
    "magicdrive.networks.map_embedder.ControlNetConditioningEmbedding":
        "dualdiff_b200.networks.map_embedder.ControlNetConditioningEmbedding",
}


def load_module(name):
    name = _ALIASES.get(name, name)
    p, m = name.rsplit(".", 1)
    return getattr(importlib.import_module(p), m)


class BEVControlNetModel(_tree.ModelBase):
    _supports_gradient_checkpointing = True

    def __init__(self, in_channels: int = 4, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 down_block_types: Tuple[str] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                                 "CrossAttnDownBlock2D", "DownBlock2D"),
                 only_cross_attention: Union[bool, Tuple[bool]] = False,
                 block_out_channels: Tuple[int] = (320, 640, 1280, 1280), layers_per_block: int = 2,
                 downsample_padding: int = 1, mid_block_scale_factor: float = 1, act_fn: str = "silu",
                 norm_num_groups: Optional[int] = 32, norm_eps: float = 1e-5, cross_attention_dim: int = 1280,
                 attention_head_dim: Union[int, Tuple[int]] = 8, use_linear_projection: bool = False,
                 class_embed_type: Optional[str] = None, num_class_embeds: Optional[int] = None,
                 upcast_attention: bool = False, resnet_time_scale_shift: str = "default",
                 projection_class_embeddings_input_dim: Optional[int] = None,
                 controlnet_conditioning_channel_order: str = "rgb",
                 conditioning_embedding_out_channels: Optional[Tuple[int]] = None,
                 global_pool_conditions: bool = False,
                 uncond_cam_in_dim: Tuple[int, int] = (3, 7), camera_in_dim: int = 189, camera_out_dim: int = 768,
                 map_embedder_cls: str = None, map_embedder_param: dict = None,
                 map_size: Tuple[int, int, int] = None, use_uncond_map: str = None, drop_cond_ratio: float = 0.0,
                 drop_cam_num: int = 1, drop_cam_with_box: bool = False, cam_embedder_param: Optional[Dict] = None,
                 bbox_embedder_cls: str = None, bbox_embedder_param: dict = None):
        super().__init__()
        cfg = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        self.config = _tree.AttrDict(cfg)
        if len(block_out_channels) != len(down_block_types):
            raise ValueError(
                f"Must provide the same number of `block_out_channels` as `down_block_types`. "
                f"`block_out_channels`: {block_out_channels}. `down_block_types`: {down_block_types}.")
        if (tuple(block_out_channels) != (320, 640, 1280, 1280) or attention_head_dim != 8 or layers_per_block != 2
                or norm_num_groups != 32 or use_linear_projection or class_embed_type is not None
                or num_class_embeds is not None or global_pool_conditions or in_channels != 4
                or controlnet_conditioning_channel_order != "rgb" or map_embedder_cls is None
                or use_uncond_map is not None):
            raise NotImplementedError("dualdiff_b200 implements the sd-controlnet-seg/config.json configuration family")
        ch, heads, temb, cad = tuple(block_out_channels), attention_head_dim, block_out_channels[0] * 4, cross_attention_dim
        self.cam2token = nn.Linear(camera_in_dim, camera_out_dim)
        if uncond_cam_in_dim:
            self.uncond_cam = nn.Embedding(1, uncond_cam_in_dim[0] * uncond_cam_in_dim[1])
            self.uncond_cam_num = uncond_cam_in_dim[1]
        self.drop_cond_ratio, self.drop_cam_num, self.drop_cam_with_box = drop_cond_ratio, drop_cam_num, drop_cam_with_box
        self.cam_embedder = get_embedder(**cam_embedder_param)
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.time_proj = _tree.Timesteps(ch[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = _tree.TimestepEmbedding(ch[0], temb)
        self.class_embedding = None
        self.controlnet_cond_embedding = load_module(map_embedder_cls)(
            conditioning_embedding_channels=ch[0], **map_embedder_param)
        self.uncond_map = None
        self.bbox_embedder = load_module(bbox_embedder_cls)(**bbox_embedder_param)
        self.down_blocks = nn.ModuleList()
        self.controlnet_down_blocks = nn.ModuleList([nn.Conv2d(ch[0], ch[0], 1)])
        oc = ch[0]
        for i in range(4):
            ic, oc = oc, ch[i]
            final = i == 3
            if not final:
                self.down_blocks.append(_tree.CrossAttnDownBlock2D(ic, oc, temb, 2, heads, cad, True, 32, norm_eps))
            else:
                self.down_blocks.append(_tree.DownBlock2D(ic, oc, temb, 2, False, 32, norm_eps))
            for _ in range(2 + (0 if final else 1)):
                self.controlnet_down_blocks.append(nn.Conv2d(oc, oc, 1))
        self.controlnet_mid_block = nn.Conv2d(ch[-1], ch[-1], 1)
        for m in list(self.controlnet_down_blocks) + [self.controlnet_mid_block]:  # zero_module (:230-281)
            nn.init.zeros_(m.weight); nn.init.zeros_(m.bias)
        self.mid_block = _tree.UNetMidBlock2DCrossAttn(ch[-1], temb, heads, cad, 32, norm_eps)
        # created by default, deleted by the caller when unused (:297-306, misc/test_utils.py:123-131)
        self.adm_proj = nn.Sequential(nn.Linear(768 + temb, temb), nn.SiLU(), nn.Linear(temb, temb))
        self.txt_con_fusion = txt_con_XFormersAttn()
        self.txt_con_fusionp = txt_con_XFormersAttn_plus()
        self._packed = None
        self._prep_cache = None

    # ---- reference helpers ---------------------------------------------------------------------------
    def uncond_cam_param(self, repeat_size: Union[List[int], int] = 1):
        if isinstance(repeat_size, int):
            repeat_size = [1, repeat_size]
        total = int(np.prod(repeat_size))
        param = self.uncond_cam.weight[0][None].expand(total, -1)
        return param.reshape(*repeat_size, -1, self.uncond_cam_num)

    def add_uncond_to_kwargs(self, camera_param, bboxes_3d_data, image, max_len=None, **kwargs):
        """uncond in the front, cond in the tail (reference :671-769)"""
        batch_size, n_cam = camera_param.shape[:2]
        ret = {"camera_param": torch.cat([self.uncond_cam_param([batch_size, n_cam]).to(camera_param), camera_param])}

        def _one(data):
            if data is None:
                if max_len is None:
                    return None
                dev = camera_param.device
                return {"bboxes": torch.zeros([batch_size * 2, n_cam, max_len, 8, 3], device=dev),
                        "classes": torch.zeros([batch_size * 2, n_cam, max_len], device=dev, dtype=torch.long),
                        "masks": torch.zeros([batch_size * 2, n_cam, max_len], device=dev, dtype=torch.bool)}
            out = {}
            for key in ("bboxes", "classes", "masks"):
                v = torch.cat([torch.zeros_like(data[key]), data[key]])
                if max_len is not None:
                    extra = max_len - v.shape[2]
                    assert extra >= 0
                    pad = torch.zeros_like(v[:, :, :1]).expand(-1, -1, extra, *v.shape[3:])
                    v = torch.cat([v, pad], dim=2)
                out[key] = v
            return out

        if isinstance(bboxes_3d_data, list):
            ret["bboxes_3d_data"] = [_one(d) for d in bboxes_3d_data]
        else:
            if bboxes_3d_data is None:
                logging.warning("Your 'bboxes_3d_data' should not be None.")
            ret["bboxes_3d_data"] = _one(bboxes_3d_data)
        ret["image"] = image
        ret.update(kwargs)
        return ret

    def add_uncond_to_emb(self, prompt_embeds, N_cam, encoder_hidden_states_with_cam):
        """reference :771-789.  Only reachable through `guess_mode` with classifier-free guidance
        (pipeline_bev_controlnet.py:452-456), which this path does not build; the reference's own body goes through a
        non-existent `self.controlnet` attribute and cannot run either."""
        raise NotImplementedError("add_uncond_to_emb belongs to guess_mode, which is not on the dual-branch path")

    def prepare(self, cfg, **kwargs):
        self.bbox_embedder.prepare(cfg, **kwargs)

    # ---- engine --------------------------------------------------------------------------------------
    def pack(self, device=None):
        from .. import engine
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: move the model to a CUDA device (sm_100a) before use")
        if getattr(self, "use_cam_in_temb", False):
            raise NotImplementedError("use_cam_in_temb: the reference itself asserts False here (:953-954)")
        sfa, sfap = bool(getattr(self, "use_txt_con_fusion", False)), bool(getattr(self, "use_txt_con_fusionp", False))
        if sfa == sfap:      # the reference asserts they are exclusive (:972) and adds nothing to the condition with neither
            raise NotImplementedError("exactly one of use_txt_con_fusion (configs/exp/dual_branch_augloss_fusion_8pts.yaml:44) and "
                                      "use_txt_con_fusionp (configs/exp/occ_bg_fusionp.yaml) must be set")
        if getattr(self, "use_box_adapter", False):
            raise NotImplementedError("use_box_adapter is incompatible with the dual branch (multiview_runner.py:240)")
        self._packed = engine.pack_controlnet(self.state_dict(), device, bool(self.use_occ_3d), "sfa_plus" if sfap else "sfa")
        self._packed_versions = self._param_versions()
        self._prep_cache = None
        return self

    def prepare_condition(self, camera_param, encoder_hidden_states, bboxes_3d_data, controlnet_cond, H, W):
        """timestep-invariant half of forward (tokens, K/V of every text cross-attention, condition embedding, SFA);
        the sampler calls it once per sample, `forward` calls it on demand."""
        from .. import engine
        self.ensure_packed(camera_param.device)
        return engine.controlnet_prepare(self._packed, camera_param.float(), encoder_hidden_states, bboxes_3d_data,
                                         controlnet_cond, H, W)

    def forward(self, sample: torch.FloatTensor, timestep: Union[torch.Tensor, float, int],
                camera_param: torch.Tensor, bboxes_3d_data: Dict[str, Any], encoder_hidden_states: torch.Tensor,
                controlnet_cond: torch.FloatTensor, encoder_hidden_states_uncond: torch.Tensor = None,
                conditioning_scale: float = 1.0, class_labels: Optional[torch.Tensor] = None,
                timestep_cond: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None, guess_mode: bool = False,
                return_dict: bool = True, **kwargs):
        from .. import engine
        self.use_aug_text = kwargs["use_aug_text"]  # required keyword, as in the reference (:812)
        rows = camera_param.shape[0] * (camera_param.shape[1] if self.use_aug_text else 1)
        if encoder_hidden_states.shape[0] != rows:   # per-view prompts with use_aug_text (:351-352), else one per scene
            raise ValueError(f"use_aug_text={bool(self.use_aug_text)}: expected {rows} prompt embeddings, got {encoder_hidden_states.shape[0]}")
        if guess_mode or attention_mask is not None or class_labels is not None or timestep_cond is not None:
            raise NotImplementedError("guess_mode / attention_mask / class_labels / timestep_cond are unused on the reference path")
        if not sample.is_cuda:
            raise RuntimeError("dualdiff_b200 has no CPU path: `sample` must be a CUDA tensor")
        P = self.ensure_packed(sample.device)
        b, n_cam, c, H, W = sample.shape
        prep = self.prepare_condition(camera_param, encoder_hidden_states, bboxes_3d_data, controlnet_cond, H, W)
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([float(t)], device=sample.device)
        t = t.reshape(-1).to(device=sample.device, dtype=torch.float32)
        lat = sample.reshape(b * n_cam, c, H, W)
        lat = lat.contiguous() if lat.dtype in (torch.float32, torch.bfloat16) else lat.float().contiguous()
        down, mid = engine.controlnet_forward(P, prep, lat, 1, b * n_cam, H, W, t, None, conditioning_scale)
        hw = [(H, W)] * 4
        h2, w2 = H, W
        sizes = [(H, W)] * 3
        for _ in range(3):
            h2, w2 = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
            sizes += [(h2, w2)] * 3
        sizes = sizes[:12]
        n = b * n_cam
        down_nchw = [nchw_view(d, n, *s) for d, s in zip(down, sizes)]
        mid_nchw = nchw_view(mid, n, *sizes[-1])
        enc = prep.enc_rows.reshape(n, prep.lk, 768)
        if not return_dict:
            return down_nchw, mid_nchw, enc
        return BEVControlNetOutput(down_block_res_samples=down_nchw, mid_block_res_sample=mid_nchw,
                                   encoder_hidden_states_with_cam=enc)
