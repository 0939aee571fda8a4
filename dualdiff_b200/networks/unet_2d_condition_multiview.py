"""UNet2DConditionModelMultiview — SDv1.5-shaped UNet whose transformer blocks are multi-view blocks.

Reference: networks/unet_2d_condition_multiview.py:44-527 (constructor kwargs :124-180, forward :327-339).
Same class name, constructor keywords, forward signature, `.config`, state-dict keys; the forward runs on
the B200 engine (dualdiff_b200.engine.unet_forward) and fails loudly without the CUDA library.
"""
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _tree
from .blocks import BasicMultiviewTransformerBlock
from .output_cls import UNet2DConditionOutput


class UNet2DConditionModelMultiview(_tree.ModelBase):
    _supports_gradient_checkpointing = True

    def __init__(self, sample_size: Optional[int] = None, in_channels: int = 4, out_channels: int = 4,
                 center_input_sample: bool = False, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 down_block_types: Tuple[str] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                                 "CrossAttnDownBlock2D", "DownBlock2D"),
                 mid_block_type: Optional[str] = "UNetMidBlock2DCrossAttn",
                 up_block_types: Tuple[str] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
                 only_cross_attention: Union[bool, Tuple[bool]] = False,
                 block_out_channels: Tuple[int] = (320, 640, 1280, 1280), layers_per_block: Union[int, Tuple[int]] = 2,
                 downsample_padding: int = 1, mid_block_scale_factor: float = 1, act_fn: str = "silu",
                 norm_num_groups: Optional[int] = 32, norm_eps: float = 1e-5,
                 cross_attention_dim: Union[int, Tuple[int]] = 1280, encoder_hid_dim: Optional[int] = None,
                 encoder_hid_dim_type: Optional[str] = None, attention_head_dim: Union[int, Tuple[int]] = 8,
                 dual_cross_attention: bool = False, use_linear_projection: bool = False,
                 class_embed_type: Optional[str] = None, addition_embed_type: Optional[str] = None,
                 num_class_embeds: Optional[int] = None, upcast_attention: bool = False,
                 resnet_time_scale_shift: str = "default", resnet_skip_time_act: bool = False,
                 resnet_out_scale_factor: int = 1.0, time_embedding_type: str = "positional",
                 time_embedding_dim: Optional[int] = None, time_embedding_act_fn: Optional[str] = None,
                 timestep_post_act: Optional[str] = None, time_cond_proj_dim: Optional[int] = None,
                 conv_in_kernel: int = 3, conv_out_kernel: int = 3,
                 projection_class_embeddings_input_dim: Optional[int] = None, class_embeddings_concat: bool = False,
                 mid_block_only_cross_attention: Optional[bool] = None, cross_attention_norm: Optional[str] = None,
                 addition_embed_type_num_heads=64,
                 trainable_state="only_new", neighboring_view_pair: Optional[dict] = None,
                 neighboring_attn_type: str = "add", zero_module_type: str = "zero_linear",
                 crossview_attn_type: str = "basic", img_size: Optional[Tuple[int, int]] = None,
                 temporal_frames: int = 0):
        super().__init__()
        cfg = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        self.config = _tree.AttrDict(cfg)
        unsupported = (tuple(down_block_types) != ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",)
                       or tuple(up_block_types) != ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3
                       or mid_block_type != "UNetMidBlock2DCrossAttn" or layers_per_block != 2
                       or tuple(block_out_channels) != (320, 640, 1280, 1280) or attention_head_dim != 8
                       or norm_num_groups != 32 or use_linear_projection or class_embed_type is not None
                       or center_input_sample or not flip_sin_to_cos or freq_shift != 0 or in_channels != 4
                       or out_channels != 4 or act_fn != "silu" or crossview_attn_type != "basic"
                       or not isinstance(cross_attention_dim, int))
        if unsupported:
            raise NotImplementedError("dualdiff_b200 implements the SDv1.5-shaped configuration of the reference "
                                      "(configs/model/SDv1.5mv_rawbox.yaml); other UNet shapes are out of scope")
        if crossview_attn_type != "basic":
            raise TypeError(f"Unknown attn type: {crossview_attn_type}")
        ch, heads, temb, cad = tuple(block_out_channels), attention_head_dim, block_out_channels[0] * 4, cross_attention_dim
        tf = dict(block_cls=BasicMultiviewTransformerBlock,
                  block_kwargs=dict(neighboring_view_pair=neighboring_view_pair,
                                    neighboring_attn_type=neighboring_attn_type, zero_module_type=zero_module_type,
                                    temporal_frames=temporal_frames))
        # video configuration (BASELINE config 5, no reference code): every multi-view block gets a temporal attention
        # over the `temporal_frames` frames of a clip; the batch is then ordered (clip, frame, view).  `frame_shard`
        # (sharding.FrameShard) marks the frames of a clip as split across ranks.
        self.temporal_frames = int(temporal_frames)
        self.frame_shard = None
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.time_proj = _tree.Timesteps(ch[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = _tree.TimestepEmbedding(ch[0], temb)
        self.class_embedding = None
        self.encoder_hid_proj = None
        self.time_embed_act = None
        self.down_blocks = nn.ModuleList()
        oc = ch[0]
        for i in range(4):
            ic, oc = oc, ch[i]
            if i < 3:
                self.down_blocks.append(_tree.CrossAttnDownBlock2D(ic, oc, temb, 2, heads, cad, True, 32, norm_eps, dict(tf)))
            else:
                self.down_blocks.append(_tree.DownBlock2D(ic, oc, temb, 2, False, 32, norm_eps))
        self.mid_block = _tree.UNetMidBlock2DCrossAttn(ch[-1], temb, heads, cad, 32, norm_eps, dict(tf))
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        oc = rev[0]
        self.num_upsamplers = 0
        for i in range(4):
            prev, oc = oc, rev[i]
            ic = rev[min(i + 1, 3)]
            add_up = i < 3
            self.num_upsamplers += int(add_up)
            if i == 0:
                self.up_blocks.append(_tree.UpBlock2D(ic, oc, prev, temb, 3, add_up, 32, norm_eps))
            else:
                self.up_blocks.append(_tree.CrossAttnUpBlock2D(ic, oc, prev, temb, 3, heads, cad, add_up, 32, norm_eps, dict(tf)))
        self.conv_norm_out = nn.GroupNorm(32, ch[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)
        self.crossview_attn_type = crossview_attn_type
        self.img_size = [int(s) for s in img_size] if img_size is not None else None
        self.trainable_state = trainable_state
        self._new_module = {}
        self.n_cam = len(neighboring_view_pair) if neighboring_view_pair is not None else 6
        self._packed = None

    # ---- reference API surface ---------------------------------------------------------------------
    @property
    def trainable_module(self) -> Dict[str, nn.Module]:
        if self.trainable_state == "all":
            return {self.__class__: self}
        if self.trainable_state == "only_new":
            return self._new_module
        raise ValueError(f"Unknown trainable_state: {self.trainable_state}")

    @property
    def trainable_parameters(self) -> List[nn.Parameter]:
        return [p for m in self.trainable_module.values() for p in m.parameters()]

    @classmethod
    def from_unet_2d_condition(cls, unet, load_weights_from_unet: bool = True, **kwargs):
        m = cls(**unet.config, **kwargs)
        if load_weights_from_unet:
            m.load_state_dict(unet.state_dict(), strict=False)
        return m

    # ---- engine ------------------------------------------------------------------------------------
    def pack(self, device=None):
        """build the kernel-layout weights (call again after load_state_dict)"""
        from .. import engine
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: move the model to a CUDA device (sm_100a) before use")
        self._packed = engine.pack_unet(self.state_dict(), device, self.config.neighboring_view_pair,
                                        self.config.neighboring_attn_type)
        self._packed_versions = self._param_versions()
        return self

    def video_ctx(self, ctx):
        """fill the clip geometry of a StepCtx (no-op for the single-frame configuration)"""
        if self.temporal_frames > 1:
            world = self.frame_shard.world if self.frame_shard is not None else 1
            ctx.n_frames, ctx.n_view, ctx.frame_shard = self.temporal_frames // world, self.n_cam, self.frame_shard
            if ctx.n % (ctx.n_frames * ctx.n_view) != 0:
                raise ValueError(f"batch {ctx.n} is not a multiple of frames x views = {ctx.n_frames} x {ctx.n_view}")
        return ctx

    def release_master_weights(self):
        """drop the fp32 diffusers-layout parameters once packed (inference-only deployments)"""
        for p in self.parameters():
            p.data = torch.empty(0, device=p.device, dtype=p.dtype)

    def forward(self, sample: torch.FloatTensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor, class_labels: Optional[torch.Tensor] = None,
                timestep_cond: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None,
                down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None, return_dict: bool = True):
        from .. import engine, ops
        if class_labels is not None or timestep_cond is not None or attention_mask is not None:
            raise NotImplementedError("class_labels / timestep_cond / attention_mask are unused on the reference path")
        if not sample.is_cuda:
            raise RuntimeError("dualdiff_b200 has no CPU path: `sample` must be a CUDA tensor")
        P = self.ensure_packed(sample.device)
        n, c, H, W = sample.shape
        if n % self.n_cam != 0:
            raise ValueError(f"batch {n} is not a multiple of n_cam={self.n_cam} ('(b n) ...' layout, blocks.py:196)")
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([float(t)], device=sample.device, dtype=torch.float32)
        t = t.reshape(-1).to(device=sample.device, dtype=torch.float32)
        if t.numel() not in (1, n):
            raise ValueError("timestep must be a scalar or have one entry per image")
        temb = engine.time_embedding(P, t)
        enc = encoder_hidden_states.to(torch.bfloat16)
        lk = enc.shape[1]
        enc_rows = enc.reshape(n * lk, enc.shape[2]).contiguous()
        ctx = engine.StepCtx(n=n, temb=temb, temb_rows_per_img_factor=n // t.numel(), lk=lk,
                             kv_map=engine.make_kv_map(n, P["view_pairs"], sample.device, P["xview_mode"]), n_nbr=P["n_nbr"],
                             xview_concat=P["xview_mode"] != "add",
                             text_kv=engine.prepare_text(P, engine.ATTN2_LAYERS_UNET, enc_rows))
        self.video_ctx(ctx)
        lat = sample.contiguous() if sample.dtype in (torch.float32, torch.bfloat16) else sample.float().contiguous()
        down = mid = None
        if down_block_additional_residuals is not None:
            down = [rows_of(r) for r in down_block_additional_residuals]
        if mid_block_additional_residual is not None:
            mid = rows_of(mid_block_additional_residual)
        eps_rows = engine.unet_forward(P, lat, 1, n, H, W, ctx, down, mid)
        out = ops.rows_to_nchw(eps_rows, n, (H, W), out_dtype=torch.float32).to(sample.dtype)
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(sample=out)


def rows_of(t: torch.Tensor) -> torch.Tensor:
    """NCHW tensor -> bf16 channels-last rows.  Zero-copy when `t` is the channels_last view the ControlNet
    mirror returns (physically [n, H, W, C]); otherwise one layout-conversion kernel."""
    from .. import ops
    n, c, h, w = t.shape
    if t.dtype == torch.bfloat16 and t.stride() == (h * w * c, 1, w * c, c):
        return t.permute(0, 2, 3, 1).reshape(n * h * w, c)
    t = t if t.dtype in (torch.float32, torch.bfloat16) else t.float()
    return ops.nchw_to_rows(t.contiguous())


def nchw_view(rows: torch.Tensor, n, h, w) -> torch.Tensor:
    """bf16 rows [n*h*w, C] -> logical NCHW tensor sharing the storage (torch channels_last strides)"""
    c = rows.shape[1]
    return rows.reshape(n, h, w, c).permute(0, 3, 1, 2)
