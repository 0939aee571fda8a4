"""UNet2DConditionModel constructor + forward (diffusers 0.17.1 models/unet_2d_condition.py semantics;
only the SDv1.5 configuration family the reference uses is supported — anything else asserts)."""
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from ..configuration_utils import ConfigMixin, register_to_config
from ..utils import BaseOutput
from .embeddings import TimestepEmbedding, Timesteps
from .modeling_utils import ModelMixin
from .unet_2d_blocks import UNetMidBlock2DCrossAttn, get_down_block, get_up_block


@dataclass
class UNet2DConditionOutput(BaseOutput):
    sample: torch.FloatTensor


class UNet2DConditionModel(ModelMixin, ConfigMixin):
    _supports_gradient_checkpointing = True

    @register_to_config
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                 mid_block_type="UNetMidBlock2DCrossAttn",
                 up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5,
                 cross_attention_dim=1280, encoder_hid_dim=None, encoder_hid_dim_type=None, attention_head_dim=8,
                 dual_cross_attention=False, use_linear_projection=False, class_embed_type=None,
                 addition_embed_type=None, num_class_embeds=None, upcast_attention=False,
                 resnet_time_scale_shift="default", resnet_skip_time_act=False, resnet_out_scale_factor=1.0,
                 time_embedding_type="positional", time_embedding_dim=None, time_embedding_act_fn=None,
                 timestep_post_act=None, time_cond_proj_dim=None, conv_in_kernel=3, conv_out_kernel=3,
                 projection_class_embeddings_input_dim=None, class_embeddings_concat=False,
                 mid_block_only_cross_attention=None, cross_attention_norm=None, addition_embed_type_num_heads=64):
        super().__init__()
        self.sample_size = sample_size
        assert time_embedding_type == "positional" and class_embed_type is None and num_class_embeds is None
        assert encoder_hid_dim is None and addition_embed_type is None and time_embedding_act_fn is None
        assert mid_block_type == "UNetMidBlock2DCrossAttn" and not dual_cross_attention
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], kernel_size=conv_in_kernel,
                                 padding=(conv_in_kernel - 1) // 2)
        time_embed_dim = time_embedding_dim or block_out_channels[0] * 4
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim, act_fn=act_fn,
                                                post_act_fn=timestep_post_act, cond_proj_dim=time_cond_proj_dim)
        self.encoder_hid_proj = None
        self.class_embedding = None
        self.time_embed_act = None
        n = len(down_block_types)
        if isinstance(only_cross_attention, bool):
            if mid_block_only_cross_attention is None:
                mid_block_only_cross_attention = only_cross_attention
            only_cross_attention = [only_cross_attention] * n
        if isinstance(attention_head_dim, int):
            attention_head_dim = (attention_head_dim,) * n
        if isinstance(cross_attention_dim, int):
            cross_attention_dim = (cross_attention_dim,) * n
        if isinstance(layers_per_block, int):
            layers_per_block = [layers_per_block] * n

        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        output_channel = block_out_channels[0]
        for i, t in enumerate(down_block_types):
            input_channel, output_channel = output_channel, block_out_channels[i]
            is_final = i == len(block_out_channels) - 1
            self.down_blocks.append(get_down_block(
                t, num_layers=layers_per_block[i], in_channels=input_channel, out_channels=output_channel,
                temb_channels=time_embed_dim, add_downsample=not is_final, resnet_eps=norm_eps,
                resnet_act_fn=act_fn, resnet_groups=norm_num_groups, cross_attention_dim=cross_attention_dim[i],
                attn_num_head_channels=attention_head_dim[i], downsample_padding=downsample_padding,
                use_linear_projection=use_linear_projection, only_cross_attention=only_cross_attention[i],
                upcast_attention=upcast_attention, resnet_time_scale_shift=resnet_time_scale_shift))
        self.mid_block = UNetMidBlock2DCrossAttn(
            in_channels=block_out_channels[-1], temb_channels=time_embed_dim, resnet_eps=norm_eps,
            resnet_act_fn=act_fn, output_scale_factor=mid_block_scale_factor,
            resnet_time_scale_shift=resnet_time_scale_shift, cross_attention_dim=cross_attention_dim[-1],
            attn_num_head_channels=attention_head_dim[-1], resnet_groups=norm_num_groups,
            use_linear_projection=use_linear_projection, upcast_attention=upcast_attention)
        self.num_upsamplers = 0
        rev_ch = list(reversed(block_out_channels))
        rev_hd = list(reversed(attention_head_dim))
        rev_lpb = list(reversed(layers_per_block))
        rev_cad = list(reversed(cross_attention_dim))
        rev_oca = list(reversed(only_cross_attention))
        output_channel = rev_ch[0]
        for i, t in enumerate(up_block_types):
            is_final = i == len(block_out_channels) - 1
            prev_output_channel, output_channel = output_channel, rev_ch[i]
            input_channel = rev_ch[min(i + 1, len(block_out_channels) - 1)]
            add_upsample = not is_final
            if add_upsample:
                self.num_upsamplers += 1
            self.up_blocks.append(get_up_block(
                t, num_layers=rev_lpb[i] + 1, in_channels=input_channel, out_channels=output_channel,
                prev_output_channel=prev_output_channel, temb_channels=time_embed_dim, add_upsample=add_upsample,
                resnet_eps=norm_eps, resnet_act_fn=act_fn, resnet_groups=norm_num_groups,
                cross_attention_dim=rev_cad[i], attn_num_head_channels=rev_hd[i],
                use_linear_projection=use_linear_projection, only_cross_attention=rev_oca[i],
                upcast_attention=upcast_attention, resnet_time_scale_shift=resnet_time_scale_shift))
        if norm_num_groups is not None:
            self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=norm_num_groups, eps=norm_eps)
            self.conv_act = nn.SiLU()
        else:
            self.conv_norm_out = None
            self.conv_act = None
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, kernel_size=conv_out_kernel,
                                  padding=(conv_out_kernel - 1) // 2)

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict=True):
        default_overall_up_factor = 2 ** self.num_upsamplers
        forward_upsample_size = any(s % default_overall_up_factor != 0 for s in sample.shape[-2:])
        upsample_size = None
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
        elif len(timesteps.shape) == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype=self.dtype), timestep_cond)
        sample = self.conv_in(sample)
        down_block_res_samples = (sample,)
        for blk in self.down_blocks:
            if getattr(blk, "has_cross_attention", False):
                sample, res = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                                  attention_mask=attention_mask, cross_attention_kwargs=cross_attention_kwargs)
            else:
                sample, res = blk(hidden_states=sample, temb=emb)
            down_block_res_samples += res
        if down_block_additional_residuals is not None:
            down_block_res_samples = tuple(a + b for a, b in zip(down_block_res_samples, down_block_additional_residuals))
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states,
                                attention_mask=attention_mask, cross_attention_kwargs=cross_attention_kwargs)
        if mid_block_additional_residual is not None:
            sample = sample + mid_block_additional_residual
        for i, blk in enumerate(self.up_blocks):
            is_final = i == len(self.up_blocks) - 1
            res = down_block_res_samples[-len(blk.resnets):]
            down_block_res_samples = down_block_res_samples[:-len(blk.resnets)]
            if not is_final and forward_upsample_size:
                upsample_size = down_block_res_samples[-1].shape[2:]
            if getattr(blk, "has_cross_attention", False):
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                             encoder_hidden_states=encoder_hidden_states,
                             cross_attention_kwargs=cross_attention_kwargs, upsample_size=upsample_size,
                             attention_mask=attention_mask)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res, upsample_size=upsample_size)
        if self.conv_norm_out:
            sample = self.conv_act(self.conv_norm_out(sample))
        sample = self.conv_out(sample)
        if not return_dict:
            return (sample,)
        return UNet2DConditionOutput(sample=sample)
