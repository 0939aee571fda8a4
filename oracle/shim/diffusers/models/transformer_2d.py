"""Transformer2DModel, continuous-input / conv-projection path (diffusers 0.17.1 semantics: GroupNorm eps 1e-6,
proj_in 1x1 conv, NCHW -> (n, HW, C), blocks, proj_out 1x1 conv, + residual)."""
from dataclasses import dataclass

import torch
import torch.nn as nn

from ..utils import BaseOutput
from .attention import BasicTransformerBlock


@dataclass
class Transformer2DModelOutput(BaseOutput):
    sample: torch.FloatTensor


class Transformer2DModel(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, num_layers=1, dropout=0.0,
                 norm_num_groups=32, cross_attention_dim=None, attention_bias=False, activation_fn="geglu",
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                 norm_type="layer_norm", norm_elementwise_affine=True, **unused):
        super().__init__()
        self.use_linear_projection = use_linear_projection
        inner_dim = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner_dim)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                  attention_bias=attention_bias, only_cross_attention=only_cross_attention,
                                  upcast_attention=upcast_attention, norm_type=norm_type,
                                  norm_elementwise_affine=norm_elementwise_affine)
            for _ in range(num_layers)])
        if use_linear_projection:
            self.proj_out = nn.Linear(inner_dim, in_channels)
        else:
            self.proj_out = nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None, return_dict=True):
        batch, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        else:
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
            hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, attention_mask=attention_mask,
                                  encoder_hidden_states=encoder_hidden_states,
                                  encoder_attention_mask=encoder_attention_mask, timestep=timestep,
                                  cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states)
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        output = hidden_states + residual
        if not return_dict:
            return (output,)
        return Transformer2DModelOutput(sample=output)
