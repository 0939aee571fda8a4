"""Attention + processors (diffusers 0.17.1 models/attention_processor.py semantics)."""
from typing import Optional, Union

import torch
import torch.nn as nn


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, cross_attention_norm=None,
                 cross_attention_norm_num_groups=32, added_kv_proj_dim=None, norm_num_groups=None,
                 spatial_norm_dim=None, out_bias=True, scale_qk=True, only_cross_attention=False, eps=1e-5,
                 rescale_output_factor=1.0, residual_connection=False, processor=None):
        super().__init__()
        inner_dim = dim_head * heads
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.sliceable_head_dim = heads
        self.added_kv_proj_dim = added_kv_proj_dim
        self.only_cross_attention = only_cross_attention
        self.group_norm = None
        self.spatial_norm = None
        self.norm_cross = None
        assert cross_attention_norm is None and added_kv_proj_dim is None and norm_num_groups is None
        self.to_q = nn.Linear(query_dim, inner_dim, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else AttnProcessor())

    def set_use_memory_efficient_attention_xformers(self, *a, **k):
        return None

    def set_processor(self, processor):
        if hasattr(self, "processor") and isinstance(self.processor, nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)

    def batch_to_head_dim(self, tensor):
        h = self.heads
        b, s, d = tensor.shape
        tensor = tensor.reshape(b // h, h, s, d)
        return tensor.permute(0, 2, 1, 3).reshape(b // h, s, d * h)

    def head_to_batch_dim(self, tensor, out_dim=3):
        h = self.heads
        b, s, d = tensor.shape
        tensor = tensor.reshape(b, s, h, d // h).permute(0, 2, 1, 3)
        if out_dim == 3:
            tensor = tensor.reshape(b * h, s, d // h)
        return tensor

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        if attention_mask is None:
            scores = self.scale * torch.bmm(query, key.transpose(-1, -2))
        else:
            scores = torch.baddbmm(attention_mask, query, key.transpose(-1, -2), beta=1, alpha=self.scale)
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size=None, out_dim=3):
        if attention_mask is None:
            return attention_mask
        head_size = self.heads
        if attention_mask.shape[-1] != target_length:
            attention_mask = torch.nn.functional.pad(attention_mask, (0, target_length), value=0.0)
        if out_dim == 3:
            if attention_mask.shape[0] < batch_size * head_size:
                attention_mask = attention_mask.repeat_interleave(head_size, dim=0)
        elif out_dim == 4:
            attention_mask = attention_mask.unsqueeze(1).repeat_interleave(head_size, dim=1)
        return attention_mask

    def norm_encoder_hidden_states(self, encoder_hidden_states):
        return encoder_hidden_states


class AttnProcessor:
    def __call__(self, attn: Attention, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b, c, h, w = hidden_states.shape
            hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
        batch_size, sequence_length, _ = (
            hidden_states.shape if encoder_hidden_states is None else encoder_hidden_states.shape)
        attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        query = attn.head_to_batch_dim(query)
        key = attn.head_to_batch_dim(key)
        value = attn.head_to_batch_dim(value)
        # batch-chunked so the dense score tensor stays below ~1 GiB at HD latents (each batch item is independent:
        # the result is identical to the one-shot evaluation)
        step = max(1, (1 << 28) // max(1, query.shape[1] * key.shape[1]))
        outs = []
        for i in range(0, query.shape[0], step):
            m = None if attention_mask is None else attention_mask[i:i + step]
            probs = attn.get_attention_scores(query[i:i + step], key[i:i + step], m)
            outs.append(torch.bmm(probs, value[i:i + step]))
        hidden_states = outs[0] if len(outs) == 1 else torch.cat(outs)
        hidden_states = attn.batch_to_head_dim(hidden_states)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        if input_ndim == 4:
            hidden_states = hidden_states.transpose(-1, -2).reshape(b, c, h, w)
        if attn.residual_connection:
            hidden_states = hidden_states + residual
        return hidden_states / attn.rescale_output_factor


class XFormersAttnProcessor(AttnProcessor):
    """same arithmetic as AttnProcessor (xformers' kernel computes softmax(QK^T*scale)V)."""

    def __init__(self, attention_op=None):
        self.attention_op = attention_op


AttentionProcessor = Union[AttnProcessor, XFormersAttnProcessor]
