"""BasicTransformerBlock / FeedForward / GEGLU / AdaLayerNorm (diffusers 0.17.1 models/attention.py semantics;
`_args` is the MagicDrive fork's addition read at networks/unet_2d_condition_multiview.py:226)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .attention_processor import Attention  # noqa: F401  (re-exported: reference imports it from here too)


class AdaLayerNorm(nn.Module):
    def __init__(self, embedding_dim, num_embeddings):
        super().__init__()
        self.emb = nn.Embedding(num_embeddings, embedding_dim)
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, embedding_dim * 2)
        self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False)

    def forward(self, x, timestep):
        emb = self.linear(self.silu(self.emb(timestep)))
        scale, shift = torch.chunk(emb, 2)
        return self.norm(x) * (1 + scale) + shift


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        assert activation_fn == "geglu"
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out)])
        if final_dropout:
            self.net.append(nn.Dropout(dropout))

    def forward(self, hidden_states):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", num_embeds_ada_norm=None, attention_bias=False,
                 only_cross_attention=False, double_self_attention=False, upcast_attention=False,
                 norm_elementwise_affine=True, norm_type="layer_norm", final_dropout=False):
        super().__init__()
        self._args = dict(dim=dim, num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                          dropout=dropout, cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                          num_embeds_ada_norm=num_embeds_ada_norm, attention_bias=attention_bias,
                          only_cross_attention=only_cross_attention, double_self_attention=double_self_attention,
                          upcast_attention=upcast_attention, norm_elementwise_affine=norm_elementwise_affine,
                          norm_type=norm_type, final_dropout=final_dropout)
        self.only_cross_attention = only_cross_attention
        self.use_ada_layer_norm_zero = (num_embeds_ada_norm is not None) and norm_type == "ada_norm_zero"
        self.use_ada_layer_norm = (num_embeds_ada_norm is not None) and norm_type == "ada_norm"
        assert not self.use_ada_layer_norm_zero
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim,
                               dropout=dropout, bias=attention_bias,
                               cross_attention_dim=cross_attention_dim if only_cross_attention else None,
                               upcast_attention=upcast_attention)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout)
        if cross_attention_dim is not None or double_self_attention:
            self.attn2 = Attention(query_dim=dim,
                                   cross_attention_dim=cross_attention_dim if not double_self_attention else None,
                                   heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                                   bias=attention_bias, upcast_attention=upcast_attention)
        else:
            self.attn2 = None
        mk = (lambda: AdaLayerNorm(dim, num_embeds_ada_norm)) if self.use_ada_layer_norm else \
            (lambda: nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine))
        self.norm1 = mk()
        self.norm2 = mk() if self.attn2 is not None else None
        self.norm3 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None):
        norm_hidden_states = self.norm1(hidden_states, timestep) if self.use_ada_layer_norm else self.norm1(hidden_states)
        cross_attention_kwargs = cross_attention_kwargs if cross_attention_kwargs is not None else {}
        attn_output = self.attn1(norm_hidden_states,
                                 encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
                                 attention_mask=attention_mask, **cross_attention_kwargs)
        hidden_states = attn_output + hidden_states
        if self.attn2 is not None:
            norm_hidden_states = self.norm2(hidden_states, timestep) if self.use_ada_layer_norm else self.norm2(hidden_states)
            attn_output = self.attn2(norm_hidden_states, encoder_hidden_states=encoder_hidden_states,
                                     attention_mask=encoder_attention_mask, **cross_attention_kwargs)
            hidden_states = attn_output + hidden_states
        norm_hidden_states = self.norm3(hidden_states)
        hidden_states = self.ff(norm_hidden_states) + hidden_states
        return hidden_states
