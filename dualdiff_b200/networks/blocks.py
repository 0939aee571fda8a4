"""BasicMultiviewTransformerBlock — parameter layout + standalone CUDA forward.

Reference: networks/blocks.py:35-238.  The block adds `norm4`, `attn4` and a zero-initialised `connector`
to diffusers' BasicTransformerBlock and, between the text cross-attention and the feed-forward, lets every
camera view attend to its two ring neighbours (`neighboring_attn_type="add"`: one attention per neighbour, outputs summed;
"concat": the neighbours' tokens under one softmax; "self": all views of the scene).
"""
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _tree


def _ensure_int_pairs(view_pair):
    return {int(k): [int(x) for x in v] for k, v in view_pair.items()}


class BasicMultiviewTransformerBlock(_tree.BasicTransformerBlock):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, dropout=0.0,
                 cross_attention_dim: Optional[int] = None, activation_fn: str = "geglu",
                 num_embeds_ada_norm: Optional[int] = None, attention_bias: bool = False,
                 only_cross_attention: bool = False, double_self_attention: bool = False,
                 upcast_attention: bool = False, norm_elementwise_affine: bool = True,
                 norm_type: str = "layer_norm", final_dropout: bool = False,
                 neighboring_view_pair: Optional[Dict[int, List[int]]] = None,
                 neighboring_attn_type: Optional[str] = "add", zero_module_type="zero_linear",
                 temporal_frames: int = 0):
        if (activation_fn != "geglu" or num_embeds_ada_norm is not None or attention_bias or only_cross_attention
                or double_self_attention or norm_type != "layer_norm" or not norm_elementwise_affine):
            raise NotImplementedError("dualdiff_b200 implements the SDv1.5 block configuration the reference uses")
        if neighboring_attn_type not in ("add", "concat", "self"):
            raise NotImplementedError(f"Unknown type: {neighboring_attn_type}")       # blocks.py:139-140
        if zero_module_type != "zero_linear":
            raise TypeError(f"Unknown zero module type: {zero_module_type}")
        super().__init__(dim, num_attention_heads, attention_head_dim, cross_attention_dim)
        self.neighboring_view_pair = _ensure_int_pairs(neighboring_view_pair)
        self.neighboring_attn_type = neighboring_attn_type
        self.norm4 = nn.LayerNorm(dim)
        self.attn4 = _tree.Attention(dim, dim, num_attention_heads, attention_head_dim)
        self.connector = nn.Linear(dim, dim)
        nn.init.zeros_(self.connector.weight)  # zero_module (blocks.py:83)
        nn.init.zeros_(self.connector.bias)
        # video configuration (BASELINE config 5; no reference code -- csrc/dd_temporal.cu defines the block):
        # temporal attention over the `temporal_frames` frames of a clip, zero-initialised output projection
        self.temporal_frames = int(temporal_frames)
        if self.temporal_frames > 1:
            self.norm_temp = nn.LayerNorm(dim)
            self.attn_temp = _tree.Attention(dim, dim, num_attention_heads, attention_head_dim)
            nn.init.zeros_(self.attn_temp.to_out[0].weight)
            nn.init.zeros_(self.attn_temp.to_out[0].bias)
        self._packed = None

    @property
    def new_module(self):
        return {"norm4": self.norm4, "attn4": self.attn4, "connector": self.connector}

    @property
    def n_cam(self):
        return len(self.neighboring_view_pair)

    def pack(self):
        from .. import engine
        pairs, n_src, concat, n_bias = engine.xview_plan(self.neighboring_view_pair, self.neighboring_attn_type)
        pk = engine.Packer({k: v for k, v in self.state_dict().items()}, next(self.parameters()).device, n_bias)
        pk.sd = {"b." + k: v for k, v in pk.sd.items()}
        pk.tblock("b", True)
        pk.out["n_nbr"] = n_src
        pk.out["xview_concat"] = concat
        self._packed = pk.out
        return self

    def _apply(self, fn, *args, **kwargs):      # .to() / .cuda(): the packed copy follows the parameters
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None, frame_shard=None):
        """hidden_states (b*n_cam, T, C), encoder_hidden_states (b*n_cam, Lk, 768) -> (b*n_cam, T, C).
        With temporal_frames > 1 the batch is ordered (clip, frame, view); `frame_shard` (sharding.FrameShard) marks
        the frames of a clip as split across ranks (this rank then holds temporal_frames / world of them)."""
        from .. import engine, ops
        if attention_mask is not None or encoder_attention_mask is not None:
            raise NotImplementedError("attention masks are not used on the reference path (attention_mask=None)")
        if self._packed is None:
            self.pack()
        P = self._packed
        n, T, C = hidden_states.shape
        h = hidden_states.to(torch.bfloat16).reshape(n * T, C).contiguous()
        enc = encoder_hidden_states.to(torch.bfloat16)
        lk = enc.shape[1]
        ctx = engine.StepCtx(n=n, temb=None, temb_rows_per_img_factor=1, lk=lk,
                             kv_map=engine.make_kv_map(n, self.neighboring_view_pair, h.device, self.neighboring_attn_type),
                             n_nbr=P["n_nbr"], xview_concat=P["xview_concat"])
        if self.temporal_frames > 1:
            ctx.n_view = self.n_cam
            ctx.frame_shard = frame_shard
            ctx.n_frames = self.temporal_frames // (frame_shard.world if frame_shard is not None else 1)
        ctx.text_kv["b.attn2"] = engine.text_kv(P, "b.attn2", enc.reshape(n * lk, enc.shape[2]).contiguous())
        out = engine.transformer_block(P, "b", h, n, T, ctx, True)
        return out.reshape(n, T, C).to(hidden_states.dtype)
