"""CLIPTextModel — the prompt encoder in front of the sampler loop, on the B200 kernels.

Reference call site: pipeline/pipeline_bev_controlnet.py:273-281 (`self._encode_prompt(...)`, inherited from diffusers'
StableDiffusionControlNetPipeline: `self.text_encoder(text_input_ids, attention_mask=None)[0]`); the network is
transformers' `CLIPTextModel` (SD-v1.5 text encoder).  This mirror keeps its name, its state-dict keys
(`text_model.embeddings.{token,position}_embedding`, `text_model.encoder.layers.N.{self_attn.{q,k,v,out}_proj,
layer_norm1, layer_norm2, mlp.fc1, mlp.fc2}`, `text_model.final_layer_norm`), `forward(input_ids, attention_mask=None)`
and the `last_hidden_state` / `pooler_output` / `[0]` result protocol, so `from_pretrained` weights load unchanged.

Per layer: dd_layernorm -> ONE fused QKV GEMM (bias in the epilogue) -> dd_seq_attention (causal, 12 heads of 64, addressed
in place in the fused projection) -> out-projection GEMM with bias + residual fused -> dd_layernorm -> fc1 GEMM ->
dd_quick_gelu -> fc2 GEMM with bias + residual fused; embedding gather and the final LayerNorm around it.  (The GEMMs
are a few tiles each -- 77 tokens per prompt -- so they take the plain tile schedule, stream_k=-1.)  The residual
stream is bf16.  Key-padding masks (`attention_mask`) are not used by the SD-v1.5 text encoder config and raise.
"""
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn as nn

from .output_cls import _Output

BF = torch.bfloat16


@dataclass
class BaseModelOutputWithPooling(_Output):
    last_hidden_state: torch.Tensor
    pooler_output: Optional[torch.Tensor] = None


class CLIPTextConfig:
    def __init__(self, vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                 num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                 eos_token_id=49407, **unused):
        if hidden_act != "quick_gelu":
            raise NotImplementedError(f"hidden_act={hidden_act!r}: only the SD-v1.5 text encoder's quick_gelu is built")
        if hidden_size != 64 * num_attention_heads:
            raise NotImplementedError("dd_seq_attention is built for head_dim 64 (hidden_size = 64 * num_attention_heads)")
        self.vocab_size, self.hidden_size, self.intermediate_size = vocab_size, hidden_size, intermediate_size
        self.num_hidden_layers, self.num_attention_heads = num_hidden_layers, num_attention_heads
        self.max_position_embeddings, self.hidden_act, self.layer_norm_eps = max_position_embeddings, hidden_act, layer_norm_eps
        self.eos_token_id = eos_token_id
        self.use_attention_mask = False


def _manifest(cfg: CLIPTextConfig) -> Dict[str, tuple]:
    C, I = cfg.hidden_size, cfg.intermediate_size
    m = {"text_model.embeddings.token_embedding.weight": (cfg.vocab_size, C),
         "text_model.embeddings.position_embedding.weight": (cfg.max_position_embeddings, C)}

    def lin(p, ci, co):
        m[p + ".weight"] = (co, ci); m[p + ".bias"] = (co,)

    def vec(p):
        m[p + ".weight"] = (C,); m[p + ".bias"] = (C,)

    for i in range(cfg.num_hidden_layers):
        p = f"text_model.encoder.layers.{i}"
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(f"{p}.self_attn.{n}", C, C)
        vec(p + ".layer_norm1")
        lin(p + ".mlp.fc1", C, I); lin(p + ".mlp.fc2", I, C)
        vec(p + ".layer_norm2")
    vec("text_model.final_layer_norm")
    return m


class _Node(nn.Module):
    """bare container: gives the flat transformers key list a module tree (state_dict / load_state_dict / .to work)"""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container only")


class CLIPTextModel(nn.Module):
    def __init__(self, config: Optional[CLIPTextConfig] = None, **kwargs):
        super().__init__()
        self.config = config if config is not None else CLIPTextConfig(**kwargs)
        for key, shape in _manifest(self.config).items():
            node = self
            *path, leaf = key.split(".")
            for part in path:
                if not hasattr(node, part):
                    node.add_module(part, _Node())
                node = getattr(node, part)
            node.register_parameter(leaf, nn.Parameter(torch.empty(shape)))
        self._packed = None

    # ---- checkpoints (transformers layout: config.json + model.safetensors | pytorch_model.bin) -------------------
    config_name = "config.json"
    weights_names = ("model.safetensors", "pytorch_model.bin")

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **kwargs):
        """load `<sd-v1-5>/text_encoder` the way `pipe_cls.from_pretrained` does for the reference (misc/test_utils.py:
        150-156).  Keys outside the text model (`position_ids` buffers of older transformers) are ignored; a missing or
        mis-shaped text-model tensor raises."""
        import json
        import os
        root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        with open(os.path.join(root, cls.config_name)) as fh:
            cfg = {k: v for k, v in json.load(fh).items() if not k.startswith("_")}
        model = cls(CLIPTextConfig(**cfg))
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
        own = model.state_dict()
        missing = [k for k in own if k not in sd or tuple(sd[k].shape) != tuple(own[k].shape)]
        if missing:
            raise KeyError(f"{root}: text-model tensors missing or mis-shaped: {missing[:4]}{' ...' if len(missing) > 4 else ''}")
        model.load_state_dict({k: sd[k].float() for k in own}, strict=True)
        return model.eval()

    def save_pretrained(self, save_directory, safe_serialization=True, **kwargs):
        import json
        import os
        os.makedirs(save_directory, exist_ok=True)
        c = self.config
        cfg = dict(architectures=["CLIPTextModel"], model_type="clip_text_model", vocab_size=c.vocab_size,
                   hidden_size=c.hidden_size, intermediate_size=c.intermediate_size, num_hidden_layers=c.num_hidden_layers,
                   num_attention_heads=c.num_attention_heads, max_position_embeddings=c.max_position_embeddings,
                   hidden_act=c.hidden_act, layer_norm_eps=c.layer_norm_eps, eos_token_id=c.eos_token_id)
        with open(os.path.join(save_directory, self.config_name), "w") as fh:
            json.dump(cfg, fh, indent=2)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, self.weights_names[0]))
        else:
            torch.save(sd, os.path.join(save_directory, self.weights_names[1]))

    @property
    def dtype(self):
        return BF

    @property
    def device(self):
        return next(self.parameters()).device

    # ---- packing ----------------------------------------------------------------------------------------
    def pack(self, device=None):
        from .. import engine
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: move the text encoder to a CUDA device before use")
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        pk = engine.Packer(sd, device)
        f = lambda k: sd[k].float()
        pk.put("tok", pk.f32(sd["text_model.embeddings.token_embedding.weight"]))
        pk.put("pos", pk.f32(sd["text_model.embeddings.position_embedding.weight"]))
        for i in range(self.config.num_hidden_layers):
            p = f"text_model.encoder.layers.{i}"
            a = p + ".self_attn"
            pk.put(p + ".qkv.w", torch.cat([f(a + ".q_proj.weight"), f(a + ".k_proj.weight"), f(a + ".v_proj.weight")], 0).to(BF))
            pk.put(p + ".qkv.b", pk.f32(torch.cat([f(a + ".q_proj.bias"), f(a + ".k_proj.bias"), f(a + ".v_proj.bias")], 0)))
            pk.lin(a + ".out_proj")
            pk.lin(p + ".mlp.fc1"); pk.lin(p + ".mlp.fc2")
            pk.norm(p + ".layer_norm1"); pk.norm(p + ".layer_norm2")
        pk.norm("text_model.final_layer_norm")
        self._packed = pk.out
        return self

    # ---- forward ----------------------------------------------------------------------------------------
    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, position_ids=None,
                output_attentions=None, output_hidden_states=None, return_dict: bool = True):
        from .. import ops
        if attention_mask is not None:
            raise NotImplementedError("attention_mask: the SD-v1.5 text encoder runs without a key-padding mask "
                                      "(`use_attention_mask` unset; diffusers passes None)")
        if position_ids is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("position_ids / output_attentions / output_hidden_states are not built")
        cfg = self.config
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be (batch, sequence)")
        n, L = input_ids.shape
        if L > cfg.max_position_embeddings:
            raise ValueError(f"sequence length {L} exceeds max_position_embeddings {cfg.max_position_embeddings}")
        if not input_ids.is_cuda:   # ids come from the tokenizer on the host: validate there, then copy
            if input_ids.numel() and (int(input_ids.min()) < 0 or int(input_ids.max()) >= cfg.vocab_size):
                raise IndexError(f"token id outside [0, {cfg.vocab_size})")
            dev = self.device
            if dev.type != "cuda":
                raise RuntimeError("dualdiff_b200 has no CPU path: move the text encoder to a CUDA device before use")
            input_ids = input_ids.to(dev)
        if self._packed is None:
            self.pack(input_ids.device)
        P = self._packed
        ids = input_ids.to(torch.int64).contiguous()
        C, H = cfg.hidden_size, cfg.num_attention_heads
        eps = cfg.layer_norm_eps
        x = ops.clip_embed(ids, P["tok"], P["pos"])
        for i in range(cfg.num_hidden_layers):
            p = f"text_model.encoder.layers.{i}"
            h = ops.layernorm(x, P[p + ".layer_norm1.g"], P[p + ".layer_norm1.b"], eps)
            qkv = ops.gemm(h, P[p + ".qkv.w"], bias=P[p + ".qkv.b"], stream_k=-1)
            a = ops.seq_attention(qkv, qkv, qkv, n_seq=n, seq_len=L, heads=H, head_dim=64, causal=True,
                                  q_col0=0, k_col0=C, v_col0=2 * C)
            x = ops.gemm(a, P[p + ".self_attn.out_proj.w"], bias=P[p + ".self_attn.out_proj.b"], res1=x, stream_k=-1)
            h = ops.layernorm(x, P[p + ".layer_norm2.g"], P[p + ".layer_norm2.b"], eps)
            h = ops.gemm(h, P[p + ".mlp.fc1.w"], bias=P[p + ".mlp.fc1.b"], stream_k=-1)
            ops.quick_gelu(h, out=h)
            x = ops.gemm(h, P[p + ".mlp.fc2.w"], bias=P[p + ".mlp.fc2.b"], res1=x, stream_k=-1)
        x = ops.layernorm(x, P["text_model.final_layer_norm.g"], P["text_model.final_layer_norm.b"], eps)
        last = x.view(n, L, C)
        # pooled output = the end-of-text position (transformers: first eos id; legacy configs: the largest id)
        if cfg.eos_token_id == 2:
            pos = ids.argmax(dim=-1)
        else:
            pos = (ids == cfg.eos_token_id).int().argmax(dim=-1)
        pooled = last[torch.arange(n, device=last.device), pos]
        if not return_dict:
            return (last, pooled)
        return BaseModelOutputWithPooling(last_hidden_state=last, pooler_output=pooled)
