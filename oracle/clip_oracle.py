"""ORACLE (test infrastructure, never on the product path): the prompt encoder of the DualDiff pipeline in fp32 on the CPU.

What it restates.  The pipeline encodes the prompt once per sample (pipeline/pipeline_bev_controlnet.py:273-281 ->
diffusers 0.17.1 `StableDiffusionControlNetPipeline._encode_prompt`): tokenizer(padding="max_length", max_length=77,
truncation=True) -> `text_encoder(input_ids, attention_mask=None)[0]` -> repeat per image -> for classifier-free guidance
the same for the negative prompt ("" when none is given) and `cat([negative, positive])`.  `text_encoder` is transformers'
`CLIPTextModel` (SD-v1.5: 12 layers, hidden 768, 12 heads, MLP 3072 with quick-GELU, 77 positions, causal mask, final
LayerNorm; `[0]` = last_hidden_state).  Neither library is vendored under /root/reference; transformers IS installed in
this image (5.5.0), so the restatement below is PINNED against the real `transformers.CLIPTextModel`:
  * tests/test_clip.py::test_oracle_matches_transformers runs both on the same random weights (max-abs diff < 2e-5),
  * tests/golden/clip_small.pt (oracle/make_golden_clip.py) holds ids / weights / output of a small CLIPTextModel produced by
    transformers itself, so the pin also holds where transformers is absent.
`encode_prompt` follows `_encode_prompt`; diffusers is absent, so that function is restated from the 0.17.1 source
(parity unpinned for the tokenizer call conventions only — the arithmetic is the pinned text model).

State-dict keys are transformers' own (`text_model.embeddings.token_embedding.weight`, `...encoder.layers.N.self_attn.
{q,k,v,out}_proj`, `layer_norm{1,2}`, `mlp.fc{1,2}`, `text_model.final_layer_norm`).
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD15 = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
            max_position_embeddings=77)


def manifest(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
             max_position_embeddings=77) -> Dict[str, Tuple[int, ...]]:
    m = {}
    C, I = hidden_size, intermediate_size
    m["text_model.embeddings.token_embedding.weight"] = (vocab_size, C)
    m["text_model.embeddings.position_embedding.weight"] = (max_position_embeddings, C)
    for i in range(num_hidden_layers):
        p = f"text_model.encoder.layers.{i}"
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            m[f"{p}.self_attn.{n}.weight"] = (C, C)
            m[f"{p}.self_attn.{n}.bias"] = (C,)
        m[f"{p}.layer_norm1.weight"] = (C,); m[f"{p}.layer_norm1.bias"] = (C,)
        m[f"{p}.mlp.fc1.weight"] = (I, C); m[f"{p}.mlp.fc1.bias"] = (I,)
        m[f"{p}.mlp.fc2.weight"] = (C, I); m[f"{p}.mlp.fc2.bias"] = (C,)
        m[f"{p}.layer_norm2.weight"] = (C,); m[f"{p}.layer_norm2.bias"] = (C,)
    m["text_model.final_layer_norm.weight"] = (C,); m["text_model.final_layer_norm.bias"] = (C,)
    return m


def n_layers(sd) -> int:
    return 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("text_model.encoder.layers."))


def quick_gelu(x):
    """transformers activations.QuickGELUActivation"""
    return x * torch.sigmoid(1.702 * x)


def text_model(sd: Dict[str, torch.Tensor], input_ids: torch.Tensor, num_heads: int = 12, eps: float = 1e-5) -> torch.Tensor:
    """CLIPTextModel(input_ids)[0]: (n, L) int64 -> (n, L, C) fp32 last_hidden_state (after final_layer_norm).
    transformers modeling_clip.py: CLIPTextEmbeddings, CLIPEncoderLayer (pre-LN, residual), CLIPAttention with the causal
    mask of CLIPTextTransformer, CLIPMLP(quick_gelu)."""
    f = lambda k: sd[k].float()
    n, L = input_ids.shape
    x = f("text_model.embeddings.token_embedding.weight")[input_ids] + f("text_model.embeddings.position_embedding.weight")[:L]
    C = x.shape[-1]
    d = C // num_heads
    mask = torch.full((L, L), float("-inf")).triu(1)            # key j visible to query i iff j <= i
    for i in range(n_layers(sd)):
        p = f"text_model.encoder.layers.{i}"
        h = F.layer_norm(x, (C,), f(p + ".layer_norm1.weight"), f(p + ".layer_norm1.bias"), eps)
        q = F.linear(h, f(p + ".self_attn.q_proj.weight"), f(p + ".self_attn.q_proj.bias"))
        k = F.linear(h, f(p + ".self_attn.k_proj.weight"), f(p + ".self_attn.k_proj.bias"))
        v = F.linear(h, f(p + ".self_attn.v_proj.weight"), f(p + ".self_attn.v_proj.bias"))
        sp = lambda t: t.reshape(n, L, num_heads, d).transpose(1, 2)
        s = sp(q) @ sp(k).transpose(-1, -2) * (d ** -0.5) + mask
        a = (s.softmax(-1) @ sp(v)).transpose(1, 2).reshape(n, L, C)
        x = x + F.linear(a, f(p + ".self_attn.out_proj.weight"), f(p + ".self_attn.out_proj.bias"))
        h = F.layer_norm(x, (C,), f(p + ".layer_norm2.weight"), f(p + ".layer_norm2.bias"), eps)
        h = quick_gelu(F.linear(h, f(p + ".mlp.fc1.weight"), f(p + ".mlp.fc1.bias")))
        x = x + F.linear(h, f(p + ".mlp.fc2.weight"), f(p + ".mlp.fc2.bias"))
    return F.layer_norm(x, (C,), f("text_model.final_layer_norm.weight"), f("text_model.final_layer_norm.bias"), eps)


def encode_prompt(sd, tokenizer, prompt, num_images_per_prompt: int = 1, do_classifier_free_guidance: bool = True,
                  negative_prompt=None, prompt_embeds: Optional[torch.Tensor] = None,
                  negative_prompt_embeds: Optional[torch.Tensor] = None, num_heads: int = 12) -> torch.Tensor:
    """diffusers 0.17.1 `_encode_prompt` (called at pipeline_bev_controlnet.py:273-281): -> (G*b*num_images, 77, C), the
    unconditional half first."""
    if prompt is not None and isinstance(prompt, str):
        prompt = [prompt]
    batch = len(prompt) if prompt is not None else prompt_embeds.shape[0]

    def enc(texts: List[str], max_length):
        ids = tokenizer(texts, padding="max_length", max_length=max_length, truncation=True, return_tensors="pt").input_ids
        return text_model(sd, ids, num_heads)

    if prompt_embeds is None:
        prompt_embeds = enc(prompt, tokenizer.model_max_length)
    b, L, C = prompt_embeds.shape
    prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, L, C)
    if do_classifier_free_guidance:
        if negative_prompt_embeds is None:
            if negative_prompt is None:
                uncond = [""] * batch
            elif isinstance(negative_prompt, str):
                uncond = [negative_prompt]
            else:
                uncond = list(negative_prompt)
            if len(uncond) != batch:
                raise ValueError(f"`negative_prompt` has batch size {len(uncond)}, but `prompt` has batch size {batch}")
            negative_prompt_embeds = enc(uncond, L)
        negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_images_per_prompt, 1).view(batch * num_images_per_prompt, L, -1)
        prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds])
    return prompt_embeds


class HashTokenizer:
    """Deterministic stand-in for CLIPTokenizer in the tests (the BPE vocabulary files cannot be fetched here): same call
    protocol (`padding="max_length"`, `max_length`, `truncation`, `return_tensors="pt"` -> `.input_ids`), CLIP's
    begin / end / pad ids (49406 / 49407 / 49407), one id per whitespace-separated word from crc32."""
    model_max_length = 77
    bos_token_id, eos_token_id, pad_token_id = 49406, 49407, 49407

    def __init__(self, vocab_size: int = 49408, model_max_length: int = 77):
        self.vocab_size, self.model_max_length = vocab_size, model_max_length
        if vocab_size < 49408:   # small test vocabularies
            self.bos_token_id, self.eos_token_id, self.pad_token_id = vocab_size - 2, vocab_size - 1, vocab_size - 1

    def __call__(self, text, padding="max_length", max_length=None, truncation=True, return_tensors="pt"):
        import zlib
        texts = [text] if isinstance(text, str) else list(text)
        L = max_length or self.model_max_length
        rows = []
        for t in texts:
            body = [zlib.crc32(w.lower().encode()) % (self.vocab_size - 2) for w in t.split()][:L - 2]
            ids = [self.bos_token_id] + body + [self.eos_token_id]
            rows.append(ids + [self.pad_token_id] * (L - len(ids)))

        class _Enc:
            pass
        e = _Enc()
        e.input_ids = torch.tensor(rows, dtype=torch.int64)
        e.attention_mask = (e.input_ids != self.pad_token_id).long()
        return e


def param_count(m: Dict[str, Tuple[int, ...]]) -> int:
    return sum(math.prod(s) for s in m.values())
