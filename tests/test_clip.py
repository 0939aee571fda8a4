"""Prompt encoder: the oracle restatement is pinned against the real transformers.CLIPTextModel and
against a golden produced by transformers; the CUDA path (GEMM + LayerNorm + dd_clip_embed / dd_seq_attention / dd_quick_gelu)
is checked against the oracle on the GPU.  Tolerance (bf16 kernels and bf16 residual stream vs fp32 oracle, stated):
cosine >= 0.999 and rel-L2 <= 2e-2 on last_hidden_state; dd_seq_attention alone: max-abs <= 2e-2 * max|ref|."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "clip_small.pt")


def _small():
    from dualdiff_b200 import synthetic as S
    from oracle import clip_oracle as CO
    g = torch.load(GOLD)
    return g, S.init_state_dict(CO.manifest(**g["config"]), seed=g["seed"])


def test_oracle_matches_transformers_golden():
    from oracle import clip_oracle as CO
    g, sd = _small()
    with torch.no_grad():
        out = CO.text_model(sd, g["ids"], num_heads=g["config"]["num_attention_heads"])
    assert (out - g["last_hidden_state"]).abs().max() < 2e-5


def test_oracle_matches_transformers_live():
    """the SD-v1.5 text encoder shape (4 of its 12 layers to keep the CPU suite short) against transformers itself"""
    tr = pytest.importorskip("transformers")
    from dualdiff_b200 import synthetic as S
    from oracle import clip_oracle as CO
    cfgd = dict(CO.SD15, num_hidden_layers=4)
    cfg = tr.CLIPTextConfig(**cfgd, hidden_act="quick_gelu", layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=49406,
                            eos_token_id=49407)
    model = tr.CLIPTextModel(cfg).eval()
    sd = S.init_state_dict(CO.manifest(**cfgd), seed=3)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all("position_ids" in k for k in res.missing_keys)
    ids = CO.HashTokenizer()(["a driving scene image at boston-seaport. rain, congestion", ""]).input_ids
    assert ids.shape == (2, 77) and ids[1, 1] == 49407
    with torch.no_grad():
        ref = model(ids)[0]
        out = CO.text_model(sd, ids)
    assert (out - ref).abs().max() < 2e-5


def test_manifest_is_the_sd15_text_encoder():
    from dualdiff_b200.networks import CLIPTextModel
    from oracle import clip_oracle as CO
    m = CO.manifest()
    assert CO.param_count(m) == 123_060_480          # openai/clip-vit-large-patch14 text model = SD-v1.5 text_encoder
    with torch.device("meta"):
        enc = CLIPTextModel()
    assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == m


def test_encode_prompt_layout():
    """_encode_prompt: negative ("" by default) half first, repeat per image"""
    from oracle import clip_oracle as CO
    g, sd = _small()
    tok = CO.HashTokenizer(vocab_size=100, model_max_length=16)
    with torch.no_grad():
        e = CO.encode_prompt(sd, tok, ["a b c", "d e"], num_images_per_prompt=2, num_heads=2)
        neg = CO.text_model(sd, tok([""]).input_ids, 2)
        pos = CO.text_model(sd, tok(["d e"]).input_ids, 2)
    assert e.shape == (8, 16, 128)
    assert torch.allclose(e[0], neg[0], atol=1e-6) and torch.allclose(e[3], neg[0], atol=1e-6)
    assert torch.allclose(e[6], pos[0], atol=1e-6) and torch.allclose(e[7], pos[0], atol=1e-6)
    with pytest.raises(ValueError):
        CO.encode_prompt(sd, tok, ["a", "b"], negative_prompt=["x"], num_heads=2)


def test_no_cpu_path():
    from dualdiff_b200.networks import CLIPTextModel
    g, sd = _small()
    enc = CLIPTextModel(**g["config"], eos_token_id=99)
    enc.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(g["ids"])
    with pytest.raises(IndexError):
        enc(torch.full((1, 16), 100))
    with pytest.raises(NotImplementedError):
        enc(g["ids"], attention_mask=torch.ones(3, 16))


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n_seq,L,heads,causal", [(3, 77, 12, True), (2, 128, 2, True), (5, 16, 2, False), (1, 1, 1, True)])
def test_seq_attention_matches_torch(n_seq, L, heads, causal):
    from dualdiff_b200 import ops
    g = torch.Generator().manual_seed(n_seq * 1000 + L)
    C = heads * 64
    qkv = (torch.randn(n_seq * L, 3 * C, generator=g) * 1.5).to(torch.bfloat16)
    out = ops.seq_attention(qkv.cuda(), qkv.cuda(), qkv.cuda(), n_seq=n_seq, seq_len=L, heads=heads, head_dim=64,
                            causal=causal, q_col0=0, k_col0=C, v_col0=2 * C).float().cpu()
    f = qkv.float().reshape(n_seq, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(f[0], f[1], f[2], is_causal=causal)
    ref = ref.permute(0, 2, 1, 3).reshape(n_seq * L, C)
    assert (out - ref).abs().max() <= 2e-2 * ref.abs().max()


@pytest.mark.gpu
def test_clip_embed_and_quick_gelu():
    from dualdiff_b200 import ops
    g = torch.Generator().manual_seed(0)
    tok, pos = torch.randn(50, 128, generator=g), torch.randn(16, 128, generator=g)
    ids = torch.randint(0, 50, (3, 16), generator=g)
    out = ops.clip_embed(ids.cuda(), tok.cuda(), pos.cuda()).float().cpu()
    ref = (tok[ids] + pos[None]).reshape(48, 128)
    assert (out - ref.to(torch.bfloat16).float()).abs().max() == 0
    bad = ids.clone(); bad[1, 2] = 50
    assert torch.isnan(ops.clip_embed(bad.cuda(), tok.cuda(), pos.cuda()).float().cpu()[16 + 2]).all()
    x = (torch.randn(64, 256, generator=g) * 3).to(torch.bfloat16)
    y = ops.quick_gelu(x.cuda()).float().cpu()
    ref = x.float() * torch.sigmoid(1.702 * x.float())
    assert (y - ref).abs().max() <= 2e-2 * ref.abs().max()


@pytest.mark.gpu
def test_clip_small_matches_transformers_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from dualdiff_b200.networks import CLIPTextModel
    g, sd = _small()
    enc = CLIPTextModel(**g["config"], eos_token_id=99)
    enc.load_state_dict(sd, strict=True)
    out = enc.to("cuda:0")(g["ids"])          # host ids, as the tokenizer produces them
    m = common.metrics(out[0].float().cpu(), g["last_hidden_state"])
    assert m["cos"] > 0.999 and m["rel_l2"] < 2e-2, m
    eos = (g["ids"] == 99).int().argmax(-1)
    assert torch.equal(out.pooler_output.cpu(), out.last_hidden_state.cpu()[torch.arange(3), eos])


@pytest.mark.gpu
def test_clip_sd15_matches_oracle():
    """the full SD-v1.5 text encoder (12 layers, 77 tokens) on 4 prompts, CUDA path vs the pinned oracle"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from dualdiff_b200 import _lib, synthetic as S
    from dualdiff_b200.networks import CLIPTextModel
    from oracle import clip_oracle as CO
    sd = S.init_state_dict(CO.manifest(), seed=3)
    with torch.device("meta"):
        enc = CLIPTextModel()
    enc.load_state_dict(sd, strict=True, assign=True)
    ids = CO.HashTokenizer()(["a driving scene image at boston-seaport. rain, congestion", "",
                              "night, difficult lighting, parked cars on the right", "x " * 100]).input_ids
    with torch.no_grad():
        ref = CO.text_model(sd, ids)
    n0 = _lib.lib().dd_launch_count()
    out = enc.to("cuda:0")(ids)[0].float().cpu()
    assert _lib.lib().dd_launch_count() - n0 == 2 + 12 * 8      # embed, 12 x (LN, QKV, attention, out, LN, fc1, gelu, fc2), final LN
    m = common.metrics(out, ref)
    assert m["cos"] > 0.999 and m["rel_l2"] < 2e-2, m
