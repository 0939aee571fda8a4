"""Writes tests/golden/clip_small.pt: ids, weights and last_hidden_state of a small `transformers.CLIPTextModel`
(2 layers, hidden 128, 2 heads of 64, MLP 256, 16 positions, vocab 100, quick_gelu) computed by transformers itself in
this container (transformers 5.5.0).  tests/test_clip.py checks oracle/clip_oracle.py against it, and the GPU test checks
the CUDA path against the same output.  Run:  python oracle/make_golden_clip.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SMALL = dict(vocab_size=100, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
             max_position_embeddings=16)


def main():
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel
    from dualdiff_b200 import synthetic as S
    from oracle import clip_oracle as CO
    cfg = CLIPTextConfig(**SMALL, hidden_act="quick_gelu", layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=98, eos_token_id=99)
    model = CLIPTextModel(cfg).eval()
    sd = S.init_state_dict(CO.manifest(**SMALL), seed=11)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 98, (3, 16), generator=g)
    ids[:, 0] = 98
    ids[0, 9:] = 99
    ids[1, 15] = 99
    ids[2, 4:] = 99
    with torch.no_grad():
        out = model(ids).last_hidden_state
    path = os.path.join(ROOT, "tests", "golden", "clip_small.pt")
    torch.save({"config": SMALL, "seed": 11, "ids": ids, "last_hidden_state": out.clone(),
                "producer": f"transformers {transformers.__version__} CLIPTextModel"}, path)
    print("wrote", path, tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
