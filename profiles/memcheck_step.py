"""One sampler step (B scenes x 6 views, latent h x w, CFG, both branches + UNet + fused scheduler kernel, eager launches) for
`compute-sanitizer --tool memcheck`: every kernel family of the hot path runs once; with EXTRAS=1 also the CLIP text encoder
and a VAE decode.  Defaults: 1 scene at 8x12 (ragged shapes).
    compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log python profiles/memcheck_step.py [B h w]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from dualdiff_b200 import synthetic as S  # noqa: E402
from dualdiff_b200.pipeline import DualDiffDenoiser  # noqa: E402

t0 = time.time()
unet, nets, _ = common.build_models()
dev = torch.device("cuda:0")
B, h, w = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 8, 12)
full = (h, w) == (28, 50)
inp = common.to_dev(S.make_inputs(B, h, w, seed=5, L_bg=28 if full else 7, L_fg=32 if full else 3), dev)
print(f"[{time.time() - t0:.1f}s] models built", file=sys.stderr, flush=True)
den = DualDiffDenoiser(unet, nets, guidance_scale=2.0, use_cuda_graph=False)
den.parallel_branches = full          # the production step runs the two branches on side streams
den.prepare(inp["latents"], inp["prompt_embeds"], inp["camera_param"], [inp["boxes_bg"], inp["boxes_fg"]],
            [inp["cond_bg"], inp["cond_fg"]], num_inference_steps=4)
torch.cuda.synchronize()
print(f"[{time.time() - t0:.1f}s] prepared", file=sys.stderr, flush=True)
den.step(0)
torch.cuda.synchronize()
print(f"[{time.time() - t0:.1f}s] step done, finite={bool(torch.isfinite(den.latents).all())}", file=sys.stderr, flush=True)
if os.environ.get("EXTRAS"):
    from dualdiff_b200.networks import AutoencoderKLDecoder, CLIPTextModel
    from oracle import clip_oracle as CO, vae_oracle as V   # manifests / tokenizer stand-in only
    with torch.device("meta"):
        enc, vae = CLIPTextModel(), AutoencoderKLDecoder()
    enc.load_state_dict(S.init_state_dict(CO.manifest(), seed=3), strict=True, assign=True)
    out = enc.to(dev)(CO.HashTokenizer()(["a driving scene image at boston-seaport. rain", "", "night"]).input_ids)[0]
    torch.cuda.synchronize()
    print(f"[{time.time() - t0:.1f}s] CLIP done, finite={bool(torch.isfinite(out.float()).all())}", file=sys.stderr, flush=True)
    vae.load_state_dict(S.init_state_dict(V.manifest(), seed=4), strict=True, assign=True)
    img = vae.to(dev).decode_latents(torch.randn(2, 4, 4, 6, device=dev))
    torch.cuda.synchronize()
    print(f"[{time.time() - t0:.1f}s] VAE done, finite={bool(torch.isfinite(img).all())}", file=sys.stderr, flush=True)
