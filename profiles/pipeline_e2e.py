"""BASELINE.json configs[1] through the pipeline mirror: prompt strings -> CLIP text encoder -> 25 UniPC steps with CFG over
8 six-view 224x400 scenes -> VAE decode of the 48 images -> numpy on the host, on one B200 (synthetic weights / inputs).
Wall clock around `pipe(...)` with a synchronize on both sides, plus the phase split from CUDA events.
    python profiles/pipeline_e2e.py [scenes] [steps] > gpurun_out/pipeline_e2e.json"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from dualdiff_b200 import _lib, scheduler as SCH, synthetic as S  # noqa: E402
from dualdiff_b200.networks import AutoencoderKLDecoder, CLIPTextModel  # noqa: E402
from dualdiff_b200.pipeline_bev_controlnet import StableDiffusionBEVControlNetPipeline  # noqa: E402
from oracle import clip_oracle as CO, vae_oracle as V  # noqa: E402  (HashTokenizer / manifests only: nothing of the oracle is timed)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 25
unet, nets, _ = common.build_models()
with torch.device("meta"):
    enc, vae = CLIPTextModel(), AutoencoderKLDecoder()
enc.load_state_dict(S.init_state_dict(CO.manifest(), seed=3), strict=True, assign=True)
vae.load_state_dict(S.init_state_dict(V.manifest(), seed=4), strict=True, assign=True)
if os.environ.get("PACK_FIRST"):      # diagnostic: pack from the host copies of the weights before moving the modules
    for m in [unet] + nets:
        m.pack(torch.device("cuda", 0))
pipe = StableDiffusionBEVControlNetPipeline(vae, enc, unet, nets, SCH.UniPCMultistepScheduler(), CO.HashTokenizer()).to("cuda:0")
pipe.set_progress_bar_config(disable=True)
inp = S.make_inputs(B, 28, 50, seed=1, L_bg=28, L_fg=32)
prompts = [f"a driving scene image at boston-seaport. scene {i}, rain, congestion" for i in range(B)]
kw = dict(prompt=prompts, image=[inp["cond_bg"], inp["cond_fg"]], camera_param=inp["camera_param"], height=224, width=400,
          num_inference_steps=STEPS, guidance_scale=2.0, output_type="np",
          bev_controlnet_kwargs={"bboxes_3d_data": [inp["boxes_bg"], inp["boxes_fg"]], "use_aug_text": False})
marks = []
pipe(**dict(kw, num_inference_steps=3), generator=torch.Generator().manual_seed(0))         # warm-up (packing, lazy init)
times = []
for rep in range(3):
    torch.cuda.synchronize()
    n0 = _lib.lib().dd_launch_count()
    t0 = time.perf_counter()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    out = pipe(**kw, generator=torch.Generator().manual_seed(rep),
               callback=lambda i, t, lat: ev[0].record() if i == 0 else (ev[1].record() if i == STEPS - 1 else None))
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    times.append((wall, ev[0].elapsed_time(ev[1]) * STEPS / max(STEPS - 1, 1), _lib.lib().dd_launch_count() - n0))
times.sort()
wall, loop_ms, launches = times[1]
img = out.images
print(json.dumps({
    "what": "StableDiffusionBEVControlNetPipeline.__call__: prompts -> CLIP -> UniPC+CFG loop -> VAE decode -> host numpy",
    "scenes": B, "steps": STEPS, "images": [int(x) for x in img.shape], "wall_s_median_of_3": round(wall, 4),
    "loop_ms": round(loop_ms, 2), "outside_loop_ms": round(wall * 1e3 - loop_ms, 2),
    "scene_steps_per_s_whole_call": round(B * STEPS / wall, 2), "scenes_per_s": round(B / wall, 3),
    "gpu_launches_counted": int(launches), "finite": bool((img == img).all()), "range": [float(img.min()), float(img.max())]}))
