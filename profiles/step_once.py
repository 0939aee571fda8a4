"""exactly one 8-scene CFG denoising step between cudaProfilerStart/Stop (for ncu --profile-from-start off).
EAGER=1 profiles the eager launch sequence instead of the CUDA-graph replay."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from dualdiff_b200 import synthetic as S
from dualdiff_b200.pipeline import DualDiffDenoiser
B = int(os.environ.get("SCENES", "8"))
dev = torch.device("cuda:0")
unet, nets, _ = common.build_models()
for m in [unet] + nets:
    m.pack(dev)
inp = common.to_dev(S.make_inputs(B, 28, 50, seed=1, L_bg=28, L_fg=32), dev)
den = DualDiffDenoiser(unet, nets, guidance_scale=2.0, use_cuda_graph=os.environ.get("EAGER", "0") != "1")
den.prepare(inp["latents"], inp["prompt_embeds"], inp["camera_param"], [inp["boxes_bg"], inp["boxes_fg"]],
            [inp["cond_bg"], inp["cond_fg"]], num_inference_steps=6)
for i in range(3):
    den.step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
den.step(3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per step:", getattr(den, "launches_per_step", None))
