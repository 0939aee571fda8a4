"""VAE decode of the bench batch: 48 latents of 28x50 (8 six-view scenes) -> 48 images of 224x400, CUDA events."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualdiff_b200 import _lib, synthetic as S
from dualdiff_b200.networks import AutoencoderKLDecoder
from oracle import vae_oracle as V
n = int(os.environ.get("N_IMG", "48"))
with torch.device("meta"):
    vae = AutoencoderKLDecoder()
vae.load_state_dict(S.init_state_dict(V.manifest(), seed=4), strict=True, assign=True)
vae = vae.to("cuda:0")
z = torch.randn(n, 4, 28, 50, device="cuda")
for _ in range(2):
    img = vae.decode_latents(z * 0.18215)
torch.cuda.synchronize()
n0 = _lib.lib().dd_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); img = vae.decode_latents(z * 0.18215); e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
# conv FLOPs of the decoder per image at latent h x w (2 * Cin * Cout * 9 * H * W per 3x3 conv)
def flops(h, w):
    f, hw = 0.0, h * w
    c3 = lambda ci, co, px: 2.0 * ci * co * 9 * px
    f += c3(4, 512, hw) + 4 * c3(512, 512, hw) + 4 * 2.0 * 512 * 512 * hw + 4.0 * hw * hw * 512      # conv_in, mid resnets, attn
    chans, prev, px = [512, 512, 256, 128], 512, hw
    for i, co in enumerate(chans):
        for j in range(3):
            ci = prev if j == 0 else co
            f += c3(ci, co, px) + c3(co, co, px) + (2.0 * ci * co * px if ci != co else 0)
        if i < 3:
            px *= 4
            f += c3(co, co, px)
        prev = co
    return f + c3(128, 3, px)
tf = flops(28, 50) * n / 1e12
print(f"VAE decode {n} x (28x50 -> 224x400): {ms:.1f} ms, {_lib.lib().dd_launch_count() - n0} launches, {tf:.1f} TFLOP -> {tf / ms * 1e3:.0f} TFLOP/s; "
      f"image range [{img.min().item():.3f}, {img.max().item():.3f}], finite={bool(torch.isfinite(img).all())}")
